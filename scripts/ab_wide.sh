python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for st in "--steps 20 --warmup 3" ""; do
python bench.py $st 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', d['steps'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['seconds_by_part'], d['e2e']['seconds'])"
done
DEMCMC_PINNED_OUT=0 python bench.py --steps 20 --warmup 3 --no-ess 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 pageable out', d['steps'], d['value'], d['e2e']['value'], d['e2e']['seconds_by_part'], d['e2e']['seconds'])"
