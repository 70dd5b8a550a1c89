python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for sp in 0 1; do
echo -n "PK_SPLIT=$sp: "
DEMCMC_PK_SPLIT=$sp python bench.py --steps 300 --warmup 5 --no-ess --no-configs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', d['value'], d['roofline']['frac'], d['e2e']['value'])"
done
python scripts/pk_timeline.py gpurun_out/pk_tl_split.csv 2>&1 | tail -9
