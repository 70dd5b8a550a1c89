for gr in 1 2; do for rep in 1 2; do
echo -n "CHUNK_GROWTH=$gr: "
DEMCMC_CHUNK_GROWTH=$gr python bench.py --steps 20 --warmup 5 --no-ess --no-configs --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', d['value'], d['roofline']['frac'], d['e2e']['value'])"
done; done
