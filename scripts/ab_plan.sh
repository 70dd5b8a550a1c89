python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for np in 1 0; do
  echo "== DEMCMC_NO_PLAN=$np"
  DEMCMC_NO_PLAN=$np python scripts/bench_configs.py c4 c5 c1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['particle_updates_per_s']), d['ms_per_iteration'])"
  DEMCMC_NO_PLAN=$np python bench.py --steps 200 --warmup 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', d['value'], d['roofline']['frac'], d['e2e']['value'])"
done
