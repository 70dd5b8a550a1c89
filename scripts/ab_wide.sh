for sh in 35 36 37; do
  echo -n "shape $sh: "
  DEMCMC_WIDE_SHAPE=$sh python scripts/bench_configs.py c4 --iters 60 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['particle_updates_per_s']), d['ms_per_iteration'])"
done
