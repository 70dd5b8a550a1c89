python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for sh in 0 1; do
  echo -n "xd_short $sh: "
  DEMCMC_XD_SHORT=$sh python scripts/bench_configs.py c4 --iters 60 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['particle_updates_per_s']), d['ms_per_iteration'])"
done
DEMCMC_LANES=1 python scripts/c4_timeline.py gpurun_out/c4_timeline_xs.csv 2>&1 | tail -10
