python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for sh in 35; do
  echo -n "shape $sh: "
  DEMCMC_WIDE_SHAPE=$sh python scripts/bench_configs.py c4 c5 --iters 60 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['particle_updates_per_s']), d['ms_per_iteration'])"
done
python bench.py --steps 200 --warmup 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', d['value'], d['roofline']['frac'], d['e2e']['value'])"
python scripts/pk_timeline.py gpurun_out/pk_tl.csv 2>&1 | tail -9
