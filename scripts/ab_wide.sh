python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for sc in 6 7 8; do
echo -n "scalar ctas $sc: "
DEMCMC_PK_SCALAR_CTAS=$sc python bench.py --steps 300 --warmup 5 --no-ess 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', d['value'], d['roofline']['frac'], d['e2e']['value'])"
done
python scripts/pk_timeline.py gpurun_out/pk_tl.csv 2>&1 | tail -9
DEMCMC_PK_SCALAR_CTAS=7 python scripts/pk_timeline.py gpurun_out/pk_tl7.csv 2>&1 | tail -9
python scripts/bench_configs.py c1 c5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['particle_updates_per_s']), d['ms_per_iteration'])"
