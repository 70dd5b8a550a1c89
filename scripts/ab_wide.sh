python -m pytest tests -x -q -m gpu -k "lba or pointwise or fused or device_math" 2>&1 | tail -2
for r in 0 1; do
echo -n "LBA_REGC=$r: "
DEMCMC_LBA_REGC=$r python scripts/bench_configs.py c3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['particle_updates_per_s']), d['ms_per_iteration'])"
done
