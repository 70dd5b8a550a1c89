python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for ns in 1 0; do
echo -n "NO_SMALL=$ns: "
DEMCMC_NO_SMALL=$ns python scripts/bench_configs.py c1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['particle_updates_per_s']), d['ms_per_iteration'], d['kernel_launches'])"
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
