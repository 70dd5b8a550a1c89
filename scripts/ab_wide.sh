python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python scripts/bench_configs.py c4 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['particle_updates_per_s']), d['ms_per_iteration'], d['roofline'])"
python scripts/c4_timeline.py gpurun_out/c4_timeline_final.csv > gpurun_out/c4_timeline_final.txt 2>&1; tail -11 gpurun_out/c4_timeline_final.txt
