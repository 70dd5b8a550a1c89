#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, a short bench, and the ncu launch list.
# Usage (from the repo root on the GPU box): bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/${TAG}_smoke.log
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -15 gpurun_out/${TAG}_pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ; echo "bench rc=$?" ; cat gpurun_out/${TAG}_bench.json ; tail -5 gpurun_out/${TAG}_bench.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1 ; echo "ncu rc=$?"
tail -3 gpurun_out/${TAG}_launches.csv
