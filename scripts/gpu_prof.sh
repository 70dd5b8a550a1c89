#!/bin/bash
# Full GPU test suite + one `ncu --set full` capture of the likelihood kernel.
TAG=${1:-r01}
mkdir -p gpurun_out
echo "== pytest -m gpu (all)" ; timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -25 gpurun_out/${TAG}_pytest_gpu.log
echo "== ncu full k_xdot"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_xdot -s 20 -c 12 -o gpurun_out/${TAG}_ssd -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1 ; echo "ncu rc=$?"
ls -la gpurun_out/
