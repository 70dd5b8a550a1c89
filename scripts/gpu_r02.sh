#!/bin/bash
# Round-2 evidence run on one B200: smoke, GPU tests, bench (both arms), launch list, ncu --set full of the persistent
# kernel and of the LBA kernel.  Usage (repo root, GPU box): bash scripts/gpu_r02.sh TAG
TAG=${1:-r02}
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
echo "== bench (driver arguments)"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_steps20.json 2> gpurun_out/${TAG}_bench_steps20.err; echo "rc=$?"
echo "== bench (defaults)"; timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$?"
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "rc=$?"
[ -n "$NO_NCU" ] && { ls -la gpurun_out/ | tail -12; exit 0; }
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 40 --warmup 3 --no-cpu --no-ess --no-configs > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full k_chunk_persist"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chunk_persist -s 4 -c 2 -o gpurun_out/${TAG}_pk -f python bench.py --steps 40 --warmup 3 --no-cpu --no-ess --no-configs > gpurun_out/${TAG}_ncu_pk.log 2>&1; echo "rc=$?"
echo "== ncu full LBA"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ll_pointwise -s 10 -c 3 -o gpurun_out/${TAG}_lba -f python scripts/bench_configs.py c3 --iters 6 > gpurun_out/${TAG}_ncu_lba.log 2>&1; echo "rc=$?"
ls -la gpurun_out/ | tail -20
