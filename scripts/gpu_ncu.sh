#!/bin/bash
# ncu --set full of a kernel (regex) inside a short bench run.  Usage: gpu_ncu.sh TAG REGEX [skip] [count]
TAG=$1; RE=$2; SKIP=${3:-30}; CNT=${4:-8}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c $CNT -o gpurun_out/${TAG} -f python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu.log 2>&1 ; echo "ncu rc=$?"
ls -la gpurun_out/${TAG}*
