#!/bin/bash
# usage: scripts/regcheck.sh differentialevolutionmcmc.jl_b200/csrc/kernels.cu
# prints the spills and the highest register index of k_chunk_persist<false>: ptxas honours the raised
# budget of the setmaxnreg region (232 for the DMMA warps) only for some code shapes -- check after every change
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -Xptxas -v -cubin -o /tmp/k.cubin -I/root/repo/differentialevolutionmcmc.jl_b200/csrc $1 2>&1 | grep -A2 "k_chunk_persistILb0" | grep -E "spill|Used"
cuobjdump -sass /tmp/k.cubin | python3 -c "
import sys,re
lines=sys.stdin.read().split('\n')
st=[i for i,l in enumerate(lines) if 'Function :' in l and 'k_chunk_persistILb0' in l][0]
en=[i for i,l in enumerate(lines) if 'Function :' in l and i>st]
en=en[0] if en else len(lines)
mx=0;ldl=0
for l in lines[st:en]:
    for m in re.finditer(r'\bR(\d+)\b', l): mx=max(mx,int(m.group(1)))
    if re.search(r'\b(LDL|STL)',l): ldl+=1
print('max reg',mx,'LDL/STL',ldl)
"
