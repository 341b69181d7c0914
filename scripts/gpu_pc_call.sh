#!/bin/bash
# split (per-row + per-cell) against per-cell triangular solves: solve time and launch shares
tag=${1:-r02q}
mkdir -p gpurun_out
export PB_NJ=720 PB_TILE=16x12 PB_TAG=$tag PB_LVL=3
PB_CFGS='[{"coarseAggregates":2000,"kspType":"idrs","idrS":8,"pcStorage":"fp32","pcSplitTri":0},
 {"coarseAggregates":2000,"kspType":"idrs","idrS":8,"pcStorage":"fp32","pcSplitTri":1}]' timeout 1200 python scripts/pc_bench.py 2>&1 | grep -v "^\[dab200\] pcSymbolic" | tail -3
IP_SPLIT=1 bash scripts/gpu_iter_profile.sh ${tag}s1 2>&1 | tail -12 | head -9
IP_SPLIT=0 bash scripts/gpu_iter_profile.sh ${tag}s0 2>&1 | tail -12 | head -6
rm -f gpurun_out/*.ncu-rep
