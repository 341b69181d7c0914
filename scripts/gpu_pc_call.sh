#!/bin/bash
# interleaved triangular solves against row-after-row; then the whole -m gpu tier
tag=${1:-r02t}
mkdir -p gpurun_out
export PB_NJ=720 PB_TILE=16x12 PB_TAG=$tag PB_LVL=3
PB_CFGS='[{"coarseAggregates":2000,"kspType":"idrs","idrS":8,"pcStorage":"fp32","pcTriInterleave":1},
 {"coarseAggregates":2000,"kspType":"idrs","idrS":8,"pcStorage":"fp32","pcTriInterleave":0}]' timeout 1200 python scripts/pc_bench.py 2>&1 | grep -v "^\[dab200\] pcSymbolic" | tail -3
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${tag}_pytest_gpu.log
