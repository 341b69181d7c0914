#!/bin/bash
# preconditioner / Krylov variants on the 1440x720 O-grid (ONE B200), pattern level 3
tag=${1:-r02j}
mkdir -p gpurun_out
export PB_NJ=720 PB_TILE=16x12 PB_TAG=$tag PB_LVL=3
PB_CFGS='[{"pcBlockCells":0,"coarseAggregates":1000,"kspType":"idrs","idrS":8},
 {"pcBlockCells":0,"coarseAggregates":1000,"kspType":"idrs","idrS":8,"pcStorage":"fp32"},
 {"pcBlockCells":0,"coarseAggregates":2000,"kspType":"idrs","idrS":8,"pcStorage":"fp32"},
 {"pcBlockCells":192,"coarseAggregates":1000,"kspType":"idrs","idrS":8,"pcStorage":"fp32"},
 {"pcBlockCells":0,"coarseAggregates":1000,"kspType":"idrs","idrS":16,"pcStorage":"fp32"},
 {"pcBlockCells":0,"coarseAggregates":1000,"kspType":"gmres","pcStorage":"fp32"}]' timeout 1200 python scripts/pc_bench.py 2>&1 | grep -v "^\[dab200\] pcSymbolic" | tail -12
