#!/bin/bash
# preconditioner / Krylov variants on the 1440x720 O-grid (ONE B200)
tag=${1:-r02g}
mkdir -p gpurun_out
export PB_NJ=720 PB_TILE=16x12 PB_TAG=$tag PB_LVL=2
PB_CFGS='[{"pcBlockCells":192,"coarseAggregates":1000,"kspType":"gmres"},
 {"pcBlockCells":192,"coarseAggregates":1000,"kspType":"idrs","idrS":8},
 {"pcBlockCells":0,"coarseAggregates":1000,"kspType":"idrs","idrS":8},
 {"pcBlockCells":192,"coarseAggregates":4000,"kspType":"idrs","idrS":8},
 {"pcBlockCells":192,"coarseAggregates":0,"kspType":"idrs","idrS":8},
 {"pcBlockCells":768,"coarseAggregates":1000,"kspType":"idrs","idrS":8},
 {"pcBlockCells":192,"coarseAggregates":1000,"kspType":"idrs","idrS":4}]' timeout 1200 python scripts/pc_bench.py 2>&1 | grep -v "^\[dab200\] pcSymbolic" | tail -12
