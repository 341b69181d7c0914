#!/bin/bash
# where the preconditioner set-up time goes (printInfo timers), after the merge-based ILU update and the pinned pattern upload
tag=${1:-r02x}
mkdir -p gpurun_out
export PB_NJ=720 PB_TILE=16x12 PB_TAG=$tag PB_LVL=3
PB_CFGS='[{"coarseAggregates":2000,"kspType":"idrs","idrS":8,"pcStorage":"fp32","printInfo":1}]' timeout 900 python scripts/pc_bench.py 2>&1 | grep -E "calcPC|pcSymbolic|coarse space|cells" | cut -c1-330
