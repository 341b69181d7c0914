#!/bin/bash
# triangular solves (slot constants, 4 load chains), sparse A*P: the bench-default IDR(8) solve + launch shares
tag=${1:-r02m}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_adjoint_solve.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
export PB_NJ=720 PB_TILE=16x12 PB_TAG=$tag PB_LVL=3
PB_CFGS='[{"pcBlockCells":0,"coarseAggregates":2000,"kspType":"idrs","idrS":8,"pcStorage":"fp32"},
 {"pcBlockCells":0,"coarseAggregates":2000,"kspType":"idrs","idrS":8,"pcStorage":"fp32","coarseSparseAP":0},
 {"pcBlockCells":0,"coarseAggregates":4000,"kspType":"idrs","idrS":8,"pcStorage":"fp32"}]' timeout 1200 python scripts/pc_bench.py 2>&1 | grep -v "^\[dab200\] pcSymbolic" | tail -4
bash scripts/gpu_iter_profile.sh ${tag} 2>&1 | tail -14
