#!/bin/bash
# tile-kernel measurement call (ONE B200): kbench with/without tiles on the 1440x720 O-grid (16x12 tiles = 192 cells)
tag=${1:-r02c}
mkdir -p gpurun_out
export KB_QUIET=1 DAB_TILE_INFO=1 KB_NJ=720
echo "== tile 16x12, tiles on"; KB_TILE=16x12 timeout 200 python scripts/kbench.py 2>&1 | tail -2
echo "== tile 16x12 numbering, tiles off"; DAB_TILE=0 KB_TILE=16x12 timeout 200 python scripts/kbench.py 2>&1 | tail -1
echo "== tile 16x12 numbering, tiles off, boundary faces lexicographic"; KB_BFO=0 DAB_TILE=0 KB_TILE=16x12 timeout 200 python scripts/kbench.py 2>&1 | tail -1
echo "== lexicographic numbering, tiles off"; DAB_TILE=0 timeout 200 python scripts/kbench.py 2>&1 | tail -1
KB_TILE=16x12 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:tileKernel' -s 12 -c 2 -f -o gpurun_out/${tag}_tile \
    python scripts/kbench.py > gpurun_out/${tag}_tile.log 2>&1
ncu -i gpurun_out/${tag}_tile.ncu-rep --page raw --csv > gpurun_out/${tag}_tile_raw.csv 2>/dev/null
DAB_TILE=0 KB_TILE=16x12 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:Rev[ABC]' -s 6 -c 3 -f -o gpurun_out/${tag}_rev \
    python scripts/kbench.py > gpurun_out/${tag}_rev.log 2>&1
ncu -i gpurun_out/${tag}_rev.ncu-rep --page raw --csv > gpurun_out/${tag}_rev_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
