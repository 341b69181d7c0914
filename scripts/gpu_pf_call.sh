#!/bin/bash
# L2-prefetch measurement (ONE B200): kbench of the cell-per-thread kernels with and without the bulk L2 prefetch plan
tag=${1:-r02d}
mkdir -p gpurun_out
export KB_QUIET=1 KB_NJ=720 KB_TILE=16x12
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
echo "== prefetch off"; DAB_PREFETCH_L2=0 timeout 200 python scripts/kbench.py 2>&1 | tail -1
echo "== prefetch on (ahead = one wave)"; timeout 200 python scripts/kbench.py 2>&1 | tail -1
for a in 200 400 1600 3200; do echo "== prefetch ahead $a"; DAB_PF_AHEAD=$a timeout 200 python scripts/kbench.py 2>&1 | tail -1; done
echo "== lexicographic, prefetch on"; KB_TILE= timeout 200 python scripts/kbench.py 2>&1 | tail -1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:Rev[ABC]' -s 6 -c 3 -f -o gpurun_out/${tag}_rev \
    python scripts/kbench.py > gpurun_out/${tag}_rev.log 2>&1
ncu -i gpurun_out/${tag}_rev.ncu-rep --page raw --csv > gpurun_out/${tag}_rev_raw.csv 2>/dev/null
ls -la gpurun_out | tail -3
