#!/bin/bash
# tile-kernel measurement call (ONE B200): parity tests, kbench with/without tiles, ncu of the tile kernels
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${tag}_pytest_gpu.log
export KB_QUIET=1 DAB_TILE_INFO=1
for t in 14x14 7x28 28x7 10x14; do
  echo "== tile $t"; KB_TILE=$t timeout 200 python scripts/kbench.py 2>&1 | tail -2
done
echo "== tile 14x14 numbering, tiles off"; DAB_TILE=0 KB_TILE=14x14 timeout 200 python scripts/kbench.py 2>&1 | tail -1
echo "== lexicographic numbering (tiles as they fit)"; timeout 200 python scripts/kbench.py 2>&1 | tail -2
echo "== 14x14 rolled face loops (DAB_NOHEX6=1)"; DAB_NOHEX6=1 KB_TILE=14x14 timeout 200 python scripts/kbench.py 2>&1 | tail -1
KB_TILE=14x14 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:tileKernel' -s 12 -c 2 -f -o gpurun_out/${tag}_tile \
    python scripts/kbench.py > gpurun_out/${tag}_tile.log 2>&1
ncu -i gpurun_out/${tag}_tile.ncu-rep --page raw --csv > gpurun_out/${tag}_tile_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
