#!/bin/bash
# First GPU call of the next round (ONE B200, ~12 min): everything that was written after round 1's GPU budget ended and is
# therefore verified on the host build only.  usage:  gpurun --timeout 900 -- 'bash scripts/round2_first_gpu_call.sh r02a'
# Outputs under gpurun_out/<tag>_*; copy what is to be judged into profiles/.
tag=${1:-r02a}
mkdir -p gpurun_out
# 1. parity of every -m gpu test (incl. the unrun transonic / IDR(s) twins)
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${tag}_pytest_gpu.log
# 2. default bench line: GMRES adjoint solve + the IDR(4) leg, new preconditioner set-up (threaded pattern, coloured probing)
timeout 300 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
grep -E "pcSymbolic|coarse space|Main iteration" gpurun_out/${tag}_bench_n1.err | tail -14
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d.get("adjoint_solve"))
PY
# 3. Krylov / ordering alternatives on the same 980k-cell system (one process, PC re-assembled per radius)
CB_CELLS=980000 CB_RADII=0,4 CB_KSP=gmres,idrs:4,idrs:8 CB_RESTART=1500 timeout 420 python scripts/colour_bench.py > gpurun_out/${tag}_colour_bench.log 2>&1
cat gpurun_out/${tag}_colour_bench.log
# 4. config 3 (DARhoSimpleFoam, 2M cells): first timing of the compressible kernels
timeout 300 python bench.py --solver DARhoSimpleFoam --cells 2000000 --no-cpu-baseline > gpurun_out/${tag}_bench_rhosimple_2m.json 2> gpurun_out/${tag}_bench_rhosimple_2m.err
tail -c 1500 gpurun_out/${tag}_bench_rhosimple_2m.json
# 5. lane-per-face pilot of RevA (8 lanes per cell + butterfly reduction) against the cell-per-thread kernel, and its parity on the device
KB_QUIET=1 timeout 120 python scripts/kbench.py > gpurun_out/${tag}_kbench_cells.txt 2>&1; tail -1 gpurun_out/${tag}_kbench_cells.txt
KB_QUIET=1 DAB_LANES=1 timeout 120 python scripts/kbench.py > gpurun_out/${tag}_kbench_lanes.txt 2>&1; tail -1 gpurun_out/${tag}_kbench_lanes.txt
DAB_LANES=1 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
