#!/bin/bash
# default workload, product + assembly + IDR(8) solve only (no GMRES leg, no CPU arms): one number for a solver-side change
tag=${1:-r02G}
mkdir -p gpurun_out
DAB_SETUP_INFO=1 timeout 900 python bench.py --no-gmres --no-cpu-baseline > gpurun_out/${tag}_bench_n1_idr.json 2> gpurun_out/${tag}_bench_n1_idr.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_n1_idr.json").read().strip().splitlines()[-1])
a = d["adjoint_solve"]
print({k: d[k] for k in ("value", "ms_per_step")}, {k: a.get(k) for k in ("pc_s", "wall_s", "solve_s", "iterations", "n_matvec", "fail")})
PY
grep "calcPC" gpurun_out/${tag}_bench_n1_idr.err | tail -4
