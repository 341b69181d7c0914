#!/bin/bash
# the adjoint legs after N SIMPLE iterations of the device primal (the bench state is synthetic)
tag=${1:-r02v}
mkdir -p gpurun_out
timeout 1200 python bench.py --primal-iters 2000 --no-gmres --no-cpu-baseline --steps 20 > gpurun_out/${tag}_bench_primal2000.json 2> gpurun_out/${tag}_bench_primal2000.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_primal2000.json").read().strip().splitlines()[-1])
a = d["adjoint_solve"]
print(d["primal_solve"])
print({k: d[k] for k in ("value", "ms_per_step")}, {k: a.get(k) for k in ("pc_s", "solve_s", "iterations", "fail", "rel_residual", "error")})
PY
grep -E "Error|error" gpurun_out/${tag}_bench_primal2000.err | tail -3
