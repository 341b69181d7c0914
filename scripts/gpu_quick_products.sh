#!/bin/bash
# product / forward / SIMPLE timings of the default workload and of the config-5 passage, no adjoint solve (a kernel change measured in one call)
tag=${1:-r02B}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 30 --warmup 5 --no-solve --no-cpu-baseline > gpurun_out/${tag}_n1_nosolve.json 2> gpurun_out/${tag}_n1_nosolve.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-solve --no-cpu-baseline --mesh passage --solver DATurboFoam --cells 1000000 --primal-iters 100 \
  > gpurun_out/${tag}_cfg5_nosolve.json 2> gpurun_out/${tag}_cfg5_nosolve.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-solve --no-cpu-baseline --primal-iters 100 > gpurun_out/${tag}_n1_primal100.json 2> gpurun_out/${tag}_n1_primal100.err
python - <<PY
import json
for f in ("n1_nosolve", "cfg5_nosolve", "n1_primal100"):
    try:
        d = json.loads(open("gpurun_out/${tag}_%s.json" % f).read().strip().splitlines()[-1])
        p = d.get("primal_solve") or {}
        print(f, "%.4f ms" % d["ms_per_step"], "frac %.4f" % d["roofline"]["frac"], d["roofline"]["kernels_ms"], "fwd %.4f" % d["roofline"]["forward_R_ms"], "e2e %.3f" % d["e2e"]["ms_per_step"], p.get("ms_per_iteration"))
    except Exception as e:
        print(f, "failed", e)
PY
