#!/bin/bash
# BASELINE config 5's shape on one GPU: DATurboFoam + SA, one passage of an annular rotor row (cyclic sides, MRF zone); cyclic parity first
tag=${1:-r02y}
cells=${2:-1000000}
mkdir -p gpurun_out
[ -z "$SKIP_TESTS" ] && timeout 600 python -m pytest tests/test_cyclic.py -m gpu -x -q > gpurun_out/${tag}_cyclic_tests.log 2>&1
[ -z "$SKIP_TESTS" ] && tail -3 gpurun_out/${tag}_cyclic_tests.log
DAB_SETUP_INFO=1 timeout 900 python bench.py --steps 30 --warmup 5 --mesh passage --solver DATurboFoam --cells $cells --no-gmres --no-cpu-baseline \
  > gpurun_out/${tag}_bench_cfg5_n1.json 2> gpurun_out/${tag}_bench_cfg5_n1.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_cfg5_n1.json").read().strip().splitlines()[-1])
    a = d.get("adjoint_solve") or {}
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["roofline"]["frac"], d["roofline"]["kernels_ms"], d["config"]["workload"][:160], "setup %.1f" % d["config"]["setup_s"],
          {k: a.get(k) for k in ("pc_s", "wall_s", "solve_s", "iterations", "n_matvec", "fail", "error")}, d.get("primal_solve"))
except Exception as e:
    print("failed", e)
PY
grep -E "Error|error|Traceback" gpurun_out/${tag}_bench_cfg5_n1.err | sort | uniq -c | tail -5
tail -5 gpurun_out/${tag}_bench_cfg5_n1.err
