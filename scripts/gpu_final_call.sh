#!/bin/bash
# end-of-round check on ONE B200: the whole -m gpu tier, smoke(), the default bench line and config 3
tag=${1:-r02D}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -4 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
DAB_SETUP_INFO=1 timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_n1.json").read().strip().splitlines()[-1])
a = d["adjoint_solve"]
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["roofline"]["kernels_ms"], d["e2e"]["value"], d["config"]["setup_s"])
print({k: a.get(k) for k in ("pc_s", "wall_s", "solve_s", "iterations", "method")}, a.get("gmres"))
for k in ("cpu_baseline", "cpu_baseline_handcoded", "cpu_baseline_1core"):
    print(k, {x: d.get(k, {}).get(x) for x in ("value", "cores", "slowdown_all_vs_alone", "error")})
PY
timeout 900 python bench.py --solver DARhoSimpleFoam --cells 2000000 --no-gmres --no-cpu-baseline --steps 20 > gpurun_out/${tag}_bench_cfg3.json 2> gpurun_out/${tag}_bench_cfg3.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_cfg3.json").read().strip().splitlines()[-1])
    a = d.get("adjoint_solve") or {}
    print("cfg3", {k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["roofline"]["kernels_ms"], {k: a.get(k) for k in ("pc_s", "wall_s", "solve_s", "iterations", "fail", "error")})
except Exception as e:
    print("cfg3 failed", e)
PY
