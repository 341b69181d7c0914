#!/bin/bash
# default bench line on ONE B200 (what the driver runs) + the reference arm + config 3 + the round's ncu evidence
tag=${1:-r02u}
mkdir -p gpurun_out
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
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
tail -c 300 gpurun_out/${tag}_bench_ref.json
timeout 900 python bench.py --solver DARhoSimpleFoam --cells 2000000 --no-gmres --no-cpu-baseline --steps 20 > gpurun_out/${tag}_bench_cfg3.json 2> gpurun_out/${tag}_bench_cfg3.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_cfg3.json").read().strip().splitlines()[-1])
    a = d.get("adjoint_solve") or {}
    print("cfg3", {k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["roofline"]["kernels_ms"], {k: a.get(k) for k in ("pc_s", "solve_s", "iterations", "fail", "error")})
except Exception as e:
    print("cfg3 failed", e)
PY
bash scripts/ncu_capture.sh ${tag}
rm -f gpurun_out/*.ncu-rep
