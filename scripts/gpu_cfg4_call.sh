#!/bin/bash
# BASELINE config 4: DASimpleFoam + SA, swept tapered wing, 5M fully 3-D hexahedra on 4 GPUs (peer-memory halos)
tag=${1:-r02s}
mkdir -p gpurun_out
DAB_SETUP_INFO=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 4 --steps 30 --warmup 5 --mesh wing3d --cells 5000000 --scaling strong --no-gmres --no-cpu-baseline \
  > gpurun_out/${tag}_bench_cfg4_n4.json 2> gpurun_out/${tag}_bench_cfg4_n4.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_cfg4_n4.json").read().strip().splitlines()[-1])
    a = d.get("adjoint_solve") or {}
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["roofline"]["frac"], d["roofline"]["kernels_ms"], d["config"]["workload"][:120], "setup %.1f" % d["config"]["setup_s"],
          {k: a.get(k) for k in ("pc_s", "wall_s", "solve_s", "iterations", "fail", "error")})
except Exception as e:
    print("failed", e)
PY
grep -E "Error|error|Traceback|tiles" gpurun_out/${tag}_bench_cfg4_n4.err | sort | uniq -c | tail -5
