#!/bin/bash
# multi-GPU call: NCCL/P2P parity tests + bench at N GPUs, weak and strong (usage: gpu_multi_call.sh <tag> <N> [solve])
tag=${1:-r02i}; N=${2:-2}; SOLVE=${3:-solve}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_multirank.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
  DAB_P2P=0 timeout 600 python -m pytest tests/test_multirank.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -1
fi
run() { # name, extra env, args
  env $2 DAB_SETUP_INFO=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-gmres --no-cpu-baseline $3 > gpurun_out/${tag}_$1.json 2> gpurun_out/${tag}_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_$1.json").read().strip().splitlines()[-1])
    a = d.get("adjoint_solve") or {}
    print("$1", {k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["roofline"]["kernels_ms"], "setup %.1f" % d["config"]["setup_s"],
          {k: a.get(k) for k in ("pc_s", "wall_s", "solve_s", "iterations", "fail", "error")})
except Exception as e:
    print("$1 failed", e)
PY
  grep -E "halo exchange|Error|error|Traceback" gpurun_out/${tag}_$1.err | sort | uniq -c | tail -5
}
if [ "$N" = "2" ]; then
  run n${N}_weak_nccl DAB_P2P=0 "--scaling weak --no-solve"
  run n${N}_weak_p2p DAB_P2P=1 "--scaling weak --no-solve"
fi
if [ "$SOLVE" = "solve" ]; then
  run n${N}_weak DAB_P2P=1 "--scaling weak"
  run n${N}_strong DAB_P2P=1 "--scaling strong"
fi
