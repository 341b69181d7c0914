#!/bin/bash
# multi-GPU call: NCCL parity tests + bench at N GPUs, weak and strong (usage: gpu_multi_call.sh <tag> <N>)
tag=${1:-r02i}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multirank.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for sc in weak strong; do
  DAB_SETUP_INFO=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 30 --warmup 5 --scaling $sc --no-gmres --no-cpu-baseline > gpurun_out/${tag}_bench_n${N}_${sc}.json 2> gpurun_out/${tag}_bench_n${N}_${sc}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_n${N}_${sc}.json").read().strip().splitlines()[-1])
    print("$sc", {k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["roofline"]["kernels_ms"], d["config"]["setup_s"], d.get("adjoint_solve"))
except Exception as e:
    print("$sc failed", e)
PY
  grep -E "Error|error|Traceback" gpurun_out/${tag}_bench_n${N}_${sc}.err | tail -5
done
