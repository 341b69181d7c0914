#!/bin/bash
# end-of-round check on 2 GPUs: the NCCL / peer-memory parity tests (incl. the cyclic passage) and the driver's N = 2 bench command
tag=${1:-r02C}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multirank.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${tag}_n2_tests.log 2>&1
tail -3 gpurun_out/${tag}_n2_tests.log
DAB_P2P=0 timeout 600 python -m pytest tests/test_multirank.py -m gpu -q -x -p no:cacheprovider -k passage 2>&1 | tail -1
DAB_SETUP_INFO=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_n2.json").read().strip().splitlines()[-1])
    a = d.get("adjoint_solve") or {}
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling")}, d["roofline"]["kernels_ms"], "setup %.1f" % d["config"]["setup_s"],
          {k: a.get(k) for k in ("pc_s", "wall_s", "solve_s", "iterations", "n_matvec", "fail", "error")})
except Exception as e:
    print("failed", e)
PY
grep -E "halo exchange|Error|error|Traceback" gpurun_out/${tag}_bench_n2.err | sort | uniq -c | tail -5
