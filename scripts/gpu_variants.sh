#!/bin/bash
# occupancy variants of the compressible reverse kernels (variants/libdab200_v*.so, built with -DDAB_CREV?_MINBLOCKS=...): config-5
# product timings per kernel; the snapshot's own library is the base line
tag=${1:-r02H}
mkdir -p gpurun_out
cp dafoam_b200/libdab200.so /tmp/libdab200_base.so
for v in base v1 v2; do
  [ "$v" != "base" ] && cp variants/libdab200_$v.so dafoam_b200/libdab200.so
  timeout 600 python bench.py --steps 30 --warmup 5 --no-solve --no-cpu-baseline --mesh passage --solver DATurboFoam --cells 1000000 --primal-iters 0 \
    > gpurun_out/${tag}_cfg5_$v.json 2> gpurun_out/${tag}_cfg5_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_cfg5_$v.json").read().strip().splitlines()[-1])
    print("$v", "%.4f ms" % d["ms_per_step"], d["roofline"]["kernels_ms"])
except Exception as e:
    print("$v failed", e)
PY
done
cp /tmp/libdab200_base.so dafoam_b200/libdab200.so
