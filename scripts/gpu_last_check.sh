#!/bin/bash
# last call of the round: the compressible / cyclic GPU tests and the product timings of configs 3 and 5 with the final library
tag=${1:-r02I}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_compressible.py tests/test_cyclic.py tests/test_mrf.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
for w in "cfg5 --mesh passage --solver DATurboFoam --cells 1000000 --primal-iters 0" "cfg3 --solver DARhoSimpleFoam --cells 2000000"; do
  set -- $w; name=$1; shift
  timeout 300 python bench.py --steps 30 --warmup 5 --no-solve --no-cpu-baseline "$@" > gpurun_out/${tag}_${name}_nosolve.json 2> gpurun_out/${tag}_${name}_nosolve.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_${name}_nosolve.json").read().strip().splitlines()[-1])
    print("$name", "%.4f ms" % d["ms_per_step"], "%.3f GCells/s" % d["value"], "frac %.4f" % d["roofline"]["frac"], d["roofline"]["kernels_ms"])
except Exception as e:
    print("$name failed", e)
PY
done
