#!/bin/bash
# default bench line on ONE B200 (what the driver runs) + the reference arm
tag=${1:-r02h}
mkdir -p gpurun_out
DAB_SETUP_INFO=1 timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
tail -c 6000 gpurun_out/${tag}_bench_n1.json
grep -E "setup|Main iteration|WARNING|Error|error" gpurun_out/${tag}_bench_n1.err | tail -20
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
tail -c 1500 gpurun_out/${tag}_bench_ref.json
