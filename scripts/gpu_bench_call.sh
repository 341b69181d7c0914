#!/bin/bash
# default bench line on ONE B200 (what the driver runs) + the reference arm + the round's ncu evidence
tag=${1:-r02p}
mkdir -p gpurun_out
DAB_SETUP_INFO=1 timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
tail -c 4500 gpurun_out/${tag}_bench_n1.json
grep -E "Main iteration|WARNING|Error|error|sparse A" gpurun_out/${tag}_bench_n1.err | tail -8
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
tail -c 600 gpurun_out/${tag}_bench_ref.json
bash scripts/ncu_capture.sh ${tag}
