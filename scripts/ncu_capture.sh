#!/bin/bash
# ncu evidence for one round (run under gpurun on ONE GPU): launch list of a short bench + one --set full capture of the
# reverse and the forward kernels on the bench mesh; raw pages exported to CSV next to the reports.  usage: scripts/ncu_capture.sh r02
tag=${1:-rXX}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_steps3.csv \
    python bench.py --steps 3 --warmup 1 --no-solve --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
export KB_QUIET=1 KB_NJ=720 KB_TILE=16x12
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:Rev[ABC]' -s 6 -c 3 -f -o gpurun_out/${tag}_rev \
    python scripts/kbench.py > gpurun_out/${tag}_rev.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:Fwd[ABC]' -s 3 -c 3 -f -o gpurun_out/${tag}_fwd \
    python scripts/kbench.py > gpurun_out/${tag}_fwd.log 2>&1
for r in rev fwd; do
  ncu -i gpurun_out/${tag}_${r}.ncu-rep --page raw --csv > gpurun_out/${tag}_${r}_raw.csv 2>/dev/null
done
rm -f gpurun_out/${tag}_*.ncu-rep  # the reports are 35 MB each: gpurun_out/ travels back only below 64 MiB
tail -1 gpurun_out/${tag}_rev.log
ls -la gpurun_out | tail -6
