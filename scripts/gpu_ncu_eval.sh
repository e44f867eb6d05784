#!/bin/bash
# one ncu --set full capture of the evaluation kernel (after the same command has run clean without ncu)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --curves 2000 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fit_eval -s 10 -c 2 -f -o gpurun_out/prof_eval $CMD > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full.log | cut -c1-300
