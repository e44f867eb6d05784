#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --curves 2000 --no-cpu-baseline --no-library-baseline"
$CMD > gpurun_out/prof_plain6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fit_eval -s 10 -c 1 -f -o gpurun_out/prof_eval_r2c $CMD > gpurun_out/ncu_full6.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full6.log | cut -c1-200
