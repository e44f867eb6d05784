#!/bin/bash
# ncu evidence for the evaluation kernel: a launch list (durations) and one --set full capture.  Run under gpurun, 1 GPU.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --curves ${PROF_CURVES:-1000} --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fit_eval_kernel -s 10 -c 2 -o gpurun_out/prof_eval -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/prof_plain.log
