#!/bin/bash
# ncu evidence: a launch list (durations of every kernel of one step) and --set full captures of the two heavy kernels.
# Run under gpurun, 1 GPU.  Each ncu run is preceded by the same command without ncu.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --curves ${PROF_CURVES:-1000} --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fit_eval_kernel -s 10 -c 2 -o gpurun_out/prof_eval -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "eval capture exit $?"
$CMD > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 1 -c 1 -o gpurun_out/prof_predict -f $CMD > gpurun_out/ncu_full2.log 2>&1
echo "predict capture exit $?"
tail -3 gpurun_out/prof_plain.log
