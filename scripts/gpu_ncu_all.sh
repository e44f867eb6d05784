#!/bin/bash
# profiles for profiles/: launch list of a short bench run + ncu --set full of the evaluation and prediction kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --curves 2000 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fit_eval -s 10 -c 2 -f -o gpurun_out/prof_eval $CMD > gpurun_out/ncu_full.log 2>&1
echo "eval rc=$?"
$CMD > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 1 -c 1 -f -o gpurun_out/prof_predict $CMD > gpurun_out/ncu_full2.log 2>&1
echo "predict rc=$?"
