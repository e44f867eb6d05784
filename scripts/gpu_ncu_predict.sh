#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --curves 2000 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 1 -c 1 -f -o gpurun_out/prof_predict $CMD > gpurun_out/ncu_full2.log 2>&1
echo "predict rc=$?"
