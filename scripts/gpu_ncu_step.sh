#!/bin/bash
# ncu --set full of the optimiser-step kernel (launch 20 of a 2,000-curve bench run)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --curves 2000 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fit_step -s 20 -c 1 -f -o gpurun_out/prof_step $CMD > gpurun_out/ncu_full3.log 2>&1
echo "step rc=$?"
