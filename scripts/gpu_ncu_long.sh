#!/bin/bash
# ncu --set full of the long-curve evaluation (N = 1024 and 2048, matrix in a global slab), after the same command ran without ncu
mkdir -p gpurun_out
CMD="python scripts/long_curve_probe.py 1024 2048"
$CMD > gpurun_out/long_plain.log 2>&1; echo "plain rc=$?"; cat gpurun_out/long_plain.log | tail -3
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lml_kernel -s 1 -c 1 -f -o gpurun_out/prof_long1024 python scripts/long_curve_probe.py 1024 > gpurun_out/ncu_long1.log 2>&1; echo "ncu1024 rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lml_kernel -s 1 -c 1 -f -o gpurun_out/prof_long2048 python scripts/long_curve_probe.py 2048 > gpurun_out/ncu_long2.log 2>&1; echo "ncu2048 rc=$?"
