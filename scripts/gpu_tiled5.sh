#!/bin/bash
mkdir -p gpurun_out
for v in v1 v2; do
FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_$v.so timeout 300 python scripts/long_curve_probe.py 1024 2048 > gpurun_out/t5_probe_$v.log 2>&1; echo "probe $v rc=$?"; cat gpurun_out/t5_probe_$v.log | tail -3
done
CMD="python scripts/long_curve_probe.py 2048"
timeout 300 $CMD > gpurun_out/t5_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:lml_kernel -s 1 -c 1 -f -o gpurun_out/prof_tiled2048 $CMD > gpurun_out/t5_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/t5_ncu.log; cat gpurun_out/t5_plain.log
