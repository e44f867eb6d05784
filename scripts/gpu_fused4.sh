#!/bin/bash
mkdir -p gpurun_out
P="python scripts/fused_probe.py 10000 3"
echo "== rounds"; FULU_GP_FUSED=0 timeout 200 $P 2>&1 | tail -1
echo "== fused, slots 16"; timeout 200 $P 2>&1 | tail -1
echo "== fused, slots 32"; FULU_GP_FUSED_SLOTS=32 timeout 200 $P 2>&1 | tail -1
echo "== fused stats, slots 32"; FULU_GP_FUSED_SLOTS=32 FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_stats.so timeout 200 python scripts/fused_probe.py 10000 1 2>&1 | grep "cta 1" | sort | uniq | tail -6
FULU_GP_FUSED_SLOTS=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fit_fused -s 1 -c 1 -f -o gpurun_out/prof_fused python scripts/fused_probe.py 2000 1 > gpurun_out/ncu_fused.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_fused.log | cut -c1-200
