#!/bin/bash
mkdir -p gpurun_out
P="python scripts/fused_probe.py 10000 3"
U=$PWD/fulu_b200/_lib/libfulu_gp_unr.so
echo "== rounds, rolled";   FULU_GP_FUSED=0 timeout 200 $P 2>&1 | tail -1
echo "== rounds, inner loops unrolled"; FULU_GP_FUSED=0 FULU_GP_LIB=$U timeout 200 $P 2>&1 | tail -1
for st in 1 2 4; do for sl in 16 32; do
echo "== fused rolled, steppers $st slots $sl"; FULU_GP_FUSED_STEPPERS=$st FULU_GP_FUSED_SLOTS=$sl timeout 200 $P 2>&1 | tail -1
done; done
echo "== fused unrolled, steppers 4 slots 32"; FULU_GP_LIB=$U timeout 200 $P 2>&1 | tail -1
echo "== fused unrolled, steppers 2 slots 32"; FULU_GP_FUSED_STEPPERS=2 FULU_GP_LIB=$U timeout 200 $P 2>&1 | tail -1
echo "== fused stats (rolled, 4 x 8)"; FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_stats.so timeout 200 python scripts/fused_probe.py 10000 1 2>&1 | grep "cta 1" | sort | uniq | tail -12
