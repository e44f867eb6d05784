#!/bin/bash
P="python scripts/fused_probe.py 10000 3"
U=$PWD/fulu_b200/_lib/libfulu_gp_unr.so
echo "== rounds"; FULU_GP_FUSED=0 timeout 200 $P 2>&1 | tail -1
echo "== rounds, state advanced in global memory"; FULU_GP_FUSED=0 FULU_GP_LIB=$U timeout 200 $P 2>&1 | tail -1
for st in 1 2 4; do
echo "== fused, local copy, steppers $st"; FULU_GP_FUSED_STEPPERS=$st timeout 200 $P 2>&1 | tail -1
echo "== fused, state in global memory, steppers $st"; FULU_GP_FUSED_STEPPERS=$st FULU_GP_LIB=$U timeout 200 $P 2>&1 | tail -1
done
