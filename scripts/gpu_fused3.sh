#!/bin/bash
P="python scripts/fused_probe.py 10000 3"
echo "== rounds, default build"; FULU_GP_FUSED=0 timeout 200 $P 2>&1 | tail -1
echo "== rounds, 96 registers"; FULU_GP_FUSED=0 FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_r96.so timeout 200 $P 2>&1 | tail -1
echo "== fused, slots 16"; timeout 200 $P 2>&1 | tail -1
echo "== fused, slots 32"; FULU_GP_FUSED_SLOTS=32 timeout 200 $P 2>&1 | tail -1
echo "== fused stats, slots 32"; FULU_GP_FUSED_SLOTS=32 FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_stats.so timeout 200 python scripts/fused_probe.py 10000 1 2>&1 | grep "cta 1" | sort | uniq | tail -6
