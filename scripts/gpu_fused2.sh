#!/bin/bash
for s in 16 32; do
echo "=== stats build, slots $s"
FULU_GP_FUSED_SLOTS=$s FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_stats.so timeout 300 python scripts/fused_probe.py 10000 1 2>&1 | sort | uniq | tail -14
done
