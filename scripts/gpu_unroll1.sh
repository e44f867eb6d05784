#!/bin/bash
P="python scripts/fused_probe.py 10000 3"
echo "== compiler's choice"; timeout 200 $P 2>&1 | tail -1
for u in 1 2 4; do echo "== unroll $u"; FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_u$u.so timeout 200 $P 2>&1 | tail -1; done
echo "== compiler's choice again"; timeout 200 $P 2>&1 | tail -1
