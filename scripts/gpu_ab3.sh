#!/bin/bash
P="python scripts/fused_probe.py 10000 3"
echo "== groups of 2 (generic code)"; timeout 200 $P 2>&1 | tail -1
echo "== groups of 4"; FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_g4.so timeout 200 $P 2>&1 | tail -1
echo "== groups of 2"; timeout 200 $P 2>&1 | tail -1
echo "== groups of 4"; FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_g4.so timeout 200 $P 2>&1 | tail -1
