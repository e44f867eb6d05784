#!/bin/bash
P="python scripts/fused_probe.py 10000 3"
echo "== new"; timeout 200 $P 2>&1 | tail -1
echo "== pointer loops"; FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_pl.so timeout 200 $P 2>&1 | tail -1
echo "== new"; timeout 200 $P 2>&1 | tail -1
