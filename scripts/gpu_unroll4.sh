#!/bin/bash
P="python scripts/fused_probe.py 10000 3"
echo "== default (pairs + assembly fast path)"; timeout 200 $P 2>&1 | tail -1
for v in tf sn; do echo "== $v"; FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_$v.so timeout 200 $P 2>&1 | tail -1; done
echo "== default again"; timeout 200 $P 2>&1 | tail -1
