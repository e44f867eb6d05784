#!/bin/bash
P="python scripts/fused_probe.py 10000 3"
echo "== trtri with the pre-pass"; timeout 200 $P 2>&1 | tail -1
echo "== trtri with a W pass per block column"; FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_old.so timeout 200 $P 2>&1 | tail -1
echo "== pre-pass"; timeout 200 $P 2>&1 | tail -1
echo "== ztf pre-pass"; PROBE_WORKLOAD=ztf timeout 300 python scripts/fused_probe.py 20000 2 2>&1 | tail -1
echo "== ztf W pass"; PROBE_WORKLOAD=ztf FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_old.so timeout 300 python scripts/fused_probe.py 20000 2 2>&1 | tail -1
