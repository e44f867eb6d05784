#!/bin/bash
P="python scripts/fused_probe.py 10000 3"
echo "== n compile-time"; timeout 200 $P 2>&1 | tail -1
echo "== n run-time"; FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_rn.so timeout 200 $P 2>&1 | tail -1
echo "== n compile-time"; timeout 200 $P 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py tests/test_gpu_families.py -x -q -m gpu 2>&1 | tail -2
