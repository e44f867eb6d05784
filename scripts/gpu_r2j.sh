#!/bin/bash
mkdir -p gpurun_out
echo "library in tree (before the cross-product look-ahead)"; timeout 300 python scripts/long_curve_probe.py 600 1024 1536 2048 2>&1 | tail -4
echo "with the look-ahead running on into the next product"; FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_c1.so timeout 300 python scripts/long_curve_probe.py 600 1024 1536 2048 2>&1 | tail -4
FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_c1.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "long or config5" 2>&1 | tail -2
