#!/bin/bash
echo "== tiled: 64x64 block code rolled"; timeout 300 python scripts/long_curve_probe.py 1024 2048 2>&1 | tail -2
echo "== tiled: 64x64 block code as the compiler unrolls it"; FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_tu.so timeout 300 python scripts/long_curve_probe.py 1024 2048 2>&1 | tail -2
echo "== rolled again"; timeout 300 python scripts/long_curve_probe.py 1024 2048 2>&1 | tail -2
