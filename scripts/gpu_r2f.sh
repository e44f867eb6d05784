#!/bin/bash
mkdir -p gpurun_out
echo "slab path (default), 296 evaluations"; PROBE_CURVES=296 timeout 300 python scripts/long_curve_probe.py 240 300 400 500 2>&1 | tail -4
echo "tiled path from np > 224, 296 evaluations"; FULU_GP_TILED_ABOVE=224 PROBE_CURVES=296 timeout 300 python scripts/long_curve_probe.py 240 300 400 500 2>&1 | tail -4
