#!/bin/bash
P="python scripts/fused_probe.py 10000 3"
echo "== lean chol8"; timeout 200 $P 2>&1 | tail -1
echo "== previous commit"; FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_old.so timeout 200 $P 2>&1 | tail -1
echo "== lean chol8 + optimiser state advanced in global memory"; FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_sg.so timeout 200 $P 2>&1 | tail -1
echo "== lean chol8"; timeout 200 $P 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -2
