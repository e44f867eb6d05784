#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/long_curve_probe.py 600 1024 1536 2048 > gpurun_out/d_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/d_probe.log | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "long or config5" > gpurun_out/d_pytest_long.log 2>&1; echo "pytest-long rc=$?"; tail -2 gpurun_out/d_pytest_long.log
