#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/long_curve_probe.py 600 1024 1536 2048 > gpurun_out/t9_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/t9_probe.log | tail -6
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_abi.py -x -q -m gpu > gpurun_out/t9_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t9_pytest.log
