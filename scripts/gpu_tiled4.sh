#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "long or config5" > gpurun_out/t4_pytest_long.log 2>&1; echo "pytest-long rc=$?"; tail -3 gpurun_out/t4_pytest_long.log
timeout 300 python scripts/long_curve_probe.py 300 500 1024 1536 2048 > gpurun_out/t4_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/t4_probe.log | tail -6
FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_ns3.so timeout 300 python scripts/long_curve_probe.py 1024 2048 > gpurun_out/t4_probe_ns3.log 2>&1; echo "probe ns3 rc=$?"; cat gpurun_out/t4_probe_ns3.log | tail -6
