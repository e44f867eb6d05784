#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/run_lengths_probe.py > gpurun_out/i_run_lengths.log 2>&1; echo "rc=$?"; cat gpurun_out/i_run_lengths.log
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_families.py tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/i_pytest.log
