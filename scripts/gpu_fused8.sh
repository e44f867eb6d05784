#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/f8_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/f8_pytest_gpu.log
FULU_GP_FUSED=1 timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/f8_pytest_gpu_fused.log 2>&1; echo "pytest (fused on) rc=$?"; tail -3 gpurun_out/f8_pytest_gpu_fused.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/f8_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/f8_bench.log | cut -c1-400
