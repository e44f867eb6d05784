#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s2_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/s2_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/s2_bench.log | cut -c1-330
timeout 600 python bench.py --workload ztf > gpurun_out/s2_bench_ztf.log 2>&1; echo "ztf rc=$?"; tail -1 gpurun_out/s2_bench_ztf.log | cut -c1-330
timeout 900 python bench.py --workload ddf > gpurun_out/s2_bench_ddf.log 2>&1; echo "ddf rc=$?"; tail -1 gpurun_out/s2_bench_ddf.log | cut -c1-330
