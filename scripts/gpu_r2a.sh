#!/bin/bash
# round 2, checkpoint A: the whole GPU suite, the three workloads' bench lines with their CPU baselines, the 1M-curve status scan
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/a_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/a_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/a_bench.log | cut -c1-300
timeout 900 python bench.py --workload ztf --steps 2 --warmup 1 > gpurun_out/a_bench_ztf.log 2>&1; echo "ztf rc=$?"; tail -1 gpurun_out/a_bench_ztf.log | cut -c1-200
timeout 1200 python bench.py --workload ddf --steps 1 --warmup 1 > gpurun_out/a_bench_ddf.log 2>&1; echo "ddf rc=$?"; tail -1 gpurun_out/a_bench_ddf.log | cut -c1-200
timeout 900 python scripts/ztf_flagged_scan.py 1000000 100000 > gpurun_out/a_scan.log 2>&1; echo "scan rc=$?"; tail -2 gpurun_out/a_scan.log | cut -c1-600
