#!/bin/bash
# first look at the tiled long-curve path: parity tests of the long classes, one evaluation per CTA at N = 1024 / 2048, the ddf bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "long or config5" > gpurun_out/t1_pytest_long.log 2>&1; echo "pytest-long rc=$?"; tail -5 gpurun_out/t1_pytest_long.log
timeout 300 python scripts/long_curve_probe.py 1024 2048 > gpurun_out/t1_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/t1_probe.log | tail -4
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t1_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t1_pytest_gpu.log
timeout 600 python bench.py --workload ddf --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/t1_bench_ddf.log 2>&1; echo "ddf rc=$?"; tail -1 gpurun_out/t1_bench_ddf.log
