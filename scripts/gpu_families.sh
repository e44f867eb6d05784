#!/bin/bash
# family tests first (all failures shown), then the whole GPU suite and a bench of the default family (regression check)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_families.py -m gpu -q > gpurun_out/pytest_families.log 2>&1; echo "families rc=$?"; tail -25 gpurun_out/pytest_families.log
python -m pytest tests -m gpu -q --deselect tests/test_gpu_families.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cur.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_cur.log
