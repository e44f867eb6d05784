#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_plasticc.log 2>&1; echo "plasticc rc=$?"; tail -1 gpurun_out/bench_plasticc.log | cut -c1-400
python bench.py --workload ztf --steps 2 --warmup 1 > gpurun_out/bench_ztf.log 2>&1; echo "ztf rc=$?"; tail -1 gpurun_out/bench_ztf.log
python bench.py --workload ddf --steps 1 --warmup 1 > gpurun_out/bench_ddf.log 2>&1; echo "ddf rc=$?"; tail -1 gpurun_out/bench_ddf.log
