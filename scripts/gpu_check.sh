#!/bin/bash
# parity tests, phase cycles, step timeline and the bench on one B200
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
bash scripts/gpu_phases.sh 2>&1 | tail -2
bash scripts/gpu_timeline.sh 2>&1 | head -6
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cur.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_cur.log
