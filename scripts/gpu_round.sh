#!/bin/bash
# One GPU-box session: peaks probe, GPU tests, smoke, short bench.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python - > gpurun_out/peaks.log 2>&1 <<'PY'
import fulu_b200, torch, json
e = fulu_b200.GPEngine()
out = {"dfma_tflops": e.fp64_peak(0), "dmma_tflops": e.fp64_peak(1), "dfma_long": e.fp64_peak(0, 100000), "dmma_long": e.fp64_peak(1, 100000)}
print(json.dumps(out))
PY
cat gpurun_out/peaks.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 2 --warmup 1 --curves ${BENCH_CURVES:-2000} > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -5 gpurun_out/bench.log
