#!/bin/bash
# v5 check: whole GPU suite, phase cycles, bench of the default family
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
bash scripts/gpu_phases.sh 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cur.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_cur.log | cut -c1-300; tail -1 gpurun_out/bench_cur.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('value',d['value'],'e2e',d['e2e']['value'],'frac',r['frac'],'eval_ms',r['eval_kernel_ms_per_step'],'step_ms',r['step_kernel_ms_per_step'],'evals',r['evals_per_curve'],d['parity'])"
