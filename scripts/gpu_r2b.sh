#!/bin/bash
# the stream-ordered fit: whole GPU suite, headline + ztf bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/b_pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/b_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/b_bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['rounds_per_step'], d['roofline']['eval_kernel_ms_per_step'], d['roofline']['step_kernel_ms_per_step'])"
timeout 900 python bench.py --workload ztf --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_bench_ztf.log 2>&1; echo "ztf rc=$?"; tail -1 gpurun_out/b_bench_ztf.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'])"
