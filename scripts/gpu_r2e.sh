#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/long_curve_probe.py 600 1024 1536 2048 > gpurun_out/e_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/e_probe.log | tail -5
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/e_pytest_gpu.log
timeout 900 python bench.py --workload ddf --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/e_bench_ddf.log 2>&1; echo "ddf rc=$?"; tail -1 gpurun_out/e_bench_ddf.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])"
timeout 900 python bench.py --workload ddf --curves 192 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/e_bench_ddf192.log 2>&1; echo "ddf192 rc=$?"; tail -1 gpurun_out/e_bench_ddf192.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])"
