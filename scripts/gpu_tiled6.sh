#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "long or config5" > gpurun_out/t6_pytest_long.log 2>&1; echo "pytest-long rc=$?"; tail -3 gpurun_out/t6_pytest_long.log
timeout 300 python scripts/long_curve_probe.py 600 1024 1536 2048 > gpurun_out/t6_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/t6_probe.log | tail -6
timeout 900 python bench.py --workload ddf --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/t6_bench_ddf.log 2>&1; echo "ddf rc=$?"; tail -1 gpurun_out/t6_bench_ddf.log | cut -c1-400; tail -1 gpurun_out/t6_bench_ddf.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])"
timeout 900 python bench.py --workload ztf --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/t6_bench_ztf.log 2>&1; echo "ztf rc=$?"; tail -1 gpurun_out/t6_bench_ztf.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac']); [print(c) for c in d['roofline']['launch_classes']]"
