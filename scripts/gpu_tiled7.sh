#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/long_curve_probe.py 1024 2048 > gpurun_out/t7_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/t7_probe.log | tail -6
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu > gpurun_out/t7_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t7_pytest.log
timeout 900 python bench.py --workload ddf --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/t7_bench_ddf.log 2>&1; echo "ddf rc=$?"; tail -1 gpurun_out/t7_bench_ddf.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['launch_classes'])"
