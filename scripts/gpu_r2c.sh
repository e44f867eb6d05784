#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload ztf --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c_bench_ztf.log 2>&1; echo "ztf rc=$?"; tail -1 gpurun_out/c_bench_ztf.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['parity'])"
FULU_GP_SERIAL_CLASSES=1 timeout 900 python bench.py --workload ztf --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c_bench_ztf_serial.log 2>&1; echo "ztf serial rc=$?"; tail -1 gpurun_out/c_bench_ztf_serial.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'])"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c_pytest_gpu.log
FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_nocopy.so timeout 300 python scripts/long_curve_probe.py 1024 2048 > gpurun_out/c_probe_nocopy.log 2>&1; echo "nocopy rc=$?"; cat gpurun_out/c_probe_nocopy.log | tail -3
