#!/bin/bash
# round 2 final lines: GPU suite, smoke, the three workloads' bench lines with CPU baselines on the final build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/h_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/h_smoke.log
timeout 900 python bench.py > gpurun_out/h_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/h_bench.log | cut -c1-200
timeout 900 python bench.py --impl reference > gpurun_out/h_bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/h_bench_ref.log | cut -c1-200
timeout 900 python bench.py --workload ztf --steps 2 --warmup 1 > gpurun_out/h_bench_ztf.log 2>&1; echo "ztf rc=$?"; tail -1 gpurun_out/h_bench_ztf.log | cut -c1-200
timeout 1200 python bench.py --workload ddf --steps 1 --warmup 1 > gpurun_out/h_bench_ddf.log 2>&1; echo "ddf rc=$?"; tail -1 gpurun_out/h_bench_ddf.log | cut -c1-200
timeout 1200 python bench.py --workload ddf --curves 192 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/h_bench_ddf192.log 2>&1; echo "ddf192 rc=$?"; tail -1 gpurun_out/h_bench_ddf192.log | cut -c1-200
