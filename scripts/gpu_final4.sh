#!/bin/bash
mkdir -p gpurun_out
for k in rbf_matern notebook rbf_rq; do
  timeout 600 python bench.py --kernel $k --no-library-baseline > gpurun_out/z_bench_$k.log 2>&1; echo "$k rc=$?"; tail -1 gpurun_out/z_bench_$k.log | cut -c1-160
done
timeout 200 python scripts/fused_probe.py 10000 3 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_families.py -x -q -m gpu 2>&1 | tail -2
