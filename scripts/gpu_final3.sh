#!/bin/bash
# the non-default kernels on config #3's curves, final round-2 build
mkdir -p gpurun_out
for k in rbf_matern notebook rbf_rq; do
  timeout 600 python bench.py --kernel $k > gpurun_out/z_bench_$k.log 2>&1; echo "$k rc=$?"; tail -1 gpurun_out/z_bench_$k.log | cut -c1-160
done
