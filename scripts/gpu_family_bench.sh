#!/bin/bash
# phase cycles of one evaluation + bench of the non-default kernels on config #3's curves (with a bounded CPU baseline each)
mkdir -p gpurun_out
bash scripts/gpu_phases.sh 2>&1 | tail -3
for k in rbf_matern notebook rbf_rq; do
  python bench.py --steps 2 --warmup 3 --kernel $k --ref-curves 64 > gpurun_out/bench_$k.log 2>&1; echo "bench $k rc=$?"; tail -1 gpurun_out/bench_$k.log | cut -c1-1500
done
