#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload plasticc_ragged > gpurun_out/z_bench_ragged.log 2>&1; echo "ragged rc=$?"; tail -1 gpurun_out/z_bench_ragged.log | cut -c1-300
