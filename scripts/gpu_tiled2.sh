#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/long_curve_probe.py 300 500 1024 1536 2048 > gpurun_out/t2_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/t2_probe.log | tail -6
timeout 600 python scripts/ddf_fit_probe.py 1024 2048 > gpurun_out/t2_fitprobe.log 2>&1; echo "fitprobe rc=$?"; cat gpurun_out/t2_fitprobe.log | tail -6
