#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pipeline.py -m gpu -q > gpurun_out/pytest_pipeline.log 2>&1; echo "pipeline rc=$?"; tail -30 gpurun_out/pytest_pipeline.log
