#!/bin/bash
# the scratch-free layout of the N = 104..127 / 136..159 classes: parity (api + parity + pipeline suites), then A/B on the ragged workloads
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/ns_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/ns_pytest.log
for v in 1 0; do
echo "== FULU_GP_NO_SCRATCH=$v"
FULU_GP_NO_SCRATCH=$v PROBE_WORKLOAD=plasticc_ragged timeout 100 python scripts/fused_probe.py 10000 2 2>&1 | tail -1
FULU_GP_NO_SCRATCH=$v PROBE_WORKLOAD=ztf timeout 100 python scripts/fused_probe.py 20000 2 2>&1 | tail -1
done
