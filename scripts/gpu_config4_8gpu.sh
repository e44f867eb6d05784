#!/bin/bash
# BASELINE config #4 at full size: 1,000,000 ZTF-like ragged two-band curves sharded over 8 GPUs (125,000 per rank), and the
# headline config #3 at 8 GPUs
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_8.log 2>&1
echo "plasticc x8 rc=$?"; grep '^{' gpurun_out/scale_8.log | tail -1 | cut -c1-220
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --workload ztf --curves 125000 --steps 1 --warmup 1 > gpurun_out/ztf_1m_8gpu.log 2>&1
echo "ztf 1M x8 rc=$?"; grep '^{' gpurun_out/ztf_1m_8gpu.log | tail -1 | cut -c1-400
