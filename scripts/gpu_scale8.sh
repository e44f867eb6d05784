#!/bin/bash
# weak scaling of the headline bench on one box: N = 1, 2, 4, 8 ranks (run under gpurun --gpus 8)
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_1.log 2>&1; tail -1 gpurun_out/scale_1.log | cut -c1-200
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.log 2>&1
  echo "n=$n rc=$?"; grep '^{' gpurun_out/scale_$n.log | tail -1 | cut -c1-200
done
