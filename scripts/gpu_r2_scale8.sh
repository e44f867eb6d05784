#!/bin/bash
# run under gpurun --gpus 8: BASELINE config #4 as ONE global batch of 160,000 ragged curves, LPT-sharded over 1/2/4/8 GPUs with the gather
# inside the timed call (strong scaling); the sharded entry point against one GPU bit for bit; the headline weak-scaling points
mkdir -p gpurun_out
run() { n=$1; shift; if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n "$@"; fi; }
for n in 8 4 2 1; do
  run $n --workload ztf --scaling strong --curves 160000 --steps 2 --warmup 1 > gpurun_out/strong_$n.log 2>&1
  echo "strong n=$n rc=$?"; grep '^{' gpurun_out/strong_$n.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['sharding'], d['gpu_launches'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29790 scripts/sharded_check.py > gpurun_out/sharded_check.log 2>&1; echo "sharded_check rc=$?"; tail -1 gpurun_out/sharded_check.log
for n in 2 8; do
run $n --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/weak_$n.log 2>&1; echo "weak $n rc=$?"; grep '^{' gpurun_out/weak_$n.log | tail -1 | cut -c1-200
done
