"""torchrun --nproc-per-node N scripts/sharded_check.py : fit_augment_sharded over N GPUs (NCCL) against the same batch on one GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fulu_b200
from fulu_b200 import datasets, sharding

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hb = datasets.ztf_like(600, first=0)
eng = fulu_b200.GPEngine(torch.device("cuda", local))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sharding.fit_augment_sharded(hb, n_obs=32, engine=eng)   # warm-up
dist.barrier()
e0.record()
res = sharding.fit_augment_sharded(hb, n_obs=32, engine=eng)
e1.record()
torch.cuda.synchronize()
ms = sharding.max_over_ranks(e0.elapsed_time(e1), device=eng.device)
if dist.get_rank() == 0:
    ref = fulu_b200.fit_augment_batch(hb, n_obs=32, engine=eng)
    same = all(np.array_equal(getattr(res, k), getattr(ref, k), equal_nan=True) for k in ("flux_aug", "flux_err_aug", "theta", "lml", "n_eval", "status"))
    print(f"sharded over {dist.get_world_size()} GPUs: {ms:.1f} ms (max over ranks), bit-identical to one GPU: {same}", flush=True)
    assert same
dist.barrier()
dist.destroy_process_group()
