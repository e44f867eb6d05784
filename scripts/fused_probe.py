#!/usr/bin/env python
"""One resident fit + augmentation of the headline batch (config #3), timed with CUDA events; with a library built with
-DGP_FUSED_STATS (FULU_GP_LIB=...) the fused kernel prints where its groups and its stepper warp spend their cycles.
usage: fused_probe.py [curves] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import fulu_b200
from fulu_b200.engine import DeviceBatch, fit_augment_device, kernel_spec_from_sklearn

B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
workload = os.environ.get("PROBE_WORKLOAD", "plasticc")
engine = fulu_b200.GPEngine(torch.device("cuda", 0))
spec = kernel_spec_from_sklearn(bench.make_kernel("default"))
host = bench.make_workload(workload, B, 0)
lengths = host.lengths()
dev = DeviceBatch.from_host(host, engine.device, family=spec.family)
n_obs = 128
out = fit_augment_device(engine, dev, lengths, n_obs, spec=spec)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(reps):
    e0.record()
    out = fit_augment_device(engine, dev, lengths, n_obs, spec=spec)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
lml = out["lml"].cpu().numpy()
print(f"{workload} {B} curves: fit+augment {min(ts):.2f} ms (best of {reps}) = {B / min(ts) * 1e3:.0f} curves/s; evals/curve {out['n_eval'].double().mean().item():.3f}; "
      f"mean lml {np.mean(lml[np.isfinite(lml)]):.9f}")
