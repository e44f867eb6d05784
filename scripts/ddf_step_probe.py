"""One step of bench.py's config #5 workload (48 curves, N = 1024 / 1536 / 2048, fit + augmentation, batch resident), timed with CUDA
events after one warm-up step: the short form of `bench.py --workload ddf` for A/B runs (a full bench line is five such steps)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fulu_b200 import DeviceBatch, GPEngine, fit_augment_device  # noqa: E402

eng = GPEngine()
host = bench.make_workload("ddf", int(sys.argv[1]) if len(sys.argv) > 1 else 48, 0)
dev = DeviceBatch.from_host(host, eng.device)
lengths = host.lengths()
for it in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    out = fit_augment_device(eng, dev, lengths, 128, profile="counts")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ev = out["n_eval"].cpu().numpy().astype(np.float64)
    flops = float(np.dot(ev, lengths.astype(np.float64) ** 3 + 12.0 * lengths.astype(np.float64) ** 2))
    print(f"{'warm-up' if it == 0 else 'timed'}: {ms:.1f} ms per step, {host.n_curves / ms * 1e3:.2f} curves/s, {flops / ms / 1e9:.2f} TFLOP/s algorithmic, "
          f"{ev.mean():.1f} evaluations per curve, mean LML {float(out['lml'].mean()):.6f}, flagged {int((out['status'].cpu().numpy() & ~8 != 0).sum())}, "
          f"classes {[(int(p['bound']), int(p['curves']), int(p['rounds']), int(p['launches']), int(p['window'])) for p in out['profile']]}", flush=True)
