"""One LML + gradient evaluation per CTA on long curves (config #5's path: the packed tiles live in a per-CTA slab in global
memory), timed with CUDA events: python scripts/long_curve_probe.py [N ...].  Used under ncu by scripts/gpu_ncu_long.sh to get the
counters of the long-curve evaluation (the whole fit at these sizes takes seconds per launch group, too long to replay)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fulu_b200 import DeviceBatch, GPEngine, datasets  # noqa: E402

eng = GPEngine()
for n in [int(a) for a in sys.argv[1:]] or [1024, 2048]:
    hb = datasets.ddf_like(int(os.environ.get("PROBE_CURVES", "148")), n_points=n)
    db = DeviceBatch.from_host(hb, eng.device)
    theta = torch.zeros(hb.n_curves, 4, dtype=torch.float64, device=eng.device)
    theta[:, 3] = -2.0
    eng.lml(db, theta)   # warm-up (module load, workspace)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    out = eng.lml(db, theta)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    flops = hb.n_curves * (float(n) ** 3 + 12.0 * n * n)
    print(f"N={n}: {hb.n_curves} evaluations (one per CTA per wave) in {ms:.2f} ms incl. prep = {flops / ms / 1e9:.2f} TFLOP/s algorithmic; "
          f"lml[0]={float(out['lml'][0]):.6f} finite={bool(torch.isfinite(out['lml']).all())}", flush=True)
