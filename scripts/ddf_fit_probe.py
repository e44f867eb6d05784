"""Where does a long-curve fit spend its time?  148 curves of one length, ONE start each (148 tasks = one wave of CTAs), fitted by the
streamed fused kernel (default) and on rounds of two launches (FULU_GP_NO_TAIL=1): wall time against the largest evaluation count of
a task gives the time per evaluation inside the fit, to compare with scripts/long_curve_probe.py (one evaluation per CTA).
python scripts/ddf_fit_probe.py [N ...]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fulu_b200 import DeviceBatch, GPEngine, datasets  # noqa: E402
from fulu_b200.engine import KernelSpec  # noqa: E402

eng = GPEngine()
spec = KernelSpec(n_restarts=0)
for n in [int(a) for a in sys.argv[1:]] or [1024, 2048]:
    hb = datasets.ddf_like(148, n_points=n)
    db = DeviceBatch.from_host(hb, eng.device)
    for mode in ("streamed", "rounds"):
        if mode == "rounds":
            os.environ["FULU_GP_NO_TAIL"] = "1"
        else:
            os.environ.pop("FULU_GP_NO_TAIL", None)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = eng.fit(db, spec, maxiter=12)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ne = out["n_eval"].cpu().numpy()
        flops = float(ne.sum()) * (float(n) ** 3 + 12.0 * n * n)
        print(f"N={n} {mode}: {dt * 1e3:.0f} ms, evaluations per task min {ne.min()} mean {ne.mean():.1f} max {ne.max()}, "
              f"{dt * 1e3 / ne.max():.1f} ms per evaluation of the longest task, {flops / dt / 1e12:.2f} TFLOP/s", flush=True)
