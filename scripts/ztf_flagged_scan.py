"""BASELINE config #4 at full size on one GPU, in chunks: which curves end with a status other than "theta on a bound", and with which
bits (the 1M-curve run flagged 5 curves without saying how).  python scripts/ztf_flagged_scan.py [n_curves] [chunk]
Writes gpurun_out/ztf_flagged.json: the flagged curves (global index of datasets.ztf_like, N, status bits, LML, evaluations) and the
status histogram of the whole run."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fulu_b200  # noqa: E402
from fulu_b200 import datasets  # noqa: E402

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
eng = fulu_b200.GPEngine()
flagged, hist, gpu_s = [], {}, 0.0
t_all = time.perf_counter()
for first in range(0, n_total, chunk):
    hb = datasets.ztf_like(min(chunk, n_total - first), first=first)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = fulu_b200.fit_augment_batch(hb, n_obs=128, engine=eng, reduce="peak")
    gpu_s += time.perf_counter() - t0
    for k, v in res.status_histogram().items():
        hist[k] = hist.get(k, 0) + v
    L = hb.lengths()
    for i in np.nonzero(res.status & ~8)[0]:
        flagged.append({"index": int(first + i), "n": int(L[i]), "status": int(res.status[i]), "lml": float(res.lml[i]), "n_eval": int(res.n_eval[i]),
                        "theta": [float(x) for x in res.theta[i]]})
    print(f"{first + hb.n_curves} curves, {len(flagged)} flagged so far, {gpu_s:.1f} s in the library", flush=True)
out = {"curves": n_total, "status_histogram": hist, "flagged": flagged, "seconds_in_fit_augment_batch": gpu_s, "curves_per_s": n_total / gpu_s,
       "seconds_total_with_generation": time.perf_counter() - t_all}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ztf_flagged.json", "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "flagged"}))
print(json.dumps(flagged))
