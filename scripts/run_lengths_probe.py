"""Evaluations per optimiser run, by start index, on config #5's curves: is the length of a run predictable from its start point?
(The tail of a long-curve fit is its longest runs; a task order that starts them first would shorten it.)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fulu_b200 import DeviceBatch, GPEngine, datasets  # noqa: E402

eng = GPEngine()
for n, B in ((1024, 48), (100, 2000)):
    hb = datasets.ddf_like(B, n_points=n) if n > 300 else datasets.plasticc_like(B, n_points=n)
    db = DeviceBatch.from_host(hb, eng.device)
    out = eng.fit(db, return_runs=True)
    ne = out["run_n_eval"].cpu().numpy()
    best = np.argmax(out["run_lml"].cpu().numpy(), axis=1)
    print(f"N={n}, {B} curves: evaluations per run by start index: mean {np.round(ne.mean(0), 1)}, p10 {np.percentile(ne, 10, axis=0)}, "
          f"p90 {np.percentile(ne, 90, axis=0)}, max {ne.max(0)}; winning start histogram {np.bincount(best, minlength=6)}", flush=True)
