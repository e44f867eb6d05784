#!/bin/bash
# where does a resident step spend its time: fit (with and without library-side profiling events), predict, flush
mkdir -p gpurun_out
python - > gpurun_out/timeline.log 2>&1 <<'PY'
import torch, numpy as np, json, time, fulu_b200
from fulu_b200 import datasets, DeviceBatch
eng = fulu_b200.GPEngine()
B = 10000
hb = datasets.plasticc_like(B, first=0)
db = DeviceBatch.from_host(hb, eng.device)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); w = time.perf_counter() - t0
        out.append((round(e0.elapsed_time(e1), 2), round(w * 1e3, 2)))
    return out, r
o, fit = timed(lambda: eng.fit(db)); print("fit noprofile (event ms, wall ms)", o)
o, fitp = timed(lambda: eng.fit(db, profile=True)); print("fit profile", o, fitp["profile"])
o, _ = timed(lambda: eng.predict_grid(db, fit["theta"], 128)); print("predict", o)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=eng.device)
o, _ = timed(lambda: flush.zero_()); print("flush", o)
for w in (4096, 8192, 32768, 60000):
    o, f = timed(lambda: eng.fit(db, window=w, profile=True), reps=2); print("window", w, o, f["profile"])
print("evals/curve", float(fit["n_eval"].double().mean()))
PY
cat gpurun_out/timeline.log
