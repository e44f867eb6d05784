#!/bin/bash
# config #4 (ragged 20-300, 2 bands) by length class and config #5 (N = 1024) single evaluations: where do we stand
mkdir -p gpurun_out
python - > gpurun_out/configs45.log 2>&1 <<'PY'
import time, numpy as np, torch, fulu_b200
from fulu_b200 import datasets, DeviceBatch
from fulu_b200.engine import bucket_curves
eng = fulu_b200.GPEngine()
t0 = time.time(); hb = datasets.ztf_like(6000, first=0); print("gen", round(time.time() - t0, 1), "s", flush=True)
L = hb.lengths()
db = DeviceBatch.from_host(hb, eng.device)
for bound, idx in bucket_curves(L):
    sub = db.select(idx.astype(np.int32), n_max=bound)
    torch.cuda.synchronize(); t0 = time.time()
    f = eng.fit(sub, profile=True)
    p = eng.predict_grid(sub, f["theta"], 128)
    torch.cuda.synchronize(); dt = time.time() - t0
    ne = f["n_eval"].cpu().numpy()[idx]
    flops = float(np.sum(ne * (L[idx].astype(float) ** 3 + 12 * L[idx].astype(float) ** 2)))
    pr = f["profile"]
    print(f"class <= {bound}: {len(idx)} curves, mean N {L[idx].mean():.0f}, {dt*1e3:.0f} ms, {len(idx)/dt:.0f} curves/s, evals/curve {ne.mean():.0f}, "
          f"eval kernel {pr['eval_ms']:.0f} ms = {flops/pr['eval_ms']/1e9:.2f} TFLOP/s, threads {pr['threads_eval']:.0f} x {pr['ctas_per_sm']:.0f} CTAs/SM, global {pr['global_a']:.0f}", flush=True)
torch.cuda.synchronize(); t0 = time.time()
r = fulu_b200.fit_augment_batch(hb, n_obs=128, engine=eng)
dt = time.time() - t0
print(f"config #4 sample: 6000 ragged curves end to end {dt:.2f} s = {6000/dt:.0f} curves/s; flagged {(r.status & ~8 != 0).sum()}", flush=True)
for n in (1024, 2048):
    hb5 = datasets.ddf_like(4, n_points=n, first=0)
    d5 = DeviceBatch.from_host(hb5, eng.device)
    th = torch.tensor(np.tile([0.5, -1.0, 1.0, -3.0], (4, 1)))
    eng.lml(d5, th); torch.cuda.synchronize(); t0 = time.time()
    o = eng.lml(d5, th); torch.cuda.synchronize(); dt = time.time() - t0
    print(f"N={n}: 4 LML+grad evaluations in {dt*1e3:.1f} ms -> {dt/4*1e3:.1f} ms each on one CTA ({(n**3+12*n*n)/(dt)/1e9:.1f} GFLOP/s per CTA); lml {o['lml'].cpu().numpy()}", flush=True)
PY
cat gpurun_out/configs45.log
