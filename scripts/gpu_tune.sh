#!/bin/bash
# A/B of a runtime tuning knob on the 10,000-curve fit: bash scripts/gpu_tune.sh VAR v1 v2 ...
mkdir -p gpurun_out
var=$1; shift
for v in "$@"; do
  echo "== $var=$v"
  env $var=$v python - <<'PY'
import torch, fulu_b200
from fulu_b200 import datasets, DeviceBatch
eng = fulu_b200.GPEngine()
hb = datasets.plasticc_like(10000, first=0)
db = DeviceBatch.from_host(hb, eng.device)
eng.fit(db); torch.cuda.synchronize()
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f = eng.fit(db); e1.record(); torch.cuda.synchronize()
    print("fit ms (no per-launch events)", round(e0.elapsed_time(e1), 2), "lml sum", float(f["lml"].sum()), "evals", int(f["n_eval"].sum()))
PY
done 2>&1 | tee gpurun_out/tune.log
