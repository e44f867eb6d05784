#!/bin/bash
# A/B: the same fit on differently tuned builds of the library (FULU_GP_LIB)
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib"
  FULU_GP_LIB=$PWD/fulu_b200/_lib/$lib python - <<'PY'
import torch, time, fulu_b200
from fulu_b200 import datasets, DeviceBatch
eng = fulu_b200.GPEngine()
hb = datasets.plasticc_like(10000, first=0)
db = DeviceBatch.from_host(hb, eng.device)
eng.fit(db); torch.cuda.synchronize()
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f = eng.fit(db, profile=True); e1.record(); torch.cuda.synchronize()
    print("fit ms", round(e0.elapsed_time(e1), 2), {k: round(f["profile"][k], 2) for k in ("eval_ms", "step_ms", "threads_eval", "ctas_per_sm")}, "lml sum", float(f["lml"].sum()))
PY
done 2>&1 | tee gpurun_out/ab.log
