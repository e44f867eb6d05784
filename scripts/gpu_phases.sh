#!/bin/bash
mkdir -p gpurun_out
python - > gpurun_out/phases.log 2>&1 <<'PY'
import torch, numpy as np, json, fulu_b200
from fulu_b200 import datasets, DeviceBatch
eng = fulu_b200.GPEngine()
for B in (1, 2000):
    hb = datasets.plasticc_like(B, first=77)
    db = DeviceBatch.from_host(hb, eng.device)
    th = torch.tensor(np.tile([0.2, -1.5, 2.5, -4.5], (B, 1)))
    eng.phase_cycles(db, th)
    out = eng.phase_cycles(db, th)
    print(B, json.dumps(out), "total", sum(v for k, v in out.items() if isinstance(v, int) and k != "diag_tiles_in_cholesky"))
PY
cat gpurun_out/phases.log
