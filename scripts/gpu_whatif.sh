#!/bin/bash
# marginal cost of each phase of one evaluation under load: the fixed-theta LML kernel over 11,840 curves (20 per CTA of the
# 592-CTA grid) with one phase compiled out at a time (libraries built with -DGP_WHATIF=<bit>, see gp_eval_lml_grad)
mkdir -p gpurun_out
for lib in "$@"; do
  FULU_GP_LIB=$PWD/fulu_b200/_lib/$lib python - "$lib" <<'PY'
import sys, torch, numpy as np, fulu_b200
from fulu_b200 import datasets, DeviceBatch
eng = fulu_b200.GPEngine()
B = 11840
hb = datasets.plasticc_like(B, first=0)
db = DeviceBatch.from_host(hb, eng.device)
th = torch.tensor(np.tile([0.2, -1.5, 2.5, -4.5], (B, 1)), device=eng.device)
for _ in range(3): eng.lml(db, th)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.lml(db, th); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(sys.argv[1], "lml call ms (prep + kernel), min of 5:", round(min(ts), 4), "=> us per evaluation per SM:", round(min(ts) * 1e3 / (B / 148), 3))
PY
done 2>&1 | tee gpurun_out/whatif.log
