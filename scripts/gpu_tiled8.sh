#!/bin/bash
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/t8_phases.log 2>&1 <<'PY'
import torch, numpy as np, json, fulu_b200
from fulu_b200 import datasets, DeviceBatch
eng = fulu_b200.GPEngine()
for n in (600, 1024, 2048):
    hb = datasets.ddf_like(148, n_points=n)
    db = DeviceBatch.from_host(hb, eng.device)
    th = torch.tensor(np.tile([0.0, 0.0, 0.0, -2.0], (148, 1)))
    eng.phase_cycles(db, th)
    out = eng.phase_cycles(db, th)
    tot = sum(v for k, v in out.items() if isinstance(v, int) and k != "diag_tiles_in_cholesky")
    ideal = n ** 3 / 3 / 128
    print(n, json.dumps(out), "total", tot, "ideal cycles per N^3/3 phase", int(ideal), "efficiency chol/inv/traces",
          [round(ideal / max(out[k], 1), 3) for k in ("cholesky", "inverse", "traces")])
PY
cat gpurun_out/t8_phases.log
CMD="python scripts/long_curve_probe.py 1024"
timeout 300 $CMD > gpurun_out/t8_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:lml_kernel -s 1 -c 1 -f -o gpurun_out/prof_tiled1024_kc4 $CMD > gpurun_out/t8_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/t8_ncu.log
