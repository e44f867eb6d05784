import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, warnings, numpy as np, torch, fulu_b200
from fulu_b200 import datasets
hb = datasets.plasticc_like(4, n_points=100, first=11)
p2l = {i: float(v) for i, v in enumerate(hb.loglam)}
warnings.simplefilter("ignore")
for i in range(4):
    t, f, e, b = hb.curve(i)
    aug = fulu_b200.GaussianProcessesAugmentation(p2l)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    aug.fit(t, f, e, b); t1 = time.perf_counter()
    aug.augmentation(t.min(), t.max(), 128); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"curve {i}: fit {1e3*(t1-t0):.1f} ms, augmentation {1e3*(t2-t1):.1f} ms, evals {aug.reg.n_eval_}")
