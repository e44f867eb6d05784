"""A survey table as it comes (one row per observation, rows of an object contiguous) -> GP fit + augmentation of every object,
with the ragged batch built on the device and only the peak of the band-summed curve read back.

    python examples/survey_table.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fulu_b200 import datasets, fit_augment_table

batch = datasets.plasticc_like(1000, n_points=100)                       # stand-in for a PLAsTiCC-like table
object_id = np.repeat(np.arange(batch.n_curves) + 34299, batch.lengths())
passband2lam = {i: float(v) for i, v in enumerate(batch.loglam)}
res, ids = fit_augment_table(object_id, batch.t, batch.flux, batch.flux_err, batch.band, passband2lam, n_obs=128, reduce="peak")
print(f"{len(ids)} objects; first peaks (mjd): {np.round(res.peak_t[:5], 2)}; mean evaluations per object {res.n_eval.mean():.0f}; "
      f"{res.d2h_bytes / len(ids):.0f} bytes read back per object")
