"""The reference's README example (fulu README.md:40-65, BASELINE config #1) with the drop-in class: one synthetic 3-band curve,
10 points per band, GP fit + 100-point augmentation.  Needs a B200 (there is no CPU path):

    python examples/readme_example.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fulu_b200 import GaussianProcessesAugmentation          # instead of: from fulu import GaussianProcessesAugmentation

passband2lam = {"u": np.log10(3751.36), "g": np.log10(4741.64), "r": np.log10(6173.23)}
n = 10 * len(passband2lam)
t = np.linspace(0, n - 1, n)
flux = 10 + np.sin(2 * t) + np.random.RandomState(0).normal(0, 0.1, n)
flux_err = np.full(n, 0.1)
passbands = np.array(list(passband2lam) * 10)

aug = GaussianProcessesAugmentation(passband2lam)
aug.fit(t, flux, flux_err, passbands)
t_aug, flux_aug, flux_err_aug, passband_aug = aug.augmentation(t.min(), t.max(), n_obs=100)

print("theta* =", aug.reg.kernel_.theta, " LML* =", aug.reg.log_marginal_likelihood_value_)
for band in passband2lam:
    sel = passband_aug == band
    print(f"band {band}: flux in [{flux_aug[sel].min():.3f}, {flux_aug[sel].max():.3f}], mean error {flux_err_aug[sel].mean():.3f}")

try:   # the reference's README ends with the visualisation (README.md:60-64): same calls, drawn only where matplotlib is installed
    import matplotlib  # noqa: F401
except ImportError:
    print("matplotlib is not installed: skipping LcPlotter.plot_one_graph_all(...)")
else:
    from fulu_b200 import LcPlotter                         # instead of: fulu.LcPlotter

    plotic = LcPlotter(passband2lam)
    plotic.plot_one_graph_all(t=t, flux=flux, flux_err=flux_err, passbands=passbands, t_approx=t_aug, flux_approx=flux_aug,
                              flux_err_approx=flux_err_aug, passband_approx=passband_aug, save="readme_example.png")
