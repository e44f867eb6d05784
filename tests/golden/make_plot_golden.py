"""Record what the REAL reference draws (fulu/plotting.py LcPlotter, fulu/_base_aug.py BaseAugmentation.plot) into
tests/golden/plot_calls.json, through the recording matplotlib stand-in tests/fakempl.py (matplotlib is absent from this image).

Run in the build container only (``python tests/golden/make_plot_golden.py``): /root/reference does not exist on the GPU box.
The reference source is executed where it lies; nothing is copied.  The fixture holds the inputs (the README example's curve with
a seeded noise) and, per scenario, the list of axes calls [method, positional arrays, keywords].
"""
import json
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference"

import fakempl  # noqa: E402


def readme_curve():
    """README.md:45-58 with RandomState(0) noise."""
    p2l = {"u": np.log10(3751.36), "g": np.log10(4741.64), "r": np.log10(6173.23)}
    t = np.arange(0, 30, dtype=float)
    flux = 10.0 + np.sin(2.0 * t) + np.random.RandomState(0).normal(0, 0.1, len(t))
    flux_err = np.full_like(t, 0.1)
    passband = np.array(["u", "g", "r"] * 10)
    return p2l, t, flux, flux_err, passband


def main():
    plt = fakempl.install()
    pkg = types.ModuleType("fulu")
    pkg.__path__ = [os.path.join(REF, "fulu")]
    sys.modules["fulu"] = pkg
    import fulu.gp_aug as gp_aug
    import fulu.plotting as plotting

    p2l, t, flux, flux_err, passband = readme_curve()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        aug = gp_aug.GaussianProcessesAugmentation(p2l).fit(t, flux, flux_err, passband)
    n_approx = 24
    approx = aug.augmentation(t.min(), t.max(), n_approx)
    out = dict(inputs=dict(t=t.tolist(), flux=flux.tolist(), flux_err=flux_err.tolist(), passband=passband.tolist(),
                           passband2lam={k: float(v) for k, v in p2l.items()}, n_approx=n_approx),
               approx=dict(t=approx[0].tolist(), flux=approx[1].tolist(), flux_err=approx[2].tolist(), passband=approx[3].tolist()),
               scenarios={})

    def record(name, fn):
        ax = fakempl.RecordingAxes()
        fn(ax)
        out["scenarios"][name] = ax.log

    pl = plotting.LcPlotter(p2l)
    record("errorbar_passband_g", lambda ax: pl.errorbar_passband(t, flux, flux_err, passband, "g", ax=ax, title="obs", label="train"))
    record("plot_approx_r", lambda ax: pl.plot_approx(*approx, passband="r", ax=ax))
    record("plot_sum_passbands", lambda ax: pl.plot_sum_passbands(*approx, ax=ax))
    record("plot_peak", lambda ax: pl.plot_peak(*approx, ax=ax, title="peak"))
    record("plot_true_peak", lambda ax: pl.plot_true_peak(true_peak=12.5, ax=ax))
    record("plot_one_graph_passband_u", lambda ax: pl.plot_one_graph_passband(
        t=t, flux=flux, flux_err=flux_err, passbands=passband, passband="u", ax=ax, t_approx=approx[0], flux_approx=approx[1],
        flux_err_approx=approx[2], passband_approx=approx[3], title="u"))
    record("plot_one_graph_all_peak", lambda ax: pl.plot_one_graph_all(
        t=t, flux=flux, flux_err=flux_err, passbands=passband, ax=ax, t_approx=approx[0], flux_approx=approx[1],
        flux_err_approx=approx[2], passband_approx=approx[3], plot_peak=True, title="all"))
    # the class method (fulu/_base_aug.py:160-230): approximation from the fitted sklearn model
    record("aug_plot_default", lambda ax: aug.plot(n_approx=n_approx, ax=ax))
    record("aug_plot_peak_g", lambda ax: aug.plot(n_approx=n_approx, ax=ax, passband="g", true_peak=12.5, plot_peak=True, title="g"))
    record("aug_plot_no_approx", lambda ax: aug.plot(n_approx=n_approx, ax=ax, plot_approx=False))
    with open(os.path.join(HERE, "plot_calls.json"), "w") as f:
        json.dump(out, f)
    print({k: len(v) for k, v in out["scenarios"].items()}, "saved", plt.saved)


if __name__ == "__main__":
    main()
