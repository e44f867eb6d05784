"""the LcPlotter-compatible layer draws what the reference draws.

tests/golden/plot_calls.json is the list of axes calls the REAL reference made (fulu/plotting.py, fulu/_base_aug.py:160-230),
recorded through tests/fakempl.py by tests/golden/make_plot_golden.py; here fulu_b200.plotting replays the same inputs against
the same recording axes and must produce the same calls, arrays included."""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fakempl  # noqa: E402
from conftest import GOLDEN  # noqa: E402


@pytest.fixture()
def plot_golden():
    with open(os.path.join(GOLDEN, "plot_calls.json")) as f:
        g = json.load(f)
    plt = fakempl.install()
    yield g, plt
    fakempl.uninstall()


def same_calls(got, want, rtol=0.0, atol=0.0):
    assert [c[0] for c in got] == [c[0] for c in want]
    for (name, a, k), (_, a2, k2) in zip(got, want):
        assert len(a) == len(a2) and sorted(k) == sorted(k2), (name, k, k2)
        for x, y in list(zip(a, a2)) + [(k[q], k2[q]) for q in k]:
            if isinstance(y, (list, float)) and not isinstance(y, str):
                np.testing.assert_allclose(np.asarray(x, float), np.asarray(y, float), rtol=rtol, atol=atol, err_msg=name)
            else:
                assert x == y, (name, x, y)


def _inputs(g):
    i, a = g["inputs"], g["approx"]
    train = tuple(np.array(i[k]) for k in ("t", "flux", "flux_err", "passband"))
    approx = tuple(np.array(a[k]) for k in ("t", "flux", "flux_err", "passband"))
    return i["passband2lam"], train, approx


def test_lcplotter_draws_what_the_reference_draws(plot_golden):
    from fulu_b200.plotting import LcPlotter

    g, _ = plot_golden
    p2l, (t, flux, flux_err, passband), approx = _inputs(g)
    pl = LcPlotter(p2l)
    cases = {
        "errorbar_passband_g": lambda ax: pl.errorbar_passband(t, flux, flux_err, passband, "g", ax=ax, title="obs", label="train"),
        "plot_approx_r": lambda ax: pl.plot_approx(*approx, passband="r", ax=ax),
        "plot_sum_passbands": lambda ax: pl.plot_sum_passbands(*approx, ax=ax),
        "plot_peak": lambda ax: pl.plot_peak(*approx, ax=ax, title="peak"),
        "plot_true_peak": lambda ax: pl.plot_true_peak(true_peak=12.5, ax=ax),
        "plot_one_graph_passband_u": lambda ax: pl.plot_one_graph_passband(
            t=t, flux=flux, flux_err=flux_err, passbands=passband, passband="u", ax=ax, t_approx=approx[0], flux_approx=approx[1],
            flux_err_approx=approx[2], passband_approx=approx[3], title="u"),
        "plot_one_graph_all_peak": lambda ax: pl.plot_one_graph_all(
            t=t, flux=flux, flux_err=flux_err, passbands=passband, ax=ax, t_approx=approx[0], flux_approx=approx[1],
            flux_err_approx=approx[2], passband_approx=approx[3], plot_peak=True, title="all"),
    }
    for name, fn in cases.items():
        ax = fakempl.RecordingAxes()
        assert fn(ax) is ax
        same_calls(ax.log, g["scenarios"][name], rtol=1e-14)   # band sums: summation order (pandas groupby) is worth an ulp


def test_band_sum_and_peak_match_pandas(plot_golden):
    import pandas as pd

    from fulu_b200.plotting import band_sum, peak_of_sum

    rng = np.random.default_rng(3)
    t = np.tile(np.linspace(5.0, 9.0, 17), 4)
    flux = rng.normal(size=t.size)
    flux[[3, 20]] = 50.0   # two equal maxima of the sum are possible only by construction; the first one wins
    frame = pd.DataFrame(dict(time=t, flux=flux))
    curve = frame[["time", "flux"]].groupby("time", as_index=False).sum()
    times, total = band_sum(t, flux)
    np.testing.assert_array_equal(times, curve["time"].values)
    np.testing.assert_allclose(total, curve["flux"].values, rtol=1e-15, atol=1e-15)
    assert peak_of_sum(t, flux)[0] == curve["time"][curve["flux"].argmax()]


def test_new_figure_save_and_argument_errors(plot_golden):
    from fulu_b200.plotting import LcPlotter

    g, plt = plot_golden
    p2l, (t, flux, flux_err, passband), approx = _inputs(g)
    pl = LcPlotter(p2l)
    ax = pl.plot_true_peak(true_peak=3.0, save="peak.pdf")       # no axes given: a new figure (plotting.py:27-41), then savefig
    assert ax is plt.axes[-1] and plt.saved == ["peak.pdf"]
    with pytest.raises(ValueError, match="numerical arrays"):    # plotting.py:19-20
        pl.errorbar_passband(t, None, flux_err, passband, "g", ax=ax)
    n = len(ax.log)
    assert pl.plot_peak(None, None, None, None, ax=ax) is ax and len(ax.log) == n   # nothing to draw, nothing drawn (plotting.py:106-111)
    assert list(pl.colors) == list(p2l) and pl.colors["u"] == "tab:blue"            # plotting.py:11


def test_importing_the_package_does_not_import_matplotlib():
    fakempl.uninstall()
    import fulu_b200  # noqa: F401
    from fulu_b200.plotting import LcPlotter

    assert "matplotlib" not in sys.modules
    pl = LcPlotter({0: 3.5})
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="matplotlib"):
            pl.plot_true_peak(true_peak=1.0)


def _aug_scenarios():
    return {
        "aug_plot_default": dict(),
        "aug_plot_peak_g": dict(passband="g", true_peak=12.5, plot_peak=True, title="g"),
        "aug_plot_no_approx": dict(plot_approx=False),
    }


def test_class_plot_call_sequence_with_the_device_step_replaced(plot_golden, monkeypatch):
    """GaussianProcessesAugmentation.plot draws what BaseAugmentation.plot draws (fulu/_base_aug.py:160-230) when its device
    step returns the reference's own augmentation (the device step itself is the GPU test below)."""
    from fulu_b200 import GaussianProcessesAugmentation
    from fulu_b200.plotting import band_sum, peak_of_sum

    g, _ = plot_golden
    p2l, (t, flux, flux_err, passband), approx = _inputs(g)
    aug = GaussianProcessesAugmentation(p2l)
    aug.t_train, aug.flux_train, aug.flux_err_train, aug.passband_train = t, flux, flux_err, passband
    seen = []

    def fake(t_min, t_max, n_obs):
        seen.append((t_min, t_max, n_obs))
        return approx, band_sum(approx[0], approx[1]), peak_of_sum(approx[0], approx[1])[0]

    monkeypatch.setattr(aug, "_peak_on_device", fake)
    for name, kw in _aug_scenarios().items():
        ax = fakempl.RecordingAxes()
        assert aug.plot(n_approx=g["inputs"]["n_approx"], ax=ax, **kw) is ax
        same_calls(ax.log, g["scenarios"][name], rtol=1e-14)
    assert seen[0] == (t.min(), t.max(), g["inputs"]["n_approx"])


@pytest.mark.gpu
def test_class_plot_on_the_device_matches_the_reference(plot_golden):
    """The same scenarios end to end: fit on the GPU, approximation through the grid kernel, band sum and peak through
    fulu_gp_peak_batch.  After optimisation theta agrees with sklearn's to ~1e-5, hence the curve tolerance (as in test_gpu_api)."""
    import warnings

    from fulu_b200 import GaussianProcessesAugmentation

    g, _ = plot_golden
    p2l, (t, flux, flux_err, passband), _ = _inputs(g)
    aug = GaussianProcessesAugmentation(p2l)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        aug.fit(t, flux, flux_err, passband)
    scale = float(np.std(flux))
    for name, kw in _aug_scenarios().items():
        ax = fakempl.RecordingAxes()
        aug.plot(n_approx=g["inputs"]["n_approx"], ax=ax, **kw)
        same_calls(ax.log, g["scenarios"][name], rtol=1e-3, atol=1e-3 * scale)
    # the device's band sum / peak against the host's on the device's own curves (exact grid time, sums to rounding)
    from fulu_b200.plotting import band_sum, peak_of_sum

    approx, (t_sum, f_sum), peak_t = aug._peak_on_device(t.min(), t.max(), 24)
    times, total = band_sum(approx[0], approx[1])
    np.testing.assert_array_equal(t_sum, times)
    np.testing.assert_allclose(f_sum, total, rtol=1e-14)
    assert peak_t == peak_of_sum(approx[0], approx[1])[0]
    a2 = aug.augmentation(t.min(), t.max(), 24)            # grid kernel vs the points kernel the public method uses
    np.testing.assert_allclose(approx[1], a2[1], rtol=1e-12, atol=1e-12 * scale)
    np.testing.assert_allclose(approx[2], a2[2], rtol=1e-9)
