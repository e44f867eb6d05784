"""LcPlotter-compatible presentation layer for the drop-in class (reference: fulu/plotting.py:6-330,
fulu/_base_aug.py:116-230).

Same method names, keyword arguments and drawing order as the reference's ``LcPlotter``, so ``aug.plot(...)`` and
``aug.plotter.<method>(...)`` of ``fulu_b200.GaussianProcessesAugmentation`` keep working.  Two differences in how it is built:
matplotlib is imported when something is drawn, not when the module is imported (the GP path itself never needs it, and a
batch node has no display stack), and the table work is plain numpy on the arrays as they are (stable sort by time, band-sum
by grouping equal times) instead of pandas frames.  ``GaussianProcessesAugmentation.plot(plot_peak=True)`` takes the band-summed
curve and its peak from the device kernel (fulu_gp_peak_batch) rather than from here.
"""
from __future__ import annotations

import numpy as np

_STYLE = dict(legend=dict(loc="best", ncol=3, fontsize=20), title=dict(size=35, pad=15))


def _pyplot():
    try:
        import matplotlib.pyplot as plt
    except ImportError as e:  # pragma: no cover - depends on the environment
        raise ImportError("plotting needs matplotlib (the GP path of fulu_b200 does not); install it to use .plot() / LcPlotter") from e
    return plt


def _columns(t, flux, flux_err, passband):
    """The reference's _make_dataframe check (plotting.py:17-25) on plain arrays."""
    if t is None or flux is None or flux_err is None or passband is None:
        raise ValueError("All values must be numerical arrays")
    return np.asarray(t), np.asarray(flux), np.asarray(flux_err), np.asarray(passband)


def band_sum(t, flux):
    """Flux summed over the passbands at each distinct time, times ascending (groupby("time").sum(), plotting.py:93,114)."""
    t, flux = np.asarray(t, np.float64), np.asarray(flux, np.float64)
    times, inverse = np.unique(t, return_inverse=True)
    total = np.zeros(len(times))
    np.add.at(total, inverse.reshape(-1), flux)
    return times, total


def peak_of_sum(t, flux):
    """(time, flux) of the first maximum of the band-summed curve (plotting.py:114-115)."""
    times, total = band_sum(t, flux)
    k = int(np.argmax(total))
    return times[k], total[k]


class LcPlotter:
    """Drawing helpers with the reference's LcPlotter surface (fulu/plotting.py:8-11)."""

    def __init__(self, passband2lam):
        self.passband2lam = passband2lam
        self._colors = None

    @property
    def colors(self):
        if self._colors is None:
            import matplotlib.colors as mcolors  # deferred, like every matplotlib use here
            self._colors = dict(zip(self.passband2lam.keys(), mcolors.TABLEAU_COLORS.keys()))
        return self._colors

    # -- figure plumbing (plotting.py:27-46)
    @staticmethod
    def _ax_adjust():
        plt = _pyplot()
        fig = plt.figure(figsize=(20, 10), dpi=100)
        plt.rcParams.update({"font.size": 30})
        fig.subplots_adjust(left=0.07, right=0.98, top=0.90, bottom=0.15)
        ax = plt.subplot(1, 1, 1)
        for side in ("bottom", "top", "left", "right"):
            ax.spines[side].set_linewidth(3)
        ax.set_xlabel("Time", size=30)
        ax.set_ylabel("Flux", size=30)
        ax.grid(linewidth=2)
        return ax

    @staticmethod
    def _save_fig(save_path):
        _pyplot().savefig(save_path)
        print("Your graph was saved into {}".format(save_path))

    def _finish(self, ax, title, save, legend=True, title_first=False):
        if title_first:   # the composed graphs set the title before the legend (plotting.py:326-327, _base_aug.py:224-225)
            ax.set_title(title, **_STYLE["title"])
        if legend:
            ax.legend(**_STYLE["legend"])
        if not title_first:
            ax.set_title(title, **_STYLE["title"])
        if save is not None:
            self._save_fig(save)
        return ax

    @staticmethod
    def _band_sorted(t, flux, flux_err, passbands, passband):
        t, flux, flux_err, passbands = _columns(t, flux, flux_err, passbands)
        keep = np.nonzero(passbands == passband)[0]
        keep = keep[np.argsort(t[keep], kind="stable")]
        return t[keep], flux[keep], flux_err[keep]

    # -- single layers
    def errorbar_passband(self, t, flux, flux_err, passbands, passband, *, ax=None, title="", label="", marker="^", save=None):
        """Observations of one passband with error bars (plotting.py:48-76)."""
        if ax is None:
            ax = self._ax_adjust()
        x, y, e = self._band_sorted(t, flux, flux_err, passbands, passband)
        ax.errorbar(x, y, yerr=e, linewidth=3.5, fmt=marker, elinewidth=1.7, markersize=14.50, markeredgecolor="black",
                    markeredgewidth=1.50, color=self.colors[passband], label="{} {}".format(passband, label))
        return self._finish(ax, title, save)

    def plot_approx(self, t_approx, flux_approx, flux_err_approx, passband_approx, *, passband, ax=None, title="", save=None):
        """Approximation curve of one passband with its one-sigma band (plotting.py:125-159)."""
        if ax is None:
            ax = self._ax_adjust()
        self._draw_approx(ax, t_approx, flux_approx, flux_err_approx, passband_approx, passband)
        return self._finish(ax, title, save)

    def _draw_approx(self, ax, t_approx, flux_approx, flux_err_approx, passband_approx, passband):
        x, y, e = self._band_sorted(t_approx, flux_approx, flux_err_approx, passband_approx, passband)
        ax.plot(x, y, linewidth=3.5, color=self.colors[passband], label="{} approx flux".format(passband), zorder=10)
        ax.fill_between(x, y - e, y + e, color=self.colors[passband], alpha=0.2, label="{} approx sigma".format(passband))

    def plot_sum_passbands(self, t_approx, flux_approx, flux_err_approx, passband_approx, *, ax=None, title="", save=None):
        """The band-summed approximation (plotting.py:78-100)."""
        if ax is None:
            ax = self._ax_adjust()
        have = not any(v is None for v in (t_approx, flux_approx, flux_err_approx, passband_approx))
        if have:
            self._draw_sum(ax, *band_sum(t_approx, flux_approx))
        return self._finish(ax, title, save) if have else self._save_only(ax, save)

    def plot_peak(self, t_approx, flux_approx, flux_err_approx, passband_approx, *, ax=None, title="", save=None):
        """A vertical line at the band-summed approximation's maximum (plotting.py:102-123)."""
        if ax is None:
            ax = self._ax_adjust()
        have = not any(v is None for v in (t_approx, flux_approx, flux_err_approx, passband_approx))
        if have:
            self._draw_peak(ax, peak_of_sum(t_approx, flux_approx)[0])
        return self._finish(ax, title, save) if have else self._save_only(ax, save)

    @staticmethod
    def _draw_sum(ax, times, total):
        ax.plot(times, total, label="sum", linewidth=5.5, color="pink")

    @staticmethod
    def _draw_peak(ax, pred_peak):
        ax.axvline(pred_peak, label="pred peak", color="red", linestyle="--", linewidth=5.5)

    def _save_only(self, ax, save):
        if save is not None:
            self._save_fig(save)
        return ax

    def plot_true_peak(self, *, true_peak, ax=None, title="", save=None):
        """A vertical line at the known peak time (plotting.py:208-218)."""
        if ax is None:
            ax = self._ax_adjust()
        ax.axvline(true_peak, label="true peak", color="black", linewidth=5.5)
        return self._finish(ax, title, save, legend=False)

    # -- composed graphs
    def plot_one_graph_passband(self, *, t, flux, flux_err, passbands, passband, ax=None, t_test=None, flux_test=None, flux_err_test=None,
                                passband_test=None, t_approx=None, flux_approx=None, flux_err_approx=None, passband_approx=None, title="",
                                save=None):
        """Train points, optional test points and optional approximation of one passband (plotting.py:161-206)."""
        if ax is None:
            ax = self._ax_adjust()
        self.errorbar_passband(t, flux, flux_err, passbands, passband, ax=ax)
        if not any(v is None for v in (t_test, flux_test, flux_err_test, passband_test)):
            # the reference's call here omits the passband argument and raises TypeError (plotting.py:185); drawn as intended
            self.errorbar_passband(t_test, flux_test, flux_err_test, passband_test, passband, ax=ax)
        if not any(v is None for v in (t_approx, flux_approx, flux_err_approx, passband_approx)):
            self.plot_approx(t_approx, flux_approx, flux_err_approx, passband_approx, passband=passband, ax=ax)
        return self._finish(ax, title, save)

    def plot_one_graph_all(self, *, t, flux, flux_err, passbands, t_approx=None, flux_approx=None, flux_err_approx=None, passband_approx=None,
                           passband=None, ax=None, true_peak=None, plot_peak=False, title="", save=None):
        """All passbands (or one) on one graph, optional peak markers (plotting.py:220-330)."""
        if ax is None:
            ax = self._ax_adjust()
        approx = dict(t_approx=t_approx, flux_approx=flux_approx, flux_err_approx=flux_err_approx, passband_approx=passband_approx)
        for band in (self.passband2lam.keys() if passband is None else (passband,)):
            self.plot_one_graph_passband(t=t, flux=flux, flux_err=flux_err, passbands=passbands, passband=band, ax=ax, **approx)
        if true_peak is not None:
            # the reference passes true_peak positionally to a keyword-only parameter and raises TypeError (plotting.py:303)
            self.plot_true_peak(true_peak=true_peak, ax=ax)
        if plot_peak:
            self.plot_sum_passbands(t_approx, flux_approx, flux_err_approx, passband_approx, ax=ax)
            self.plot_peak(t_approx, flux_approx, flux_err_approx, passband_approx, ax=ax)
        return self._finish(ax, title, save, title_first=True)
