"""A recording stand-in for matplotlib (absent from this image): axes whose drawing methods append (method, arrays, keywords)
to a log.  Used by tests/golden/make_plot_golden.py to record what the REAL reference LcPlotter / BaseAugmentation.plot draw,
and by the tests to record what fulu_b200 draws for the same inputs."""
import sys
import types

import numpy as np

TABLEAU = ["tab:blue", "tab:orange", "tab:green", "tab:red", "tab:purple", "tab:brown", "tab:pink", "tab:gray", "tab:olive", "tab:cyan"]


def _plain(v):
    if isinstance(v, np.ndarray):
        return [float(x) for x in v.reshape(-1)]
    if isinstance(v, (np.floating, np.integer)):
        return float(v)
    return v


class _Spine:
    def set_linewidth(self, w):
        pass


class RecordingAxes:
    def __init__(self):
        self.log = []
        self.spines = {k: _Spine() for k in ("bottom", "top", "left", "right")}

    def _rec(self, name, args, kw):
        self.log.append([name, [_plain(np.asarray(a)) if np.ndim(a) else _plain(a) for a in args], {k: _plain(np.asarray(v)) if np.ndim(v) else _plain(v) for k, v in sorted(kw.items())}])

    def errorbar(self, *a, **k):
        self._rec("errorbar", a, k)

    def plot(self, *a, **k):
        self._rec("plot", a, k)

    def fill_between(self, *a, **k):
        self._rec("fill_between", a, k)

    def axvline(self, *a, **k):
        self._rec("axvline", a, k)

    def legend(self, *a, **k):
        self._rec("legend", a, k)

    def set_title(self, *a, **k):
        self._rec("set_title", a, k)

    def set_xlabel(self, *a, **k):
        pass

    def set_ylabel(self, *a, **k):
        pass

    def grid(self, *a, **k):
        pass


class _Figure:
    def subplots_adjust(self, **k):
        pass


def install():
    """Register the stand-in as matplotlib / matplotlib.pyplot / matplotlib.colors; returns the pyplot module (``.saved``: paths
    passed to savefig, ``.axes``: every axes subplot() made)."""
    mpl, plt, colors = (types.ModuleType(m) for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors"))
    colors.TABLEAU_COLORS = {k: k for k in TABLEAU}
    plt.saved, plt.axes, plt.rcParams = [], [], {}
    plt.figure = lambda **k: _Figure()

    def subplot(*a):
        plt.axes.append(RecordingAxes())
        return plt.axes[-1]

    plt.subplot = subplot
    plt.savefig = lambda path: plt.saved.append(path)
    mpl.pyplot, mpl.colors = plt, colors
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.colors": colors})
    return plt


def uninstall():
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors"):
        sys.modules.pop(m, None)
