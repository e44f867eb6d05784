"""Synthetic light-curve generators for the BASELINE.json configurations.

Every generator returns a ragged CSR batch::

    offsets int64[B+1], t f64[sum N], flux f64[sum N], flux_err f64[sum N], band int32[sum N]

plus the passband table ``loglam f64[n_bands]`` (log10 of the effective wavelength, Angstrom).  Seeds
are fixed so the CPU reference and the GPU path see identical arrays.
"""
from __future__ import annotations

import numpy as np

LSST_LAMBDA = (3751.36, 4741.64, 6173.23, 7501.62, 8679.19, 9711.53)  # u g r i z y  (reference tests/test_aug.py:15-22)
LSST_KEYS = ("u", "g", "r", "i", "z", "y")
ZTF_LAMBDA = (4741.64, 6173.23)  # g r


def lsst_passband2lam(keys=LSST_KEYS, lams=LSST_LAMBDA):
    return {k: float(np.log10(v)) for k, v in zip(keys, lams)}


class CurveBatch:
    """Ragged batch of light curves in CSR layout."""

    def __init__(self, offsets, t, flux, flux_err, band, loglam):
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.t = np.ascontiguousarray(t, dtype=np.float64)
        self.flux = np.ascontiguousarray(flux, dtype=np.float64)
        self.flux_err = np.ascontiguousarray(flux_err, dtype=np.float64)
        self.band = np.ascontiguousarray(band, dtype=np.int32)
        self.loglam = np.ascontiguousarray(loglam, dtype=np.float64)

    @property
    def n_curves(self):
        return len(self.offsets) - 1

    @property
    def n_bands(self):
        return len(self.loglam)

    def curve(self, i):
        a, b = int(self.offsets[i]), int(self.offsets[i + 1])
        return self.t[a:b], self.flux[a:b], self.flux_err[a:b], self.band[a:b]

    def lengths(self):
        return np.diff(self.offsets)

    def slice(self, lo, hi):
        a, b = int(self.offsets[lo]), int(self.offsets[hi])
        return CurveBatch(self.offsets[lo : hi + 1] - a, self.t[a:b], self.flux[a:b], self.flux_err[a:b], self.band[a:b], self.loglam)

    def take(self, idx):
        idx = np.asarray(idx, dtype=np.int64)
        lens = self.lengths()[idx]
        offs = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        # source position of every gathered observation, without a Python loop over the curves (a shard of 10^5 curves is taken per call)
        gather = np.repeat(self.offsets[idx] - offs[:-1], lens) + np.arange(offs[-1], dtype=np.int64) if len(idx) else np.zeros(0, np.int64)
        return CurveBatch(offs, self.t[gather], self.flux[gather], self.flux_err[gather], self.band[gather], self.loglam)

    @staticmethod
    def from_curves(curves, loglam):
        """curves: iterable of (t, flux, flux_err, band_index)."""
        curves = list(curves)
        offs = np.zeros(len(curves) + 1, dtype=np.int64)
        np.cumsum([len(c[0]) for c in curves], out=offs[1:])
        cat = lambda k, dt: np.concatenate([np.asarray(c[k], dtype=dt) for c in curves]) if curves else np.zeros(0, dt)
        return CurveBatch(offs, cat(0, np.float64), cat(1, np.float64), cat(2, np.float64), cat(3, np.int32), loglam)


def _bazin(t, lam, A, t0, tr, tf):
    colour = 1.0 + 0.15 * (lam / 6000.0 - 1.0)
    with np.errstate(over="ignore"):   # exp overflow far before the rise: the quotient is exactly 0 there
        return colour * A * np.exp(-(t - t0) / tf) / (1.0 + np.exp(-(t - t0) / tr))


def _bazin_curve(rng, n, lams, span, t_start=59000.0, n_bumps=1, const=0.0):
    lams = np.asarray(lams)
    band = rng.integers(0, len(lams), size=n)
    t = t_start + np.sort(rng.uniform(0.0, span, size=n))
    flux = np.full(n, const)
    for _ in range(n_bumps):
        A = rng.lognormal(np.log(300.0), 0.5)
        t0 = t_start + rng.uniform(0.2 * span, 0.45 * span) if n_bumps == 1 else t_start + rng.uniform(0.05 * span, 0.9 * span)
        tr = rng.uniform(2.0, 6.0)
        tf = rng.uniform(15.0, 50.0)
        flux = flux + _bazin(t, lams[band], A, t0, tr, tf)
    err = rng.uniform(1.5, 8.0, size=n)
    flux = flux + rng.normal(0.0, err)
    return t, flux, err, band.astype(np.int32)


def plasticc_like(n_curves, n_points=100, first=0, ragged=False):
    """Config #3: 6 LSST bands, N=100 (or clip(Poisson(100),64,160)), span 200 d, seed [12345, i]."""
    curves = []
    for i in range(first, first + n_curves):
        rng = np.random.default_rng([12345, i])
        n = int(np.clip(rng.poisson(n_points), 64, 160)) if ragged else n_points
        curves.append(_bazin_curve(rng, n, LSST_LAMBDA, 200.0))
    return CurveBatch.from_curves(curves, np.log10(LSST_LAMBDA))


def ztf_like(n_curves, first=0, n_lo=20, n_hi=300):
    """Config #4: 2 ZTF bands, N=round(exp(U(ln 20, ln 300))), span 300 d, seed [12345, 4, i]."""
    curves = []
    for i in range(first, first + n_curves):
        rng = np.random.default_rng([12345, 4, i])
        n = int(round(np.exp(rng.uniform(np.log(n_lo), np.log(n_hi)))))
        curves.append(_bazin_curve(rng, n, ZTF_LAMBDA, 300.0))
    return CurveBatch.from_curves(curves, np.log10(ZTF_LAMBDA))


def ddf_like(n_curves, n_points=1024, first=0):
    """Config #5: 6 bands, long curves, span 3650 d, 3-6 Bazin bumps + constant, seed [12345, 5, i]."""
    curves = []
    for i in range(first, first + n_curves):
        rng = np.random.default_rng([12345, 5, i])
        curves.append(_bazin_curve(rng, n_points, LSST_LAMBDA, 3650.0, n_bumps=int(rng.integers(3, 7)), const=rng.uniform(20.0, 100.0)))
    return CurveBatch.from_curves(curves, np.log10(LSST_LAMBDA))


def readme_example(seed=0):
    """Config #1: reference README.md:45-58 with the unseeded noise replaced by RandomState(seed)."""
    lam = LSST_LAMBDA[:3]
    keys = LSST_KEYS[:3]
    n_per_band = 10
    n = n_per_band * 3
    t = np.linspace(0.0, n - 1, n)
    flux = 10.0 + np.sin(2 * t) + np.random.RandomState(seed).normal(0, 0.1, len(t))
    flux_err = 0.1 * np.ones_like(flux)
    passbands = np.tile(list(keys), n_per_band)
    return lsst_passband2lam(keys, lam), t, flux, flux_err, passbands


def reference_test_fixture():
    """The sample_data() of the reference's own test (tests/test_aug.py:14-29)."""
    p2l = lsst_passband2lam()
    n_per_band = 10
    n = n_per_band * len(p2l)
    t = np.linspace(0.0, n - 1, n)
    flux = 10.0 + np.sin(t)
    flux_err = np.ones_like(flux)
    passbands = np.tile(list(p2l), n_per_band)
    return p2l, t, flux, flux_err, passbands
