"""The fulu GP adapter re-typed on top of the installed scikit-learn.  TEST INFRASTRUCTURE ONLY.

This is the "live" oracle and the CPU baseline of ``bench.py``: it performs exactly the library
calls ``/root/reference/fulu/gp_aug.py:54-77,99-108`` and ``/root/reference/fulu/_base_aug.py:9-20,
111-112`` perform (StandardScaler -> GaussianProcessRegressor(optimizer="fmin_l_bfgs_b",
n_restarts_optimizer=5, normalize_y=True, random_state=42[, alpha]) -> predict(return_std=True)),
so every floating-point operation is executed by scikit-learn 1.9.0 / SciPy 1.18.1 / OpenBLAS — the
third-party code in which the reference's arithmetic lives (``/root/reference`` does not travel to
the GPU box; scikit-learn is part of the image on both boxes).  ``tests/test_oracle.py`` checks this
port against fixtures produced by the real reference class (``tests/golden/make_golden.py``).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may use it.
"""
from __future__ import annotations

import numpy as np
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, WhiteKernel
from sklearn.preprocessing import StandardScaler


def default_kernel():
    # fulu/gp_aug.py:29
    return ConstantKernel(1.0) * RBF([1.0, 1.0]) + WhiteKernel()


class SklearnGPAugmentation:
    """Same constructor / fit / predict / augmentation surface as fulu.GaussianProcessesAugmentation."""

    def __init__(self, passband2lam, kernel=None, use_err=False):
        self.passband2lam = passband2lam
        self.kernel = default_kernel() if kernel is None else kernel
        self.use_err = use_err
        self.ss = None
        self.reg = None

    def _features(self, t, passband):
        t = np.array(t)
        log_lam = np.array([self.passband2lam[i] for i in np.array(passband)])  # _base_aug.py:9-11
        return np.concatenate((t.reshape(-1, 1), log_lam.reshape(-1, 1)), axis=1)  # gp_aug.py:62

    def fit(self, t, flux, flux_err, passband):
        flux = np.array(flux)
        flux_err = np.array(flux_err)
        X = self._features(t, passband)
        with np.errstate(divide="ignore", invalid="ignore"):
            X_error = (flux_err / np.std(flux)).reshape(-1)  # gp_aug.py:64
        self.ss = StandardScaler()
        X_ss = self.ss.fit_transform(X)  # gp_aug.py:65-66
        kw = dict(kernel=self.kernel, optimizer="fmin_l_bfgs_b", n_restarts_optimizer=5, normalize_y=True, random_state=42)
        if self.use_err:
            kw["alpha"] = X_error**2  # gp_aug.py:72-73
        self.reg = GaussianProcessRegressor(**kw)
        self.reg.fit(X_ss, flux)  # gp_aug.py:77
        return self

    def predict(self, t, passband):
        X_ss = self.ss.transform(self._features(t, passband))  # gp_aug.py:101-104
        return self.reg.predict(X_ss, return_std=True)  # gp_aug.py:106

    def augmentation(self, t_min, t_max, n_obs=100):
        t, pb = [], []
        for band in tuple(self.passband2lam):  # _base_aug.py:14-20,111
            t += list(np.linspace(t_min, t_max, n_obs))
            pb += [band] * n_obs
        t, pb = np.array(t), np.array(pb)
        flux, err = self.predict(t, pb)
        return t, flux, err, pb


def count_evals(aug, t, flux, flux_err, passband):
    """Fit while counting log_marginal_likelihood(eval_gradient=True) calls (E of SURVEY 8d)."""
    import sklearn.gaussian_process._gpr as gpr_mod

    n = [0]
    orig = gpr_mod.GaussianProcessRegressor.log_marginal_likelihood

    def counted(self, theta=None, eval_gradient=False, clone_kernel=True):
        if eval_gradient:
            n[0] += 1
        return orig(self, theta, eval_gradient, clone_kernel)

    gpr_mod.GaussianProcessRegressor.log_marginal_likelihood = counted
    try:
        aug.fit(t, flux, flux_err, passband)
    finally:
        gpr_mod.GaussianProcessRegressor.log_marginal_likelihood = orig
    return n[0]
