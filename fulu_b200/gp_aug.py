"""Drop-in for fulu.GaussianProcessesAugmentation (reference: fulu/gp_aug.py:10-108, fulu/_base_aug.py:23-114).

Same constructor, ``fit`` / ``predict`` / ``augmentation`` signatures, argument meaning, return types and error behaviour; the
work runs on the GPU through libfulu_gp (one-curve batches).  For many curves use :func:`fulu_b200.fit_augment_batch`.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch
from sklearn.gaussian_process.kernels import RBF, ConstantKernel as C, WhiteKernel

from .datasets import CurveBatch
from .plotting import LcPlotter
from .engine import (STATUS_AT_BOUND, STATUS_BAD_INPUT, STATUS_NEG_VARIANCE, STATUS_NOT_CONVERGED, STATUS_NOT_PD, DeviceBatch, GPEngine,
                     kernel_spec_from_sklearn)


class ConvergenceWarning(UserWarning):
    """Raised where sklearn raises sklearn.exceptions.ConvergenceWarning (optimiser ended abnormally / theta on a bound)."""


def add_log_lam(passband, passband2lam):
    """fulu/_base_aug.py:9-11 (raises KeyError for an unknown passband, like the reference)."""
    return np.array([passband2lam[i] for i in passband])


def create_aug_data(t_min, t_max, passband_ids, n_obs=1000):
    """Band-major augmentation grid, fulu/_base_aug.py:14-20."""
    t = []
    passband = []
    for band in passband_ids:
        t += list(np.linspace(t_min, t_max, n_obs))
        passband += [band] * n_obs
    return np.array(t), np.array(passband)


class _Scaler:
    """What the reference exposes as ``aug.ss`` (a fitted StandardScaler): mean_, scale_, transform."""

    def __init__(self, mean, scale):
        self.mean_, self.scale_ = np.asarray(mean), np.asarray(scale)
        self.var_ = self.scale_ ** 2

    def transform(self, X):
        return (np.asarray(X, np.float64) - self.mean_) / self.scale_


class _FittedGP:
    """What the reference exposes as ``aug.reg``: kernel_.theta, log_marginal_likelihood_value_, log_marginal_likelihood()."""

    def __init__(self, owner, theta, lml, n_eval):
        self._owner = owner
        # a real sklearn kernel object, as in the reference (reg.kernel_ = clone of the prior with the fitted theta, _gpr.py:335-340);
        # theta is in the kernel object's own (sklearn) order
        self.kernel_ = owner.kernel.clone_with_theta(np.asarray(theta, np.float64))
        self.kernel = owner.kernel
        self.log_marginal_likelihood_value_ = float(lml)
        self.n_eval_ = int(n_eval)

    def log_marginal_likelihood(self, theta=None, eval_gradient=False, clone_kernel=True):
        if theta is None:
            if eval_gradient:
                raise ValueError("Gradient can only be evaluated for theta!=None")
            return self.log_marginal_likelihood_value_
        o = self._owner
        th = torch.as_tensor(np.ascontiguousarray(o._spec.from_sklearn(np.asarray(theta, np.float64).reshape(1, -1))))
        out = o._engine.lml(o._batch, th)
        lml = float(out["lml"].cpu()[0])
        return (lml, o._spec.to_sklearn(out["grad"].cpu().numpy()[0])) if eval_gradient else lml


class GaussianProcessesAugmentation:
    """
    Light Curve Augmentation based on Gaussian Processes Regression (GPU, batched kernels underneath).

    Parameters:
    -----------
    passband2lam : dict
        A dictionary, where key is a passband ID and value is Log10 of its wave length.
    kernel : kernel object from sklearn
        ``C * [Matern] * RBF([l_t, l_lambda]) + [Matern | RationalQuadratic] + WhiteKernel`` in any order (the reference's default
        ``C(1.0)*RBF([1.0, 1.0]) + WhiteKernel()``, its docstring kernel ``C(1.0)*RBF([1, 1]) + Matern() + WhiteKernel()``, the
        notebook's ``C*Matern()*RBF([1, 1]) + Matern() + WhiteKernel()``, ...; see ``engine.kernel_spec_from_sklearn``).  Initial
        values and bounds are honoured; other kernels raise NotImplementedError (there is no CPU fallback).
    use_err : bool
        Flag responsible for using flux error by GP Regressor.

    The constructor is the reference's, argument for argument (fulu/gp_aug.py:29; tests/test_abi.py compares the signatures).  The GPU
    is the process's current CUDA device; set the class or instance attribute ``device`` before ``fit`` to choose another one.
    """

    device = None   # torch device (or None = current CUDA device) the engine of this object is created on

    def __init__(self, passband2lam, kernel=C(1.0) * RBF([1.0, 1.0]) + WhiteKernel(), use_err=False):
        self.passband2lam = passband2lam
        self.plotter = LcPlotter(passband2lam)   # fulu/_base_aug.py:37; matplotlib is imported when something is drawn
        self.t_train = None
        self.flux_train = None
        self.flux_err_train = None
        self.passband_train = None
        self.ss = None
        self.reg = None
        self.kernel = kernel
        self.use_err = use_err
        self._spec = kernel_spec_from_sklearn(kernel)
        self._engine = None
        self._batch = None
        self._theta = None

    # -- helpers
    def _band_index(self, passband):
        lut = {k: i for i, k in enumerate(self.passband2lam)}
        return np.array([lut[b] for b in passband], dtype=np.int32)

    def fit(self, t, flux, flux_err, passband):
        """Fit an augmentation model (same arguments as fulu/gp_aug.py:37-78)."""
        self.t_train = np.asarray(t)
        self.flux_train = np.asarray(flux)
        self.flux_err_train = np.asarray(flux_err)
        self.passband_train = np.asarray(passband)
        t = np.array(t, dtype=np.float64).reshape(-1)
        flux = np.array(flux, dtype=np.float64).reshape(-1)
        flux_err = np.array(flux_err, dtype=np.float64).reshape(-1)
        add_log_lam(np.array(passband), self.passband2lam)  # KeyError for unknown bands, as in the reference
        band = self._band_index(np.array(passband))
        if not (np.all(np.isfinite(t)) and np.all(np.isfinite(flux))):
            raise ValueError("Input contains NaN or infinity.")  # sklearn validate_data (_gpr.py:257-265)
        loglam = np.array([self.passband2lam[k] for k in self.passband2lam], dtype=np.float64)
        if self._engine is None:
            self._engine = GPEngine(self.device)
        host = CurveBatch(np.array([0, len(t)], np.int64), t, flux, flux_err, band, loglam)
        self._batch = DeviceBatch.from_host(host, self._engine.device, use_err=self.use_err, family=self._spec.family)
        fit = self._engine.fit(self._batch, self._spec)
        status = int(fit["status"].cpu()[0])
        if status & STATUS_BAD_INPUT:
            raise ValueError("Input contains NaN, infinity or an unusable noise term.")
        if status & STATUS_NOT_PD:   # the final factorisation of sklearn's fit fails the same way (_gpr.py:349-361)
            raise np.linalg.LinAlgError("The kernel is not returning a positive definite matrix. Try gradually increasing the 'alpha' "
                                        "parameter of your GaussianProcessRegressor estimator.")
        theta = self._spec.to_sklearn(fit["theta"].cpu().numpy()[0])   # reg.kernel_.theta is in the kernel object's order
        self._theta = fit["theta"]
        scaler = fit["scaler"].cpu().numpy()[0]     # the fit's own outputs (fulu_gp_fit_batch): no second call
        self.ss = _Scaler(scaler[[0, 2]], scaler[[1, 3]])
        self.reg = _FittedGP(self, theta, float(fit["lml"].cpu()[0]), int(fit["n_eval"].cpu()[0]))
        self.reg._y_train_mean, self.reg._y_train_std = (float(v) for v in fit["ynorm"].cpu().numpy()[0])
        if status & STATUS_NOT_CONVERGED:
            warnings.warn("lbfgs failed to converge: abnormal termination in the line search or iteration limit reached.", ConvergenceWarning)
        if status & STATUS_AT_BOUND:
            warnings.warn("The optimal value found for a hyperparameter is close to its bound; changing the bound might give a better fit.",
                          ConvergenceWarning)
        return self

    def predict(self, t, passband):
        """Apply the augmentation model to the given observation time moments (fulu/gp_aug.py:80-108)."""
        if self._batch is None:
            raise RuntimeError("fit must be called before predict")
        t = np.array(t, dtype=np.float64).reshape(-1)
        passband = np.array(passband)
        add_log_lam(passband, self.passband2lam)
        band = self._band_index(passband)
        out = self._engine.predict_points(self._batch, self._theta, np.array([0, len(t)], np.int64), t, band)
        status = int(out["status"].cpu()[0])
        if status & STATUS_NOT_PD:
            raise np.linalg.LinAlgError("The kernel is not returning a positive definite matrix. Try gradually increasing the 'alpha' "
                                        "parameter of your GaussianProcessRegressor estimator.")  # _gpr.py:353-361
        if status & STATUS_NEG_VARIANCE:
            warnings.warn("Predicted variances smaller than 0. Setting those variances to 0.")  # _gpr.py:485-491
        return out["mean"].cpu().numpy(), out["std"].cpu().numpy()

    def augmentation(self, t_min, t_max, n_obs=100):
        """The light curve augmentation (fulu/_base_aug.py:88-114): band-major grid of n_obs points per passband."""
        t_aug, passband_aug = create_aug_data(t_min, t_max, tuple(self.passband2lam), n_obs)
        flux_aug, flux_err_aug = self.predict(t_aug, passband_aug)
        return t_aug, flux_aug, flux_err_aug, passband_aug

    def _peak_on_device(self, t_min, t_max, n_obs):
        """Grid augmentation, band-summed curve and its peak through the batched grid / peak kernels
        (fulu_gp_predict_grid_batch, fulu_gp_peak_batch) -> (t_aug, flux_aug, flux_err_aug, passband_aug), (t_grid, sum), peak_t."""
        if self._batch is None:
            raise RuntimeError("fit must be called before plot")
        lo = torch.tensor([float(t_min)], dtype=torch.float64)
        hi = torch.tensor([float(t_max)], dtype=torch.float64)
        grid = self._engine.predict_grid(self._batch, self._theta, n_obs, lo, hi)
        if int(grid["status"].cpu()[0]) & STATUS_NOT_PD:
            raise np.linalg.LinAlgError("The kernel is not returning a positive definite matrix.")
        pk = self._engine.peak(self._batch, grid["mean"], lo, hi, want_sum=True)
        t_aug, passband_aug = create_aug_data(t_min, t_max, tuple(self.passband2lam), n_obs)
        approx = (t_aug, grid["mean"].cpu().numpy().reshape(-1), grid["std"].cpu().numpy().reshape(-1), passband_aug)
        return approx, (t_aug[:n_obs], pk["sum_flux"].cpu().numpy()[0]), float(pk["peak_t"].cpu()[0])

    def _plot_passband(self, passband, ax, approx=None):
        """Train points of one passband and, if given, its approximation with the one-sigma band (fulu/_base_aug.py:116-158)."""
        self.plotter.errorbar_passband(self.t_train, self.flux_train, self.flux_err_train, self.passband_train, passband, ax=ax)
        if approx:
            self.plotter._draw_approx(ax, *approx, passband)

    def plot(self, *, plot_approx=True, n_approx=1000, passband=None, ax=None, true_peak=None, plot_peak=False, title="", save=None):
        """Train points with errors for all passbands (or one) on one graph, the approximation curves, optionally the true and the
        predicted peak; same keywords and drawing order as fulu/_base_aug.py:160-230.  The approximation, the band-summed curve
        and its peak come from the device (grid and peak kernels); matplotlib only draws."""
        if ax is None:
            ax = self.plotter._ax_adjust()
        approx, (t_sum, f_sum), peak_t = self._peak_on_device(np.min(self.t_train), np.max(self.t_train), n_approx)
        for band in (self.passband2lam if passband is None else (passband,)):
            self._plot_passband(band, ax, approx if plot_approx else None)
        if true_peak is not None:
            self.plotter.plot_true_peak(true_peak=true_peak, ax=ax)
        if plot_peak:
            self.plotter._draw_sum(ax, t_sum, f_sum)    # LcPlotter.plot_sum_passbands (plotting.py:78-100) with the device's band sum
            self.plotter._finish(ax, "", None)
            self.plotter._draw_peak(ax, peak_t)         # LcPlotter.plot_peak (plotting.py:102-123) with the device's arg-max
            self.plotter._finish(ax, "", None)
        return self.plotter._finish(ax, title, save, title_first=True)
