"""CPU oracle for the fulu GP-augmentation hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product (``fulu_b200``) never does: it
fails loudly if its CUDA library is missing.

What this restates
------------------
The reference path is ``fulu.GaussianProcessesAugmentation.fit/.predict/.augmentation``
(``/root/reference/fulu/gp_aug.py:37-108``, ``/root/reference/fulu/_base_aug.py:9-20,88-114``).
fulu itself is a thin adapter; the arithmetic lives in third-party code that is NOT vendored
under ``/root/reference``:

* scikit-learn (unpinned in ``/root/reference/setup.cfg:25``; the image has **1.9.0**):
  ``GaussianProcessRegressor`` (``sklearn/gaussian_process/_gpr.py``), kernels
  (``sklearn/gaussian_process/kernels.py``), ``StandardScaler`` (``sklearn/preprocessing/_data.py``),
* SciPy **1.18.1**: ``scipy.optimize.minimize(method="L-BFGS-B")`` (L-BFGS-B 3.0), LAPACK
  ``potrf/potrs/trtrs`` through ``scipy.linalg``.

Two oracles live here:

1. the *closed-form* NumPy restatement (``standardize`` … ``predict``) of the published algorithm
   (Rasmussen & Williams Alg. 2.1 / Eq. 5.9 as implemented at the sklearn lines cited on each
   function).  It is the spec the CUDA kernels are written against.
2. ``oracle.fulu_sklearn_port`` — the fulu adapter re-typed call-for-call on top of the *installed*
   scikit-learn, i.e. live execution of the very library the reference executes.

Pinning: the reference's own test (``/root/reference/tests/test_aug.py:52-67``) pins only shapes and
signs, so numerical parity is pinned against outputs of the reference itself run in the build
container: ``tests/golden/make_golden.py`` imports ``/root/reference/fulu/gp_aug.py`` and stores
inputs/outputs under ``tests/golden/*.npz``; ``tests/test_oracle.py`` checks both oracles against
those fixtures (and against live sklearn, which is part of the image on every box).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular
from scipy.optimize import minimize

LOG_BOUND = float(np.log(1e5))  # sklearn default hyperparameter bounds (1e-5, 1e5) in log space

# Kernel families (same integer codes as include/fulu_gp.h: FULU_GP_KERNEL(f1, ex) = f1 | ex << 4).
#   K = c * [M1] * RBF([lt, ll]) + [EX] + White;  f1 / ex in {0 none, 1 Matern 1/2, 2 Matern 3/2, 3 Matern 5/2, 4 RationalQuadratic}
#   theta in CANONICAL order: [c, (l_1), lt, ll, (EX: l | alpha, l), s2]   (log-transformed)
F_NONE, F_MAT05, F_MAT15, F_MAT25, F_RQ = 0, 1, 2, 3, 4


def family_code(f1=F_NONE, ex=F_NONE):
    return f1 | (ex << 4)


KERNEL_RBF_WHITE = family_code()  # C(1.0) * RBF([1, 1]) + WhiteKernel()                      theta = [c, lt, ll, s2]
KERNEL_RBF_MATERN_WHITE = family_code(ex=F_MAT15)  # C * RBF([1, 1]) + Matern() + White       theta = [c, lt, ll, lm, s2]
KERNEL_MATERN_RBF_MATERN_WHITE = family_code(F_MAT15, F_MAT15)  # C * Matern() * RBF([1, 1]) + Matern() + White  [c, l1, lt, ll, lm, s2]
KERNEL_RBF_RQ_WHITE = family_code(ex=F_RQ)  # C * RBF([1, 1]) + RationalQuadratic() + White   theta = [c, lt, ll, alpha, lq, s2]


def family_parts(kind):
    return kind & 15, kind >> 4


def n_params(kind):
    f1, ex = family_parts(kind)
    return 4 + (1 if f1 else 0) + (2 if ex == F_RQ else 1 if ex else 0)


class _NParams(dict):
    def __missing__(self, kind):
        return n_params(kind)


KERNEL_NPARAMS = _NParams()


# --------------------------------------------------------------------------------------------------
# a1/a2/a3: features, StandardScaler, y normalisation, noise diagonal
# --------------------------------------------------------------------------------------------------
def add_log_lam(passband, passband2lam):
    """Per-observation lookup of log10(lambda).  fulu/_base_aug.py:9-11."""
    return np.array([passband2lam[b] for b in passband], dtype=np.float64)


def scaler_fit(col):
    """StandardScaler statistics of one feature column.

    sklearn/utils/extmath.py:1203-1262 (corrected two-pass variance, ddof=0) and
    sklearn/preprocessing/_data.py:83-98,1073-1078 (near-constant column -> scale 1).
    """
    col = np.asarray(col, dtype=np.float64)
    n = col.shape[0]
    s = np.sum(col)
    mean = s / n
    temp = col - s / n
    correction = np.sum(temp)
    unnorm = np.sum(temp * temp) - correction**2 / n
    var = unnorm / n
    eps = np.finfo(np.float64).eps
    constant = var <= n * eps * var + (n * mean * eps) ** 2
    scale = np.sqrt(var)
    if constant or scale < 10 * eps:
        scale = 1.0
    return float(mean), float(scale)


def standardize(t, log_lam):
    """X = [t, log_lam] -> StandardScaler().fit_transform(X).  fulu/gp_aug.py:62,65-66."""
    mt, st = scaler_fit(t)
    ml, sl = scaler_fit(log_lam)
    X = np.stack([(np.asarray(t, np.float64) - mt) / st, (np.asarray(log_lam, np.float64) - ml) / sl], axis=1)
    return X, np.array([mt, ml]), np.array([st, sl])


def normalize_y(flux):
    """normalize_y=True branch of GaussianProcessRegressor.fit.  sklearn/gaussian_process/_gpr.py:275-280."""
    flux = np.asarray(flux, dtype=np.float64)
    mean = float(np.mean(flux))
    std = float(np.std(flux))
    if std == 0.0:
        std = 1.0
    return (flux - mean) / std, mean, std


def noise_diag(flux, flux_err, use_err):
    """alpha of the regressor: 1e-10 by default, (flux_err/std(flux))**2 with use_err.

    fulu/gp_aug.py:64,72-73; sklearn/gaussian_process/_gpr.py:215.
    """
    n = len(flux)
    if not use_err:
        return np.full(n, 1e-10)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (np.asarray(flux_err, np.float64) / np.std(np.asarray(flux, np.float64))) ** 2


# --------------------------------------------------------------------------------------------------
# a4: kernel and its gradient
# --------------------------------------------------------------------------------------------------
def _sqdist(A, B):
    d = A[:, None, :] - B[None, :, :]
    return np.einsum("ijk,ijk->ij", d, d)


def _iso_factor(kind, D2, params, eval_gradient):
    """One isotropic stationary factor from squared RAW distances D2: value and d/dlog(hyperparameter) list.

    Matern nu = 1/2, 3/2, 5/2: sklearn/gaussian_process/kernels.py:1716-1729 (value), :1757-1774 (gradient);
    RationalQuadratic: :1917-1947 (theta order alpha, length_scale: hyperparameters sort by name, :280-288).
    """
    if kind == F_RQ:
        alpha, ell = params
        tmp = D2 / (2.0 * alpha * ell**2)
        base = 1.0 + tmp
        K = base**-alpha
        if not eval_gradient:
            return K, []
        g_len = D2 * K / (ell**2 * base)
        g_alpha = K * (-alpha * np.log(base) + D2 / (2.0 * ell**2 * base))
        return K, [g_alpha, g_len]
    (ell,) = params
    S = D2 / ell**2
    r = np.sqrt(S)
    if kind == F_MAT05:
        K = np.exp(-r)
        return K, ([K * r] if eval_gradient else [])
    if kind == F_MAT15:
        a = np.sqrt(3.0) * r
        K = (1.0 + a) * np.exp(-a)
        return K, ([3.0 * S * np.exp(-np.sqrt(3.0 * S))] if eval_gradient else [])
    if kind == F_MAT25:
        a = np.sqrt(5.0) * r
        K = (1.0 + a + a**2 / 3.0) * np.exp(-a)
        tmp = np.sqrt(5.0 * S)
        return K, ([5.0 / 3.0 * S * (tmp + 1.0) * np.exp(-tmp)] if eval_gradient else [])
    raise NotImplementedError(kind)


def _split_theta(theta, kind):
    f1, ex = family_parts(kind)
    th = np.exp(np.asarray(theta, np.float64))
    k = 1
    p1 = None
    if f1:
        p1 = (th[k],)
        k += 1
    ell = th[k:k + 2]
    k += 2
    pe = None
    if ex == F_RQ:
        pe = (th[k], th[k + 1])
        k += 2
    elif ex:
        pe = (th[k],)
        k += 1
    assert k == len(th) - 1, "theta length does not match the kernel family"
    return f1, ex, th[0], p1, ell, pe, th[-1]


def _kernel_pairs(XA, XB, theta, kind, eval_gradient):
    """Noise-free covariance between two point sets and (optionally) its gradient planes in canonical theta order (no White)."""
    f1, ex, c, p1, ell, pe, _ = _split_theta(theta, kind)
    E = np.exp(-0.5 * _sqdist(XA / ell, XB / ell))
    D2 = _sqdist(XA, XB) if (f1 or ex) else None
    grads = []
    if f1:
        M, gM = _iso_factor(f1, D2, p1, eval_gradient)
        T = c * E * M
    else:
        M, gM = 1.0, []
        T = c * E
    K = T
    if eval_gradient:
        grads.append(T)  # d/dlog c (kernels.py:1290-1292 with the product rule :967-969)
        grads += [c * E * g for g in gM]
        for d in range(2):  # anisotropic RBF (kernels.py:1581-1584)
            D = (XA[:, d][:, None] - XB[:, d][None, :]) ** 2 / ell[d] ** 2
            grads.append(T * D)
    if ex:
        A, gA = _iso_factor(ex, D2, pe, eval_gradient)
        K = K + A
        grads += gA
    return K, grads


def kernel_train(X, theta, kind=KERNEL_RBF_WHITE, eval_gradient=False):
    """K(X, X) (without the regressor's alpha) and dK/dtheta_k, canonical theta order.

    sklearn/gaussian_process/kernels.py:1558-1585 (anisotropic RBF), :1278-1292 (Constant), :1403-1413 (White),
    :964-969 (Product rule), :866-869 (Sum rule), :1716-1784 (Matern), :1917-1947 (RationalQuadratic).
    """
    n = X.shape[0]
    s2 = np.exp(np.asarray(theta, np.float64)[-1])
    K, grads = _kernel_pairs(X, X, theta, kind, eval_gradient)
    # pdist/squareform + fill_diagonal(K, 1) in sklearn: exact ones on the diagonal of every stationary factor, exact zeros on
    # the diagonal of its gradient
    c = np.exp(np.asarray(theta, np.float64)[0])
    K = K.copy()
    K[np.diag_indices(n)] = c + 1.0 if family_parts(kind)[1] else c
    if eval_gradient:
        for k, G in enumerate(grads):
            G[np.diag_indices(n)] = c if k == 0 else 0.0
    K = K + s2 * np.eye(n)
    if eval_gradient:
        grads.append(s2 * np.eye(n))
        return K, grads
    return K


def kernel_cross(Xq, X, theta, kind=KERNEL_RBF_WHITE):
    """K(X*, X): WhiteKernel contributes zeros.  kernels.py:1569-1570, :1418-1419."""
    return _kernel_pairs(Xq, X, theta, kind, False)[0]


def kernel_diag(theta, m, kind=KERNEL_RBF_WHITE):
    """kernel_.diag(X*) = c + noise_level (+1 for an added Matern / RQ term).  kernels.py:890,990,1315-1319,1438-1440."""
    theta = np.asarray(theta, np.float64)
    v = np.exp(theta[0]) + np.exp(theta[-1])
    if family_parts(kind)[1]:
        v = v + 1.0
    return np.full(m, v)


# ---- sklearn kernel object <-> family (used by the tests to run live sklearn on the same kernel)
def sklearn_kernel(kind, theta=None):
    """The sklearn kernel object of a family, hyperparameters in CANONICAL order == sklearn's tree order for this construction."""
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, RationalQuadratic, WhiteKernel

    f1, ex = family_parts(kind)
    nu = {F_MAT05: 0.5, F_MAT15: 1.5, F_MAT25: 2.5}
    k = ConstantKernel(1.0)
    if f1:
        k = k * Matern(nu=nu[f1])
    k = k * RBF([1.0, 1.0])
    if ex == F_RQ:
        k = k + RationalQuadratic()
    elif ex:
        k = k + Matern(nu=nu[ex])
    k = k + WhiteKernel()
    return k if theta is None else k.clone_with_theta(np.asarray(theta, np.float64))


# --------------------------------------------------------------------------------------------------
# a5: log marginal likelihood and gradient
# --------------------------------------------------------------------------------------------------
def lml_and_grad(X, y, alpha, theta, kind=KERNEL_RBF_WHITE, return_scale=False):
    """log p(y|X,theta) and d/dtheta.  sklearn/gaussian_process/_gpr.py:583-651.

    Non-PD K returns (-inf, zeros) exactly like _gpr.py:590-593.
    """
    n = X.shape[0]
    K, dK = kernel_train(X, theta, kind, eval_gradient=True)
    K[np.diag_indices(n)] += alpha
    try:
        L = cholesky(K, lower=True, check_finite=False)
    except np.linalg.LinAlgError:
        z = np.zeros(len(theta))
        return (-np.inf, z, z) if return_scale else (-np.inf, z)
    a = cho_solve((L, True), y, check_finite=False)
    lml = -0.5 * float(y @ a) - float(np.log(np.diag(L)).sum()) - n / 2.0 * np.log(2.0 * np.pi)
    W = np.outer(a, a) - cho_solve((L, True), np.eye(n), check_finite=False)
    grad = np.array([0.5 * np.einsum("ij,ji->", W, G) for G in dK])
    if return_scale:
        # condition-aware magnitude of each gradient component: the sum of |terms| of the trace
        return lml, grad, np.array([0.5 * np.abs(W * G.T).sum() for G in dK])
    return lml, grad


# --------------------------------------------------------------------------------------------------
# a6: restarts + L-BFGS-B
# --------------------------------------------------------------------------------------------------
def start_points(nparams, n_restarts=5, seed=42, theta0=None):
    """Start 0 = kernel defaults (theta 0), then RandomState(42).uniform over the log bounds.

    sklearn/gaussian_process/_gpr.py:251,312-333.
    """
    rng = np.random.RandomState(seed)
    lo = np.full(nparams, -LOG_BOUND)
    hi = np.full(nparams, LOG_BOUND)
    pts = [np.zeros(nparams) if theta0 is None else np.asarray(theta0, np.float64)]
    for _ in range(n_restarts):
        pts.append(rng.uniform(lo, hi))
    return np.array(pts)


def fit_theta(X, y, alpha, kind=KERNEL_RBF_WHITE, starts=None, return_all=False):
    """Six bound-constrained L-BFGS-B runs, best -LML wins.  _gpr.py:299-340,658-668."""
    p = n_params(kind)
    if starts is None:
        starts = start_points(p)
    bounds = [(-LOG_BOUND, LOG_BOUND)] * p
    n_eval = [0]

    def obj(th):
        n_eval[0] += 1
        lml, g = lml_and_grad(X, y, alpha, th, kind)
        return -lml, -g

    runs = []
    for s in starts:
        res = minimize(obj, s, method="L-BFGS-B", jac=True, bounds=bounds)
        runs.append((res.x, res.fun, res.nfev, res.status))
    best = int(np.argmin([r[1] for r in runs]))
    out = (runs[best][0], -runs[best][1], n_eval[0])
    if return_all:
        return out + (runs,)
    return out


# --------------------------------------------------------------------------------------------------
# a7/a9: final factorisation and prediction
# --------------------------------------------------------------------------------------------------
def factor(X, y, alpha, theta, kind=KERNEL_RBF_WHITE):
    """L_ = chol(K(theta*) + diag(alpha)), alpha_ = K^-1 y.  _gpr.py:349-367."""
    K = kernel_train(X, theta, kind)
    K[np.diag_indices(X.shape[0])] += alpha
    L = cholesky(K, lower=True, check_finite=False)
    return L, cho_solve((L, True), y, check_finite=False)


def predict(Xq, X, L, a, theta, y_mean, y_std, kind=KERNEL_RBF_WHITE):
    """Predictive mean and std in flux units.  _gpr.py:444-500."""
    Ks = kernel_cross(Xq, X, theta, kind)
    mean = y_std * (Ks @ a) + y_mean
    V = solve_triangular(L, Ks.T, lower=True, check_finite=False)
    var = kernel_diag(theta, Xq.shape[0], kind) - np.einsum("ij,ij->j", V, V)
    var = np.where(var < 0, 0.0, var)
    return mean, np.sqrt(var * y_std**2)


def create_aug_data(t_min, t_max, passband_ids, n_obs):
    """Band-major augmentation grid.  fulu/_base_aug.py:14-20."""
    t, pb = [], []
    for band in passband_ids:
        t += list(np.linspace(t_min, t_max, n_obs))
        pb += [band] * n_obs
    return np.array(t), np.array(pb)


class ClosedFormGP:
    """The whole path for one curve with the closed-form NumPy oracle (same surface as fulu's class)."""

    def __init__(self, passband2lam, kind=KERNEL_RBF_WHITE, use_err=False):
        self.passband2lam = passband2lam
        self.kind = kind
        self.use_err = use_err

    def prepare(self, t, flux, flux_err, passband):
        self.X, self.x_mean, self.x_scale = standardize(np.asarray(t, np.float64), add_log_lam(passband, self.passband2lam))
        self.y, self.y_mean, self.y_std = normalize_y(flux)
        self.alpha = noise_diag(flux, flux_err, self.use_err)
        return self

    def fit(self, t, flux, flux_err, passband, theta=None):
        self.prepare(t, flux, flux_err, passband)
        if theta is None:
            self.theta, self.lml, self.n_eval = fit_theta(self.X, self.y, self.alpha, self.kind)
        else:
            self.theta = np.asarray(theta, np.float64)
            self.lml = lml_and_grad(self.X, self.y, self.alpha, self.theta, self.kind)[0]
            self.n_eval = 1
        self.L, self.a = factor(self.X, self.y, self.alpha, self.theta, self.kind)
        return self

    def predict(self, t, passband):
        ll = add_log_lam(passband, self.passband2lam)
        Xq = np.stack([(np.asarray(t, np.float64) - self.x_mean[0]) / self.x_scale[0], (ll - self.x_mean[1]) / self.x_scale[1]], axis=1)
        return predict(Xq, self.X, self.L, self.a, self.theta, self.y_mean, self.y_std, self.kind)

    def augmentation(self, t_min, t_max, n_obs=100):
        t_aug, pb_aug = create_aug_data(t_min, t_max, tuple(self.passband2lam), n_obs)
        m, s = self.predict(t_aug, pb_aug)
        return t_aug, m, s, pb_aug


# --------------------------------------------------------------------------------------------------
# 8f-3: the steps either side of the path (checkers for fulu_gp_ingest_table / fulu_gp_peak_batch)
# --------------------------------------------------------------------------------------------------
def table_to_csr(object_id):
    """Rows grouped by object -> (offsets, ids): what a pandas groupby(sort=False) on the id column yields for such a table."""
    object_id = np.asarray(object_id, np.int64)
    n = len(object_id)
    if n == 0:
        return np.zeros(1, np.int64), np.zeros(0, np.int64)
    heads = np.flatnonzero(np.r_[True, object_id[1:] != object_id[:-1]])
    return np.r_[heads, n].astype(np.int64), object_id[heads]


def band_sum_peak(t_aug, flux_aug, passband_aug):
    """LcPlotter.plot_sum_passbands / plot_peak (fulu/plotting.py:90-94, 111-115): the augmented curve summed over the
    passbands per time, and the time of its maximum.  Executes the same pandas calls."""
    import pandas as pd

    frame = pd.DataFrame({"time": np.asarray(t_aug), "flux": np.asarray(flux_aug), "passband": np.asarray(passband_aug)})  # plotting.py:14-23
    curve = frame[["time", "flux"]].groupby("time", as_index=False).sum()  # plotting.py:114
    peak = curve["time"][curve["flux"].argmax()]  # plotting.py:115
    return curve["time"].values, curve["flux"].values, float(peak)
