"""Host side of the batched GP path: device-resident ragged batches and the calls into libfulu_gp.

PyTorch is plumbing here (device memory, streams, pinned host buffers, torch.distributed for sharding); every
floating-point operation of the path runs in the hand-written CUDA kernels behind the C ABI (include/fulu_gp.h).
"""
from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _capi
from .datasets import CurveBatch

LOG_BOUND = math.log(1e5)  # sklearn's default (1e-5, 1e5) hyperparameter bounds, log space (kernels.py:1236,1366,1511)

# Covariance families (include/fulu_gp.h): K = c * [M1] * RBF([lt, ll]) + [EX] + White, code = f1 | ex << 4
F_NONE, F_MAT05, F_MAT15, F_MAT25, F_RQ = 0, 1, 2, 3, 4


def family_code(f1=F_NONE, ex=F_NONE):
    return f1 | (ex << 4)


KERNEL_RBF_WHITE = family_code()  # C * RBF([lt, ll]) + White                           (fulu/gp_aug.py:29 default)
KERNEL_RBF_MATERN_WHITE = family_code(ex=F_MAT15)  # C * RBF + Matern() + White         (fulu/gp_aug.py:24, tests/test_aug.py:37)
KERNEL_MATERN_RBF_MATERN_WHITE = family_code(F_MAT15, F_MAT15)  # C * Matern() * RBF + Matern() + White  (plotting.ipynb:33)
KERNEL_RBF_RQ_WHITE = family_code(ex=F_RQ)  # C * RBF + RationalQuadratic() + White     (fulu/gp_aug.py:5)
# families compiled into libfulu_gp (fulu_b200/build.py FAMILIES)
COMPILED_FAMILIES = {family_code(*fe) for fe in ((0, 0), (0, 2), (2, 2), (0, 4), (0, 1), (0, 3), (2, 0))}

STATUS_BAD_INPUT, STATUS_NOT_PD, STATUS_NOT_CONVERGED, STATUS_AT_BOUND, STATUS_NEG_VARIANCE = 1, 2, 4, 8, 16


def family_n_params(family):
    f1, ex = family & 15, family >> 4
    return 4 + (1 if f1 else 0) + (2 if ex == F_RQ else 1 if ex else 0)


@dataclass
class KernelSpec:
    """The covariance family and its optimiser configuration (what sklearn derives from the kernel object).

    theta0 / lower / upper are in the library's CANONICAL order [c, (l_1), lt, ll, (ex: l | alpha, l), s2]; `perm` maps it onto
    the order sklearn uses for the kernel object the spec was built from: theta_canonical = theta_sklearn[perm] (identity for
    kernels written as C*[Matern*]RBF + [Matern|RQ] + White).  The restart points are drawn in sklearn's order, as sklearn
    draws them, and then permuted."""

    family: int = KERNEL_RBF_WHITE
    theta0: np.ndarray = None  # start 0 = the kernel's own hyperparameters (log)
    lower: np.ndarray = None
    upper: np.ndarray = None
    n_restarts: int = 5  # fulu/gp_aug.py:69
    seed: int = 42  # fulu/gp_aug.py:69 random_state
    perm: np.ndarray = None
    polish: int = None  # continuation runs from the best point (see GPEngine.fit); default 0 for the default family, 1 otherwise
    extra_starts: int = 0  # insurance starts beyond sklearn's 1 + n_restarts, drawn from RandomState(seed + 1): the first 1 + n_restarts
    #                        runs stay exactly sklearn's, the extra ones can only raise the final LML (multi-modal 5/6-parameter kernels)

    def __post_init__(self):
        if self.family not in COMPILED_FAMILIES:
            raise NotImplementedError(f"kernel family {self.family} is not compiled into libfulu_gp")
        p = family_n_params(self.family)
        if self.polish is None:
            self.polish = 0 if self.family == KERNEL_RBF_WHITE else 1
        if self.theta0 is None:
            self.theta0 = np.zeros(p)
        if self.lower is None:
            self.lower = np.full(p, -LOG_BOUND)
        if self.upper is None:
            self.upper = np.full(p, LOG_BOUND)
        if self.perm is None:
            self.perm = np.arange(p)
        self.theta0 = np.asarray(self.theta0, np.float64)
        self.lower = np.asarray(self.lower, np.float64)
        self.upper = np.asarray(self.upper, np.float64)
        self.perm = np.asarray(self.perm, np.int64)
        if not (len(self.theta0) == len(self.lower) == len(self.upper) == len(self.perm) == p):
            raise ValueError(f"kernel family {self.family} has {p} hyperparameters")

    @property
    def n_params(self):
        return len(self.theta0)

    def to_sklearn(self, theta):
        """Canonical -> sklearn order along the last axis."""
        theta = np.asarray(theta)
        out = np.empty_like(theta)
        out[..., self.perm] = theta
        return out

    def from_sklearn(self, theta):
        """sklearn -> canonical order along the last axis."""
        return np.asarray(theta)[..., self.perm]

    def start_points(self):
        """theta0, then n_restarts draws of RandomState(seed).uniform(lower, upper) in sklearn's hyperparameter order
        (_gpr.py:251,312-333); returned in canonical order."""
        rng = np.random.RandomState(self.seed)
        lo, hi = self.to_sklearn(self.lower), self.to_sklearn(self.upper)
        pts = [self.theta0.copy()]
        for _ in range(self.n_restarts):
            pts.append(self.from_sklearn(rng.uniform(lo, hi)))
        rng2 = np.random.RandomState(self.seed + 1)
        for _ in range(self.extra_starts):
            pts.append(self.from_sklearn(rng2.uniform(lo, hi)))
        return np.ascontiguousarray(np.array(pts), np.float64)


def _flatten(kernel, op):
    """Leaves of a (left- or right-nested) Sum / Product tree, left to right."""
    if type(kernel).__name__ == op:
        return _flatten(kernel.k1, op) + _flatten(kernel.k2, op)
    return [kernel]


def kernel_spec_from_sklearn(kernel):
    """Map an sklearn kernel object onto a compiled family; anything else is refused (there is no CPU fallback).

    Accepted: any ordering of  ConstantKernel * [Matern(nu in {0.5, 1.5, 2.5})] * RBF([l_t, l_lambda])  +  [Matern | RationalQuadratic]
    +  WhiteKernel  whose (product factor, added term) pair is compiled (COMPILED_FAMILIES) - this covers the reference's default
    (fulu/gp_aug.py:29), its docstring/test kernel C*RBF+Matern()+White (gp_aug.py:24, tests/test_aug.py:37), the notebook's
    C*Matern()*RBF+Matern()+White (plotting.ipynb:33) and the RationalQuadratic the reference imports (gp_aug.py:5).
    Initial values and bounds of the kernel object are honoured; fixed hyperparameters are refused."""
    if kernel is None:
        return KernelSpec()
    name = lambda k: type(k).__name__
    nu_code = {0.5: F_MAT05, 1.5: F_MAT15, 2.5: F_MAT25}

    def refuse(why):
        raise NotImplementedError(
            "fulu_b200 runs kernels of the form ConstantKernel * [Matern] * RBF([l_t, l_lambda]) + [Matern | RationalQuadratic] + "
            f"WhiteKernel on the GPU path ({why}); got {kernel!r}")

    # leaves in sklearn's theta order (KernelOperator.theta = k1.theta ++ k2.theta, kernels.py:739-752)
    offset = 0
    terms = []
    for term in _flatten(kernel, "Sum"):
        leaves = []
        for leaf in _flatten(term, "Product"):
            if name(leaf) in ("Sum", "Exponentiation", "CompoundKernel"):
                refuse("nested sums / exponentiation are not supported")
            if any(h.fixed for h in leaf.hyperparameters):
                refuse("fixed hyperparameters are not supported")
            leaves.append((leaf, offset))
            offset += leaf.n_dims
        terms.append(leaves)

    def iso_matern(leaf):
        return name(leaf) == "Matern" and np.ndim(leaf.length_scale) == 0 and leaf.nu in nu_code

    white = [t for t in terms if len(t) == 1 and name(t[0][0]) == "WhiteKernel"]
    extra = [t for t in terms if len(t) == 1 and (iso_matern(t[0][0]) or name(t[0][0]) == "RationalQuadratic")]
    main = [t for t in terms if t not in white and t not in extra]
    if len(white) != 1 or len(main) != 1 or len(extra) > 1:
        refuse("need exactly one WhiteKernel term, one ConstantKernel*RBF term and at most one added Matern/RationalQuadratic term")
    consts = [(l, o) for l, o in main[0] if name(l) == "ConstantKernel"]
    rbfs = [(l, o) for l, o in main[0] if name(l) == "RBF" and np.size(l.length_scale) == 2]
    mats = [(l, o) for l, o in main[0] if iso_matern(l)]
    if len(consts) != 1 or len(rbfs) != 1 or len(mats) > 1 or len(consts) + len(rbfs) + len(mats) != len(main[0]):
        refuse("the product term must be ConstantKernel * [isotropic Matern, nu in {0.5, 1.5, 2.5}] * RBF with two length scales")
    f1 = nu_code[mats[0][0].nu] if mats else F_NONE
    perm = [consts[0][1]] + ([mats[0][1]] if mats else []) + [rbfs[0][1], rbfs[0][1] + 1]
    ex = F_NONE
    if extra:
        leaf, o = extra[0][0]
        if name(leaf) == "RationalQuadratic":
            ex = F_RQ
            perm += [o, o + 1]  # alpha, length_scale (hyperparameters sort by name, kernels.py:280-288)
        else:
            ex = nu_code[leaf.nu]
            perm += [o]
    perm.append(white[0][0][1])
    family = family_code(f1, ex)
    if family not in COMPILED_FAMILIES:
        refuse(f"the family (product factor {f1}, added term {ex}) is not compiled into libfulu_gp")
    perm = np.asarray(perm, np.int64)
    b = np.asarray(kernel.bounds, np.float64)
    return KernelSpec(family, np.asarray(kernel.theta, np.float64)[perm], b[perm, 0].copy(), b[perm, 1].copy(), perm=perm)


class _nvtx:
    """NVTX range around a stage of the path (ingest / upload / fit / predict / download), visible in Nsight timelines; free when no
    profiler is attached."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        try:
            torch.cuda.nvtx.range_push(self.name)
            self.on = True
        except Exception:   # NVTX missing from this torch build
            self.on = False

    def __exit__(self, *exc):
        if self.on:
            torch.cuda.nvtx.range_pop()
        return False


def _require_cuda(device=None):
    if not torch.cuda.is_available():
        raise _capi.FuluGpError("fulu_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


class DeviceBatch:
    """A ragged CSR batch of light curves resident in HBM (layout of SURVEY 8b / include/fulu_gp.h)."""

    def __init__(self, offsets, t, flux, flux_err, band, loglam, n_max, use_err=False, alpha0=1e-10, family=KERNEL_RBF_WHITE):
        self.offsets, self.t, self.flux, self.flux_err, self.band, self.loglam = offsets, t, flux, flux_err, band, loglam
        self.n_curves = int(offsets.numel() - 1)
        self.n_obs_total = int(t.numel())
        self.n_max = int(n_max)
        self.n_bands = int(loglam.numel())
        self.use_err = bool(use_err)
        self.device = t.device
        c = _capi.Batch()
        c.n_curves, c.n_obs_total, c.n_max, c.n_bands = self.n_curves, self.n_obs_total, max(self.n_max, 1), self.n_bands
        c.offsets, c.t, c.flux, c.flux_err = offsets.data_ptr(), t.data_ptr(), flux.data_ptr(), flux_err.data_ptr()
        c.band, c.loglam = band.data_ptr(), loglam.data_ptr()
        c.kernel, c.use_err, c.alpha0 = family, int(use_err), alpha0
        c.order, c.n_order = None, 0
        self.c = c
        self.order = None
        self.family, self.alpha0 = family, alpha0

    def with_family(self, family):
        """The same batch (nothing copied) read under another covariance family."""
        if family == self.family:
            return self
        sub = DeviceBatch(self.offsets, self.t, self.flux, self.flux_err, self.band, self.loglam, self.n_max, self.use_err, self.alpha0, family)
        sub.order = self.order
        sub.c.order, sub.c.n_order = self.c.order, self.c.n_order
        return sub

    def select(self, order, n_max=None):
        """The same batch restricted to the curves `order` (int32 curve ids, processing order), with launch geometry for curves
        of up to `n_max` observations.  Nothing is copied: the library follows the index list (fulu_gp_batch.order)."""
        sub = DeviceBatch(self.offsets, self.t, self.flux, self.flux_err, self.band, self.loglam, self.n_max if n_max is None else n_max,
                          self.use_err, self.alpha0, self.family)
        if order is not None:
            sub.order = (order if torch.is_tensor(order) else torch.as_tensor(order, dtype=torch.int32)).to(self.device, torch.int32).contiguous()
            sub.c.order, sub.c.n_order = sub.order.data_ptr(), int(sub.order.numel())
        return sub

    @staticmethod
    def from_host(b: CurveBatch, device=None, use_err=False, pinned=None, n_max=None, family=KERNEL_RBF_WHITE):
        """H2D copy of a host batch.  `pinned`: optional dict of pinned staging tensors (see HostStaging).  `n_max`: an upper
        bound on the curve length (default: the batch's own maximum); it fixes the launch geometry."""
        device = _require_cuda(device)
        own = int(b.lengths().max()) if b.n_curves else 0
        n_max = own if n_max is None else max(int(n_max), own)
        if pinned is None:
            src = dict(offsets=torch.from_numpy(b.offsets), t=torch.from_numpy(b.t), flux=torch.from_numpy(b.flux),
                       flux_err=torch.from_numpy(b.flux_err), band=torch.from_numpy(b.band), loglam=torch.from_numpy(b.loglam))
        else:
            src = pinned
        dev = {k: v.to(device, non_blocking=True) for k, v in src.items()}
        return DeviceBatch(dev["offsets"], dev["t"], dev["flux"], dev["flux_err"], dev["band"], dev["loglam"], n_max, use_err, family=family)

    def h2d_bytes(self):
        return sum(x.numel() * x.element_size() for x in (self.offsets, self.t, self.flux, self.flux_err, self.band, self.loglam))


def pin_batch(b: CurveBatch):
    """Pinned host staging copy of a CurveBatch (what a production ingest thread would fill)."""
    out = {}
    for k in ("offsets", "t", "flux", "flux_err", "band", "loglam"):
        a = torch.from_numpy(getattr(b, k))
        p = torch.empty(a.shape, dtype=a.dtype, pin_memory=True)
        p.copy_(a)
        out[k] = p
    return out


class GPEngine:
    """Thin owner of the library handle, the workspace and the stream for one GPU."""

    def __init__(self, device=None):
        self.lib = _capi.load()
        self.device = _require_cuda(device)
        self._ws = None
        self.lane = 0       # which cached workspace (and side stream) the next call uses
        self._streams = {}
        self.launches = 0  # kernels this engine has launched through the library (the library reports a fit's count; the other calls are fixed)

    # ---- plumbing
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _workspace(self, batch, n_params, n_starts, window):
        """The device workspace of a call, cached per lane (`self.lane`: calls that may run concurrently on different streams -
        the length classes of fit_augment_device - must not share one)."""
        need = ctypes.c_size_t()
        _capi.check(self.lib.fulu_gp_workspace_bytes(ctypes.byref(batch.c), n_params, n_starts, window, ctypes.byref(need)))
        if self._ws is None:
            self._ws = {}
        ws = self._ws.get(self.lane)
        if ws is None or ws.numel() < need.value:
            self._ws[self.lane] = None
            ws = self._ws[self.lane] = torch.empty(int(need.value * 1.1) + 4096, dtype=torch.uint8, device=self.device)
        return ws

    def side_stream(self, lane):
        """A CUDA stream per lane (created once, reused): the library is stream-ordered, so independent calls overlap on the GPU."""
        st = self._streams.get(lane)
        if st is None:
            st = self._streams[lane] = torch.cuda.Stream(device=self.device)
        return st

    @staticmethod
    def _p(x):
        return None if x is None else ctypes.c_void_p(x.data_ptr())

    def _new(self, *shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, device=self.device)

    # ---- a5: LML + gradient at fixed hyperparameters
    def lml(self, batch: DeviceBatch, theta):
        """theta [B,P] (device) -> dict(lml [B], grad [B,P], scaler [B,4], ynorm [B,2], status [B])."""
        with torch.cuda.device(self.device):
            B, P = batch.n_curves, theta.shape[1]
            theta = theta.to(self.device, torch.float64).contiguous()
            out = dict(lml=self._new(B), grad=self._new(B, P), scaler=self._new(B, 4), ynorm=self._new(B, 2), status=self._new(B, dtype=torch.int32))
            ws = self._workspace(batch, P, 1, 1)
            _capi.check(self.lib.fulu_gp_lml_batch(ctypes.byref(batch.c), P, self._p(theta), self._p(out["lml"]), self._p(out["grad"]),
                                                   self._p(out["scaler"]), self._p(out["ynorm"]), self._p(out["status"]), self._p(ws),
                                                   ws.numel(), self._stream()))
            self.launches += 4   # prep, evaluation, the two exports
            return out

    def phase_cycles(self, batch: DeviceBatch, theta):
        """Diagnostics: SM-clock cycles CTA 0 spent in each phase of its last LML+gradient evaluation (under a full grid)."""
        with torch.cuda.device(self.device):
            B, P = batch.n_curves, theta.shape[1]
            theta = theta.to(self.device, torch.float64).contiguous()
            lml, grad = self._new(B), self._new(B, P)
            marks = torch.zeros(16, dtype=torch.int64, device=self.device)
            ws = self._workspace(batch, P, 1, 1)
            _capi.check(self.lib.fulu_gp_debug_phase_cycles(ctypes.byref(batch.c), P, self._p(theta), self._p(lml), self._p(grad),
                                                            self._p(marks), self._p(ws), ws.numel(), self._stream()))
            m = marks.cpu().numpy()
            names = ("load", "assemble", "cholesky", "collect_z", "inverse", "alpha", "traces", "reduce")
            out = {n: int(m[i + 1] - m[i]) for i, n in enumerate(names)}
            out["diag_tiles_in_cholesky"] = int(m[9])
            return out

    # ---- a6: optimiser
    def fit(self, batch: DeviceBatch, spec: KernelSpec = None, window=0, max_rounds=0, factr=1e7, pgtol=1e-5, maxiter=15000,
            maxfun=15000, maxls=20, m=10, return_runs=False, profile=False, out=None, polish=None):
        """S = 1 + n_restarts L-BFGS-B runs per curve -> dict(theta [B,P], lml [B], n_eval [B], status [B], scaler [B,4], ynorm [B,2]).
        `out`: a dict of such tensors to write into (rows of the curves `batch` selects; the others are left alone).

        `polish` (default spec.polish) continuation runs: a fresh L-BFGS-B run per curve started at its best point, kept where
        it improves the LML.  The kernels with 5-6 hyperparameters have nearly flat valleys (a Matern length or the RQ alpha
        drifting to its bound) in which L-BFGS-B stops on its relative-reduction test at a point that depends on rounding;
        sklearn's own runs stop at a different point of the same valley, sometimes 1e-6..1e-4 higher in LML.  One continuation
        run (a few dozen evaluations, counted in n_eval) removes that dependence on the stopping accident."""
        spec = spec or KernelSpec(batch.family)
        batch = batch.with_family(spec.family)
        with torch.cuda.device(self.device):
            B, P = batch.n_curves, spec.n_params
            starts_h = spec.start_points()
            S = starts_h.shape[0]
            starts = torch.from_numpy(starts_h).to(self.device)
            o = _capi.FitOptions()
            o.n_params, o.n_starts, o.m, o.maxiter, o.maxfun, o.maxls, o.window, o.max_rounds = P, S, m, maxiter, maxfun, maxls, window, max_rounds
            o.factr, o.pgtol = factr, pgtol
            for k in range(P):
                o.lower[k], o.upper[k] = spec.lower[k], spec.upper[k]
            if out is None:
                out = dict(theta=self._new(B, P), lml=self._new(B), n_eval=self._new(B, dtype=torch.int32), status=self._new(B, dtype=torch.int32),
                           scaler=self._new(B, 4), ynorm=self._new(B, 2))
            else:
                out = dict(out)
            if return_runs:
                out["run_lml"] = self._new(B, S)
                out["run_theta"] = self._new(B, S, P)
                out["run_n_eval"] = self._new(B, S, dtype=torch.int32)
            # True: per-launch CUDA events (eval_ms, step_ms); "counts" (and by default, for the launch counter): only rounds / launches /
            # geometry, no events - the same launches as an unprofiled call
            prof = (ctypes.c_double * 12)()
            o.profile = ctypes.cast(prof, ctypes.POINTER(ctypes.c_double))
            o.profile_counts_only = 0 if profile is True else 1
            ws = self._workspace(batch, P, S, window)
            _capi.check(self.lib.fulu_gp_fit_batch(ctypes.byref(batch.c), ctypes.byref(o), self._p(starts), self._p(out["theta"]),
                                                   self._p(out["lml"]), self._p(out.get("scaler")), self._p(out.get("ynorm")),
                                                   self._p(out["n_eval"]), self._p(out["status"]),
                                                   self._p(out.get("run_lml")), self._p(out.get("run_theta")), self._p(out.get("run_n_eval")),
                                                   self._p(ws), ws.numel(), self._stream()))
            self.launches += int(prof[3])
            if profile:
                keys = ("rounds", "eval_ms", "step_ms", "launches", "window", "grid_eval", "smem_eval", "global_a", "threads_eval", "ctas_per_sm",
                        "tail_ms", "place")
                out["profile"] = {k: float(prof[i]) for i, k in enumerate(keys)}
            for _ in range(spec.polish if polish is None else polish):
                # rows the batch does not select hold copies of the current values, so they compare equal and stay
                o.n_starts, o.starts_per_curve = 1, 1
                cont = dict(theta=out["theta"].clone(), lml=out["lml"].clone(), n_eval=torch.zeros_like(out["n_eval"]), status=out["status"].clone())
                starts2 = out["theta"].contiguous()
                prof2 = (ctypes.c_double * 12)()
                o.profile, o.profile_counts_only = ctypes.cast(prof2, ctypes.POINTER(ctypes.c_double)), 1
                _capi.check(self.lib.fulu_gp_fit_batch(ctypes.byref(batch.c), ctypes.byref(o), self._p(starts2), self._p(cont["theta"]),
                                                       self._p(cont["lml"]), None, None, self._p(cont["n_eval"]), self._p(cont["status"]),
                                                       None, None, None, self._p(ws), ws.numel(), self._stream()))
                self.launches += int(prof2[3])
                better = cont["lml"] > out["lml"]
                out["theta"].copy_(torch.where(better[:, None], cont["theta"], out["theta"]))
                out["status"].copy_(torch.where(better, cont["status"], out["status"]))
                out["lml"].copy_(torch.where(better, cont["lml"], out["lml"]))
                out["n_eval"].add_(cont["n_eval"])
            return out

    # ---- a7-a9: augmentation grid / arbitrary points
    def predict_grid(self, batch: DeviceBatch, theta, n_obs, t_min=None, t_max=None, out=None):
        """-> dict(mean [B,n_bands,n_obs], std [...], lml [B], status [B]); t_min/t_max [B] default to each curve's span.
        `out`: a dict of such tensors to write into (rows of the curves `batch` selects)."""
        with torch.cuda.device(self.device):
            B, P = batch.n_curves, theta.shape[1]
            theta = theta.to(self.device, torch.float64).contiguous()
            if (t_min is None) != (t_max is None):
                raise ValueError("give both t_min and t_max or neither")
            if t_min is not None:
                t_min = torch.as_tensor(t_min, dtype=torch.float64).to(self.device).contiguous().reshape(B)
                t_max = torch.as_tensor(t_max, dtype=torch.float64).to(self.device).contiguous().reshape(B)
            if out is None:
                out = dict(mean=self._new(B, batch.n_bands, n_obs), std=self._new(B, batch.n_bands, n_obs), lml=self._new(B),
                           status=self._new(B, dtype=torch.int32))
            ws = self._workspace(batch, P, 1, 1)
            _capi.check(self.lib.fulu_gp_predict_grid_batch(ctypes.byref(batch.c), P, self._p(theta), self._p(t_min), self._p(t_max), n_obs,
                                                            self._p(out["mean"]), self._p(out["std"]), self._p(out["lml"]),
                                                            self._p(out["status"]), self._p(ws), ws.numel(), self._stream()))
            self.launches += 2   # prep + predict
            return out

    def predict_points(self, batch: DeviceBatch, theta, q_offsets, q_t, q_band):
        """Ragged queries: curve i at q_t/q_band[q_offsets[i]:q_offsets[i+1]] -> dict(mean [sum M], std [sum M], status [B])."""
        with torch.cuda.device(self.device):
            B, P = batch.n_curves, theta.shape[1]
            theta = theta.to(self.device, torch.float64).contiguous()
            q_offsets = torch.as_tensor(q_offsets, dtype=torch.int64).to(self.device).contiguous()
            q_t = torch.as_tensor(q_t, dtype=torch.float64).to(self.device).contiguous()
            q_band = torch.as_tensor(q_band, dtype=torch.int32).to(self.device).contiguous()
            M = int(q_t.numel())
            out = dict(mean=self._new(M), std=self._new(M), status=self._new(B, dtype=torch.int32))
            ws = self._workspace(batch, P, 1, 1)
            _capi.check(self.lib.fulu_gp_predict_points_batch(ctypes.byref(batch.c), P, self._p(theta), self._p(q_offsets), self._p(q_t),
                                                              self._p(q_band), 0, self._p(out["mean"]), self._p(out["std"]),
                                                              self._p(out["status"]), self._p(ws), ws.numel(), self._stream()))
            self.launches += 2
            return out

    # ---- 8f-3: downstream consumer and upstream ingest on the device
    def peak(self, batch: DeviceBatch, mean, t_min=None, t_max=None, want_sum=False):
        """Band-summed augmented curve and its first maximum per curve (LcPlotter.plot_sum_passbands / plot_peak,
        fulu/plotting.py:93,114-115), computed where the curves lie.  mean [B,n_bands,n_obs] (device) ->
        dict(peak_t [B], peak_flux [B], peak_index [B] (+ sum_flux [B,n_obs]))."""
        with torch.cuda.device(self.device):
            B, n_obs = batch.n_curves, int(mean.shape[2])
            if (t_min is None) != (t_max is None):
                raise ValueError("give both t_min and t_max or neither")
            if t_min is not None:
                t_min = torch.as_tensor(t_min, dtype=torch.float64).to(self.device).contiguous().reshape(B)
                t_max = torch.as_tensor(t_max, dtype=torch.float64).to(self.device).contiguous().reshape(B)
            mean = mean.contiguous()
            out = dict(peak_t=self._new(B), peak_flux=self._new(B), peak_index=self._new(B, dtype=torch.int32))
            if want_sum:
                out["sum_flux"] = self._new(B, n_obs)
            _capi.check(self.lib.fulu_gp_peak_batch(ctypes.byref(batch.c), self._p(mean), self._p(t_min), self._p(t_max), n_obs,
                                                    self._p(out.get("sum_flux")), self._p(out["peak_t"]), self._p(out["peak_flux"]),
                                                    self._p(out["peak_index"]), self._stream()))
            self.launches += 1
            return out

    def ingest_table(self, object_id, t, flux, flux_err, passband, passband2lam, use_err=False, family=KERNEL_RBF_WHITE):
        """A survey table grouped by object (columns as in notebook_examples/object_34299.csv plus an object id) -> a DeviceBatch,
        built on the device: CSR offsets from the id column, band indices through a lookup table of the (integer) passband ids
        (fulu/_base_aug.py:9-11).  The t / flux / flux_err columns are used where they lie.  Returns (batch, None, object_ids):
        the lengths are not read back (None tells fit_augment_device to build the length classes on the device),
        object_ids (host) = the id of each curve.
        Raises KeyError for a passband that is not in passband2lam, like the reference."""
        with torch.cuda.device(self.device):
            dev = self.device
            oid = torch.as_tensor(object_id, dtype=torch.int64).to(dev).contiguous()
            cols = [torch.as_tensor(c, dtype=torch.float64).to(dev).contiguous() for c in (t, flux, flux_err)]
            pb = torch.as_tensor(passband, dtype=torch.int32).to(dev).contiguous()
            n_rows = int(oid.numel())
            keys = list(passband2lam)
            if not all(isinstance(k, (int, np.integer)) and k >= 0 for k in keys):
                raise TypeError("ingest_table needs non-negative integer passband ids (map string keys on the host first)")
            lut_h = np.full(max(keys) + 1, -1, np.int32)
            lut_h[np.asarray(keys, np.int64)] = np.arange(len(keys), dtype=np.int32)
            lut = torch.from_numpy(lut_h).to(dev)
            loglam = torch.tensor([float(passband2lam[k]) for k in keys], dtype=torch.float64, device=dev)
            need = ctypes.c_size_t()
            _capi.check(self.lib.fulu_gp_ingest_workspace_bytes(n_rows, ctypes.byref(need)))
            ws = torch.empty(max(int(need.value), 8), dtype=torch.uint8, device=dev)
            offsets = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)   # at most one object per row
            n_obj = torch.zeros(1, dtype=torch.int64, device=dev)
            band = torch.empty(n_rows, dtype=torch.int32, device=dev)
            unknown = torch.zeros(1, dtype=torch.int32, device=dev)
            _capi.check(self.lib.fulu_gp_ingest_table(n_rows, self._p(oid), self._p(pb), self._p(lut), int(lut.numel()), n_rows,
                                                      self._p(offsets), self._p(n_obj), self._p(band), self._p(unknown), self._p(ws),
                                                      ws.numel(), self._stream()))
            B = int(n_obj.cpu()[0])
            if int(unknown.cpu()[0]):
                bad = pb[band < 0][0].item()
                raise KeyError(bad)
            offsets = offsets[: B + 1].contiguous()
            ids = oid[offsets[:-1]].cpu().numpy() if B else np.zeros(0, np.int64)
            # the offsets stay on the device: fit_augment_device builds the length classes and their index lists there
            # (class_orders_device); n_max of the batch itself = its longest curve, one integer
            n_max = int((offsets[1:] - offsets[:-1]).max().item()) if B else 0
            batch = DeviceBatch(offsets, cols[0], cols[1], cols[2], band, loglam, n_max, use_err, family=family)
            return batch, None, ids

    def fp64_peak(self, mode=0, iters=20000, reps=4):
        """Measured FP64 TFLOP/s of a register-resident DFMA (mode 0) or DMMA (mode 1) loop on this GPU: the library enqueues
        the probe (fulu_gp_fp64_probe), the timing is CUDA events on the launching stream here; best of `reps` after a warm-up."""
        with torch.cuda.device(self.device):
            sms = torch.cuda.get_device_properties(self.device).multi_processor_count
            scratch = torch.empty(sms * 8 * 256, dtype=torch.float64, device=self.device)
            flop = ctypes.c_double()
            best = 0.0
            for rep in range(reps + 1):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _capi.check(self.lib.fulu_gp_fp64_probe(mode, iters, self._p(scratch), scratch.numel() * 8, ctypes.byref(flop), self._stream()))
                e1.record()
                e1.synchronize()
                if rep > 0:
                    best = max(best, flop.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
            return best


# ------------------------------------------------------------------------------------------------ batched entry point
# upper bounds of the length classes one launch covers: the class bound (not the batch's own maximum) fixes the launch geometry
# (shared memory, CTAs per SM, threads per CTA), so a curve's result does not depend on what else is in the batch
# (bounds are 8k - 1: the factored matrix is the bordered one of order N + 1, padded to a multiple of 8; N <= 223 lives in shared
# memory; 1031 / 1543 / 2055 keep config #5's N = 1024 / 1536 / 2048 in tight classes)
N_BUCKETS = (31, 63, 103, 127, 159, 223, 303, 511, 1031, 1543, 2055, 4095)


@dataclass
class AugmentResult:
    flux_aug: np.ndarray  # [B, n_bands, n_obs]
    flux_err_aug: np.ndarray  # [B, n_bands, n_obs]
    theta: np.ndarray  # [B, P]
    lml: np.ndarray  # [B]
    n_eval: np.ndarray  # [B]
    status: np.ndarray  # [B]
    h2d_bytes: int = 0
    d2h_bytes: int = 0
    peak_t: np.ndarray = None  # [B] grid time of the band-summed curve's maximum (reduce="peak")
    peak_flux: np.ndarray = None  # [B]
    peak_index: np.ndarray = None  # [B]
    lpt_imbalance: float = None  # sharded runs: largest / mean modelled load of the ranks (sharding.fit_augment_sharded)

    def status_histogram(self):
        """How many curves carry each FULU_GP_STATUS_* bit (a curve can carry several): the per-batch health record."""
        names = {"bad_input": STATUS_BAD_INPUT, "not_pd": STATUS_NOT_PD, "not_converged": STATUS_NOT_CONVERGED, "at_bound": STATUS_AT_BOUND,
                 "neg_variance": STATUS_NEG_VARIANCE}
        st = np.asarray(self.status)
        out = {k: int(np.count_nonzero(st & bit)) for k, bit in names.items()}
        out["ok"] = int(np.count_nonzero(st == 0))
        out["curves"] = int(st.size)
        return out


MERGE_ABOVE = 303     # classes above this bound share one launch geometry (one 512-thread CTA per SM, the matrix in a global slab streamed
                      # through shared-memory stages) and are merged into ONE class, run longest curves first: a fit there is one fused
                      # launch fed by a task queue, and every class run on its own would end in its own tail of a few long optimiser runs


def bucket_curves(lengths, buckets=N_BUCKETS):
    """Group curve indices by padded length so one kernel launch sees similar N (cost ~ N^3).  The long-curve classes
    (bound > MERGE_ABOVE) all run with the same geometry and are merged into one class bounded by the longest (results do not
    depend on the class there: only the slab stride and the size of the shared-memory vectors do)."""
    lengths = np.asarray(lengths)
    which = np.searchsorted(np.asarray(buckets), lengths, side="left")
    # beyond the last bucket: one class bounded by its own longest curve (8k - 1 like the others), never a bound below a member
    beyond = (int(lengths.max()) + 8) // 8 * 8 - 1 if len(lengths) else 0
    groups = [[int(buckets[k]) if k < len(buckets) else beyond, np.nonzero(which == k)[0]] for k in np.unique(which)]
    out = []
    for bound, idx in groups:   # ascending bounds
        if out and out[-1][0] > MERGE_ABOVE:
            prev = out.pop()
            idx = np.sort(np.concatenate([prev[1], idx]))
        out.append([bound, idx])
    return [(b, i) for b, i in out]


def class_orders_device(offsets, buckets=N_BUCKETS):
    """bucket_curves for a batch whose offsets live on the device (an ingested table): the curve ids sorted by (length class,
    longest first) stay on the device as ONE int32 index array; only the class sizes and the longest length - a dozen integers - come
    back to the host, which needs them to size the launches.  Returns [(bound, count, order slice)] in ascending bounds, the
    long-curve classes (bound > MERGE_ABOVE) merged into one, like bucket_curves."""
    dev = offsets.device
    lengths = offsets[1:] - offsets[:-1]
    bounds = torch.tensor(buckets, dtype=torch.int64, device=dev)
    cls = torch.bucketize(lengths, bounds, right=False)              # first class whose bound >= length
    first_long = int(np.searchsorted(np.asarray(buckets), MERGE_ABOVE, side="right"))
    grp = torch.clamp(cls, max=first_long)                          # every class above MERGE_ABOVE is one group
    key = grp * (1 << 32) - lengths                                  # group ascending, longest first inside a group
    order = torch.argsort(key, stable=True).to(torch.int32)
    summary = torch.cat([torch.bincount(grp, minlength=first_long + 1), lengths.max().reshape(1) if lengths.numel() else lengths.new_zeros(1)]).cpu().numpy()
    counts, longest = summary[:-1], int(summary[-1])
    out, start = [], 0
    for g, c in enumerate(counts):
        c = int(c)
        if c:
            bound = int(buckets[g]) if g < first_long else max((longest + 8) // 8 * 8 - 1, int(buckets[min(g, len(buckets) - 1)]))
            if g == first_long:   # the merged long class is bounded by its own longest curve, in 8k - 1 steps like the table
                bound = next((int(b) for b in buckets if b >= longest), (longest + 8) // 8 * 8 - 1)
            out.append((bound, c, order[start:start + c]))
        start += c
    return out


def pin_results(n_curves, n_bands, n_obs, n_params=4):  # n_params: KernelSpec.n_params of the family that will be fitted
    """Pinned host buffers for the results of fit_augment_batch (reusable across calls: what a production loop double-buffers)."""
    mk = lambda *shape, dtype=torch.float64: torch.empty(shape, dtype=dtype, pin_memory=True)
    return dict(mean=mk(n_curves, n_bands, n_obs), std=mk(n_curves, n_bands, n_obs), theta=mk(n_curves, n_params), lml=mk(n_curves),
                n_eval=mk(n_curves, dtype=torch.int32), status=mk(n_curves, dtype=torch.int32))


def fit_augment_device(engine: "GPEngine", db: DeviceBatch, lengths, n_obs=128, t_min=None, t_max=None, spec: KernelSpec = None, window=0,
                       profile=False):
    """Fit + augmentation of a batch that is already resident in HBM: one pass of the library per length class (N_BUCKETS) over
    an index list of the class's curves, longest class first; every pass writes its rows of the same result tensors.
    Returns device tensors dict(mean, std [B,n_bands,n_obs], theta [B,P], lml, n_eval, status [B]) (+ "profile": per-class list)."""
    spec = spec or KernelSpec(db.family)
    db = db.with_family(spec.family)
    B, nb, P = db.n_curves, db.n_bands, spec.n_params
    if lengths is None:   # offsets on the device only (an ingested table): classes and index lists are built there
        groups = [(bound, count, order) for bound, count, order in class_orders_device(db.offsets)]
        on_device = True
    else:
        lengths = np.asarray(lengths)
        groups = [(bound, len(idx), idx) for bound, idx in bucket_curves(lengths)]
        on_device = False
    single = len(groups) == 1
    new = engine._new
    fit = dict(theta=new(B, P), lml=new(B), n_eval=new(B, dtype=torch.int32), status=new(B, dtype=torch.int32))
    pred = dict(mean=new(B, nb, n_obs), std=new(B, nb, n_obs), lml=new(B), status=new(B, dtype=torch.int32))
    profiles = []
    # Every class is a stream-ordered sequence of launches (fit, then prediction) that ends in a tail - the last, longest optimiser
    # runs on a few CTAs.  The classes therefore go to streams of their own, longest curves first: the tail of one class runs under
    # the bulk of the next instead of leaving the GPU idle.  (With per-launch timing on, one class at a time on the caller's stream.)
    concurrent = (profile is not True) and len(groups) > 1 and os.environ.get("FULU_GP_SERIAL_CLASSES") is None
    main = torch.cuda.current_stream(engine.device)
    used = []
    for lane, (bound, count, idx) in enumerate(reversed(groups)):   # longest curves first
        order = None
        if on_device:
            order = idx     # (already longest first, already on the device)
        elif not single or lengths.min() != lengths.max():
            # within a class, longest first too: the tail of the optimiser's task queue is then made of the cheapest curves
            order = idx[np.argsort(-lengths[idx], kind="stable")].astype(np.int32)
        if concurrent:
            st = engine.side_stream(lane)
            st.wait_stream(main)
            engine.lane = lane
            used.append(st)
        else:
            st = main
        with torch.cuda.stream(st):
            sub = db.select(order, n_max=bound)   # (the index list is uploaded on the class's own stream)
            with _nvtx(f"fulu_gp.fit n<={bound} x{count}"):
                f = engine.fit(sub, spec, window=window, out=fit, profile=profile)
            with _nvtx(f"fulu_gp.predict n<={bound} x{count}"):
                engine.predict_grid(sub, fit["theta"], n_obs, t_min, t_max, out=pred)
        if profile:
            profiles.append(dict(f["profile"], bound=bound, curves=count))
    engine.lane = 0
    for st in used:
        main.wait_stream(st)
    out = dict(mean=pred["mean"], std=pred["std"], theta=fit["theta"], lml=fit["lml"], n_eval=fit["n_eval"], status=fit["status"] | pred["status"])
    if profile:
        out["profile"] = profiles
    return out


def fit_augment_batch(curves: CurveBatch, n_obs=128, t_min=None, t_max=None, use_err=False, spec: KernelSpec = None, engine: GPEngine = None,
                      device=None, window=0, pinned=None, out=None, reduce=None):
    """The batched ragged entry point: GP fit + augmentation for every curve of `curves` (host arrays in, host arrays out).

    Equivalent to, for each curve i:  aug = fulu.GaussianProcessesAugmentation(passband2lam, use_err=use_err).fit(t, flux, flux_err, band);
    aug.augmentation(t_min[i], t_max[i], n_obs)  (t_min/t_max default to the curve's own span).

    The batch is uploaded once and stays where it is: each length class (N_BUCKETS) is one pass of the library over an index list
    of its curves, longest class first, and every pass writes its rows of the same device result tensors, which come back in one
    copy.  `pinned`: pinned staging of the inputs (pin_batch); `out`: pinned result buffers (pin_results) - the returned arrays
    are then views of them.  `spec`: the covariance family (KernelSpec / kernel_spec_from_sklearn); `theta` comes back in the
    order of the sklearn kernel the spec was built from (KernelSpec.perm).  `reduce="peak"`: the augmented curves stay on the
    device; the band-summed curve's peak (time, flux, grid index; LcPlotter.plot_peak, fulu/plotting.py:114-115) comes back instead
    (flux_aug / flux_err_aug are None, 20 bytes per curve instead of 16 n_bands n_obs).
    """
    engine = engine or GPEngine(device)
    spec = spec or KernelSpec()
    B, nb, P = curves.n_curves, curves.n_bands, spec.n_params
    if B == 0:
        return AugmentResult(np.empty((0, nb, n_obs)), np.empty((0, nb, n_obs)), np.empty((0, P)), np.empty(0), np.zeros(0, np.int32), np.zeros(0, np.int32))
    with torch.cuda.device(engine.device):
        with _nvtx("fulu_gp.h2d"):
            db = DeviceBatch.from_host(curves, engine.device, use_err=use_err, pinned=pinned, family=spec.family)
            tmn = None if t_min is None else torch.as_tensor(np.asarray(t_min, np.float64)).to(engine.device)
            tmx = None if t_max is None else torch.as_tensor(np.asarray(t_max, np.float64)).to(engine.device)
        dev = fit_augment_device(engine, db, curves.lengths(), n_obs, tmn, tmx, spec, window)
        if reduce == "peak":
            pk = engine.peak(db, dev["mean"], tmn, tmx)
            dev = {k: v for k, v in dev.items() if k not in ("mean", "std")}
            dev.update(pk)
            out = None
        elif reduce is not None:
            raise ValueError("reduce must be None or 'peak'")
        with _nvtx("fulu_gp.d2h"):
            if out is None:
                host = {k: v.cpu() for k, v in dev.items()}
            else:
                host = {k: out[k][:B] for k in dev}
                for k, v in dev.items():
                    host[k].copy_(v, non_blocking=True)
                torch.cuda.current_stream(engine.device).synchronize()
    theta = host["theta"].numpy()
    if not np.array_equal(spec.perm, np.arange(P)):
        theta = spec.to_sklearn(theta)
    res = AugmentResult(host["mean"].numpy() if "mean" in host else None, host["std"].numpy() if "std" in host else None, theta,
                        host["lml"].numpy(), host["n_eval"].numpy(), host["status"].numpy())
    if reduce == "peak":
        res.peak_t, res.peak_flux, res.peak_index = host["peak_t"].numpy(), host["peak_flux"].numpy(), host["peak_index"].numpy()
    res.h2d_bytes = db.h2d_bytes()
    res.d2h_bytes = sum(int(v.numel() * v.element_size()) for v in dev.values())
    return res


def fit_augment_table(object_id, t, flux, flux_err, passband, passband2lam, n_obs=128, use_err=False, spec: KernelSpec = None,
                      engine: GPEngine = None, device=None, reduce=None):
    """GP fit + augmentation of every object of a survey table grouped by object (one row per observation: object id, time, flux,
    flux error, integer passband id).  The table is uploaded as it is and turned into the ragged batch on the device
    (GPEngine.ingest_table); grids span each object's own observations.  Returns (AugmentResult, object_ids)."""
    engine = engine or GPEngine(device)
    spec = spec or KernelSpec()
    with torch.cuda.device(engine.device):
        with _nvtx("fulu_gp.ingest"):
            db, lengths, ids = engine.ingest_table(object_id, t, flux, flux_err, passband, passband2lam, use_err=use_err, family=spec.family)
        B, P = db.n_curves, spec.n_params
        if B == 0:
            nb = db.n_bands
            return AugmentResult(np.empty((0, nb, n_obs)), np.empty((0, nb, n_obs)), np.empty((0, P)), np.empty(0), np.zeros(0, np.int32),
                                 np.zeros(0, np.int32)), ids
        dev = fit_augment_device(engine, db, lengths, n_obs, None, None, spec)
        if reduce == "peak":
            pk = engine.peak(db, dev["mean"])
            dev = {k: v for k, v in dev.items() if k not in ("mean", "std")}
            dev.update(pk)
        elif reduce is not None:
            raise ValueError("reduce must be None or 'peak'")
        host = {k: v.cpu() for k, v in dev.items()}
    theta = host["theta"].numpy()
    if not np.array_equal(spec.perm, np.arange(P)):
        theta = spec.to_sklearn(theta)
    res = AugmentResult(host["mean"].numpy() if "mean" in host else None, host["std"].numpy() if "std" in host else None, theta,
                        host["lml"].numpy(), host["n_eval"].numpy(), host["status"].numpy())
    if reduce == "peak":
        res.peak_t, res.peak_flux, res.peak_index = host["peak_t"].numpy(), host["peak_flux"].numpy(), host["peak_index"].numpy()
    res.h2d_bytes = db.h2d_bytes() + 12 * db.n_obs_total
    res.d2h_bytes = sum(int(v.numel() * v.element_size()) for v in dev.values())
    return res, ids
