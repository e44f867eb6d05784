"""ctypes wrapper around the host-thread emulation of the CUDA block algorithm (tests/_host/host_gp_block.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "_host", "host_gp_block.cpp")
LIB = os.path.join(HERE, "_build", "host_gp_block.so")
HDR = os.path.join(os.path.dirname(HERE), "fulu_b200", "csrc", "gp_block.cuh")
HDR2 = os.path.join(os.path.dirname(HERE), "fulu_b200", "csrc", "gp_tiled.cuh")
_dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
_ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _lib():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR), os.path.getmtime(HDR2)):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++20", "-shared", "-fPIC", "-pthread", "-x", "c++", SRC, "-o", LIB])
    return ctypes.CDLL(LIB)


class Prepared:
    pass


def prepare(t, flux, flux_err, band, loglam, use_err, alpha0=1e-10, nthr=32):
    lib = _lib()
    n = len(t)
    t, flux, flux_err = (np.ascontiguousarray(a, np.float64) for a in (t, flux, flux_err))
    band = np.ascontiguousarray(band, np.int32)
    loglam = np.ascontiguousarray(loglam, np.float64)
    p = Prepared()
    p.n, p.band = n, band
    p.xs, p.y, p.al = np.zeros(n), np.zeros(n), np.zeros(n)
    p.lam, p.stats = np.zeros(len(loglam)), np.zeros(6)
    p.status = lib.host_gp_prepare(n, _dp(t), _dp(flux), _dp(flux_err), _ip(band), _dp(loglam), len(loglam), int(use_err),
                                   ctypes.c_double(alpha0), nthr, _dp(p.xs), _dp(p.y), _dp(p.al), _dp(p.lam), _dp(p.stats))
    return p


def lml_grad(p, theta, nthr=32, family=0):
    lib = _lib()
    theta = np.ascontiguousarray(theta, np.float64)
    lml = ctypes.c_double()
    grad = np.zeros(len(theta))
    ok = lib.host_gp_lml(family, p.n, _dp(p.xs), _ip(p.band), _dp(p.y), _dp(p.al), _dp(p.lam), _dp(theta), nthr, ctypes.byref(lml), _dp(grad))
    return lml.value, grad, bool(ok)


def lml_grad_tiled(p, theta, nthr=128, family=0, reps=1):
    """The long-curve path (gp_tiled.cuh) on the same prepared curve."""
    lib = _lib()
    theta = np.ascontiguousarray(theta, np.float64)
    lml = ctypes.c_double()
    grad = np.zeros(len(theta))
    ok = lib.host_gp_lml_tiled(family, p.n, _dp(p.xs), _ip(p.band), _dp(p.y), _dp(p.al), _dp(p.lam), _dp(theta), nthr, reps, ctypes.byref(lml),
                               _dp(grad))
    return lml.value, grad, bool(ok)


def predict_grid_tiled(p, theta, n_bands, n_obs, t_min, t_max, nthr=128, family=0):
    lib = _lib()
    theta = np.ascontiguousarray(theta, np.float64)
    m = n_bands * n_obs
    mean, std = np.zeros(m), np.zeros(m)
    lml = ctypes.c_double()
    st = lib.host_gp_predict_tiled(family, p.n, _dp(p.xs), _ip(p.band), _dp(p.y), _dp(p.al), _dp(p.lam), _dp(theta), _dp(p.stats), m, n_obs,
                                   ctypes.c_double(t_min), ctypes.c_double(t_max), None, None, nthr, _dp(mean), _dp(std), ctypes.byref(lml))
    return mean, std, lml.value, st


def predict_grid(p, theta, n_bands, n_obs, t_min, t_max, nthr=32, family=0):
    lib = _lib()
    theta = np.ascontiguousarray(theta, np.float64)
    m = n_bands * n_obs
    mean, std = np.zeros(m), np.zeros(m)
    lml = ctypes.c_double()
    st = lib.host_gp_predict(family, p.n, _dp(p.xs), _ip(p.band), _dp(p.y), _dp(p.al), _dp(p.lam), _dp(theta), _dp(p.stats), m, n_obs,
                             ctypes.c_double(t_min), ctypes.c_double(t_max), None, None, nthr, _dp(mean), _dp(std), ctypes.byref(lml))
    return mean, std, lml.value, st


def predict_points(p, theta, qt, qband, nthr=32, family=0):
    lib = _lib()
    theta = np.ascontiguousarray(theta, np.float64)
    qt = np.ascontiguousarray(qt, np.float64)
    qband = np.ascontiguousarray(qband, np.int32)
    m = len(qt)
    mean, std = np.zeros(m), np.zeros(m)
    lml = ctypes.c_double()
    st = lib.host_gp_predict(family, p.n, _dp(p.xs), _ip(p.band), _dp(p.y), _dp(p.al), _dp(p.lam), _dp(theta), _dp(p.stats), m, 0,
                             ctypes.c_double(0), ctypes.c_double(0), _dp(qt), _ip(qband), nthr, _dp(mean), _dp(std), ctypes.byref(lml))
    return mean, std, lml.value, st


def exp_neg(x):
    lib = _lib()
    x = np.ascontiguousarray(x, np.float64)
    out = np.empty_like(x)
    lib.host_gp_exp_neg(len(x), _dp(x), _dp(out))
    return out
