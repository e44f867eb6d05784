"""ctypes wrapper around the host build of fulu_b200/csrc/lbfgsb.cuh (test harness, see tests/_host)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "_host", "host_lbfgsb.cpp")
LIB = os.path.join(HERE, "_build", "host_lbfgsb.so")
HDR = os.path.join(os.path.dirname(HERE), "fulu_b200", "csrc", "lbfgsb.cuh")

OBJ = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double))


def _lib():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", SRC, "-o", LIB])
    lib = ctypes.CDLL(LIB)
    lib.host_lbfgsb_minimize.restype = ctypes.c_int
    return lib


def minimize(fun, x0, lo, hi, factr=1e7, pgtol=1e-5, maxiter=15000, maxfun=15000, maxls=20, max_traj=4096):
    lib = _lib()
    n = len(x0)
    x0 = np.ascontiguousarray(x0, np.float64)
    lo = np.ascontiguousarray(lo, np.float64)
    hi = np.ascontiguousarray(hi, np.float64)

    def cb(xp, fp, gp):
        x = np.array([xp[i] for i in range(n)])
        f, g = fun(x)
        fp[0] = f
        for i in range(n):
            gp[i] = g[i]

    cbo = OBJ(cb)
    x = np.zeros(n)
    f = ctypes.c_double()
    nfev = ctypes.c_int()
    nit = ctypes.c_int()
    traj = np.zeros((max_traj, n + 1))
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    task = lib.host_lbfgsb_minimize(n, dp(x0), dp(lo), dp(hi), ctypes.c_double(factr), ctypes.c_double(pgtol), maxiter, maxfun, maxls, cbo,
                                    dp(x), ctypes.byref(f), ctypes.byref(nfev), ctypes.byref(nit), dp(traj), max_traj)
    return dict(x=x, fun=f.value, nfev=nfev.value, nit=nit.value, task=task, traj=traj[: nfev.value])
