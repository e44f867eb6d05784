"""Host build of the device optimiser (fulu_b200/csrc/lbfgsb.cuh) against scipy's L-BFGS-B on the oracle objective."""
import warnings

import numpy as np
import pytest
from scipy.optimize import minimize

import hostopt
from conftest import load_golden
from oracle import gp_oracle as O


@pytest.mark.parametrize("name", ["readme", "testfixture", "csv34299", "csv34299_err", "plasticc_0", "ztf_2"])
def test_same_trajectory_and_optimum_as_scipy(name):
    g = load_golden(name)
    p2l = {i: float(v) for i, v in enumerate(g["loglam"])}
    gp = O.ClosedFormGP(p2l, use_err=bool(g["use_err"])).prepare(g["t"], g["flux"], g["flux_err"], g["band"])

    def obj(th):
        l, gr = O.lml_and_grad(gp.X, gp.y, gp.alpha, th)
        return -l, -gr

    B = O.LOG_BOUND
    best_s, best_m, same_nfev = np.inf, np.inf, 0
    starts = O.start_points(4)
    for s0 in starts:
        traj = []

        def rec(th):
            f, gr = obj(th)
            traj.append(th.copy())
            return f, gr

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rs = minimize(rec, s0, method="L-BFGS-B", jac=True, bounds=[(-B, B)] * 4)
        rm = hostopt.minimize(obj, s0, [-B] * 4, [B] * 4)
        # the first evaluations follow scipy's path exactly (identical algorithm), later ones to rounding
        k = min(8, len(traj), len(rm["traj"]))
        np.testing.assert_allclose(rm["traj"][:k, :4], np.array(traj)[:k], rtol=0, atol=1e-7)
        assert abs(rm["fun"] - rs.fun) < 1e-5 * max(1.0, abs(rs.fun))
        same_nfev += rm["nfev"] == rs.nfev
        best_s, best_m = min(best_s, rs.fun), min(best_m, rm["fun"])
    assert best_m <= best_s + 1e-6
    assert -best_m >= float(g["lml_star"]) - 1e-6
    assert same_nfev >= 4


def test_bounds_and_quadratic():
    # convex quadratic whose unconstrained minimiser is outside the box: solution must sit on the bound
    A = np.array([[3.0, 0.5, 0.0], [0.5, 2.0, 0.3], [0.0, 0.3, 1.0]])
    b = np.array([10.0, -1.0, 0.5])

    def obj(x):
        return 0.5 * x @ A @ x - b @ x, A @ x - b

    r = hostopt.minimize(obj, [0.0, 0.0, 0.0], [-1.0] * 3, [1.0] * 3)
    rs = minimize(obj, [0.0, 0.0, 0.0], method="L-BFGS-B", jac=True, bounds=[(-1, 1)] * 3)
    np.testing.assert_allclose(r["x"], rs.x, atol=1e-6)
    assert r["x"][0] == 1.0


def test_start_at_optimum_converges_immediately():
    def obj(x):
        return float(x @ x), 2 * x

    r = hostopt.minimize(obj, [0.0, 0.0], [-1.0] * 2, [1.0] * 2)
    assert r["nfev"] == 1 and r["task"] == 2
