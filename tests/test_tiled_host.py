"""The long-curve path (fulu_b200/csrc/gp_tiled.cuh: matrix in a global slab, operand tiles streamed through shared-memory stages by
bulk copies, 64-column panels) executed on host threads - same harness as test_block_host.py, with the copy engine and its
barriers emulated (tests/_host/host_gp_block.cpp) - against the golden vectors of the real reference, the closed-form oracle and the
shared-memory path.  On the GPU the same code serves every curve longer than 223 points; here it is also run on short curves so the
reference's own fixtures apply."""
import numpy as np
import pytest

import hostblock as H
from conftest import load_golden, relerr
from fulu_b200 import datasets
from oracle import gp_oracle as O


@pytest.mark.parametrize("name,nthr", [("csv34299", 128), ("csv34299_err", 64), ("plasticc_1", 256), ("ztf_2", 128), ("readme_err", 128)])
def test_tiled_lml_grad_predict_vs_reference_fixtures(name, nthr):
    g = load_golden(name)
    p = H.prepare(g["t"], g["flux"], g["flux_err"], g["band"], g["loglam"], bool(g["use_err"]), nthr=64)
    p2l = {i: float(v) for i, v in enumerate(g["loglam"])}
    gp = O.ClosedFormGP(p2l, use_err=bool(g["use_err"])).prepare(g["t"], g["flux"], g["flux_err"], g["band"])
    for k in (0, 3, 6):
        th = g["thetas"][k]
        lml, grad, ok = H.lml_grad_tiled(p, th, nthr=nthr)
        assert ok
        assert abs(lml - g["lml"][k]) <= 1e-9 * max(1.0, abs(g["lml"][k]))
        _, _, gs = O.lml_and_grad(gp.X, gp.y, gp.alpha, th, return_scale=True)
        den = np.maximum(np.maximum(np.abs(g["grad"][k]), 1e-2 * gs), 1e-30)
        assert np.max(np.abs(grad - g["grad"][k]) / den) < 1e-9
    nb, n_obs = len(g["loglam"]), int(g["n_obs"])
    m, s, lml, st = H.predict_grid_tiled(p, g["theta_star"], nb, n_obs, float(g["t_min"]), float(g["t_max"]), nthr=nthr)
    assert st == 0
    assert relerr(m, g["flux_aug"], floor=1e-3 * float(g["y_std"])) < 1e-9
    assert relerr(s, g["flux_err_aug"]) < 1e-9
    assert abs(lml - float(g["lml_star"])) <= 1e-9 * max(1.0, abs(float(g["lml_star"])))


@pytest.mark.parametrize("n,nthr", [(1, 64), (7, 64), (8, 64), (63, 128), (64, 128), (65, 64), (127, 256), (128, 64), (129, 128), (200, 128),
                                    (224, 256), (300, 128), (300, 512)])
def test_tiled_sizes_vs_oracle(n, nthr):
    """Panel and block boundaries: one panel / several, a last panel of one tile (n = 64, 128: the border opens a tile-row of its
    own), row blocks with fewer rows than warps, 2 to 16 warps."""
    b = datasets.ztf_like(1, first=3, n_lo=max(n, 2), n_hi=max(n, 2))
    t, f, e, bd = (a[:n] for a in b.curve(0))
    p = H.prepare(t, f, e, bd, b.loglam, False, nthr=64)
    p2l = {i: float(v) for i, v in enumerate(b.loglam)}
    gp = O.ClosedFormGP(p2l).prepare(t, f, e, bd)
    th = np.array([0.5, -1.0, 1.5, -3.0])
    lml, grad, ok = H.lml_grad_tiled(p, th, nthr=nthr)
    l2, g2, gs = O.lml_and_grad(gp.X, gp.y, gp.alpha, th, return_scale=True)
    assert ok and abs(lml - l2) <= 1e-10 * max(1.0, abs(l2))
    assert np.max(np.abs(grad - g2) / np.maximum(np.maximum(np.abs(g2), 1e-2 * gs), 1e-30)) < 1e-9


@pytest.mark.parametrize("kind", [0x20, 0x22, 0x40, 0x10, 0x02])
def test_tiled_families_vs_oracle(kind):
    b = datasets.plasticc_like(1, n_points=150, first=9)
    t, f, e, bd = b.curve(0)
    p = H.prepare(t, f, e, bd, b.loglam, True, nthr=64)
    p2l = {i: float(v) for i, v in enumerate(b.loglam)}
    gp = O.ClosedFormGP(p2l, use_err=True, kind=kind).prepare(t, f, e, bd)
    P = O.n_params(kind)
    th = np.linspace(-0.5, 0.5, P)
    th[-1] = -3.0
    lml, grad, ok = H.lml_grad_tiled(p, th, nthr=128, family=kind)
    l2, g2, gs = O.lml_and_grad(gp.X, gp.y, gp.alpha, th, kind, return_scale=True)
    assert ok and abs(lml - l2) <= 1e-10 * max(1.0, abs(l2))
    assert np.max(np.abs(grad - g2) / np.maximum(np.maximum(np.abs(g2), 1e-2 * gs), 1e-30)) < 1e-9
    gp.fit(t, f, e, bd, theta=th)
    _, mo, so, _ = gp.augmentation(t.min(), t.max(), 8)
    m, s, l3, st = H.predict_grid_tiled(p, th, 6, 8, float(t.min()), float(t.max()), nthr=128, family=kind)
    assert st == 0 and abs(l3 - l2) <= 1e-10 * max(1.0, abs(l2))
    assert relerr(m, mo, floor=1e-3 * gp.y_std) < 1e-9 and relerr(s, so, floor=1e-6 * gp.y_std) < 1e-8


def test_tiled_evaluations_in_a_row_are_bit_identical():
    """The ring of stages and its barriers carry on from one evaluation to the next inside a kernel: the third evaluation of a CTA
    returns the bits of the first."""
    g = load_golden("csv34299")
    p = H.prepare(g["t"], g["flux"], g["flux_err"], g["band"], g["loglam"], False, nthr=64)
    a = H.lml_grad_tiled(p, g["thetas"][2], nthr=128, reps=1)
    b = H.lml_grad_tiled(p, g["thetas"][2], nthr=128, reps=3)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


def test_tiled_equals_shared_memory_path_to_rounding():
    g = load_golden("plasticc_3")
    p = H.prepare(g["t"], g["flux"], g["flux_err"], g["band"], g["loglam"], bool(g["use_err"]), nthr=64)
    for k in (1, 4):
        a = H.lml_grad(p, g["thetas"][k], nthr=128)
        b = H.lml_grad_tiled(p, g["thetas"][k], nthr=128)
        assert abs(a[0] - b[0]) <= 1e-12 * max(1.0, abs(a[0]))


def test_tiled_not_positive_definite_is_reported():
    t = np.array([0.0, 0.0, 1.0, 1.0, 2.0, 2.0, 3.0, 3.0])
    p = H.prepare(t, np.sin(t), np.ones(8), np.zeros(8, np.int32), np.array([3.5]), False, alpha0=0.0, nthr=64)
    lml, grad, ok = H.lml_grad_tiled(p, np.array([0.0, 0.0, 0.0, -800.0]), nthr=64)
    assert not ok and lml == -np.inf and np.all(grad == 0.0)
