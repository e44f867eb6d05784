"""The CUDA block algorithm (fulu_b200/csrc/gp_block.cuh) executed on host threads (tests/_host/host_gp_block.cpp: barriers for
__syncthreads/__syncwarp, emulated DMMA and shuffles) against the golden vectors of the real reference.  Same tolerances as the
GPU parity tests; lets the device code be checked where there is no GPU."""
import numpy as np
import pytest

import hostblock as H
from conftest import load_golden, relerr
from oracle import gp_oracle as O

CASES = ["readme", "readme_err", "testfixture", "csv34299", "csv34299_err", "csv34299_band2", "plasticc_0", "plasticc_1", "ztf_2", "ztf_5"]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("nthr", [32, 128, 160, 256])
def test_block_lml_grad_predict(name, nthr):
    if nthr == 256 and name not in ("csv34299", "ztf_2", "readme"):
        pytest.skip("8-warp emulation only on a few cases (slow)")
    if nthr in (32, 160) and name not in ("csv34299", "plasticc_0", "readme_err", "ztf_5"):
        pytest.skip("1- and 5-warp emulation only on a few cases")
    g = load_golden(name)
    p = H.prepare(g["t"], g["flux"], g["flux_err"], g["band"], g["loglam"], bool(g["use_err"]), nthr=nthr)
    assert p.status == 0
    np.testing.assert_allclose([p.stats[0], p.stats[2]], g["x_mean"], rtol=1e-14)
    np.testing.assert_allclose([p.stats[1], p.stats[3]], g["x_scale"], rtol=1e-12)
    np.testing.assert_allclose(p.stats[4:6], [g["y_mean"], g["y_std"]], rtol=1e-13)
    p2l = {i: float(v) for i, v in enumerate(g["loglam"])}
    gp = O.ClosedFormGP(p2l, use_err=bool(g["use_err"])).prepare(g["t"], g["flux"], g["flux_err"], g["band"])
    for k in (0, 2, 4, 6):
        th = g["thetas"][k]
        lml, grad, ok = H.lml_grad(p, th, nthr=nthr)
        assert ok
        assert abs(lml - g["lml"][k]) <= 1e-9 * max(1.0, abs(g["lml"][k]))
        _, _, gs = O.lml_and_grad(gp.X, gp.y, gp.alpha, th, return_scale=True)
        den = np.maximum(np.maximum(np.abs(g["grad"][k]), 1e-2 * gs), 1e-30)
        assert np.max(np.abs(grad - g["grad"][k]) / den) < 1e-9
    nb, n_obs = len(g["loglam"]), int(g["n_obs"])
    m, s, lml, st = H.predict_grid(p, g["theta_star"], nb, n_obs, float(g["t_min"]), float(g["t_max"]), nthr=nthr)
    assert st == 0
    assert relerr(m, g["flux_aug"], floor=1e-3 * float(g["y_std"])) < 1e-9
    assert relerr(s, g["flux_err_aug"]) < 1e-9
    assert abs(lml - float(g["lml_star"])) <= 1e-9 * max(1.0, abs(float(g["lml_star"])))


@pytest.mark.parametrize("name", ["csv34299", "readme_err", "ztf_5"])
def test_block_stored_k_gradient_equals_recomputed(name):
    """Default family: the gradient pass fed from the K slab the assembly filled (what the GPU path does, there in L2) against the
    pass that recomputes K_ij - same LML bit for bit, gradient to rounding; and against the reference fixture."""
    g = load_golden(name)
    p = H.prepare(g["t"], g["flux"], g["flux_err"], g["band"], g["loglam"], bool(g["use_err"]), nthr=128)
    for k in (0, 2, 5, 6):
        a = H.lml_grad(p, g["thetas"][k], nthr=128, family=0)
        b = H.lml_grad(p, g["thetas"][k], nthr=128, family=0x200)
        assert a[2] and b[2] and a[0] == b[0]
        np.testing.assert_allclose(b[1], a[1], rtol=1e-11, atol=1e-12 * np.abs(a[1]).max())
        assert abs(b[0] - g["lml"][k]) <= 1e-9 * max(1.0, abs(g["lml"][k]))


def test_block_not_positive_definite_is_reported():
    # duplicate observations with (almost) no noise: K is numerically singular
    t = np.array([0.0, 0.0, 1.0, 1.0, 2.0, 2.0, 3.0, 3.0])
    p = H.prepare(t, np.sin(t), np.ones(8), np.zeros(8, np.int32), np.array([3.5]), False, alpha0=0.0, nthr=64)
    lml, grad, ok = H.lml_grad(p, np.array([0.0, 0.0, 0.0, -800.0]), nthr=64)
    assert not ok and lml == -np.inf and np.all(grad == 0.0)


def test_block_predict_points_equals_grid():
    g = load_golden("ztf_0")
    p = H.prepare(g["t"], g["flux"], g["flux_err"], g["band"], g["loglam"], False, nthr=64)
    nb, n_obs = len(g["loglam"]), 16
    m, s, _, _ = H.predict_grid(p, g["theta_star"], nb, n_obs, float(g["t_min"]), float(g["t_max"]), nthr=64)
    qt = np.tile(np.linspace(float(g["t_min"]), float(g["t_max"]), n_obs), nb)
    qb = np.repeat(np.arange(nb, dtype=np.int32), n_obs)
    m2, s2, _, _ = H.predict_points(p, g["theta_star"], qt, qb, nthr=64)
    np.testing.assert_array_equal(m, m2)
    np.testing.assert_array_equal(s, s2)


def test_tile_swizzle_is_conflict_free():
    """Every fragment access pattern of gp_block.cuh hits 16 distinct 64-bit banks per half-warp (the layout comment's claim)."""
    swz = lambda c: (6 if c & 2 else 0) ^ (c & 4)
    el = lambda r, c: c * 8 + (r ^ swz(c))
    perm = lambda g: 4 + (g >> 1) if g & 1 else g >> 1
    assert sorted(el(r, c) for r in range(8) for c in range(8)) == list(range(64))
    pats = [lambda g, t: el(g, t), lambda g, t: el(g, 4 + t), lambda g, t: el(perm(g), t), lambda g, t: el(perm(g), 4 + t),
            lambda g, t: el(t, g), lambda g, t: el(4 + t, g), lambda g, t: el(t, perm(g)), lambda g, t: el(4 + t, perm(g))]
    for f in pats:
        for half in (0, 1):
            banks = {}
            for lane in range(16 * half, 16 * half + 16):
                a = f(lane >> 2, lane & 3)
                banks.setdefault(a % 16, set()).add(a)
            assert max(len(v) for v in banks.values()) == 1


def test_exp_accuracy():
    """gp_exp_tab (the kernel-matrix exponential: 16-entry table + degree-7 polynomial) against a long-double reference:
    <= 1.6 ulp on [-690, 0], exact at 0, 0 below the cut-off."""
    rng = np.random.default_rng(0)
    x = -(10.0 ** rng.uniform(-8, np.log10(690.0), 400000))
    got = H.exp_neg(x)
    ref = np.exp(x.astype(np.longdouble))
    err = np.abs((got - ref) / ref).astype(np.float64) / 2.0 ** -52
    assert err.max() < 1.6
    np.testing.assert_array_equal(H.exp_neg(np.array([0.0, -0.0, -1e-300, -690.5, -709.0, -1e6, -np.inf])), [1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0])


@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 15, 16, 17, 24, 31, 32, 33])
def test_block_sizes_around_tile_boundaries(n):
    """The bordered layout at every position of the border row inside its tile (n mod 8 = 0 opens a tile-row of its own) and at
    one and two tile-rows: LML, gradient and prediction against the closed-form oracle, default kernel and the 6-parameter one."""
    g = load_golden("csv34299")
    rng = np.random.default_rng(n)
    t, f, e, b = g["t"][:n], g["flux"][:n] + rng.normal(size=n), g["flux_err"][:n], g["band"][:n]
    p2l = {i: float(v) for i, v in enumerate(g["loglam"])}
    p = H.prepare(t, f, e, b, g["loglam"], False, nthr=64)
    for kind in (0, 34):
        gp = O.ClosedFormGP(p2l, kind=kind).prepare(t, f, e, b)
        th = rng.uniform(-1, 1, O.n_params(kind))
        lml, grad, ok = H.lml_grad(p, th, nthr=64, family=kind)
        l2, g2, gs = O.lml_and_grad(gp.X, gp.y, gp.alpha, th, kind, return_scale=True)
        assert ok and abs(lml - l2) <= 1e-10 * max(1.0, abs(l2))
        assert np.max(np.abs(grad - g2) / np.maximum(np.maximum(np.abs(g2), 1e-2 * gs), 1e-30)) < 1e-9
        gp.fit(t, f, e, b, theta=th)
        _, mo, so, _ = gp.augmentation(t.min(), t.max() + 1.0, 5)
        m, s, l3, st = H.predict_grid(p, th, len(g["loglam"]), 5, float(t.min()), float(t.max() + 1.0), nthr=64, family=kind)
        assert st == 0 and abs(l3 - l2) <= 1e-10 * max(1.0, abs(l2))
        assert relerr(m, mo, floor=1e-3 * gp.y_std) < 1e-9 and relerr(s, so, floor=1e-6 * gp.y_std) < 1e-8


@pytest.mark.parametrize("name,nthr", [("csv34299", 160), ("csv34299", 256), ("plasticc_0", 128), ("readme_err", 32), ("ztf_5", 64)])
def test_block_scratch_free_layout_is_bit_identical(name, nthr):
    """The scratch-free shared-memory layout (gp_block.cuh: no scratch column, the triangular inverse forms its W tiles in place and holds
    a block column's rows in registers across a barrier; the noise diagonal and y staged behind the vectors) - what the N = 104..127 and
    136..159 length classes use to hold one more CTA per SM - gives the bits of the scratch layout: LML and every gradient entry, with the
    K slab and without, at the thread counts those classes run with (160, 256) and others.  The harness allocates exactly the
    scratch-free footprint."""
    g = load_golden(name)
    p = H.prepare(g["t"], g["flux"], g["flux_err"], g["band"], g["loglam"], bool(g["use_err"]), nthr=128)
    if (p.n + 8) // 8 - 1 > 4 * (nthr // 32):
        pytest.skip("more than four rows of a block column per warp: plan() keeps the scratch layout there")
    for k in (0, 3, 6):
        for fam in (0, 0x200):
            a = H.lml_grad(p, g["thetas"][k], nthr=nthr, family=fam)
            b = H.lml_grad(p, g["thetas"][k], nthr=nthr, family=fam | 0x400)
            assert a[2] and b[2]
            assert a[0] == b[0]
            np.testing.assert_array_equal(a[1], b[1])
