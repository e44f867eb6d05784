"""The drop-in class and the batched entry point: the reference's own test (tests/test_aug.py:52-67) plus edge cases."""
import warnings

import numpy as np
import pytest
import torch

from conftest import load_golden, relerr

pytestmark = pytest.mark.gpu


def test_reference_test_aug_with_sample_data():
    """Port of /root/reference/tests/test_aug.py:52-67 for the GP parametrisations (default and use_err=True)."""
    from numpy.testing import assert_allclose

    from fulu_b200 import GaussianProcessesAugmentation, datasets

    for init_kwargs in ({}, dict(use_err=True)):
        n_aug = 100
        passband2lam, *lc = datasets.reference_test_fixture()
        n_band = len(passband2lam)
        t, *_ = lc
        aug = GaussianProcessesAugmentation(passband2lam, **init_kwargs)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            aug_fit = aug.fit(*lc)
        assert aug_fit is aug, ".fit() must return self"
        t_aug, flux_aug, flux_err_aug, passband_aug = aug.augmentation(t.min(), t.max(), n_aug)
        assert_allclose(t_aug, np.tile(np.linspace(t.min(), t.max(), n_aug), n_band))
        assert flux_aug.size == n_aug * n_band
        assert np.all(flux_aug >= 0.0)
        assert flux_err_aug.size == n_aug * n_band
        assert np.all(flux_err_aug >= 0.0)
        assert passband_aug.size == n_aug * n_band
        assert set(passband_aug) == set(passband2lam)


@pytest.mark.parametrize("name", ["readme", "csv34299", "csv34299_err", "testfixture"])
def test_drop_in_class_matches_reference_outputs(name):
    from fulu_b200 import ConvergenceWarning, GaussianProcessesAugmentation

    g = load_golden(name)
    p2l = {i: float(v) for i, v in enumerate(g["loglam"])}
    aug = GaussianProcessesAugmentation(p2l, use_err=bool(g["use_err"]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", ConvergenceWarning)
        aug.fit(g["t"], g["flux"], g["flux_err"], g["band"])
    assert aug.reg.log_marginal_likelihood_value_ >= float(g["lml_star"]) - 1e-6
    np.testing.assert_allclose(aug.ss.mean_, g["x_mean"], rtol=1e-14)
    np.testing.assert_allclose(aug.ss.scale_, g["x_scale"], rtol=1e-12)
    t_aug, f, e, pb = aug.augmentation(float(g["t_min"]), float(g["t_max"]), int(g["n_obs"]))
    np.testing.assert_array_equal(t_aug, g["t_aug"])
    np.testing.assert_array_equal(pb, g["band_aug"])
    # after optimisation theta agrees to ~1e-5, so the curves agree to ~1e-4 of the flux scale
    ystd = float(g["y_std"])
    if np.max(np.abs(aug.reg.kernel_.theta - g["theta_star"])) < 1e-3:
        assert relerr(f, g["flux_aug"], floor=ystd) < 1e-3
        assert relerr(e, g["flux_err_aug"], floor=1e-2 * ystd) < 1e-2
    # reg shim: LML at a given theta, with gradient
    l, gr = aug.reg.log_marginal_likelihood(g["thetas"][1], eval_gradient=True)
    assert abs(l - g["lml"][1]) <= 1e-9 * abs(g["lml"][1])
    assert gr.shape == (4,)
    # predict at arbitrary points == the same points of the grid
    fp, ep = aug.predict(t_aug[::7], pb[::7])
    np.testing.assert_allclose(fp, f[::7], rtol=1e-12, atol=1e-12 * ystd)
    np.testing.assert_allclose(ep, e[::7], rtol=1e-10)


def test_string_passband_keys_and_key_error():
    from fulu_b200 import GaussianProcessesAugmentation, datasets

    p2l, t, flux, err, pb = datasets.readme_example()
    aug = GaussianProcessesAugmentation(p2l)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        aug.fit(t, flux, err, pb)
    t_aug, f, e, pba = aug.augmentation(t.min(), t.max(), 10)
    assert list(pba[:10]) == ["u"] * 10 and f.shape == (30,)
    with pytest.raises(KeyError):
        aug.fit(t, flux, err, np.array(["q"] * len(t)))
    with pytest.raises(ValueError):
        aug.fit(t, np.where(np.arange(len(t)) == 3, np.nan, flux), err, pb)


def test_batched_entry_point_equals_per_curve_and_is_batch_independent():
    import fulu_b200
    from fulu_b200 import datasets

    eng = fulu_b200.GPEngine()
    hb = datasets.ztf_like(40, first=200, n_lo=20, n_hi=180)  # ragged: several N buckets incl. the global-memory one
    res = fulu_b200.fit_augment_batch(hb, n_obs=32, engine=eng)
    assert res.flux_aug.shape == (40, 2, 32) and np.all(np.isfinite(res.flux_aug)) and np.all(res.flux_err_aug >= 0)
    # every curve alone gives bit-identical results: no cross-curve coupling, deterministic reductions
    for i in (0, 7, 23, 39):
        one = fulu_b200.fit_augment_batch(hb.slice(i, i + 1), n_obs=32, engine=eng)
        np.testing.assert_array_equal(one.theta[0], res.theta[i])
        np.testing.assert_array_equal(one.flux_aug[0], res.flux_aug[i])
        np.testing.assert_array_equal(one.flux_err_aug[0], res.flux_err_aug[i])
        assert one.n_eval[0] == res.n_eval[i]
    # small window (forces slot recycling) does not change anything either
    res2 = fulu_b200.fit_augment_batch(hb, n_obs=32, engine=eng, window=7)
    np.testing.assert_array_equal(res2.theta, res.theta)
    np.testing.assert_array_equal(res2.lml, res.lml)


def test_order_list_processes_only_the_listed_curves_in_place():
    """fulu_gp_batch.order: a launch over an index list reads the batch where it lies and writes only the listed rows."""
    import fulu_b200
    from fulu_b200 import DeviceBatch, datasets

    eng = fulu_b200.GPEngine()
    hb = datasets.ztf_like(24, first=900, n_lo=20, n_hi=100)
    db = DeviceBatch.from_host(hb, eng.device)
    theta = torch.tensor(np.tile([0.1, -1.0, 1.5, -3.0], (24, 1)))
    full = eng.lml(db, theta)
    order = np.array([17, 3, 8, 21], np.int32)
    sub = db.select(order, n_max=104)
    part = eng.lml(sub, theta)
    np.testing.assert_allclose(part["lml"].cpu().numpy()[order], full["lml"].cpu().numpy()[order], rtol=1e-13)
    np.testing.assert_allclose(part["grad"].cpu().numpy()[order], full["grad"].cpu().numpy()[order], rtol=1e-9, atol=1e-9)
    sentinel = dict(theta=torch.full((24, 4), -7.0, dtype=torch.float64, device=eng.device), lml=torch.full((24,), -7.0, dtype=torch.float64, device=eng.device),
                    n_eval=torch.full((24,), -7, dtype=torch.int32, device=eng.device), status=torch.full((24,), -7, dtype=torch.int32, device=eng.device))
    eng.fit(sub, out=sentinel)
    rest = np.setdiff1d(np.arange(24), order)
    assert np.all(sentinel["lml"].cpu().numpy()[rest] == -7.0) and np.all(sentinel["n_eval"].cpu().numpy()[rest] == -7)
    assert np.all(sentinel["n_eval"].cpu().numpy()[order] > 0) and np.all(np.isfinite(sentinel["lml"].cpu().numpy()[order]))
    # the same curves fitted as their own batch (same length class => same launch geometry) give the same bits
    own = eng.fit(DeviceBatch.from_host(hb.take(order), eng.device, n_max=104))
    np.testing.assert_array_equal(own["theta"].cpu().numpy(), sentinel["theta"].cpu().numpy()[order])


def test_curve_longer_than_n_max_is_flagged_not_run():
    """fulu_gp_batch.n_max sizes the shared memory / slab of every kernel: a listed curve longer than the caller said is flagged
    BAD_INPUT and skipped by every consumer instead of overrunning the block's memory; the others are unaffected."""
    import fulu_b200
    from fulu_b200 import DeviceBatch, datasets

    eng = fulu_b200.GPEngine()
    hb = datasets.ztf_like(8, first=900, n_lo=20, n_hi=100)          # lengths 49 92 41 42 71 51 28 95
    long_ = hb.lengths() > 63
    assert long_.sum() == 3
    db = DeviceBatch.from_host(hb, eng.device)
    theta = torch.tensor(np.tile([0.1, -1.0, 1.5, -3.0], (8, 1)))
    full = eng.lml(db, theta)
    sub = db.select(np.arange(8, dtype=np.int32), n_max=63)
    part = eng.lml(sub, theta)
    st = part["status"].cpu().numpy()
    assert np.all(st[long_] & 1) and np.all(st[~long_] == 0)
    assert np.all(np.isnan(part["lml"].cpu().numpy()[long_]))
    np.testing.assert_allclose(part["lml"].cpu().numpy()[~long_], full["lml"].cpu().numpy()[~long_], rtol=1e-10)   # another length class: rounding only
    fit = eng.fit(sub)
    pred = eng.predict_grid(sub, fit["theta"], 8)
    assert np.all(fit["status"].cpu().numpy()[long_] & 1) and np.all(pred["status"].cpu().numpy()[long_] & 1)
    assert np.all(np.isfinite(pred["mean"].cpu().numpy()[~long_])) and np.all(np.isnan(pred["mean"].cpu().numpy()[long_]))


def test_edge_cases():
    import fulu_b200
    from fulu_b200 import CurveBatch, DeviceBatch

    eng = fulu_b200.GPEngine()
    loglam = np.log10([3751.36, 4741.64, 6173.23])
    rng = np.random.default_rng(3)
    curves = [
        (np.array([5.0]), np.array([2.0]), np.array([0.1]), np.array([1])),                           # N = 1
        (np.array([1.0, 2.0]), np.array([2.0, 3.0]), np.array([0.1, 0.1]), np.array([0, 2])),         # N = 2
        (np.arange(12.0), 5 + np.sin(np.arange(12.0)), np.full(12, 0.2), np.full(12, 1)),             # single band
        (np.repeat(np.arange(6.0), 2), np.repeat(3 + np.cos(np.arange(6.0)), 2) + rng.normal(0, 0.01, 12), np.full(12, 0.1),
         np.tile([0, 0], 6)),                                                                         # duplicate (t, band) rows
        (np.arange(10.0), np.full(10, 7.0), np.full(10, 0.3), rng.integers(0, 3, 10)),                # constant flux
        (np.arange(8.0), np.array([1, 2, np.nan, 4, 5, 6, 7, 8.0]), np.full(8, 0.1), rng.integers(0, 3, 8)),  # NaN flux
        (np.arange(8.0), -50 + 10 * np.sin(np.arange(8.0)), np.zeros(8), rng.integers(0, 3, 8)),      # negative flux, zero errors
    ]
    hb = CurveBatch.from_curves(curves, loglam)
    res = fulu_b200.fit_augment_batch(hb, n_obs=5, engine=eng)
    assert res.status[5] & 1 and np.all(np.isnan(res.flux_aug[5]))                 # bad input flagged, batch not poisoned
    ok = [0, 1, 2, 3, 4, 6]
    assert np.all(np.isfinite(res.flux_aug[ok])) and np.all(res.flux_err_aug[ok] >= 0)
    np.testing.assert_allclose(res.flux_aug[4], 7.0, atol=1e-6)                    # constant flux predicts the constant
    assert np.any(res.flux_aug[6] < 0)                                             # GP output is not clipped at zero
    assert abs(res.theta[2][2]) < 1e-12                                            # single band: wavelength scale is a flat direction
    res_e = fulu_b200.fit_augment_batch(hb, n_obs=5, use_err=True, engine=eng)
    assert res_e.status[4] & 1                                                     # constant flux + use_err: alpha = inf -> flagged
    assert np.all(np.isfinite(res_e.flux_aug[[0 + 1, 2, 3, 6]]))
    # empty batch
    empty = CurveBatch.from_curves([], loglam)
    r0 = fulu_b200.fit_augment_batch(empty, n_obs=5, engine=eng)
    assert r0.flux_aug.shape == (0, 3, 5)
    # compare the small cases with sklearn
    from oracle.fulu_sklearn_port import SklearnGPAugmentation

    p2l = {i: float(v) for i, v in enumerate(loglam)}
    for i in (1, 2, 3, 6):
        t, f, e, b = hb.curve(i)
        ref = SklearnGPAugmentation(p2l)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref.fit(t, f, e, b)
        assert res.lml[i] >= ref.reg.log_marginal_likelihood_value_ - 1e-6, i


def test_invariances_at_full_size():
    """Size-independent properties on the headline shape (N=100, 6 bands, M=768): observation order and a time shift of the
    whole curve do not change LML / predictions beyond rounding."""
    import fulu_b200
    from fulu_b200 import CurveBatch, DeviceBatch, datasets

    eng = fulu_b200.GPEngine()
    hb = datasets.plasticc_like(256, first=5000)
    rng = np.random.default_rng(0)
    perm_curves, shift_curves = [], []
    for i in range(hb.n_curves):
        t, f, e, b = hb.curve(i)
        p = rng.permutation(len(t))
        perm_curves.append((t[p], f[p], e[p], b[p]))
        shift_curves.append((t + 1000.0, f, e, b))
    theta = torch.tensor(np.tile([0.2, -1.5, 2.5, -4.5], (hb.n_curves, 1)))
    base = DeviceBatch.from_host(hb, eng.device)
    a = eng.lml(base, theta)
    pa = eng.predict_grid(base, theta, 128)
    b_ = eng.lml(DeviceBatch.from_host(CurveBatch.from_curves(perm_curves, hb.loglam), eng.device), theta)
    pb_ = eng.predict_grid(DeviceBatch.from_host(CurveBatch.from_curves(perm_curves, hb.loglam), eng.device), theta, 128)
    c_ = eng.lml(DeviceBatch.from_host(CurveBatch.from_curves(shift_curves, hb.loglam), eng.device), theta)
    np.testing.assert_allclose(b_["lml"].cpu().numpy(), a["lml"].cpu().numpy(), rtol=1e-11)
    np.testing.assert_allclose(c_["lml"].cpu().numpy(), a["lml"].cpu().numpy(), rtol=1e-9)
    np.testing.assert_allclose(pb_["std"].cpu().numpy(), pa["std"].cpu().numpy(), rtol=1e-9)
    ystd = a["ynorm"].cpu().numpy()[:, 1][:, None, None]
    assert np.max(np.abs(pb_["mean"].cpu().numpy() - pa["mean"].cpu().numpy()) / ystd) < 1e-10
    # gradient vs central finite differences of the LML (sklearn's own test_lml_gradient pattern)
    h = 1e-5
    for k in range(4):
        d = torch.zeros(4, dtype=torch.float64)
        d[k] = h
        up = eng.lml(base, theta + d)["lml"].cpu().numpy()
        dn = eng.lml(base, theta - d)["lml"].cpu().numpy()
        fd = (up - dn) / (2 * h)
        np.testing.assert_allclose(a["grad"].cpu().numpy()[:, k], fd, rtol=2e-5, atol=2e-5)


def test_more_long_curves_than_sms_in_one_class():
    """a class of long curves (matrix in a per-CTA global slab) holding more curves than the GPU has SMs.  The
    prediction kernel runs up to two CTAs per SM, each with a slab of its own: the workspace must hold one slab per CTA that can be at
    work, whichever kernel asks for more.  2.5 curves per SM of N = 310 (tiled placement) and of N = 250 (slab placement) through the
    batched entry point: every curve finite, and a sample of them bit-identical to the same curve fitted alone."""
    import fulu_b200
    from fulu_b200 import datasets

    eng = fulu_b200.GPEngine()
    sms = torch.cuda.get_device_properties(eng.device).multi_processor_count
    for n in (310, 250):
        hb = datasets.ddf_like(int(2.5 * sms), n_points=n, first=500)
        res = fulu_b200.fit_augment_batch(hb, n_obs=8, engine=eng)
        assert np.all(np.isfinite(res.lml)) and np.all(np.isfinite(res.flux_aug)) and np.all(np.isfinite(res.flux_err_aug))
        assert not np.any(res.status & ~8)
        for i in (0, sms - 1, sms, 2 * sms + 3, hb.n_curves - 1):
            one = fulu_b200.fit_augment_batch(hb.slice(i, i + 1), n_obs=8, engine=eng)
            assert one.lml[0] == res.lml[i] and np.array_equal(one.flux_aug[0], res.flux_aug[i]) and np.array_equal(one.theta[0], res.theta[i]), (n, i)


def test_query_band_outside_the_table_is_flagged_not_indexed():
    """fulu_gp_predict_points_batch validates the query bands (the header promises per-curve problems in status[]);
    an empty curve at the end of the CSR arrays is flagged without its first observation being read."""
    import fulu_b200
    from fulu_b200 import CurveBatch, DeviceBatch, datasets

    eng = fulu_b200.GPEngine()
    hb = datasets.plasticc_like(2, n_points=40, first=9)
    db = DeviceBatch.from_host(hb, eng.device)
    theta = torch.zeros(2, 4, dtype=torch.float64)
    qt = np.array([59010.0, 59020.0, 59030.0, 59040.0])
    out = eng.predict_points(db, theta, np.array([0, 2, 4], np.int64), qt, np.array([0, 5, 1, 6], np.int32))   # band 6 of 6: out of range
    st = out["status"].cpu().numpy()
    assert st[0] == 0 and st[1] & 1
    assert np.all(np.isfinite(out["mean"].cpu().numpy()[:2]))
    curves = [hb.curve(0), (np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0, np.int32))]          # the last curve is empty
    db2 = DeviceBatch.from_host(CurveBatch.from_curves(curves, hb.loglam), eng.device, n_max=40)
    g = eng.predict_grid(db2, theta, 4)
    assert int(g["status"][0]) == 0 and int(g["status"][1]) & 1 and bool(torch.isnan(g["mean"][1]).all())


def test_fused_fit_is_bit_identical_to_the_rounds(monkeypatch):
    """FULU_GP_FUSED=1 runs a fit as ONE persistent launch (evaluation groups + stepper warps per SM, csrc/gp_device.cuh
    fit_fused_kernel) instead of rounds of two kernels: same block code, same geometry per group - the same bits, whatever the number
    of slots and stepper warps, for shared-memory classes of several residencies and the slab class."""
    import fulu_b200
    from fulu_b200 import datasets

    eng = fulu_b200.GPEngine()
    hb = datasets.ztf_like(700, first=5000, n_lo=20, n_hi=300)   # every length class of config #4: the short ones stay below one wave of CTAs (not fused), the others fuse
    monkeypatch.setenv("FULU_GP_FUSED", "0")
    ref = fulu_b200.fit_augment_batch(hb, n_obs=16, engine=eng)
    for slots, steppers in (("32", "1"), ("12", "3"), ("5", "4")):
        monkeypatch.setenv("FULU_GP_FUSED", "1")
        monkeypatch.setenv("FULU_GP_FUSED_SLOTS", slots)
        monkeypatch.setenv("FULU_GP_FUSED_STEPPERS", steppers)
        res = fulu_b200.fit_augment_batch(hb, n_obs=16, engine=eng)
        np.testing.assert_array_equal(res.theta, ref.theta)
        np.testing.assert_array_equal(res.lml, ref.lml)
        np.testing.assert_array_equal(res.n_eval, ref.n_eval)
        np.testing.assert_array_equal(res.status, ref.status)
        np.testing.assert_array_equal(res.flux_aug, ref.flux_aug)
    # the headline class (four groups per SM) with more runs than slots: continuous batching inside the CTAs
    hp = datasets.plasticc_like(1200, first=300)
    monkeypatch.setenv("FULU_GP_FUSED", "0")
    ref = fulu_b200.fit_augment_batch(hp, n_obs=8, engine=eng)
    monkeypatch.setenv("FULU_GP_FUSED", "1")
    monkeypatch.delenv("FULU_GP_FUSED_SLOTS")
    monkeypatch.delenv("FULU_GP_FUSED_STEPPERS")
    res = fulu_b200.fit_augment_batch(hp, n_obs=8, engine=eng)
    np.testing.assert_array_equal(res.theta, ref.theta)
    np.testing.assert_array_equal(res.lml, ref.lml)
    np.testing.assert_array_equal(res.n_eval, ref.n_eval)
