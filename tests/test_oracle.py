"""The oracle against the golden vectors produced by the real reference (tests/golden/make_golden.py)."""
import warnings

import numpy as np
import pytest

from conftest import relerr
from oracle import gp_oracle as O
from oracle.fulu_sklearn_port import SklearnGPAugmentation


def _prep(g):
    p2l = {i: float(v) for i, v in enumerate(g["loglam"])}
    gp = O.ClosedFormGP(p2l, use_err=bool(g["use_err"]))
    gp.prepare(g["t"], g["flux"], g["flux_err"], g["band"])
    return p2l, gp


def test_closed_form_scaler_and_norm(golden):
    _, gp = _prep(golden)
    assert relerr(gp.x_mean, golden["x_mean"]) < 4e-15   # (a few ulp: the summation order of a mean over ~300 values)
    assert relerr(gp.x_scale, golden["x_scale"]) < 1e-13
    assert abs(gp.y_mean - golden["y_mean"]) <= 1e-15 * abs(golden["y_mean"])
    assert abs(gp.y_std - golden["y_std"]) <= 1e-14 * abs(golden["y_std"])


def test_closed_form_lml_and_grad(golden):
    _, gp = _prep(golden)
    for th, lml, grad in zip(golden["thetas"], golden["lml"], golden["grad"]):
        l, g, gs = O.lml_and_grad(gp.X, gp.y, gp.alpha, th, return_scale=True)
        assert abs(l - lml) <= 1e-10 * max(1.0, abs(lml)), (golden["name"], th)
        # the gradient is a cancelling trace: relative to the larger of |g| and 1e-3 x (sum of |terms|)
        assert np.max(np.abs(g - grad) / np.maximum(np.maximum(np.abs(grad), 1e-3 * gs), 1e-30)) < 1e-9, (golden["name"], th)


def test_closed_form_predict_fixed_theta(golden):
    _, gp = _prep(golden)
    g = golden
    n_obs = int(g["n_obs"])
    for theta, fm, fs in ((g["thetas"][5], g["flux_aug_theta5"], g["flux_err_aug_theta5"]), (g["theta_star"], g["flux_aug"], g["flux_err_aug"])):
        gp.fit(g["t"], g["flux"], g["flux_err"], g["band"], theta=theta)
        t_aug, m, s, pb = gp.augmentation(float(g["t_min"]), float(g["t_max"]), n_obs)
        np.testing.assert_array_equal(t_aug, g["t_aug"])
        np.testing.assert_array_equal(pb, g["band_aug"])
        # flux crosses zero, so 'relative' uses a floor of 1e-3 x the curve's flux scatter
        assert relerr(m, fm, floor=1e-3 * gp.y_std) < 1e-9
        assert relerr(s, fs, floor=1e-6 * gp.y_std) < 1e-8  # sklearn itself is ~1.5e-10 from an 80-bit evaluation


@pytest.mark.parametrize("name", ["readme", "testfixture", "csv34299", "csv34299_err", "plasticc_0", "ztf_0"])
def test_closed_form_fit_reaches_reference_lml(name):
    from conftest import load_golden

    g = load_golden(name)
    _, gp = _prep(g)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        theta, lml, n_eval = O.fit_theta(gp.X, gp.y, gp.alpha)
    assert lml >= float(g["lml_star"]) - 1e-6
    assert abs(lml - float(g["lml_star"])) < 1e-5


@pytest.mark.parametrize("name", ["readme", "testfixture_err", "csv34299", "plasticc_1"])
def test_sklearn_port_is_the_reference(name):
    """The port performs the same library calls as fulu/gp_aug.py, so it reproduces the fixture bit for bit."""
    from conftest import load_golden

    g = load_golden(name)
    p2l = {i: float(v) for i, v in enumerate(g["loglam"])}
    aug = SklearnGPAugmentation(p2l, use_err=bool(g["use_err"]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert aug.fit(g["t"], g["flux"], g["flux_err"], g["band"]) is aug
        t_aug, f, e, pb = aug.augmentation(float(g["t_min"]), float(g["t_max"]), int(g["n_obs"]))
    np.testing.assert_array_equal(aug.reg.kernel_.theta, g["theta_star"])
    assert aug.reg.log_marginal_likelihood_value_ == float(g["lml_star"])
    np.testing.assert_array_equal(f, g["flux_aug"])
    np.testing.assert_array_equal(e, g["flux_err_aug"])
    np.testing.assert_array_equal(t_aug, g["t_aug"])


@pytest.mark.parametrize("name", ["ddf_1024", "ddf_1536"])
def test_closed_form_on_config5_fixtures(name):
    """The long curves of BASELINE config #5 as the real reference fitted them: LML and gradient at theta = 0, a restart point and
    theta*, the augmentation at theta*."""
    from conftest import load_golden

    g = load_golden(name)
    _, gp = _prep(g)
    for k in (0, 4, 6):
        l, gr, gs = O.lml_and_grad(gp.X, gp.y, gp.alpha, g["thetas"][k], return_scale=True)
        assert abs(l - g["lml"][k]) <= 1e-10 * max(1.0, abs(g["lml"][k]))
        assert np.max(np.abs(gr - g["grad"][k]) / np.maximum(np.maximum(np.abs(g["grad"][k]), 1e-3 * gs), 1e-30)) < 1e-9
    gp.fit(g["t"], g["flux"], g["flux_err"], g["band"], theta=g["theta_star"])
    t_aug, m, s, pb = gp.augmentation(float(g["t_min"]), float(g["t_max"]), int(g["n_obs"]))
    np.testing.assert_array_equal(t_aug, g["t_aug"])
    assert relerr(m, g["flux_aug"], floor=1e-3 * gp.y_std) < 1e-9
    assert relerr(s, g["flux_err_aug"], floor=1e-6 * gp.y_std) < 1e-8


def test_start_points_match_survey():
    s = O.start_points(4)
    assert s.shape == (6, 4)
    np.testing.assert_allclose(s[1], [-2.88882052, 10.37808043, 5.34185792, 2.27169555], atol=1e-8)
    np.testing.assert_allclose(s[5], [-4.50748893, 0.5700379, -1.56702386, -4.8071267], atol=1e-8)


def test_pipeline_checkers():
    """The restatements that check the device-side ingest and peak consumer."""
    off, ids = O.table_to_csr([5, 5, 5, 2, 2, 9])
    assert off.tolist() == [0, 3, 5, 6] and ids.tolist() == [5, 2, 9]
    assert O.table_to_csr([])[0].tolist() == [0]
    t = np.tile(np.linspace(0.0, 3.0, 4), 2)
    tt, ss, peak = O.band_sum_peak(t, np.array([0.0, 1.0, 5.0, 2.0, 1.0, 1.0, -3.0, 4.0]), np.repeat([0, 1], 4))
    assert ss.tolist() == [1.0, 2.0, 2.0, 6.0] and peak == 3.0


def test_staged_reference_class_is_the_cpu_arm_and_agrees_with_the_port():
    """oracle/ref_arm.py: where __graft_entry__.build() staged the reference's own files (oracle/_ref/, build container), bench.py's CPU
    arm runs the real class; it must give the port's numbers bit for bit (both drive the same scikit-learn calls)."""
    import warnings

    from fulu_b200 import datasets
    from oracle import ref_arm
    from oracle.fulu_sklearn_port import SklearnGPAugmentation

    if not ref_arm.staged() and not ref_arm.stage():
        pytest.skip("reference files not staged (no /root/reference here)")
    cls = ref_arm.load()
    assert cls is not None and cls.__module__ == "fulu.gp_aug"
    p2l, t, flux, err, pb = datasets.readme_example()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = cls(p2l).fit(t, flux, err, pb)
        b = SklearnGPAugmentation(p2l).fit(t, flux, err, pb)
        ra = a.augmentation(t.min(), t.max(), 50)
        rb = b.augmentation(t.min(), t.max(), 50)
    assert a.reg.log_marginal_likelihood_value_ == b.reg.log_marginal_likelihood_value_
    np.testing.assert_array_equal(a.reg.kernel_.theta, b.reg.kernel_.theta)
    np.testing.assert_array_equal(ra[1], rb[1])
    np.testing.assert_array_equal(ra[2], rb[2])
