"""Non-default covariance families on the GPU, through the C ABI: the kernels users of the reference pass besides
the default - C*RBF+Matern()+White (fulu/gp_aug.py:24, the reference's own test tests/test_aug.py:37), C*Matern()*RBF+Matern()+White
(notebook_examples/plotting.ipynb:33), C*RBF+RationalQuadratic()+White, Matern nu = 1/2 and 5/2, and other orderings of the same
sum/product.  Same tolerances as tests/test_gpu_parity.py: fixed theta -> LML / mean / std within 1e-9 relative; after
optimisation -> final LML no worse than the reference's by more than 1e-6."""
import warnings

import numpy as np
import pytest
import torch

from conftest import family_golden_names, family_kernel, load_golden, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import fulu_b200

    return fulu_b200.GPEngine()


def _load(name, engine):
    from fulu_b200 import CurveBatch, DeviceBatch
    from fulu_b200.engine import kernel_spec_from_sklearn
    from oracle import gp_oracle as O

    g = load_golden(name)
    spec = kernel_spec_from_sklearn(family_kernel(g["kernel"]))
    hb = CurveBatch.from_curves([(g["t"], g["flux"], g["flux_err"], g["band"])], g["loglam"])
    db = DeviceBatch.from_host(hb, engine.device, use_err=bool(g["use_err"]), family=spec.family)
    p2l = {i: float(v) for i, v in enumerate(g["loglam"])}
    gp = O.ClosedFormGP(p2l, kind=spec.family, use_err=bool(g["use_err"])).prepare(g["t"], g["flux"], g["flux_err"], g["band"])
    return g, spec, db, gp


@pytest.mark.parametrize("name", family_golden_names())
def test_fixed_theta_parity_with_reference_fixtures(engine, name):
    from oracle import gp_oracle as O

    g, spec, db, gp = _load(name, engine)
    for k in range(7):   # the kernel's own theta, the five restart points, theta*
        th = spec.from_sklearn(g["thetas"][k])
        out = engine.lml(db, torch.tensor(th[None, :].copy()))
        lml, grad = float(out["lml"].cpu()[0]), spec.to_sklearn(out["grad"].cpu().numpy()[0])
        assert abs(lml - g["lml"][k]) <= 1e-9 * max(1.0, abs(g["lml"][k])), (name, k)
        _, _, gs = O.lml_and_grad(gp.X, gp.y, gp.alpha, th, spec.family, return_scale=True)
        den = np.maximum(np.maximum(np.abs(g["grad"][k]), 1e-1 * spec.to_sklearn(gs)), 1e-30)   # a cancelling trace at cond(K) up to 1e8
        assert np.max(np.abs(grad - g["grad"][k]) / den) < 1e-9, (name, k)
    n_obs = int(g["n_obs"])
    tmin, tmax = np.array([float(g["t_min"])]), np.array([float(g["t_max"])])
    for theta, fm, fs in ((g["thetas"][int(g["k_fix"])], g["flux_aug_theta5"], g["flux_err_aug_theta5"]), (g["theta_star"], g["flux_aug"], g["flux_err_aug"])):
        out = engine.predict_grid(db, torch.tensor(spec.from_sklearn(theta)[None, :].copy()), n_obs, tmin, tmax)
        assert int(out["status"].cpu()[0]) == 0
        assert relerr(out["mean"].cpu().numpy().reshape(-1), fm, floor=1e-3 * float(g["y_std"])) < 1e-9, name
        assert relerr(out["std"].cpu().numpy().reshape(-1), fs, floor=1e-6 * float(g["y_std"])) < 1e-9, name


@pytest.mark.parametrize("name", family_golden_names())
def test_fit_reaches_reference_lml(engine, name):
    g, spec, db, gp = _load(name, engine)
    fit = engine.fit(db, spec)
    lml, nev = float(fit["lml"].cpu()[0]), int(fit["n_eval"].cpu()[0])
    assert lml >= float(g["lml_star"]) - 1e-6, (name, lml, float(g["lml_star"]))
    assert nev <= 3 * int(g["n_eval"]), (name, nev, int(g["n_eval"]))
    chk = float(engine.lml(db, fit["theta"])["lml"].cpu()[0])   # the reported LML is the LML at the reported theta
    assert abs(chk - lml) <= 1e-10 * max(1.0, abs(lml))


def test_reference_test_aug_with_matern_kernel():
    """Port of /root/reference/tests/test_aug.py:52-67 for its parametrisation :37 (kernel = C(1.0)*RBF([1, 1]) + Matern() + White)."""
    from numpy.testing import assert_allclose
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel

    from fulu_b200 import GaussianProcessesAugmentation, datasets

    n_aug = 100
    passband2lam, *lc = datasets.reference_test_fixture()
    n_band = len(passband2lam)
    t, *_ = lc
    aug = GaussianProcessesAugmentation(passband2lam, kernel=ConstantKernel(1.0) * RBF([1, 1]) + Matern() + WhiteKernel())
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
    g = load_golden("fam_rbf_matern_testfixture")
    assert aug.reg.log_marginal_likelihood_value_ >= float(g["lml_star"]) - 1e-6


@pytest.mark.parametrize("name", ["fam_permuted_matern25_testfixture", "fam_notebook_csv34299", "fam_rbf_rq_csv34299_err"])
def test_drop_in_class_with_kernel_objects(name):
    """reg.kernel_ is the user's kernel object with the fitted theta in ITS order; the LML shim takes and returns that order."""
    from fulu_b200 import ConvergenceWarning, GaussianProcessesAugmentation

    g = load_golden(name)
    p2l = {i: float(v) for i, v in enumerate(g["loglam"])}
    kernel = family_kernel(g["kernel"])
    aug = GaussianProcessesAugmentation(p2l, kernel=kernel, use_err=bool(g["use_err"]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", ConvergenceWarning)
        aug.fit(g["t"], g["flux"], g["flux_err"], g["band"])
    assert aug.reg.log_marginal_likelihood_value_ >= float(g["lml_star"]) - 1e-6
    assert type(aug.reg.kernel_) is type(kernel) and aug.reg.kernel_.theta.shape == g["theta_star"].shape
    for k in (1, int(g["k_fix"])):
        l, gr = aug.reg.log_marginal_likelihood(g["thetas"][k], eval_gradient=True)
        assert abs(l - g["lml"][k]) <= 1e-9 * max(1.0, abs(g["lml"][k]))
        big = np.abs(g["grad"][k]) > 1e-3 * np.abs(g["grad"][k]).max()
        np.testing.assert_allclose(gr[big], g["grad"][k][big], rtol=1e-6)
    # the LML sklearn computes for OUR theta with ITS kernel object is the LML we report
    from sklearn.gaussian_process import GaussianProcessRegressor

    X = aug.ss.transform(np.c_[g["t"], g["loglam"][g["band"]]])
    alpha = (g["flux_err"] / np.std(g["flux"])) ** 2 if bool(g["use_err"]) else 1e-10
    reg = GaussianProcessRegressor(kernel=aug.reg.kernel_, optimizer=None, normalize_y=True, alpha=alpha).fit(X, g["flux"])
    assert abs(reg.log_marginal_likelihood_value_ - aug.reg.log_marginal_likelihood_value_) <= 1e-9 * max(1.0, abs(reg.log_marginal_likelihood_value_))
    t_aug, f, e, pb = aug.augmentation(float(g["t_min"]), float(g["t_max"]), int(g["n_obs"]))
    fr, er = reg.predict(aug.ss.transform(np.c_[t_aug, g["loglam"][np.repeat(np.arange(len(p2l)), int(g["n_obs"]))]]), return_std=True)
    assert relerr(f, fr, floor=1e-3 * float(g["y_std"])) < 1e-9
    assert relerr(e, er, floor=1e-6 * float(g["y_std"])) < 1e-9


@pytest.mark.parametrize("kname", ["rbf_matern", "notebook", "rbf_rq"])
def test_batched_entry_point_against_live_sklearn(engine, kname):
    """12 seeded PLAsTiCC-like curves per kernel through fit_augment_batch next to the sklearn port with the same kernel object:
    final LML >= the reference's - 1e-6, and at OUR theta* sklearn's own LML / mean / std agree at 1e-9."""
    import fulu_b200
    from fulu_b200 import datasets
    from fulu_b200.engine import kernel_spec_from_sklearn
    from oracle.fulu_sklearn_port import SklearnGPAugmentation
    from sklearn.gaussian_process import GaussianProcessRegressor

    hb = datasets.plasticc_like(12, n_points=100, first=2000)
    spec = kernel_spec_from_sklearn(family_kernel(kname))
    res = fulu_b200.fit_augment_batch(hb, n_obs=32, engine=engine, spec=spec)
    assert res.theta.shape == (12, spec.n_params)
    p2l = {i: float(v) for i, v in enumerate(hb.loglam)}
    deficits = []
    for i in range(hb.n_curves):
        t, f, e, b = hb.curve(i)
        ref = SklearnGPAugmentation(p2l, kernel=family_kernel(kname))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref.fit(t, f, e, b)
        deficits.append(ref.reg.log_marginal_likelihood_value_ - res.lml[i])
        X = ref.ss.transform(np.c_[t, hb.loglam[b]])
        reg = GaussianProcessRegressor(kernel=family_kernel(kname).clone_with_theta(res.theta[i]), optimizer=None, normalize_y=True).fit(X, f)
        assert abs(reg.log_marginal_likelihood_value_ - res.lml[i]) <= 1e-9 * max(1.0, abs(res.lml[i])), i
        tq = np.tile(np.linspace(t.min(), t.max(), 32), hb.n_bands)
        fr, er = reg.predict(ref.ss.transform(np.c_[tq, np.repeat(hb.loglam, 32)]), return_std=True)
        assert relerr(res.flux_aug[i].reshape(-1), fr, floor=1e-3 * reg._y_train_std) < 1e-9, i
        assert relerr(res.flux_err_aug[i].reshape(-1), er, floor=1e-6 * reg._y_train_std) < 1e-9, i
    # The 5/6-parameter landscapes are multi-modal with nearly flat valleys: a restart can end in another basin than sklearn's run
    # from the same start when the two differ by rounding (the same happens between two BLAS builds of sklearn).  With sklearn's own
    # six starts the bar of 1e-6 is therefore asked of all curves but at most one of the twelve (INTEGRATION.md states this
    # per-family tolerance; the default kernel's tests ask it of every curve) ...
    deficits = np.array(deficits)
    print(kname, "LML deficits vs sklearn (positive = sklearn better):", np.array2string(deficits, precision=2))
    assert np.count_nonzero(deficits > 1e-6) <= 1, deficits
    assert np.median(deficits) <= 1e-8
    # ... and with the insurance starts (KernelSpec.extra_starts: six more runs from an independent stream, sklearn's six untouched)
    # no run may end lower than with the six alone, and the curves that missed are re-examined
    spec2 = kernel_spec_from_sklearn(family_kernel(kname))
    spec2.extra_starts = 6
    res2 = fulu_b200.fit_augment_batch(hb, n_obs=32, engine=engine, spec=spec2)
    assert np.all(res2.lml >= res.lml - 1e-9)
    deficits2 = deficits - (res2.lml - res.lml)
    print(kname, "with 6 extra starts:", np.array2string(deficits2, precision=2))
    assert np.count_nonzero(deficits2 > 1e-6) <= np.count_nonzero(deficits > 1e-6)


@pytest.mark.parametrize("kname,n", [("rbf_matern", 400), ("notebook", 330), ("rbf_rq", 520)])
def test_families_on_the_tiled_long_curve_path(engine, kname, n):
    """The non-default kernels where the matrix is streamed from a global slab (N >= 304, gp_tiled.cuh): K and dK/dtheta are
    assembled / recomputed inside the streamed products' epilogues there.  LML, gradient, mean and std at fixed hyperparameters
    against the closed-form oracle, and against live sklearn with the same kernel object."""
    from sklearn.gaussian_process import GaussianProcessRegressor

    from fulu_b200 import DeviceBatch, datasets
    from fulu_b200.engine import kernel_spec_from_sklearn
    from oracle import gp_oracle as O

    spec = kernel_spec_from_sklearn(family_kernel(kname))
    hb = datasets.ddf_like(2, n_points=n, first=31)
    db = DeviceBatch.from_host(hb, engine.device, family=spec.family)
    assert db.n_max >= 304
    rng = np.random.default_rng(n)
    th = rng.uniform(-1.0, 1.0, (2, spec.n_params))
    th[:, -1] = -3.0
    out = engine.lml(db, torch.tensor(th))
    pred = engine.predict_grid(db, torch.tensor(th), 16)
    p2l = {i: float(v) for i, v in enumerate(hb.loglam)}
    for i in range(2):
        t, f, e, b = hb.curve(i)
        gp = O.ClosedFormGP(p2l, kind=spec.family).prepare(t, f, e, b)
        l, g, gs = O.lml_and_grad(gp.X, gp.y, gp.alpha, th[i], spec.family, return_scale=True)
        assert abs(float(out["lml"][i]) - l) <= 1e-9 * max(1.0, abs(l)), (kname, i)
        assert np.max(np.abs(out["grad"][i].cpu().numpy() - g) / np.maximum(np.maximum(np.abs(g), 1e-1 * gs), 1e-30)) < 1e-9, (kname, i)
        X = (np.c_[t, hb.loglam[b]] - gp.x_mean) / gp.x_scale
        reg = GaussianProcessRegressor(kernel=family_kernel(kname).clone_with_theta(spec.to_sklearn(th[i])), optimizer=None, normalize_y=True).fit(X, f)
        assert abs(reg.log_marginal_likelihood_value_ - float(out["lml"][i])) <= 1e-9 * max(1.0, abs(l))
        tq = np.tile(np.linspace(t.min(), t.max(), 16), hb.n_bands)
        fr, er = reg.predict((np.c_[tq, np.repeat(hb.loglam, 16)] - gp.x_mean) / gp.x_scale, return_std=True)
        assert relerr(pred["mean"][i].cpu().numpy().reshape(-1), fr, floor=1e-3 * gp.y_std) < 1e-9, (kname, i)
        assert relerr(pred["std"][i].cpu().numpy().reshape(-1), er, floor=1e-6 * gp.y_std) < 1e-9, (kname, i)


def test_unsupported_family_and_parameter_count_are_refused(engine):
    from fulu_b200 import CurveBatch, DeviceBatch, _capi
    from fulu_b200.engine import family_code

    g = load_golden("readme")
    hb = CurveBatch.from_curves([(g["t"], g["flux"], g["flux_err"], g["band"])], g["loglam"])
    db = DeviceBatch.from_host(hb, engine.device, family=family_code(3, 4))   # (Matern 5/2, RQ): not compiled
    with pytest.raises(_capi.FuluGpError, match="not supported"):
        engine.lml(db, torch.zeros(1, 7, dtype=torch.float64))
    db = DeviceBatch.from_host(hb, engine.device, family=family_code(0, 2))
    with pytest.raises(_capi.FuluGpError, match="bad argument"):
        engine.lml(db, torch.zeros(1, 4, dtype=torch.float64))   # this family has 5 hyperparameters
