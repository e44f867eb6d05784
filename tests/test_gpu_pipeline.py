"""The steps either side of the path, on the device: ingest of an object-grouped survey table into the ragged batch,
and the band-sum / peak consumer of the augmented curves (LcPlotter.plot_sum_passbands / plot_peak, fulu/plotting.py:78-115)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import fulu_b200

    return fulu_b200.GPEngine()


def _table(hb, ids):
    """Flatten a CurveBatch into table columns (object id per row)."""
    oid = np.repeat(np.asarray(ids, np.int64), hb.lengths())
    return oid, hb.t, hb.flux, hb.flux_err, hb.band


def test_ingest_matches_numpy_groupby(engine):
    from fulu_b200 import datasets
    from oracle import gp_oracle as O

    hb = datasets.ztf_like(3000, first=100)     # ragged 20..300 points: chunk boundaries (1024 rows) fall everywhere
    rng = np.random.default_rng(0)
    ids = rng.permutation(10_000_000)[: hb.n_curves]          # arbitrary, unsorted object ids
    oid, t, f, e, b = _table(hb, ids)
    p2l = {1: float(hb.loglam[0]), 2: float(hb.loglam[1])}    # passband ids 1, 2 -> band index 0, 1
    batch, lengths, got_ids = engine.ingest_table(oid, t, f, e, b + 1, p2l)
    off_ref, ids_ref = O.table_to_csr(oid)
    np.testing.assert_array_equal(batch.offsets.cpu().numpy(), off_ref)
    np.testing.assert_array_equal(got_ids, ids_ref)
    assert lengths is None       # not read back: the length classes are built on the device (engine.class_orders_device)
    from fulu_b200.engine import bucket_curves, class_orders_device

    dev_classes = class_orders_device(batch.offsets)
    ref_classes = bucket_curves(hb.lengths())
    assert [(b_, c_) for b_, c_, _ in dev_classes] == [(b_, len(i_)) for b_, i_ in ref_classes]
    for (_, _, order), (_, idx) in zip(dev_classes, ref_classes):
        o = order.cpu().numpy()
        assert sorted(o.tolist()) == sorted(idx.tolist()) and np.all(np.diff(hb.lengths()[o]) <= 0)    # same curves, longest first
    np.testing.assert_array_equal(batch.band.cpu().numpy(), hb.band)
    assert batch.n_curves == hb.n_curves and batch.n_max == int(hb.lengths().max())
    # sizes around the chunk size, a single object, a single row, an empty table
    for n_rows, n_obj in ((1, 1), (1023, 1), (1024, 1024), (1025, 3), (4096, 4), (0, 0)):
        o = np.sort(rng.integers(0, max(n_obj, 1), n_rows)) * 7 + 3
        if n_rows:
            o[-1] = o.max()
        z = np.zeros(n_rows)
        bt, ln, gi = engine.ingest_table(o, z, z, z, np.ones(n_rows, np.int32), p2l)
        off_ref, ids_ref = O.table_to_csr(o)
        np.testing.assert_array_equal(bt.offsets.cpu().numpy(), off_ref)
        np.testing.assert_array_equal(gi, ids_ref)
    with pytest.raises(KeyError):
        engine.ingest_table(oid[:50], t[:50], f[:50], e[:50], np.full(50, 7, np.int32), p2l)     # the reference: KeyError (_base_aug.py:10)


def test_table_entry_point_equals_batched_entry_point(engine):
    import fulu_b200
    from fulu_b200 import datasets

    hb = datasets.plasticc_like(40, n_points=100, first=300)
    oid, t, f, e, b = _table(hb, 1000 + np.arange(40))
    p2l = {i: float(v) for i, v in enumerate(hb.loglam)}
    res_t, ids = fulu_b200.fit_augment_table(oid, t, f, e, b, p2l, n_obs=32, engine=engine)
    res_b = fulu_b200.fit_augment_batch(hb, n_obs=32, engine=engine)
    np.testing.assert_array_equal(ids, 1000 + np.arange(40))
    np.testing.assert_array_equal(res_t.flux_aug, res_b.flux_aug)      # same kernels, same launch geometry: bit-identical
    np.testing.assert_array_equal(res_t.flux_err_aug, res_b.flux_err_aug)
    np.testing.assert_array_equal(res_t.lml, res_b.lml)


def test_band_sum_and_peak_match_pandas(engine):
    """The consumer against the reference's own pandas calls (plotting.py:114-115) on real augmentation output."""
    import fulu_b200
    from fulu_b200 import DeviceBatch, create_aug_data, datasets
    from oracle import gp_oracle as O

    hb = datasets.plasticc_like(24, n_points=100, first=500)
    n_obs = 64
    res = fulu_b200.fit_augment_batch(hb, n_obs=n_obs, engine=engine)
    db = DeviceBatch.from_host(hb, engine.device)
    out = engine.peak(db, torch.from_numpy(res.flux_aug).to(engine.device), want_sum=True)
    pk = fulu_b200.fit_augment_batch(hb, n_obs=n_obs, engine=engine, reduce="peak")
    assert pk.flux_aug is None and pk.d2h_bytes < 200 * hb.n_curves
    for i in range(hb.n_curves):
        t = hb.curve(i)[0]
        t_aug, pb_aug = create_aug_data(t.min(), t.max(), tuple(range(hb.n_bands)), n_obs)
        tt, ss, peak = O.band_sum_peak(t_aug, res.flux_aug[i].reshape(-1), pb_aug)
        np.testing.assert_array_equal(tt, np.linspace(t.min(), t.max(), n_obs))
        np.testing.assert_allclose(out["sum_flux"][i].cpu().numpy(), ss, rtol=1e-14, atol=1e-12)   # pandas sums with Kahan compensation
        assert float(out["peak_t"][i]) == peak, i
        assert int(out["peak_index"][i]) == int(np.argmax(ss))
        assert pk.peak_t[i] == peak and pk.peak_index[i] == int(np.argmax(ss))
        np.testing.assert_allclose(pk.peak_flux[i], ss.max(), rtol=1e-14)
    # explicit grids, ties (first maximum wins) and NaN curves
    mean = torch.zeros(3, 2, 40, dtype=torch.float64, device=engine.device)
    mean[0, 0, 7] = 1.0
    mean[0, 1, 30] = 1.0                      # two equal maxima of the sum: index 7 wins
    mean[1] = float("nan")
    mean[2, :, 39] = 2.0
    hb3 = datasets.ztf_like(3, first=1)
    db3 = DeviceBatch.from_host(hb3, engine.device)
    o = engine.peak(db3, mean, t_min=np.array([0.0, 0.0, 10.0]), t_max=np.array([39.0, 1.0, 20.0]))
    assert o["peak_index"].cpu().tolist() == [7, -1, 39]
    assert float(o["peak_t"][0]) == 7.0 and np.isnan(float(o["peak_t"][1])) and float(o["peak_t"][2]) == 20.0
    assert float(o["peak_flux"][2]) == 4.0


def test_csv_table_through_ingest(engine):
    """notebook_examples/object_34299.csv as a one-object table (columns mjd, flux, flux_err, passband): same result as the class."""
    import fulu_b200

    g = load_golden("csv34299")
    p2l = {i: float(v) for i, v in enumerate(g["loglam"])}
    res, ids = fulu_b200.fit_augment_table(np.full(len(g["t"]), 34299), g["t"], g["flux"], g["flux_err"], g["band"], p2l, n_obs=int(g["n_obs"]),
                                           engine=engine)
    assert ids.tolist() == [34299]
    assert res.lml[0] >= float(g["lml_star"]) - 1e-6
