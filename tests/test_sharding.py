"""Host logic of the N>1 path on CPU: world_size-2 gloo processes shard a ragged batch, "process" their curves, and the results are
gathered once at the end in the original order.  (The GPU work itself is per-curve and needs no exchange.)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fulu_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, lengths, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = sharding.lpt_partition(lengths, world)
    mine = parts[rank]
    # stand-in for the per-curve GPU work: a deterministic function of the curve id and its length
    rows = torch.tensor(np.stack([np.array([i, lengths[i], i * 0.5]) for i in mine])) if len(mine) else torch.zeros((0, 3), dtype=torch.float64)
    full = sharding.gather_rows(rows, mine, len(lengths), dst=0)
    slow = sharding.max_over_ranks(10.0 + rank)
    if rank == 0:
        np.save(out_path, np.concatenate([full.numpy().reshape(-1), [slow]]))
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    rng = np.random.default_rng(0)
    lengths = np.round(np.exp(rng.uniform(np.log(20), np.log(300), 37))).astype(int)
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), lengths, out), nprocs=2, join=True)
    got = np.load(out)
    rows, slow = got[:-1].reshape(37, 3), got[-1]
    np.testing.assert_array_equal(rows[:, 0], np.arange(37))
    np.testing.assert_array_equal(rows[:, 1], lengths)
    assert slow == 11.0


def test_lpt_partition_balances_cubic_cost():
    rng = np.random.default_rng(1)
    lengths = np.round(np.exp(rng.uniform(np.log(20), np.log(300), 4000))).astype(int)
    for world in (2, 4, 8):
        parts = sharding.lpt_partition(lengths, world)
        allidx = np.sort(np.concatenate(parts))
        np.testing.assert_array_equal(allidx, np.arange(len(lengths)))   # a partition: every curve exactly once
        loads = np.array([sharding.curve_cost(lengths[p]).sum() for p in parts])
        assert loads.max() / loads.mean() < 1.01


def test_contiguous_shard_covers_everything():
    for n, world in ((10, 3), (8, 8), (5, 8), (1_000_000, 8)):
        spans = [sharding.contiguous_shard(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _fake_fit(batch, n_obs=128, **kw):
    """Stand-in for fit_augment_batch on a machine without a GPU: deterministic per-curve outputs."""
    from fulu_b200.engine import AugmentResult

    B, nb = batch.n_curves, batch.n_bands
    key = np.array([batch.curve(i)[0].sum() for i in range(B)])          # a function of the curve's data only
    mean = key[:, None, None] + np.arange(nb)[None, :, None] + 0.001 * np.arange(n_obs)[None, None, :]
    return AugmentResult(mean, 2 * mean, np.stack([key, -key, key * 2, key * 3], 1), key * 7, batch.lengths().astype(np.int32),
                         np.zeros(B, np.int32))


def _sharded_worker(rank, world, port, out_path):
    from fulu_b200 import datasets

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hb = datasets.ztf_like(23, first=0)
    res = sharding.fit_augment_sharded(hb, n_obs=4, fit_fn=_fake_fit)
    if rank == 0:
        ref = _fake_fit(hb, n_obs=4)
        ok = all(np.array_equal(getattr(res, k), getattr(ref, k)) for k in ("flux_aug", "flux_err_aug", "theta", "lml", "n_eval", "status"))
        np.save(out_path, np.array([float(ok)]))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_entry_point_reassembles_the_batch(tmp_path):
    """fit_augment_sharded on two gloo ranks: LPT split, per-rank work, one gather - same rows, original order."""
    out = str(tmp_path / "ok.npy")
    mp.spawn(_sharded_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert np.load(out)[0] == 1.0
