"""Multi-GPU plumbing: curves are independent GPs, so the batch is sharded by curve, one process per GPU, and nothing is
exchanged on the data path.  torch.distributed is used only for the rendezvous, the max-over-ranks timing, and (optionally)
one gather of the finished results at the end."""
from __future__ import annotations

import numpy as np
import torch


def curve_cost(lengths, n_eval=135, m=768):
    """Flop model of SURVEY 8d: E*(N^3 + 12 N^2) + N^3/3 + 2 N^2 + N^2 M + 14 N M."""
    n = np.asarray(lengths, np.float64)
    return n_eval * (n**3 + 12 * n**2) + n**3 / 3 + 2 * n**2 + n**2 * m + 14 * n * m


def contiguous_shard(n_items, rank, world):
    """[lo, hi) of an even contiguous split (weak-scaling benchmarks: every rank generates its own slice)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def lpt_partition(lengths, world, n_eval=135, m=768):
    """Greedy longest-processing-time partition of ragged curves over `world` GPUs; returns one sorted index array per rank."""
    cost = curve_cost(lengths, n_eval, m)
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world)
    parts = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        parts[r].append(int(i))
        load[r] += cost[i]
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def max_over_ranks(value, device=None):
    """Max of a per-rank scalar (timings are always reported as the slowest rank)."""
    import torch.distributed as dist

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(local_rows, local_index, n_total, dst=0, device=None):
    """Gather per-curve result rows ([n_local, ...] tensor, global indices [n_local]) onto rank `dst` in original curve order.

    One variable-size gather at the end of the run; returns the [n_total, ...] tensor on `dst`, None elsewhere.
    """
    import torch.distributed as dist

    local_rows = torch.as_tensor(local_rows)
    local_index = torch.as_tensor(local_index, dtype=torch.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = torch.empty((n_total,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype)
        out[local_index] = local_rows.cpu()
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else local_rows.device
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=dev))
    nmax = int(max(int(c.item()) for c in counts))
    tail = tuple(local_rows.shape[1:])
    pad_rows = torch.zeros((nmax,) + tail, dtype=local_rows.dtype, device=dev)
    pad_idx = torch.full((nmax,), -1, dtype=torch.int64, device=dev)
    pad_rows[: local_rows.shape[0]] = local_rows.to(dev)
    pad_idx[: local_index.shape[0]] = local_index.to(dev)
    rows = [torch.empty_like(pad_rows) for _ in range(world)] if rank == dst else None
    idxs = [torch.empty_like(pad_idx) for _ in range(world)] if rank == dst else None
    dist.gather(pad_rows, rows, dst=dst)
    dist.gather(pad_idx, idxs, dst=dst)
    if rank != dst:
        return None
    out = torch.empty((n_total,) + tail, dtype=local_rows.dtype)
    for r in range(world):
        k = int(counts[r].item())
        out[idxs[r][:k].cpu()] = rows[r][:k].cpu()
    return out


def fit_augment_sharded(curves, n_obs=128, dst=0, fit_fn=None, **kw):
    """GP fit + augmentation of a ragged batch over all ranks of the default process group (one process per GPU).

    Every rank calls this with the same host batch (or a batch whose curves it can index): the curves are split by the LPT
    partition on the cubic cost model (`lpt_partition`), every rank runs `fit_augment_batch` on its own curves on its own GPU - no
    collective on the data path - and the results are gathered ONCE at the end onto rank `dst`, in the original curve order.
    Returns an AugmentResult on `dst`, None elsewhere.  `fit_fn` (tests) replaces `fit_augment_batch`.
    """
    import torch.distributed as dist

    from .engine import AugmentResult, fit_augment_batch

    fit_fn = fit_fn or fit_augment_batch
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    lengths = curves.lengths()
    mine = lpt_partition(lengths, world, m=curves.n_bands * n_obs)[rank]
    res = fit_fn(curves.take(mine), n_obs=n_obs, **kw)
    backend = dist.get_backend() if world > 1 else None
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    fields = {}
    for name in ("flux_aug", "flux_err_aug", "theta", "lml", "n_eval", "status"):
        fields[name] = gather_rows(torch.as_tensor(getattr(res, name)), mine, curves.n_curves, dst=dst, device=dev)
    if rank != dst:
        return None
    out = {k: v.numpy() for k, v in fields.items()}
    return AugmentResult(out["flux_aug"], out["flux_err_aug"], out["theta"], out["lml"], out["n_eval"], out["status"])
