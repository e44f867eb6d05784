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


def lpt_partition(lengths, world, n_eval=135, m=768, exact_below=1024):
    """Longest-processing-time partition of ragged curves over `world` GPUs; returns one sorted index array per rank.

    Curves in decreasing cost order; the greedy rule (next curve to the least loaded rank) is applied one by one to the last
    `exact_below` curves only - the cheap ones, which level the loads - while the bulk before them is dealt in serpentine order
    (0..w-1, w-1..0, ...), which is what the greedy rule does on slowly varying costs and is a single vectorised pass: a million
    curves partition in well under a second (the bench times this inside the end-to-end call)."""
    cost = curve_cost(lengths, n_eval, m)
    n = len(cost)
    order = np.argsort(-cost, kind="stable")
    n_bulk = max(0, n - exact_below)
    n_bulk -= n_bulk % (2 * world)
    owner = np.empty(n, dtype=np.int64)
    load = np.zeros(world)
    if n_bulk:
        pos = np.arange(n_bulk) % (2 * world)
        r = np.where(pos < world, pos, 2 * world - 1 - pos)
        owner[order[:n_bulk]] = r
        load += np.bincount(r, weights=cost[order[:n_bulk]], minlength=world)
    for i in order[n_bulk:]:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += cost[i]
    return [np.nonzero(owner == r)[0].astype(np.int64) for r in range(world)]


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


def gather_fields(local, parts, n_total, dst=0):
    """One gather of per-curve result tensors at the end of a sharded run, straight from where they lie (device tensors over
    NCCL, host tensors over gloo): rank r sends its rows as they are - no padding, no host bounce - into rank r's slice of one
    buffer per field on `dst`, and one indexed copy there puts the rows into the original curve order.  `parts[r]` = the global curve
    ids of rank r's rows, known on every rank (the partition is a pure function of the lengths), so no index travels.
    local: dict name -> tensor [len(parts[rank]), ...].  Returns dict name -> tensor [n_total, ...] on `dst`, None elsewhere."""
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    counts = [len(p) for p in parts]
    offs = np.concatenate([[0], np.cumsum(counts)])
    perm = torch.as_tensor(np.concatenate(parts).astype(np.int64)) if n_total else torch.zeros(0, dtype=torch.int64)
    out = {}
    if world == 1:
        for k, v in local.items():
            o = torch.empty((n_total,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
            o[perm.to(v.device)] = v
            out[k] = o
        return out
    ops, bufs = [], {}
    for k in sorted(local):   # same order on every rank
        v = local[k].contiguous()
        if rank == dst:
            buf = torch.empty((n_total,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
            bufs[k] = buf
            buf[offs[dst]:offs[dst + 1]] = v
            for r in range(world):
                if r != dst and counts[r]:
                    ops.append(dist.P2POp(dist.irecv, buf[offs[r]:offs[r + 1]], r))
        elif counts[rank]:
            ops.append(dist.P2POp(dist.isend, v, dst))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if rank != dst:
        return None
    for k, buf in bufs.items():
        o = torch.empty_like(buf)
        o[perm.to(buf.device)] = buf
        out[k] = o
    return out


def fit_augment_sharded(curves, n_obs=128, dst=0, fit_fn=None, t_min=None, t_max=None, use_err=False, spec=None, engine=None, device=None,
                        window=0, reduce=None, out=None):
    """GP fit + augmentation of a ragged batch over all ranks of the default process group (one process per GPU).

    Every rank calls this with the same host batch: the curves are split by the LPT partition on the cubic cost model
    (`lpt_partition`), every rank runs the batched path on its own curves on its own GPU - no collective on the data path - and the
    results are gathered ONCE at the end onto rank `dst` (`gather_fields`: device tensors over NCCL), in the original curve order,
    and copied to the host there.  Per-curve arguments (`t_min`, `t_max`, length n_curves) are restricted to each rank's curves.
    `out` (on `dst`): pinned result buffers (engine.pin_results for the whole batch) - the returned arrays are then views of them.
    Returns an AugmentResult on `dst` (with `lpt_imbalance` = largest / mean modelled load), None elsewhere.
    `fit_fn(batch, n_obs=...)` (tests) replaces the GPU work by a host function returning an AugmentResult."""
    import torch.distributed as dist

    from .engine import AugmentResult, DeviceBatch, GPEngine, KernelSpec, fit_augment_device

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    lengths = curves.lengths()
    m = curves.n_bands * n_obs
    parts = lpt_partition(lengths, world, m=m)
    mine = parts[rank]
    loads = np.array([curve_cost(lengths[p], m=m).sum() for p in parts])
    sub = curves.take(mine)
    tmn = None if t_min is None else np.asarray(t_min, np.float64).reshape(-1)[mine]
    tmx = None if t_max is None else np.asarray(t_max, np.float64).reshape(-1)[mine]
    if reduce not in (None, "peak"):
        raise ValueError("reduce must be None or 'peak'")
    perm_theta = None
    if fit_fn is not None:
        res = fit_fn(sub, n_obs=n_obs)
        local = {k: torch.as_tensor(getattr(res, k)) for k in ("flux_aug", "flux_err_aug", "theta", "lml", "n_eval", "status")}
        h2d = 0
    else:
        engine = engine or GPEngine(device)
        spec = spec or KernelSpec()
        nb, P = curves.n_bands, spec.n_params
        with torch.cuda.device(engine.device):
            if sub.n_curves:
                db = DeviceBatch.from_host(sub, engine.device, use_err=use_err, family=spec.family)
                dt0 = None if tmn is None else torch.as_tensor(tmn).to(engine.device)
                dt1 = None if tmx is None else torch.as_tensor(tmx).to(engine.device)
                dev = fit_augment_device(engine, db, sub.lengths(), n_obs, dt0, dt1, spec, window)
                h2d = db.h2d_bytes()
                if reduce == "peak":
                    dev.update(engine.peak(db, dev["mean"], dt0, dt1))
            else:   # a rank without curves still takes part in the gather
                new = engine._new
                dev = dict(mean=new(0, nb, n_obs), std=new(0, nb, n_obs), theta=new(0, P), lml=new(0), n_eval=new(0, dtype=torch.int32),
                           status=new(0, dtype=torch.int32))
                if reduce == "peak":
                    dev.update(peak_t=new(0), peak_flux=new(0), peak_index=new(0, dtype=torch.int32))
                h2d = 0
        local = dict(theta=dev["theta"], lml=dev["lml"], n_eval=dev["n_eval"], status=dev["status"])
        if reduce == "peak":
            local.update(peak_t=dev["peak_t"], peak_flux=dev["peak_flux"], peak_index=dev["peak_index"])
        else:
            local.update(flux_aug=dev["mean"], flux_err_aug=dev["std"])
        if not np.array_equal(spec.perm, np.arange(spec.n_params)):
            perm_theta = spec
    full = gather_fields(local, parts, curves.n_curves, dst=dst)
    if rank != dst:
        return None
    names = {"flux_aug": "mean", "flux_err_aug": "std"}
    if out is not None and all(names.get(k, k) in out for k in full):
        host = {}
        for k, v in full.items():
            dstbuf = out[names.get(k, k)][: curves.n_curves]
            dstbuf.copy_(v, non_blocking=True)
            host[k] = dstbuf
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        host = {k: v.numpy() for k, v in host.items()}
    else:
        host = {k: v.cpu().numpy() for k, v in full.items()}
    theta = host["theta"] if perm_theta is None else perm_theta.to_sklearn(host["theta"])
    out = AugmentResult(host.get("flux_aug"), host.get("flux_err_aug"), theta, host["lml"], host["n_eval"], host["status"])
    if reduce == "peak":
        out.peak_t, out.peak_flux, out.peak_index = host["peak_t"], host["peak_flux"], host["peak_index"]
    out.h2d_bytes = h2d
    out.d2h_bytes = sum(int(v.numel() * v.element_size()) for v in full.values())
    out.lpt_imbalance = float(loads.max() / max(loads.mean(), 1e-300))
    return out
