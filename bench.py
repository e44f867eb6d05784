#!/usr/bin/env python
"""Benchmark of the GP fit + augmentation hot path (BASELINE.json metric: light curves/sec, 6 bands, N=100).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--curves B] [--impl reference]

One "step" = one pass of the hot path (6 x L-BFGS-B fit + 128-point x 6-band augmentation) over one batch of
synthetic PLAsTiCC-like curves (config #3 of BASELINE.json: B=10,000 per GPU, N=100, M=768).  With N>1 (torchrun, one rank per
GPU) every rank owns its own 10,000 curves (weak scaling); there is no collective on the data path.  Rank 0 prints ONE JSON line.

  value         curves/s with the batch already resident in HBM (CUDA events around K steps, max over ranks)
  e2e           the same through the public API fit_augment_batch(): pinned host arrays in, host arrays out, H2D+D2H timed
  roofline      the evaluation kernel (fit_eval_kernel: assemble K, Cholesky, inverse, LML+gradient) against the FP64 pipe:
                algorithmic flops N^3+12N^2 per evaluation / CUDA-event time of its launches (library-side events on
                the launching stream); peak = DFMA rate measured in this run by fulu_gp_fp64_peak (MEASURED_PEAKS.json has no FP64)
  cpu_baseline  the reference path on all host cores, bounded sample: the reference's own fulu/gp_aug.py where __graft_entry__.build() staged
                it (oracle/_ref/, kind "reference"), else the same calls re-typed (oracle/fulu_sklearn_port.py, kind "port")
  parity        curves with a status other than "theta on a bound"; vs_cpu_pass: the CPU sample is the first curves of the GPU batch, so
                the final LMLs are compared curve by curve (max deficit against sklearn, fraction within 1e-6) together with both
                evaluation counts; roofline.frac_at_sklearn_evals rescales the roofline fraction to sklearn's count
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS, N_OBS, N_BANDS = 100, 128, 6
METRIC = "light curves/sec GP fit+augment (6 bands, N=100)"


# kernels of the --kernel option: the reference's default (fulu/gp_aug.py:29) and the others its users pass
def make_kernel(name):
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel as C, Matern, RationalQuadratic, WhiteKernel

    return {
        "default": lambda: None,
        "rbf_matern": lambda: C(1.0) * RBF([1, 1]) + Matern() + WhiteKernel(),                  # fulu/gp_aug.py:24, tests/test_aug.py:37
        "notebook": lambda: C(1.0) * Matern() * RBF([1, 1]) + Matern() + WhiteKernel(),         # notebook_examples/plotting.ipynb:33
        "rbf_rq": lambda: C(1.0) * RBF([1, 1]) + RationalQuadratic() + WhiteKernel(),           # fulu/gp_aug.py:5
    }[name]()


KERNEL = "default"


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_sample_curves(workload, first, count):
    """The CPU arm's curves: the SAME curves as the head of rank 0's GPU batch of that workload (make_workload), so final LMLs compare
    curve by curve.  ddf: the N = 1024 part only (a fit at N = 2048 is ~10 minutes of one host core)."""
    from fulu_b200 import datasets

    if workload == "plasticc":
        return datasets.plasticc_like(count, n_points=N_POINTS, first=first)
    if workload == "plasticc_ragged":
        return datasets.plasticc_like(count, n_points=N_POINTS, first=first, ragged=True)
    if workload == "ztf":
        return datasets.ztf_like(count, first=first)
    return datasets.ddf_like(count, n_points=1024, first=first)


def _cpu_worker(args):
    workload, first, count, n_obs, kernel = args
    import warnings

    from oracle import ref_arm
    from oracle.fulu_sklearn_port import SklearnGPAugmentation, count_evals

    ref_cls = ref_arm.load()   # the reference's own class where its files were staged (oracle/_ref/), else the re-typed port
    hb = cpu_sample_curves(workload, first, count)
    p2l = {i: float(v) for i, v in enumerate(hb.loglam)}
    out = []
    for i in range(hb.n_curves):
        t, f, e, b = hb.curve(i)
        k = make_kernel(kernel)
        if ref_cls is not None:
            aug = ref_cls(p2l) if k is None else ref_cls(p2l, kernel=k)
        else:
            aug = SklearnGPAugmentation(p2l, kernel=k)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            n_eval = count_evals(aug, t, f, e, b)   # aug.fit(...) with the LML+gradient evaluations counted (E of SURVEY 8d)
            aug.augmentation(t.min(), t.max(), n_obs)
        out.append((float(aug.reg.log_marginal_likelihood_value_), n_eval))
    return out


def cpu_reference_throughput(n_curves, cores=None, first=0, chunk=4, workload="plasticc"):
    """curves/s of the reference CPU path (sklearn GaussianProcessRegressor driven exactly as fulu drives it) on `cores` processes."""
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor

    cores = cores or os.cpu_count() or 1
    for k in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"  # 8 procs x 1 BLAS thread beat 1 x 8 by 6x
    if workload == "ddf":
        chunk = 1
    tasks = [(workload, first + i, min(chunk, n_curves - i), N_OBS, KERNEL) for i in range(0, n_curves, chunk)]
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn")) as ex:
        list(ex.map(_cpu_worker, [("plasticc", 10_000_000 + w, 1, N_OBS, KERNEL) for w in range(cores)]))  # warm-up: imports + first fit
        t0 = time.perf_counter()
        lml = [v for part in ex.map(_cpu_worker, tasks) for v in part]
        dt = time.perf_counter() - t0
    return n_curves / dt, dt, cores, lml   # lml: [(final LML, evaluations)] per curve, in curve order


def cpu_arm_kind():
    """("reference", text) when the CPU arm runs the reference's own files (staged under oracle/_ref/ by __graft_entry__.build() in
    the build container), ("port", text) when it runs the re-typed adapter - decided the way the worker processes decide."""
    from oracle import ref_arm

    if ref_arm.staged():
        return "reference", ("the reference's own fulu/gp_aug.py + fulu/_base_aug.py (unmodified files staged under oracle/_ref/, imported "
                             "through matplotlib stubs) on the box's scikit-learn")
    return "port", "fulu adapter on scikit-learn (oracle/fulu_sklearn_port.py)"


def cpu_sample_size(workload, cores, ref_curves):
    """Bounded CPU samples: >= 256 curves of config #3 by default, a 2,000-curve sample of config #4, one N = 1024 curve
    per host core of config #5."""
    if workload == "ztf":
        return max(2000, 2 * cores)
    if workload == "ddf":
        return cores
    return max(2 * cores, min(ref_curves, 24 * cores))


CPU_SAMPLE_TEXT = {
    "plasticc": "curves of config #3 (N=100, 6 bands, M=768)",
    "plasticc_ragged": "curves of config #3's ragged variant (N = clip(Poisson(100), 64, 160), 6 bands, M=768)",
    "ztf": "curves of config #4 (ZTF-like, 2 bands, ragged N = 20..300, M=256)",
    "ddf": "curves of config #5 at N = 1024 (6 bands, M=768; the N = 1536 / 2048 curves of the GPU batch are not in the CPU sample)",
}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = cpu_sample_size(args.workload, cores, args.ref_curves) if args.workload != "ztf" else max(2 * cores, min(args.ref_curves, 24 * cores))
    vals, times = [], []
    for s in range(args.warmup + args.steps):
        v, dt, cores, _ = cpu_reference_throughput(per_step, cores, first=s * per_step, workload=args.workload)
        if s >= args.warmup:
            vals.append(v)
            times.append(dt)
    value = per_step * len(times) / sum(times)
    kind, kind_text = cpu_arm_kind()
    sample = f"{per_step} {CPU_SAMPLE_TEXT[args.workload]} per step, {cores} processes x 1 BLAS thread, {kind_text}"
    headline = args.workload == "plasticc" and args.kernel == "default"
    metric = METRIC if headline else f"light curves/sec GP fit+augment ({args.workload} workload, {args.kernel} kernel)"
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": "curves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][0], "curves_per_step": per_step, "n_bands": 2 if args.workload == "ztf" else N_BANDS,
                   "n_obs": N_OBS, "kernel": args.kernel},
        "cpu_baseline": {"value": value, "unit": "curves/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "curves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  In-process NVML on a thread (the GIL is released while the
    main thread is inside the library), because polling with an `nvidia-smi -lms` subprocess visibly slowed the measured step;
    falls back to nvidia-smi when NVML cannot be loaded."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_s=0.1):
        self.gpu, self.period, self.proc, self.path = gpu_index, period_s, None, f"/tmp/fulu_clocks_{os.getpid()}.csv"
        self.thread, self.stop_flag, self.sm, self.mx, self.reasons, self.how = None, False, [], [], set(), None
        self.active = False   # the thread exists from prepare() on (NVML initialised outside the timed region) and samples only while active

    def _nvml_loop(self, nv, h):
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        while not self.stop_flag:
            if not self.active:
                time.sleep(0.005)
                continue
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def prepare(self):
        """NVML initialisation and thread start, to be called well before the timed region: nvmlInit takes tens of milliseconds and
        driver locks, and doing it at the start of the timed region stalled the launching thread on some runs."""
        self.start(active=False)

    def start(self, active=True):
        if self.thread is not None or self.proc is not None:
            self.active = active
            return
        self.active = active
        try:
            import threading

            import pynvml as nv

            nv.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = int(vis.split(",")[self.gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.how = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.how = "nvidia-smi"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "how": "nvml thread, 100 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in open(self.path):
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "how": "nvidia-smi -lms 500"}


WORKLOADS = {
    # name: (description, default curves per GPU per step, n_obs)
    "plasticc": ("BASELINE config #3 (PLAsTiCC-like, 6 bands, N=100, 128-point grid x 6 bands), one batch per GPU per step", 10000, N_OBS),
    "plasticc_ragged": ("BASELINE config #3, the survey's ragged variant (PLAsTiCC-like, 6 bands, N = clip(Poisson(100), 64, 160), 128-point grid "
                        "x 6 bands): three length classes (N <= 103 / 127 / 159) as index-list launches over the resident batch", 10000, N_OBS),
    "ztf": ("BASELINE config #4 sample (ZTF-like, 2 bands g/r, ragged N = round(exp(U(ln 20, ln 300))), 128-point grid x 2 bands), "
            "one batch per GPU per step, length classes run as index-list launches over the resident batch", 20000, N_OBS),
    "ddf": ("BASELINE config #5 (LSST-DDF-like, 6 bands, N in {1024, 1536, 2048} in equal parts, 128-point grid x 6 bands), "
            "one CTA per (curve, start) run, the matrix in a global slab streamed through shared-memory stages by bulk copies (gp_tiled.cuh); "
            "the whole fit of the class is one fused launch fed by a task queue", 96, N_OBS),
}


def make_workload(name, B, rank):
    from fulu_b200 import CurveBatch, datasets

    if name == "plasticc":
        return datasets.plasticc_like(B, n_points=N_POINTS, first=rank * B)   # each rank: its own B curves (weak scaling)
    if name == "plasticc_ragged":
        return datasets.plasticc_like(B, n_points=N_POINTS, first=rank * B, ragged=True)
    if name == "ztf":
        return datasets.ztf_like(B, first=rank * B)
    parts = [datasets.ddf_like((B + 2 - k) // 3, n_points=n, first=rank * B + k * B) for k, n in enumerate((1024, 1536, 2048))]
    curves = [p.curve(i) for p in parts for i in range(p.n_curves)]
    return CurveBatch.from_curves(curves, parts[0].loglam)


def library_gpu_baseline(engine, host, spec, n_curves=4096, reps=3):
    """Secondary yardstick: the same LML + gradient evaluation on the same curves through the GPU LIBRARIES - padded-batch
    torch.linalg.cholesky / cholesky_solve / cholesky_inverse (cuSOLVER / cuBLAS batched kernels) and elementwise torch ops for the
    kernel matrix and the traces - timed with CUDA events against this repo's evaluation (fulu_gp_lml_batch) on the same batch.
    Not part of the product path; it only says what the hand-written kernel buys over the library route."""
    import numpy as np
    import torch

    from fulu_b200.engine import DeviceBatch

    B = min(n_curves, host.n_curves)
    sub = host.slice(0, B)
    L = sub.lengths()
    if L.min() != L.max():
        return None
    N = int(L[0])
    dev = engine.device
    db = DeviceBatch.from_host(sub, dev, family=spec.family)
    theta_h = np.tile(np.array([0.3, -0.7, 1.1, -3.0]), (B, 1))
    theta = torch.tensor(theta_h, device=dev)
    ours = engine.lml(db, theta)
    # library route: standardise like the prep kernel (StandardScaler / y normalisation), then dense batched algebra
    t = torch.tensor(sub.t.reshape(B, N), device=dev)
    lam = torch.tensor(sub.loglam, device=dev)[torch.tensor(sub.band.reshape(B, N).astype(np.int64), device=dev)]
    y = torch.tensor(sub.flux.reshape(B, N), device=dev)
    xs = (t - t.mean(1, keepdim=True)) / t.std(1, unbiased=False, keepdim=True)
    xl = (lam - lam.mean(1, keepdim=True)) / lam.std(1, unbiased=False, keepdim=True)
    yn = (y - y.mean(1, keepdim=True)) / y.std(1, unbiased=False, keepdim=True)
    c, lt, ll, s2 = (torch.exp(theta[:, k]) for k in range(4))
    eye = torch.eye(N, dtype=torch.float64, device=dev)

    def evaluate():
        du2 = ((xs[:, :, None] - xs[:, None, :]) / lt[:, None, None]) ** 2
        dv2 = ((xl[:, :, None] - xl[:, None, :]) / ll[:, None, None]) ** 2
        E = c[:, None, None] * torch.exp(-0.5 * (du2 + dv2))
        K = E + (s2[:, None, None] + 1e-10) * eye
        Lc = torch.linalg.cholesky(K)
        alpha = torch.cholesky_solve(yn[:, :, None], Lc)[:, :, 0]
        lml = -0.5 * (yn * alpha).sum(1) - torch.log(torch.diagonal(Lc, dim1=1, dim2=2)).sum(1) - 0.5 * N * math.log(2 * math.pi)
        W = alpha[:, :, None] * alpha[:, None, :] - torch.cholesky_inverse(Lc)
        g = torch.stack([0.5 * (W * E).sum((1, 2)), 0.5 * (W * E * du2).sum((1, 2)), 0.5 * (W * E * dv2).sum((1, 2)),
                         0.5 * s2 * torch.diagonal(W, dim1=1, dim2=2).sum(1)], 1)
        return lml, g

    lml_lib, g_lib = evaluate()
    torch.cuda.synchronize()

    def timed(fn):
        best = float("inf")
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    ms_lib = timed(evaluate)
    ms_ours = timed(lambda: engine.lml(db, theta))
    agree = float(torch.max(torch.abs(lml_lib - ours["lml"]) / torch.clamp(torch.abs(ours["lml"]), min=1.0)))
    return {"what": "LML + gradient evaluations/s at fixed theta on the head of the batch: torch.linalg.cholesky + cholesky_solve + cholesky_inverse "
                    "(cuSOLVER/cuBLAS batched) with elementwise torch ops, against fulu_gp_lml_batch (prep + evaluation kernel) on the same curves",
            "curves": int(B), "n_points": N, "library_evals_per_s": B / (ms_lib * 1e-3), "ours_evals_per_s": B / (ms_ours * 1e-3),
            "ours_over_library": ms_lib / ms_ours, "library_ms": ms_lib, "ours_ms": ms_ours, "max_rel_lml_difference": agree}


def run_gpu_strong(args):
    """Strong scaling of BASELINE config #4: ONE global ragged batch, the same on every rank, LPT-partitioned over the ranks by the cubic
    cost model, fitted per rank with no collective on the data path, and gathered ONCE onto rank 0 - device tensors over NCCL, then one
    D2H copy - all inside the timed call (sharding.fit_augment_sharded, the user-level entry point of a process group).
    value = curves / max-over-ranks time of that call; nothing is resident beforehand except the library and the workspace."""
    import numpy as np
    import torch

    import fulu_b200
    from fulu_b200 import sharding
    from fulu_b200.engine import kernel_spec_from_sklearn

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    engine = fulu_b200.GPEngine(torch.device("cuda", local))
    spec = kernel_spec_from_sklearn(make_kernel(args.kernel))
    B = args.curves or 160000
    desc, _, n_obs = WORKLOADS[args.workload]
    host = make_workload(args.workload, B, 0)      # the same global batch on every rank
    lengths = host.lengths()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.prepare()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    from fulu_b200.engine import pin_results

    pinned_out = pin_results(B, host.n_bands, n_obs, spec.n_params) if rank == 0 else None   # reused every step, like a production loop

    def step():
        return sharding.fit_augment_sharded(host, n_obs=n_obs, engine=engine, spec=spec, out=pinned_out)

    for _ in range(max(args.warmup, 1)):
        step()
    barrier()
    launches0 = engine.launches
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        res = step()
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([max(e0.elapsed_time(e1) * 1e-3, wall)], dtype=torch.float64, device=engine.device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        secs = float(t.item())
        value = B * args.steps / secs
        f_eval = lengths.astype(np.float64) ** 3 + 12.0 * lengths.astype(np.float64) ** 2
        flops = float(np.dot(res.n_eval.astype(np.float64), f_eval))
        line = {
            "metric": f"light curves/sec GP fit+augment ({args.workload} workload, {args.kernel} kernel)", "value": value, "unit": "curves/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc.replace("one batch per GPU per step", "ONE global batch per step, LPT-sharded over the ranks"),
                       "curves_per_step": B, "n_points": f"{int(lengths.min())}..{int(lengths.max())} (mean {lengths.mean():.0f})", "n_bands": host.n_bands,
                       "n_obs": n_obs, "kernel": args.kernel, "l2": f"inputs re-uploaded every step ({res.h2d_bytes / 1e6:.0f} MB on rank 0)",
                       "parallelism": f"curves LPT-partitioned x{world}, no collective on the data path, one gather (NCCL point-to-point of device "
                                      "tensors) onto rank 0 at the end, inside the timed call"},
            "e2e": {"value": value, "unit": "curves/s", "h2d_bytes_per_step": int(res.h2d_bytes), "d2h_bytes_per_step": int(res.d2h_bytes)},
            "gpu_launches": engine.launches - launches0, "clocks": clocks,
            "roofline": {"bound": "fp64", "achieved": flops / (secs / args.steps) / 1e12 / world, "unit": "TFLOP/s per GPU, whole call (algorithmic flops of the evaluations / wall time)"},
            "sharding": {"lpt_imbalance_max_over_mean": res.lpt_imbalance, "gathered_bytes": int(res.d2h_bytes)},
            "cpu_baseline": None,
            "parity": {"status_histogram": res.status_histogram(), "mean_lml": float(np.mean(res.lml[np.isfinite(res.lml)]))},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_gpu(args):
    import numpy as np
    import torch

    import fulu_b200
    from fulu_b200.engine import DeviceBatch, fit_augment_device, kernel_spec_from_sklearn, pin_batch, pin_results

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    desc, default_b, n_obs = WORKLOADS[args.workload]
    headline = args.workload == "plasticc" and args.kernel == "default"
    # CPU baseline first (rank 0, N=1 only), before this process's CUDA context is busy
    cpu_base, cpu_curves = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample_n = cpu_sample_size(args.workload, cores, args.ref_curves)
        v, dt, cores, cpu_curves = cpu_reference_throughput(sample_n, cores, workload=args.workload)
        kind, kind_text = cpu_arm_kind()
        cpu_base = {"value": v, "unit": "curves/s", "cores": cores, "kind": kind,
                    "sample": f"{sample_n} {CPU_SAMPLE_TEXT[args.workload]} in {dt:.1f} s, {cores} processes x 1 BLAS thread, "
                              f"{kind_text}; the same curves as the head of the GPU batch"}

    engine = fulu_b200.GPEngine(torch.device("cuda", local))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.prepare()
    spec = kernel_spec_from_sklearn(make_kernel(args.kernel))
    B = args.curves or default_b
    host = make_workload(args.workload, B, rank)
    lengths = host.lengths()
    nb = host.n_bands
    pinned = pin_batch(host)
    pinned_out = pin_results(B, nb, n_obs, spec.n_params)
    dev = DeviceBatch.from_host(host, engine.device, pinned=pinned, family=spec.family)
    in_bytes = dev.h2d_bytes()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=engine.device)   # > 126 MB L2

    def step_resident(profile=False):
        flush.zero_()   # L2 flush between timed iterations (the headline inputs are 28 MB, smaller than L2)
        return fit_augment_device(engine, dev, lengths, n_obs, spec=spec, profile=profile)

    def step_e2e():
        flush.zero_()
        return fulu_b200.fit_augment_batch(host, n_obs=n_obs, engine=engine, pinned=pinned, out=pinned_out, spec=spec)

    for _ in range(args.warmup):
        step_resident()
    peak_dfma = engine.fp64_peak(0)
    peak_dmma = engine.fp64_peak(1)
    peak = max(peak_dfma, peak_dmma)   # DFMA and DMMA share one pipe on B200: the higher reading is the pipe's rate
    library = None
    if rank == 0 and headline and not args.no_library_baseline:
        library = library_gpu_baseline(engine, host, spec)

    # ---- timed: resident inputs
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        sampler.start()
    e0.record()
    kept, timed_launches = [], 0
    for _ in range(args.steps):
        out = step_resident(profile="counts")           # counts only (no events): the launches of the timed region itself
        kept.append(out["n_eval"])                      # stays on the device: no host sync inside the timed region
        timed_launches += sum(int(pr["launches"]) + 1 for pr in out["profile"]) + 1   # + predict per class, + the L2 flush
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    f_eval_curve = lengths.astype(np.float64) ** 3 + 12.0 * lengths.astype(np.float64) ** 2      # SURVEY 8d, per evaluation
    evals = float(sum(float(x.sum().item()) for x in kept))
    flops = float(sum(float(np.dot(x.cpu().numpy().astype(np.float64), f_eval_curve)) for x in kept))
    t = torch.tensor([ms], dtype=torch.float64, device=engine.device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max * 1e-3)

    # ---- the same K steps once more with the library's per-launch CUDA events switched on (outside the timed region: creating
    # and recording three events per round costs ~10 % of a step; it also keeps the library on rounds of two kernels to the end
    # instead of finishing the tail in one fused launch): kernel durations for the roofline
    eval_ms, step_ms, rounds, classes = 0.0, 0.0, 0, []
    launches = timed_launches
    for _ in range(args.steps):
        classes = step_resident(profile=True)["profile"]
        for pr in classes:
            eval_ms += pr["eval_ms"]; step_ms += pr["step_ms"]; rounds += int(pr["rounds"])
    torch.cuda.synchronize()

    # ---- timed: end to end through the public API (pinned host in, host out)
    for _ in range(1):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        res = step_e2e()
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    t = torch.tensor([max(e0.elapsed_time(e1) * 1e-3, wall)], dtype=torch.float64, device=engine.device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())

    status = res.status
    flagged = float(np.count_nonzero(status & ~8))   # anything but "theta on a bound"
    stats = torch.tensor([evals, eval_ms, flagged, flops], dtype=torch.float64, device=engine.device)
    if dist is not None:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    evals_all, eval_ms_all, flops_all = float(stats[0]), float(stats[1]) / world, float(stats[3])
    achieved = (flops_all / world) / (eval_ms_all * 1e-3) / 1e12   # TFLOP/s of the evaluation kernel on one GPU (algorithmic flops)
    # DRAM traffic of the dominant kernel per launch: bytes per evaluation from the committed ncu --set full capture (N = 100), times
    # the evaluations an average launch of THIS run carried
    traffic, traffic_src = None, None
    if headline:
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_eval_kernel_traffic.json")))
            traffic = tj["dram_bytes_per_evaluation"] * (evals_all / world) / max(rounds, 1)
            traffic_src = ("profiles/r02_eval_kernel_traffic.json (dram__bytes_read+write per evaluation from the committed ncu --set full "
                           "capture of this kernel) x evaluations per launch of this run; not re-measured in this run")
        except (OSError, KeyError, ValueError):
            pass
    # parity next to the CPU pass: the CPU sample is the first curves of rank 0's batch, so the final LMLs compare curve by
    # curve, and the roofline gets a second figure normalised to sklearn's evaluation count (extra evaluations earn no credit)
    parity_cpu = None
    if cpu_curves:
        k = min(len(cpu_curves), B)
        ref_lml = np.array([c[0] for c in cpu_curves[:k]])
        ref_ev = np.array([c[1] for c in cpu_curves[:k]], np.float64)
        deficit = ref_lml - res.lml[:k]
        parity_cpu = {"curves": int(k), "max_lml_deficit_vs_sklearn": float(np.max(deficit)), "median_abs_lml_difference": float(np.median(np.abs(deficit))),
                      "fraction_within_1e-6": float(np.mean(deficit <= 1e-6)), "fraction_better_or_equal": float(np.mean(deficit <= 0.0)),
                      "sklearn_evals_per_curve": float(ref_ev.mean()), "gpu_evals_per_curve_same_curves": float(res.n_eval[:k].mean())}
    mean_in = in_bytes / B
    out_bytes = 16 * nb * n_obs + 8 * 4 + 16
    cfg = {"workload": desc, "curves_per_gpu_per_step": B, "n_points": N_POINTS if headline else f"{int(lengths.min())}..{int(lengths.max())} (mean {lengths.mean():.0f})",
           "n_bands": nb, "n_obs": n_obs, "n_starts": 6, "kernel": args.kernel, "n_params": spec.n_params,
           "l2": f"flushed between steps (256 MiB memset); inputs {in_bytes / 1e6:.0f} MB", "parallelism": f"curve-sharded x{world}, no collective"}
    line = {
        "metric": METRIC if headline else f"light curves/sec GP fit+augment ({args.workload} workload, {args.kernel} kernel)", "value": value, "unit": "curves/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": cfg,
        "e2e": {"value": e2e_value, "unit": "curves/s", "h2d_bytes_per_step": int(res.h2d_bytes), "d2h_bytes_per_step": int(res.d2h_bytes)},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "fp64", "kernel": "fit_eval_kernel" if rounds else "fit_tail_kernel (long-curve classes: the whole fit is one fused launch)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": (mean_in + 8 * 4 + 8 * 5) * (evals_all / world) / max(rounds, 1),
                     "peak_source": "max of the DFMA and DMMA m8n8k4 register-resident loops measured in this run (fulu_gp_fp64_probe, CUDA events; "
                                    "the two instructions share the FP64 pipe); MEASURED_PEAKS.json has no FP64 entry",
                     "peak_dfma": peak_dfma, "peak_dmma": peak_dmma, "flops_per_eval": (N_POINTS ** 3 + 12 * N_POINTS ** 2) if args.workload == "plasticc" else flops_all / max(evals_all, 1),
                     "evals_per_curve": evals_all / (world * B * args.steps),
                     "frac_at_sklearn_evals": (achieved / peak * parity_cpu["sklearn_evals_per_curve"] / parity_cpu["gpu_evals_per_curve_same_curves"])
                     if (parity_cpu and peak) else None,
                     "eval_kernel_ms_per_step": eval_ms_all / args.steps, "eval_kernel_share_of_step": eval_ms_all / ms_max,
                     "timing": "kernel durations from CUDA events on the launching stream in a separate pass of the same K steps (same launch "
                               "schedule, the fused launch that finishes a fit included); the timed region itself runs without them",
                     "step_kernel_ms_per_step": step_ms / args.steps, "rounds_per_step": rounds / args.steps,
                     "hbm_gbs_algorithmic": value / world * (mean_in + out_bytes) / 1e9,
                     "launch_classes": [{"n_max": int(c["bound"]), "curves": int(c["curves"]), "ctas": int(c["grid_eval"]), "threads": int(c["threads_eval"]),
                                         "ctas_per_sm": int(c["ctas_per_sm"]), "smem_per_cta": int(c["smem_eval"]), "matrix_in_global": bool(c["global_a"]),
                                         "eval_ms": c["eval_ms"]} for c in classes],
                     "grid": int(classes[0]["grid_eval"]), "smem_per_cta": int(classes[0]["smem_eval"])},
        "cpu_baseline": cpu_base,
        "gpu_library_baseline": library,
        "parity": {"curves_flagged": int(stats[2]), "status_histogram_rank0": res.status_histogram(),
                   "mean_lml": float(np.mean(res.lml[np.isfinite(res.lml)])), "vs_cpu_pass": parity_cpu},
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--curves", type=int, default=0, help="curves per GPU per step (default: the workload's own, 10,000 for config #3)")
    ap.add_argument("--workload", default="plasticc", choices=sorted(WORKLOADS), help="plasticc = the headline config #3 (default); ztf = config #4 sample; ddf = config #5")
    ap.add_argument("--kernel", default="default", choices=["default", "rbf_matern", "notebook", "rbf_rq"],
                    help="covariance: the reference's default C*RBF+White (headline) or another kernel its users pass (SURVEY 8f-1)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-curves", type=int, default=256, help="CPU sample size per step (bounded)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true", help="skip the cuSOLVER/torch yardstick of the headline run")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): every rank its own batch; strong: ONE global batch LPT-sharded over the ranks, results gathered "
                         "onto rank 0 inside the timed end-to-end call (sharding.fit_augment_sharded)")
    args = ap.parse_args()
    global KERNEL
    KERNEL = args.kernel
    if args.impl == "reference":
        run_reference(args)
    elif args.scaling == "strong":
        run_gpu_strong(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
