"""bench.py's contract on a box without a GPU: the reference arm (the reference's CPU path, the one place besides tests/ and smoke()
that may execute oracle/) prints one JSON line with the agreed keys, only rank 0 works under torchrun, and the product arm refuses
to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*argv, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True, cwd=ROOT, env=e, timeout=600)


def test_reference_arm_line():
    r = _bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-curves", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "curves/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    from oracle import ref_arm

    # "reference" where the reference's own files are staged for the CPU arm (oracle/_ref/: build container and what it ships), else "port"
    assert cb["kind"] == ("reference" if ref_arm.staged() else "port") and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["curves_per_step"] == 2 * cb["cores"]   # the bounded sample: never fewer than two curves per core


def test_reference_arm_other_ranks_do_nothing():
    r = _bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_product_arm_refuses_without_a_gpu():
    r = _bench("--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
