"""The REAL reference class as the CPU arm.  TEST INFRASTRUCTURE ONLY (bench.py's cpu_baseline / --impl reference, tests/).

``/root/reference`` does not travel to the GPU box, so ``__graft_entry__.build()`` - run in the build container, where the
reference tree exists - stages the hot path's own files (``fulu/gp_aug.py``, ``fulu/_base_aug.py``, ``fulu/plotting.py``: the
second imports the third) UNMODIFIED under ``oracle/_ref/fulu/``.  That directory is git-ignored (no reference source enters the
history) but not gpurun-ignored, so it rides to the GPU box like a built ``.so``.  ``load()`` imports ``fulu.gp_aug`` from there
through empty stub modules for matplotlib (absent from the image; only the plotting code touches it) and a stub ``fulu`` package,
exactly as ``tests/golden/make_golden.py`` imports it from ``/root/reference``.  Where the staged files are missing ``load()``
returns None and the callers fall back to ``oracle/fulu_sklearn_port.py`` (the same library calls re-typed; bit-for-bit equal on
the golden fixtures) and say so (``cpu_baseline.kind``: "reference" or "port").
"""
from __future__ import annotations

import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGE = os.path.join(HERE, "_ref", "fulu")
FILES = ("gp_aug.py", "_base_aug.py", "plotting.py")


def stage(reference_root="/root/reference"):
    """Copy the reference's hot-path files into oracle/_ref/fulu (build container only).  Returns True if they are in place."""
    src = os.path.join(reference_root, "fulu")
    if all(os.path.exists(os.path.join(src, f)) for f in FILES):
        os.makedirs(STAGE, exist_ok=True)
        for f in FILES:
            shutil.copyfile(os.path.join(src, f), os.path.join(STAGE, f))
    return staged()


def staged():
    return all(os.path.exists(os.path.join(STAGE, f)) for f in FILES)


def load():
    """fulu.gp_aug.GaussianProcessesAugmentation of the staged reference files, or None."""
    if not staged():
        return None
    have = sys.modules.get("fulu")
    if have is not None and getattr(have, "__path__", None) != [STAGE]:
        return None   # another `fulu` is already imported in this process: do not shadow it
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors"):
            sys.modules.setdefault(m, types.ModuleType(m))
        sys.modules["matplotlib.colors"].TABLEAU_COLORS = {}
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    if have is None:
        pkg = types.ModuleType("fulu")
        pkg.__path__ = [STAGE]
        sys.modules["fulu"] = pkg
    try:
        import fulu.gp_aug as gp_aug
    except Exception:
        return None
    return gp_aug.GaussianProcessesAugmentation
