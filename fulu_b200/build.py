"""In-tree build of libfulu_gp.so (hand-written sm_100a CUDA, see csrc/).

    python -m fulu_b200.build            # build if stale
    python -m fulu_b200.build --force

nvcc cross-compiles for sm_100a without a GPU; the resulting .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libfulu_gp.so")
SOURCES = ["gp_kernels.cu"]
HEADERS = ["gp_block.cuh", "lbfgsb.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
              "-diag-suppress", "177"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    m = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(INCLUDE, "fulu_gp.h")]
    return any(os.path.getmtime(d) > m for d in deps)


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("fulu_b200: nvcc not found; cannot build libfulu_gp.so (there is no CPU fallback)")
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + [os.path.join(CSRC, f) for f in SOURCES]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
