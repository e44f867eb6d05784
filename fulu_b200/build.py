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
SOURCES = ["gp_kernels.cu", "gp_family.cu"]
HEADERS = ["gp_block.cuh", "gp_tiled.cuh", "gp_device.cuh", "lbfgsb.cuh"]
# covariance families compiled into the library (product factor, added term; gp_block.cuh GP_F_*) - keep in step with
# GP_FAMILY_LIST in csrc/gp_device.cuh.  One translation unit each, compiled in parallel.
FAMILIES = [(0, 0), (0, 2), (2, 2), (0, 4), (0, 1), (0, 3), (2, 0)]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
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
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    extra = (["-Xptxas", "-v"] if verbose else []) + os.environ.get("FULU_GP_NVCC_EXTRA", "").split()
    jobs = [(os.path.join(obj_dir, "gp_kernels.o"), [os.path.join(CSRC, "gp_kernels.cu")])]
    for f1, ex in FAMILIES:
        jobs.append((os.path.join(obj_dir, f"gp_family_{f1}_{ex}.o"), [f"-DGP_FAM_F1={f1}", f"-DGP_FAM_EX={ex}", os.path.join(CSRC, "gp_family.cu")]))

    def compile_one(job):
        obj, args = job
        subprocess.check_call([nvcc] + NVCC_FLAGS + extra + ["-c", "-o", obj] + args, cwd=CSRC)
        return obj

    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, jobs))
    out = os.environ.get("FULU_GP_LIB_OUT", LIB_PATH)
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs, cwd=CSRC)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
