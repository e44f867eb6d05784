"""The block code (fulu_b200/csrc/gp_block.cuh) under AddressSanitizer + UBSan and under ThreadSanitizer, on the host-thread
emulation (tests/_host/sanitize_main.cpp).  compute-sanitizer is closed on the GPU pool; this is the stand-in:
  * ASan/UBSan: the "shared memory" block is a heap array of exactly gp_smem_doubles(n), the inputs arrays of exactly n, so an
    index past either end, a misaligned 16-byte access or a signed overflow aborts the run;
  * TSan: two emulated CUDA threads touching one shared-memory word without a barrier (block or warp) between them is a data
    race report - what racecheck would say.  Its reports are real; silence is not a proof (four remembered accesses per word),
    which is why the self-test below checks that dropping a barrier IS reported.
The full matrix (9 sizes x 4 block sizes x 4 families, 5.5 min) is scripts/sanitize_host.sh; this file runs a slice of it."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "_host", "sanitize_main.cpp")
DEPS = [SRC, os.path.join(HERE, "_host", "host_gp_block.cpp"), os.path.join(os.path.dirname(HERE), "fulu_b200", "csrc", "gp_block.cuh"),
        os.path.join(os.path.dirname(HERE), "fulu_b200", "csrc", "gp_tiled.cuh")]
BUILD = os.path.join(HERE, "_build")
FLAGS = {
    "asan": ["-fsanitize=address,undefined", "-fno-omit-frame-pointer"],
    "tsan": ["-fsanitize=thread"],
    "tsan_skip": ["-fsanitize=thread", "-DSAN_SKIP_SYNC"],
}


def _exe(kind):
    exe = os.path.join(BUILD, "sanitize_" + kind)
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(p) for p in DEPS):
        os.makedirs(BUILD, exist_ok=True)
        r = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-pthread", *FLAGS[kind], SRC, "-o", exe], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("sanitizer runtime not available: " + r.stderr[-300:])
    return exe


def _run(kind, n, nthr, family, kslab, n_obs=8, env=None, tiled=0):
    r = subprocess.run([_exe(kind), str(n), str(nthr), str(family), str(kslab), str(n_obs), str(tiled)], capture_output=True, text=True,
                       env=dict(os.environ, **(env or {})), timeout=600)
    return r.returncode, r.stdout + r.stderr


pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")

# (n, threads per block, family code F1 | EX << 4, K slab): the headline geometry, a border that opens its own tile-row
# (n a multiple of 8), the smallest blocks, the notebook family (two Matern factors) and RationalQuadratic
CASES = [(100, 128, 0, 1), (104, 256, 0, 0), (8, 32, 0, 1), (2, 128, 0, 0), (100, 128, 34, 0), (30, 256, 64, 0),
         # the scratch-free layout (family bit 0x400) at the geometries plan() runs it with: its footprint has no scratch column (ASan), its
         # triangular inverse holds a block column's rows in registers across a barrier between reading W and writing over it (TSan)
         (120, 160, 0x400, 1), (150, 256, 0x400, 0)]


@pytest.mark.parametrize("kind", ["asan", "tsan"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_t%d_f%d_k%d" % c)
def test_block_code_is_clean(kind, case):
    rc, out = _run(kind, *case)
    assert rc == 0 and "Sanitizer" not in out and "runtime error" not in out, out[-3000:]


# the long-curve path (gp_tiled.cuh): (n, threads per block, family): several panels with a narrow last one, a row block with fewer
# rows than warps, 2 / 4 / 8 warps, the notebook family
TILED_CASES = [(100, 128, 0), (130, 256, 0), (72, 64, 34)]


@pytest.mark.parametrize("kind", ["asan", "tsan"])
@pytest.mark.parametrize("case", TILED_CASES, ids=lambda c: "n%d_t%d_f%d" % c)
def test_tiled_path_is_clean(kind, case):
    """Slab, stages and barriers allocated at exactly their sizes (ASan); the emulated copy engine's writes into a stage against the
    warps' reads of it, ordered only by the full / empty barriers (TSan)."""
    n, nthr, fam = case
    rc, out = _run(kind, n, nthr, fam, 0, tiled=1)
    assert rc == 0 and "Sanitizer" not in out and "runtime error" not in out, out[-3000:]


def test_tsan_reports_a_stage_refilled_before_its_release():
    """Self-test of the pipeline protocol: with the producer's wait on the stage's `empty` barrier removed, the copy into a stage
    races with the slower warps still reading it - ThreadSanitizer must say so; with the wait in place it is silent."""
    rc, out = _run("tsan_skip", 130, 256, 0, 0, env={"SAN_SKIP_REUSE_WAIT": "1", "SAN_SKIP_SYNC": "0"}, tiled=1)
    assert "ThreadSanitizer: data race" in out
    rc, out = _run("tsan_skip", 130, 256, 0, 0, env={"SAN_SKIP_SYNC": "0"}, tiled=1)
    assert rc == 0 and "Sanitizer" not in out, out[-3000:]


@pytest.mark.parametrize("dropped", [1, 3, 5])
def test_tsan_reports_a_dropped_barrier(dropped):
    """Self-test: with the block barrier after the load (1), before the first panel multiplication (3) or before the second (5)
    removed from the evaluation, ThreadSanitizer must report a race in the code that barrier separates."""
    rc, out = _run("tsan_skip", 100, 128, 0, 1, env={"SAN_SKIP_SYNC": str(dropped), "SAN_SKIP_BLOCK": "1"})
    assert "ThreadSanitizer: data race" in out and "gp_block.cuh" in out
    rc, out = _run("tsan_skip", 100, 128, 0, 1, env={"SAN_SKIP_SYNC": "0"})
    assert rc == 0 and "Sanitizer" not in out


def test_optimiser_state_machine_is_clean():
    """fulu_b200/csrc/lbfgsb.cuh under ASan + UBSan (fixed-size arrays inside the state struct: an index past an end is a UBSan
    bounds report) over 1..6 variables x four objectives (interior optimum, optimum on the bounds, +inf region, flat) x three
    settings (defaults, tiny iteration / evaluation / line-search limits, a start outside the box)."""
    src = os.path.join(HERE, "_host", "sanitize_lbfgsb_main.cpp")
    deps = [src, os.path.join(HERE, "_host", "host_lbfgsb.cpp"), os.path.join(os.path.dirname(HERE), "fulu_b200", "csrc", "lbfgsb.cuh")]
    exe = os.path.join(BUILD, "sanitize_lbfgsb")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(p) for p in deps):
        os.makedirs(BUILD, exist_ok=True)
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", *FLAGS["asan"], src, "-o", exe], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("sanitizer runtime not available: " + r.stderr[-300:])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "Sanitizer" not in out and "runtime error" not in out and "0 bad" in out, out[-3000:]
