"""Build budgets of the shared-memory evaluation kernel, read from the objects `python -m fulu_b200.build` leaves behind (no GPU).

The kernel is bound by the length of each warp's own instruction stream and shares the SM's instruction cache with three more
resident copies of itself: in round 2 the six-parameter families lost a third of their throughput when extra code paths took their
evaluation kernel from ~60 KB to 99 KB of SASS, and a register spill in it costs more than any of the round's
gains.  So: every family's `fit_eval_kernel<..., GP_PLACE_SMEM>` stays under 72 KB, within 128 registers, without a stack frame."""
import glob
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "fulu_b200", "_lib", "obj")


def _objects():
    return sorted(glob.glob(os.path.join(OBJ, "gp_family_*.o")))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_evaluation_kernel_stays_inside_its_budgets():
    from fulu_b200 import build

    build.build()
    objs = _objects()
    assert len(objs) == len(build.FAMILIES)
    for o in objs:
        elf = subprocess.run(["cuobjdump", "-elf", o], capture_output=True, text=True).stdout
        sizes = [int(m.group(1), 16) for m in re.finditer(r"^\s*\w+\s+\w+\s+(\w+)\s.*PROGBITS.*\.text\.\S*fit_eval_kernel\S*ELi0EEEv", elf, re.M)]
        assert len(sizes) == 1, (o, sizes)
        assert sizes[0] <= 72 * 1024, f"{os.path.basename(o)}: evaluation kernel is {sizes[0] // 1024} KB of SASS"
        res = subprocess.run(["cuobjdump", "-res-usage", o], capture_output=True, text=True).stdout
        m = re.search(r"fit_eval_kernel\S*ELi0EEEv\S*:\s*\n\s*REG:(\d+) STACK:(\d+)", res)
        assert m, o
        assert int(m.group(1)) <= 128 and int(m.group(2)) == 0, (os.path.basename(o), m.groups())
