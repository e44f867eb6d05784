import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _all_golden():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def golden_names():
    """Fixtures of the default kernel (fulu/gp_aug.py:29), short curves."""
    return [n for n in _all_golden() if not n.startswith("fam_") and not n.startswith("ddf_")]


def long_golden_names():
    """BASELINE config #5 fixtures (N = 1024, 1536; tests/golden/make_golden.py --ddf): the long-curve path."""
    return [n for n in _all_golden() if n.startswith("ddf_")]


def family_golden_names():
    """Fixtures of the non-default kernels (tests/golden/make_golden.py --families); theta / grad in the kernel object's order."""
    return [n for n in _all_golden() if n.startswith("fam_")]


def family_kernel(name):
    """The sklearn kernel object a fam_* fixture was generated with (same table as make_golden.family_kernels)."""
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel as C, Matern, RationalQuadratic, WhiteKernel

    return {
        "rbf_matern": lambda: C(1.0) * RBF([1, 1]) + Matern() + WhiteKernel(),
        "notebook": lambda: C(1.0) * Matern() * RBF([1, 1]) + Matern() + WhiteKernel(),
        "rbf_rq": lambda: C(1.0) * RBF([1, 1]) + RationalQuadratic() + WhiteKernel(),
        "matern_rbf": lambda: C(1.0) * Matern() * RBF([1.0, 1.0]) + WhiteKernel(),
        "rbf_matern05": lambda: C(1.0) * RBF([1, 1]) + Matern(nu=0.5) + WhiteKernel(),
        "permuted_matern25": lambda: WhiteKernel(0.5) + Matern(length_scale=2.0, nu=2.5) + RBF([2.0, 0.5]) * C(4.0),
    }[str(name)]()


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(params=golden_names())
def golden(request):
    g = load_golden(request.param)
    g["name"] = request.param
    return g


def relerr(a, b, floor=0.0):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
