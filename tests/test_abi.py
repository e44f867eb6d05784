"""CPU-only checks of the C-ABI boundary: the library builds, loads, and exports every symbol include/fulu_gp.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from fulu_b200 import build

    return build.build()


def test_header_symbols_are_exported(lib_path):
    hdr = open(os.path.join(ROOT, "include", "fulu_gp.h")).read()
    declared = set(re.findall(r"\b(fulu_gp_[a-z0-9_]+)\s*\(", hdr))
    assert {"fulu_gp_fit_batch", "fulu_gp_lml_batch", "fulu_gp_predict_grid_batch", "fulu_gp_predict_points_batch"} <= declared
    lib = ctypes.CDLL(lib_path)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/fulu_gp.h but not exported by libfulu_gp.so"


def test_binding_loads_and_reports_version(lib_path):
    from fulu_b200 import _capi

    lib = _capi.load()
    assert lib.fulu_gp_abi_version() == _capi.ABI_VERSION == 5
    assert b"workspace" in lib.fulu_gp_strerror(-2)
    assert set(_capi.EXPORTS) <= {n for n in dir(lib)} | set(_capi.EXPORTS)


def test_struct_layout_matches_header():
    from fulu_b200 import _capi

    # fulu_gp_batch: 2*i64 + 2*i32 + 6 pointers + 2*i32 + f64 + pointer + i64 = 104 bytes; fit options: 10*i32 + 2*f64 + 12*f64 + pointer = 160
    assert ctypes.sizeof(_capi.Batch) == 104
    assert ctypes.sizeof(_capi.FitOptions) == 160


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import fulu_b200

    with pytest.raises(RuntimeError):
        fulu_b200.GPEngine()
    aug = fulu_b200.GaussianProcessesAugmentation({0: 3.5, 1: 3.7})
    with pytest.raises(RuntimeError):
        aug.fit([0.0, 1.0, 2.0], [1.0, 2.0, 1.5], [0.1, 0.1, 0.1], [0, 1, 0])


def _reference_class_signatures():
    """The reference's own class, imported where it lies (build container only), else the signatures recorded from it."""
    import inspect
    import sys
    import types

    ref = "/root/reference"
    recorded = {"__init__": "(self, passband2lam, kernel=1**2 * RBF(length_scale=[1, 1]) + WhiteKernel(noise_level=1), use_err=False)",
                "fit": "(self, t, flux, flux_err, passband)", "predict": "(self, t, passband)",
                "augmentation": "(self, t_min, t_max, n_obs=100)"}
    if not os.path.isdir(ref):
        return recorded, False
    saved = {m: sys.modules.get(m) for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "fulu", "fulu.gp_aug", "fulu._base_aug")}
    try:
        for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors"):
            sys.modules[m] = types.ModuleType(m)
        sys.modules["matplotlib.colors"].TABLEAU_COLORS = {}
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        pkg = types.ModuleType("fulu")
        pkg.__path__ = [os.path.join(ref, "fulu")]
        sys.modules["fulu"] = pkg
        import fulu.gp_aug as gp_aug

        cls = gp_aug.GaussianProcessesAugmentation
        live = {k: str(inspect.signature(getattr(cls, k))) for k in recorded}
    finally:
        for m, v in saved.items():
            if v is None:
                sys.modules.pop(m, None)
            else:
                sys.modules[m] = v
    assert live == recorded, "the recorded signatures are stale"
    return live, True


def test_class_signatures_are_the_reference_s():
    """SURVEY 8b "Python surface to keep (exact)": constructor, fit, predict, augmentation take the reference's arguments with the
    reference's defaults (fulu/gp_aug.py:29,37,80; fulu/_base_aug.py:88), and the default kernel is the reference's kernel object."""
    import inspect

    from sklearn.gaussian_process.kernels import RBF, ConstantKernel as C, WhiteKernel

    import fulu_b200

    ref, _ = _reference_class_signatures()
    cls = fulu_b200.GaussianProcessesAugmentation
    for name, sig in ref.items():
        assert str(inspect.signature(getattr(cls, name))) == sig, name
    aug = cls({0: 3.5, 1: 3.7})
    assert aug.kernel == C(1.0) * RBF([1.0, 1.0]) + WhiteKernel() and aug.use_err is False
    assert aug.ss is None and aug.reg is None and aug.t_train is None


def test_product_does_not_import_oracle():
    import subprocess
    import sys

    code = "import sys, fulu_b200, fulu_b200.engine, fulu_b200.gp_aug; assert not any(m.startswith('oracle') for m in sys.modules)"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for f in os.listdir(os.path.join(ROOT, "fulu_b200")):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "fulu_b200", f)).read().replace("# oracle", "")


def test_unsupported_kernel_is_refused():
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel

    import fulu_b200

    from sklearn.gaussian_process.kernels import ExpSineSquared

    fulu_b200.GaussianProcessesAugmentation({0: 3.5}, kernel=ConstantKernel(1.0) * RBF([1.0, 1.0]) + WhiteKernel())
    fulu_b200.GaussianProcessesAugmentation({0: 3.5}, kernel=ConstantKernel(1.0) * RBF([1, 1]) + Matern() + WhiteKernel())
    with pytest.raises(NotImplementedError):
        fulu_b200.GaussianProcessesAugmentation({0: 3.5}, kernel=ConstantKernel(1.0) * RBF([1, 1]) + ExpSineSquared() + WhiteKernel())
