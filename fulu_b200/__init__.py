"""fulu_b200 — B200-native Gaussian-process light-curve augmentation (the GP hot path of hse-cs/fulu).

Importing the package does not need a GPU; constructing an engine / fitting does, and fails loudly without one.
"""
from .datasets import CurveBatch  # noqa: F401
from .engine import (AugmentResult, DeviceBatch, GPEngine, KernelSpec, fit_augment_batch, fit_augment_device, fit_augment_table,  # noqa: F401
                     kernel_spec_from_sklearn, pin_results)
from .gp_aug import ConvergenceWarning, GaussianProcessesAugmentation, add_log_lam, create_aug_data  # noqa: F401
from .plotting import LcPlotter  # noqa: F401  (fulu.LcPlotter, README.md:61; imports matplotlib only when something is drawn)

__all__ = ["GaussianProcessesAugmentation", "fit_augment_batch", "GPEngine", "DeviceBatch", "CurveBatch", "KernelSpec", "AugmentResult",
           "add_log_lam", "create_aug_data", "ConvergenceWarning", "fit_augment_table", "kernel_spec_from_sklearn", "LcPlotter"]
