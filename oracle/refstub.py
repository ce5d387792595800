"""TEST INFRASTRUCTURE — makes the read-only reference (/root/reference) importable in the build
container so the oracle can be pinned against it and golden vectors generated.

The reference imports `optuna` and `bottleneck` at module top (train_SDRM.py:12, utilities.py:3);
neither is installed here, so two tiny stubs are put in sys.modules.  `pandas` must be imported
BEFORE the bottleneck stub (pandas probes sys.modules['bottleneck'] for a version).
On a CPU-only box VAE.get_l2_reg calls .cuda() unconditionally (train_SDRM.py:262-263); that is
patched to identity.  Nothing here is used by the product path, and nothing under tests -m gpu /
bench.py / smoke() may call it (the reference does not exist on the GPU box).
"""
import os
import sys
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_DIR = os.path.join(_ROOT, "baseline", "_ref")   # git-ignored copy that travels to the GPU box (stage_reference)
_FILES = ("train_SDRM.py", "utilities.py", "dataloaders.py", "svd_benchmark.py", "main.py", "hyperparameter_search.py",
          "mlp_benchmark.py", "neural_cf_benchmark_pt.py")


def _pick_dir():
    env = os.environ.get("SDRM_REFERENCE_DIR")
    for d in (env, "/root/reference", STAGED_DIR):
        if d and os.path.isfile(os.path.join(d, "train_SDRM.py")):
            return d
    return env or "/root/reference"


REFERENCE_DIR = _pick_dir()


def stage_reference(src="/root/reference"):
    """Copy the reference's (unmodified) Python files into the git-ignored baseline/_ref/ so that `bench.py --impl reference`
    can time the reference's OWN sample_ddpm on the GPU box, where /root/reference does not exist.  Never tracked by git."""
    import shutil
    if not os.path.isfile(os.path.join(src, "train_SDRM.py")):
        return False
    os.makedirs(STAGED_DIR, exist_ok=True)
    for f in _FILES:
        if os.path.isfile(os.path.join(src, f)):
            shutil.copyfile(os.path.join(src, f), os.path.join(STAGED_DIR, f))
    return True


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "train_SDRM.py"))


def import_reference():
    """Return the reference's train_SDRM module (imports utilities, dataloaders lazily)."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_DIR}")
    import numpy as np
    import pandas  # noqa: F401  (must precede the bottleneck stub)
    import torch

    if "optuna" not in sys.modules:
        optuna = types.ModuleType("optuna")

        class TrialPruned(Exception):
            pass

        optuna.TrialPruned = TrialPruned
        optuna.exceptions = types.SimpleNamespace(TrialPruned=TrialPruned)
        sys.modules["optuna"] = optuna
    if "bottleneck" not in sys.modules:
        bn = types.ModuleType("bottleneck")
        bn.argpartition = lambda a, kth, axis=-1: np.argpartition(a, kth, axis=axis)
        bn.__version__ = "1.3.7"
        sys.modules["bottleneck"] = bn
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import train_SDRM as ref  # noqa: E402

    return ref
