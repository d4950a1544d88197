"""ORACLE: import the unmodified reference (`/root/reference/robust_e_nerf`) on
top of the shims.  Only works where /root/reference exists (the build container);
the GPU box uses `oracle.restatement` and the committed golden vectors instead."""
import contextlib
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get("REN_REFERENCE_ROOT", "/root/reference")
SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "robust_e_nerf"))


def load_reference():
    """Returns the imported `robust_e_nerf` package of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for p in (SHIM_DIR, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch

    # robust_e_nerf.py:243 enters `torch.cuda.device(gpus[0])`; neutralise it on
    # CUDA-less hosts (SURVEY.md section 8b "no true CPU mode").
    if not torch.cuda.is_available():
        torch.cuda.device = lambda *a, **k: contextlib.nullcontext()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import robust_e_nerf  # noqa: F401
        import robust_e_nerf.models.robust_e_nerf  # noqa: F401
        import robust_e_nerf.data.datamodule  # noqa: F401
    return sys.modules["robust_e_nerf"]
