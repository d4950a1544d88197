"""ORACLE SHIM: inert torchmetrics (evaluation only; off the hot path)."""
from . import functional  # noqa: F401
