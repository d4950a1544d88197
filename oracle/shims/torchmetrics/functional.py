"""ORACLE SHIM: inert torchmetrics.functional."""


def _unavailable(*args, **kwargs):
    raise NotImplementedError("torchmetrics is evaluation-only; not part of the oracle")


peak_signal_noise_ratio = _unavailable
structural_similarity_index_measure = _unavailable
