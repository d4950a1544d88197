"""ORACLE SHIM: inert LPIPS (evaluation only; off the hot path,
robust_e_nerf/loss_metric/metric.py:18-19)."""
import torch


class LPIPS(torch.nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, *args, **kwargs):
        raise NotImplementedError("LPIPS is evaluation-only; not part of the oracle")
