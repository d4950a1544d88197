"""On-device training-batch production (the step before the hot path).

The reference builds every batch on the host: `IterableMapDataset` draws `randint`-with-replacement event
indices (robust_e_nerf/utils/datasets.py:20-32), three `IterableDataset` samplers draw the normalised
timestamps (robust_e_nerf/data/samplers.py:6-118, wired in data/datamodule.py:149-196), a DataLoader pins and
copies them.  Here the converted events (`events.pt` schema) live in HBM and a batch is a handful of device
ops on the current stream -- no host work, no H2D copy per step.  Same distributions and dtypes; the random
streams necessarily differ from the reference's CPU generator (statistical, not bitwise, equivalence --
`tests/test_module_cpu.py::test_device_sampler_distributions`).
"""
import math

import torch


def trunc_normal(low, high, size, mean, std, dtype=None, generator=None, device=None):
    """`samplers.trunc_normal` (data/samplers.py:33-85): inverse-CDF truncated normal."""
    def norm_cdf(x):
        return (1. + math.erf(x / math.sqrt(2.))) / 2.
    l, u = norm_cdf((low - mean) / std), norm_cdf((high - mean) / std)
    t = 2 * (u - l) * torch.rand(size, dtype=dtype, generator=generator, device=device) + (2 * l - 1)
    t.erfinv_()
    t.mul_(std * math.sqrt(2.)).add_(mean)
    return t.clamp_(min=low, max=high)


class DeviceEventSampler:
    """Yields `{"event": {...}, "normalized": {...}}` batches with the leading DataLoader dimension of 1,
    exactly what `RobustENeRF.training_step` consumes (robust_e_nerf.py:301-321).

    events: dict of tensors in the `events.pt` schema (position (M,2), start_ts, end_ts, num_pos, num_neg (M,)
    [, channel_idx]) -- e.g. the output of `ren_b200.events.queue_raw_events`.  `batch_size` may be changed at
    any time (the dynamic batch sizing of robust_e_nerf.py:907-950 writes it)."""

    def __init__(self, events, batch_size, device="cuda", seed=0, rank=0):
        self.device = torch.device(device)
        self.events = {k: v.to(self.device).contiguous() for k, v in events.items()}
        self.events["position"] = self.events["position"].to(torch.float32)       # datasets.py:60-64
        self.n = int(self.events["end_ts"].numel())
        self.batch_size = int(batch_size)
        self.generator = torch.Generator(device=self.device)
        self.generator.manual_seed(seed + rank)                                    # datamodule.py:85-89

    def __iter__(self):
        return self

    def __next__(self):
        B, g, dev = self.batch_size, self.generator, self.device
        idx = torch.randint(self.n, (B,), generator=g, device=dev)                # utils/datasets.py:20-32
        event = {k: v.index_select(0, idx).unsqueeze(0) for k, v in self.events.items()}
        normalized = {
            "ts_diff": torch.full((1, B), 1.0, dtype=torch.float64, device=dev),   # DiracDeltaSampler(center=1)
            "diff_start_ts": torch.rand((1, B), dtype=torch.float64, generator=g, device=dev),
            "grad_ts": trunc_normal(0, 1, (1, B), 0.5, 0.25, dtype=torch.float64, generator=g, device=dev),
        }
        return {"event": event, "normalized": normalized}
