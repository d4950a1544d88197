"""ORACLE SHIM (test infrastructure): the sliver of pytorch-lightning==1.4.9
(environment.yml:19) that `RobustENeRF` / `DataModule` touch when a training step
is driven by hand: `LightningModule` (save_hyperparameters, hparams, log,
all_gather, trainer, global_step, current_epoch, logger, device),
`LightningDataModule`, `seed_everything` (robust_e_nerf/models/robust_e_nerf.py:16,
162-173,375-379,489-515,916-919; robust_e_nerf/data/datamodule.py:8,56-60).
"""
import inspect
import random

import numpy as np
import torch

try:
    from easydict import EasyDict as _AttrDict
except ImportError:  # pragma: no cover
    class _AttrDict(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__


def seed_everything(seed=None, workers=False):
    seed = 0 if seed is None else int(seed)
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    return seed


class _HParamsMixin:
    def save_hyperparameters(self, *names, **kwargs):
        frame = inspect.currentframe().f_back
        local_vars = frame.f_locals
        if not hasattr(self, "_hparams"):
            object.__setattr__(self, "_hparams", _AttrDict())
        if not names:
            init_sig = inspect.signature(type(self).__init__)
            names = [n for n in init_sig.parameters if n != "self"]
        for name in names:
            self._hparams[name] = local_vars[name]

    @property
    def hparams(self):
        if not hasattr(self, "_hparams"):
            object.__setattr__(self, "_hparams", _AttrDict())
        return self._hparams


class _Trainer:
    """Attribute bag standing in for `pl.Trainer` (only what training_step reads)."""

    def __init__(self, datamodule=None, accumulate_grad_batches=1):
        self.accumulate_grad_batches = accumulate_grad_batches
        self.datamodule = datamodule
        self.is_global_zero = True
        self.log_dir = None


class LightningModule(_HParamsMixin, torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.trainer = None
        self.global_step = 0
        self.current_epoch = 0
        self.logger = None
        self.logged = {}

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        for b in self.buffers():
            return b.device
        return torch.device("cpu")

    def log(self, name, value, **kwargs):
        self.logged[name] = value

    def all_gather(self, data, group=None, sync_grads=False):
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size()
            out = [torch.zeros_like(data) for _ in range(world)]
            torch.distributed.all_gather(out, data)
            return torch.stack(out)
        return data


class LightningDataModule(_HParamsMixin):
    def __init__(self):
        super().__init__()


Trainer = _Trainer

from . import callbacks, loggers, plugins  # noqa: E402,F401
