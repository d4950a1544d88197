"""PyTorch-Lightning surface used by the task module.

When `pytorch_lightning` is importable (the reference pins 1.4.9, environment.yml:19)
`LightningModule` IS Lightning's class and `scripts/run.py` drives the module unchanged.
This image has no Lightning, so a minimal stand-in provides
exactly the attributes `RobustENeRF` touches (SURVEY.md section 8b "Python surface to keep"):
`save_hyperparameters`, `hparams`, `log`, `all_gather`, `trainer`, `global_step`,
`current_epoch`, `device`, `logger`.
"""
import inspect

import torch

try:  # pragma: no cover - not installed in this image
    import pytorch_lightning as pl
    LightningModule = pl.LightningModule
    HAVE_LIGHTNING = True
except ImportError:
    HAVE_LIGHTNING = False

    class _AttrDict(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__

    class LightningModule(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.hparams = _AttrDict()
            self.trainer = None
            self.global_step = 0
            self.current_epoch = 0
            self.logger = None
            self.logged = {}

        def save_hyperparameters(self, *names):
            frame = inspect.currentframe().f_back
            local = frame.f_locals
            for n in names:
                self.hparams[n] = local[n]

        def log(self, name, value, **kwargs):
            self.logged[name] = value

        def all_gather(self, tensor):
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                out = [torch.empty_like(tensor) for _ in range(torch.distributed.get_world_size())]
                torch.distributed.all_gather(out, tensor)
                return torch.stack(out)
            return tensor.unsqueeze(0)

        @property
        def device(self):
            for p in self.parameters():
                return p.device
            return torch.device("cpu")
