"""Data module of a sequence: the reference's `DataModule`
(robust_e_nerf/data/datamodule.py:8-272) with its constructor arguments, batch-size arithmetic, dataset ratios and
loader methods, on top of `ren_b200.events.Event`, `ren_b200.posed_images.PosedImage` and the on-device batch
producer `ren_b200.data.DeviceEventSampler` (events resident in HBM, no host work per training batch).

Without Lightning (absent in this image) it is a plain object: `setup()`, then `train_dataloader()` yields
`{"event": ..., "normalized": ...}` batches forever (what Lightning's combined loader hands `training_step`),
`val_dataloader()` / `test_dataloader()` are ordinary DataLoaders over the posed images.  `RobustENeRF.
update_train_batch_size` resizes `train_dataset.batch_size` and the three `train_normalized_sampler.datasets[i].size`
exactly as with the reference (robust_e_nerf.py:940-948); all four are views of the sampler's one batch size."""
import torch

from .data import DeviceEventSampler
from .events import Event
from .posed_images import PosedImage


class _SizeView:
    """`.size` of one normalised-timestamp sampler == the device sampler's batch size."""

    def __init__(self, sampler):
        self._sampler = sampler

    @property
    def size(self):
        return self._sampler.batch_size

    @size.setter
    def size(self, value):
        self._sampler.batch_size = int(value)


class _NormalizedSamplers:
    def __init__(self, sampler):
        self.datasets = tuple(_SizeView(sampler) for _ in range(3))       # ts_diff, diff_start_ts, grad_ts
        self.dataset_keys = ("ts_diff", "diff_start_ts", "grad_ts")


class DataModule:
    def __init__(self, seed, eval_target, num_nodes, gpus, dataset_directory, train_dataset_ratio, val_dataset_ratio,
                 test_dataset_ratio, train_dataset_perm_seed, eval_dataset_perm_seed, alpha_over_white_bg,
                 train_init_eff_batch_size, train_eff_ray_sample_batch_size, val_eff_batch_size,
                 test_eff_batch_size, num_workers_per_node, device=None):
        for r in (train_dataset_ratio, val_dataset_ratio, test_dataset_ratio):
            assert isinstance(r, int) or (isinstance(r, float) and 0.0 < r <= 1.0)
        assert num_workers_per_node == 0           # dynamic batch size needs main-process loading (datamodule.py:36-41)
        self.hparams = {"seed": seed, "train_init_eff_batch_size": train_init_eff_batch_size,
                        "train_eff_ray_sample_batch_size": train_eff_ray_sample_batch_size}
        self.eval_target = eval_target
        self.dataset_directory = dataset_directory
        self.ratio = {"train": train_dataset_ratio, "val": val_dataset_ratio, "test": test_dataset_ratio}
        self.eff_batch_size = {"train": train_init_eff_batch_size, "val": val_eff_batch_size,
                               "test": test_eff_batch_size}
        self.train_dataset_perm_seed = train_dataset_perm_seed
        self.eval_dataset_perm_seed = eval_dataset_perm_seed
        self.alpha_over_white_bg = alpha_over_white_bg
        n_gpus = 1 if gpus is None else num_nodes * len(gpus)
        self.train_batch_size = train_init_eff_batch_size // n_gpus
        self.val_batch_size = val_eff_batch_size // n_gpus
        self.test_batch_size = test_eff_batch_size // n_gpus
        self.device = torch.device(device if device is not None else ("cpu" if gpus is None else "cuda"))

    def _subset_len(self, stage, n):
        r = self.ratio[stage]
        if isinstance(r, int):                      # a number of effective batches
            k = r * self.eff_batch_size[stage]
            assert k <= n
            return k
        return int(r * n)

    def _posed(self, stage):
        if set(self.eval_target) == {"event_view"}:
            src = "train"
        elif set(self.eval_target) == {"novel_view"}:
            src = stage
        else:
            raise NotImplementedError(self.eval_target)
        ds = PosedImage(self.dataset_directory, src, self.eval_dataset_perm_seed, self.alpha_over_white_bg)
        return torch.utils.data.Subset(ds, range(self._subset_len(stage, len(ds))))

    def setup(self, stage=None):
        if stage in (None, "fit"):
            rank = torch.distributed.get_rank() if torch.distributed.is_available() and \
                torch.distributed.is_initialized() else 0
            ev = Event(self.dataset_directory, self.train_dataset_perm_seed, device=self.device)
            k = self._subset_len("train", len(ev))
            events = {key: v[:k] for key, v in ev.events.items()}
            # per-rank stream: initial seed + rank (datamodule.py:85-89)
            self.train_dataset = DeviceEventSampler(events, self.train_batch_size, device=self.device,
                                                    seed=torch.initial_seed() % (2 ** 62), rank=rank)
            self.train_normalized_sampler = _NormalizedSamplers(self.train_dataset)
            self.val_dataset = self._posed("val")
        if stage in (None, "validate"):
            self.val_dataset = self._posed("val")
        if stage in (None, "test"):
            self.test_dataset = self._posed("test")

    def train_dataloader(self):
        return self.train_dataset                   # an infinite iterator of ready (1, B, ...) device batches

    def _eval_loader(self, dataset, batch_size):
        return torch.utils.data.DataLoader(dataset, batch_size=batch_size, num_workers=0, shuffle=False,
                                           pin_memory=self.device.type == "cuda", drop_last=False)

    def val_dataloader(self):
        return self._eval_loader(self.val_dataset, self.val_batch_size)

    def test_dataloader(self):
        return self._eval_loader(self.test_dataset, self.test_batch_size)
