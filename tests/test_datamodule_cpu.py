"""`ren_b200.datamodule.DataModule` (robust_e_nerf/data/datamodule.py:8-272) on the CPU: batch-size arithmetic,
dataset ratios, batch schema, the dynamic batch-size hook, evaluation loaders."""
import os
import shutil
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import helpers as H
from oracle import synth
from test_posed_images import GOLDEN as POSED_GOLDEN, _write as write_views


def _dataset(tmp_path, n_events=5000):
    root = synth.write_dataset(str(tmp_path / "seq"))
    gen = torch.Generator().manual_seed(0)
    start = torch.randint(0, 10 ** 9, (n_events,), generator=gen)
    torch.save({"position": torch.rand(n_events, 2, generator=gen) * 100,
                "start_ts": start, "end_ts": start + torch.randint(1, 10 ** 6, (n_events,), generator=gen),
                "num_pos": torch.randint(0, 2, (n_events,), generator=gen),
                "num_neg": torch.randint(0, 2, (n_events,), generator=gen)}, os.path.join(root, "events.pt"))
    shutil.rmtree(os.path.join(root, "views"))
    src = tmp_path / "views_src"
    src.mkdir()
    write_views(str(src), "real_mono8", np.load(POSED_GOLDEN))            # stage "val", 3 views of 6 x 8
    shutil.copytree(src / "views", os.path.join(root, "views"))
    return root


def _dm(root, **kw):
    from ren_b200.datamodule import DataModule
    args = dict(seed=1, eval_target=["novel_view"], num_nodes=1, gpus=[0, 1], dataset_directory=root,
                train_dataset_ratio=0.5, val_dataset_ratio=1.0, test_dataset_ratio=1.0, train_dataset_perm_seed=4,
                eval_dataset_perm_seed=None, alpha_over_white_bg=False, train_init_eff_batch_size=256,
                train_eff_ray_sample_batch_size=1 << 16, val_eff_batch_size=2, test_eff_batch_size=2,
                num_workers_per_node=0, device="cpu")
    args.update(kw)
    return DataModule(**args)


def test_train_batches_and_dynamic_batch_size(tmp_path, built_lib):
    root = _dataset(tmp_path)
    torch.manual_seed(7)
    dm = _dm(root)
    assert dm.train_batch_size == 128 and dm.val_batch_size == 1       # effective sizes split over 2 GPUs
    dm.setup("fit")
    assert dm.train_dataset.n == 2500                                  # train_dataset_ratio 0.5 of 5000 events
    batch = next(iter(dm.train_dataloader()))
    assert set(batch) == {"event", "normalized"}
    assert set(batch["event"]) == {"position", "start_ts", "end_ts", "num_pos", "num_neg"}
    assert batch["event"]["position"].shape == (1, 128, 2) and batch["event"]["position"].dtype == torch.float32
    assert set(batch["normalized"]) == {"ts_diff", "diff_start_ts", "grad_ts"}
    for v in batch["normalized"].values():
        assert v.shape == (1, 128) and v.dtype == torch.float64 and float(v.min()) >= 0 and float(v.max()) <= 1
    # the permuted, trimmed subset is what is sampled from
    from ren_b200.events import Event
    ev = Event(root, 4, device="cpu")
    allowed = set(ev.events["end_ts"][:2500].tolist())
    assert set(batch["event"]["end_ts"].reshape(-1).tolist()) <= allowed
    # robust_e_nerf.py:940-948 through the real hook
    m = H.build_product(H.load_golden("step_plumbing"), root)
    m.trainer = NS(accumulate_grad_batches=1, datamodule=dm)
    m.update_train_batch_size(torch.tensor(64.0), 0)
    want = int(m.train_ray_sample_batch_size / 64.0)
    assert dm.train_dataset.batch_size == want and all(s.size == want for s in dm.train_normalized_sampler.datasets)
    assert next(iter(dm.train_dataloader()))["event"]["end_ts"].shape == (1, want)
    # an integer ratio counts effective batches
    dm2 = _dm(root, train_dataset_ratio=3)
    dm2.setup("fit")
    assert dm2.train_dataset.n == 3 * 256
    with pytest.raises(AssertionError):
        _dm(root, train_dataset_ratio=100).setup("fit")                # more than the sequence holds


def test_eval_loaders(tmp_path, built_lib):
    root = _dataset(tmp_path)
    dm = _dm(root, val_dataset_ratio=0.67, gpus=None)
    dm.setup("validate")
    assert len(dm.val_dataset) == 2                                    # int(0.67 * 3)
    batches = list(dm.val_dataloader())
    assert len(batches) == 1 and batches[0]["img"].shape == (2, 6, 8)
    assert batches[0]["T_wc_orientation"].shape == (2, 3, 3) and batches[0]["sample_id"].shape == (2, 16)
    with pytest.raises(FileNotFoundError):
        dm.setup("test")                                               # no transforms_test.json in this sequence
    dm_ev = _dm(root, eval_target=["event_view"])
    with pytest.raises(FileNotFoundError):
        dm_ev.setup("validate")                                        # event_view evaluates on transforms_train.json
