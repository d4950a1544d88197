"""The whole drop-in pipeline on the GPU, wired the way `scripts/run.py:38-111` wires the reference:
raw events -> `Event` (conversion on the device, `events.pt` cache) -> `DataModule` -> `RobustENeRF` training steps with
the dynamic batch size -> validation renders of the posed views -> epoch-end alignment and metrics."""
import os
import shutil
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import helpers as H
from oracle import events_prep as EP
from oracle import synth
from test_posed_images import GOLDEN as POSED_GOLDEN, _write as write_views

pytestmark = pytest.mark.gpu


def test_raw_events_to_validation_metrics(tmp_path, built_lib):
    from ren_b200.datamodule import DataModule
    from ren_b200.events import Event
    from ren_b200.trainer import FlatTrainer
    root = synth.write_dataset(str(tmp_path / "seq"))
    scene = synth.load_scene(root)
    T = scene["T_wc_timestamp"]
    raw = EP.synth_raw_events(300_000, 640, 480, seed=21)
    span = int(T[-6] - T[5])
    raw["timestamp"] = (int(T[5]) + (raw["timestamp"] - raw["timestamp"][0]) % span).astype(np.int64)
    raw["timestamp"].sort()
    np.savez(os.path.join(root, "raw_events.npz"), **raw)
    calib = dict(np.load(os.path.join(root, "camera_calibration.npz")))
    calib["distortion_params"] = np.zeros(0)                           # datasets.py:336 accepts 0 or 4 parameters
    np.savez(os.path.join(root, "camera_calibration.npz"), **calib)
    shutil.rmtree(os.path.join(root, "views"))
    src = tmp_path / "views_src"
    src.mkdir()
    write_views(str(src), "real_mono8", np.load(POSED_GOLDEN))         # 3 validation views of 6 x 8 pixels
    shutil.copytree(src / "views", os.path.join(root, "views"))

    ev = Event(root, None)                                             # builds + caches events.pt on the GPU
    # the sensor model needs every interval to exceed the refractory period (event_generation_params.py:117-131);
    # this synthetic stream is denser than the calibrated one, keep the intervals that respect it
    keep = (ev.events["end_ts"] - ev.events["start_ts"]) > 100_000
    torch.save({k: v[keep] for k, v in ev.events.items()}, os.path.join(root, "events.pt"))

    g = dict(H.load_golden("step_8x256"))
    g["rcfg"] = dict(g["rcfg"], render_step_size=4.0 / 32, far_plane=0.5 + 32.5 * (4.0 / 32))
    cfg = H.product_config(g)
    budget = 1 << 14
    cfg["train_eff_ray_sample_batch_size"] = budget
    # random synthetic events are not a consistent scene: a small step keeps the four updates well-behaved
    cfg["optimizer"] = dict(cfg["optimizer"], lr=dict(contrast_threshold=1e-3, default=1e-4))
    torch.manual_seed(0)                                               # the data module seeds its sampler from it
    dm = DataModule(seed=0, eval_target=["novel_view"], num_nodes=1, gpus=[0], dataset_directory=root,
                    train_dataset_ratio=1.0, val_dataset_ratio=1.0, test_dataset_ratio=1.0,
                    train_dataset_perm_seed=None, eval_dataset_perm_seed=None, alpha_over_white_bg=False,
                    train_init_eff_batch_size=256, train_eff_ray_sample_batch_size=budget, val_eff_batch_size=1,
                    test_eff_batch_size=1, num_workers_per_node=0)
    dm.setup("fit")
    from ren_b200 import RobustENeRF
    model = RobustENeRF(dataset_directory=root, **cfg).cuda().train()
    assert model.val_img_pixel_pos.shape == (6, 8, 2)                  # sized from the posed views
    model.trainer = NS(accumulate_grad_batches=1, datamodule=dm)
    tr = FlatTrainer(model)
    sizes, losses = [], []
    for i, batch in zip(range(4), dm.train_dataloader()):
        sizes.append(batch["event"]["end_ts"].shape[1])
        losses.append(float(tr.step(batch, i).detach()))
    assert all(np.isfinite(losses)), losses
    spp = float(model.logged["train/mean_num_samples_per_ray"])
    assert sizes[0] == 256 and sizes[-1] != 256                        # the batch was resized while training
    assert dm.train_dataset.batch_size == int(budget / spp)            # robust_e_nerf.py:940-948
    # validation epoch over the posed views
    model.eval()
    outs = []
    for i, batch in enumerate(dm.val_dataloader()):
        outs.append(model.validation_step({k: v.cuda() for k, v in batch.items()}, i))
    assert len(outs) == 3 and outs[0]["pred_intensity_img"].shape == (6, 8)
    res = model.validation_epoch_end(outs)
    assert np.isfinite(float(res["psnr"])) and np.isfinite(float(res["l1"]))
    assert res["pred_intensity_img"].shape == (3, 1, 6, 8)
