"""Synthetic 640x480 event-stream workload for `bench.py` / smoke runs:
a dataset directory in the reference's on-disk schema
(robust_e_nerf/data/datasets.py:15-34 `camera_calibration.npz`, :693-727
`camera_poses.npz`, :116-130 `max_refractory_period.pt`) and event batches shaped like
`DataModule.train_dataloader()` output (robust_e_nerf/data/datamodule.py:198-230:
`{"event": {...(1,B,...)}, "normalized": {...(1,B)}}`).  There is no network for real
datasets, so the benchmark says `"data": "synthetic"`.
"""
import math
import os

import numpy as np
import torch

IMG_W, IMG_H = 640, 480
NUM_POSES = 2001
POSE_DT_NS = 1_000_000


def _look_at_quat_xyzw(p, center):
    z = center - p
    z = z / np.linalg.norm(z)
    x = np.cross(z, np.array([0.0, 0.0, 1.0]))
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z], axis=1)
    # rotation matrix -> quaternion (w largest branch is enough for an orbit, but stay general)
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = [(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s]
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = [0.0, 0.0, 0.0, (R[k, j] - R[j, k]) / s]
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
    q = np.asarray(q)
    return q / np.linalg.norm(q)


def write_dataset(root, center=(0.0, 0.0, 0.0), scale=1.0, contrast=0.25, refractory_ns=1000,
                  max_refractory_ns=50_000, num_poses=NUM_POSES):
    """1 kHz orbit of radius 3*scale about `center`, camera z-axis toward the centre."""
    os.makedirs(root, exist_ok=True)
    K = np.array([[500.0, 0.0, 319.5], [0.0, 500.0, 239.5], [0.0, 0.0, 1.0]], dtype=np.float32)
    np.savez(os.path.join(root, "camera_calibration.npz"), intrinsics=K,
             distortion_model=np.array("plumb_bob"), distortion_params=np.zeros(5),
             img_height=np.array(IMG_H, dtype=np.uint16), img_width=np.array(IMG_W, dtype=np.uint16),
             pos_contrast_threshold=np.array(contrast, dtype=np.float32),
             neg_contrast_threshold=np.array(contrast, dtype=np.float32),
             refractory_period=np.array(refractory_ns), bayer_pattern=np.array(""))
    center = np.asarray(center, dtype=np.float64)
    phi = np.linspace(0.0, 2.0 * math.pi, num_poses)
    pos = np.stack([3.0 * np.cos(phi), 3.0 * np.sin(phi), 0.2 * np.sin(3.0 * phi)], axis=1) * scale + center
    quats, prev = [], None
    for p in pos:
        q = _look_at_quat_xyzw(p, center)
        if prev is not None and np.dot(prev, q) < 0:
            q = -q
        prev = q
        quats.append(q)
    np.savez(os.path.join(root, "camera_poses.npz"), T_wc_position=pos.astype(np.float32),
             T_wc_orientation=np.stack(quats).astype(np.float32),
             T_wc_timestamp=np.arange(num_poses, dtype=np.int64) * POSE_DT_NS)
    torch.save(torch.tensor(max_refractory_ns, dtype=torch.int64),
               os.path.join(root, "max_refractory_period.pt"))
    return root


def _trunc_normal(low, high, size, mean, std, generator):
    """Inverse-CDF truncated normal (robust_e_nerf/data/samplers.py:33-85)."""
    cdf = lambda x: (1.0 + math.erf(x / math.sqrt(2.0))) / 2.0
    l, u = cdf((low - mean) / std), cdf((high - mean) / std)
    t = 2 * (u - l) * torch.rand(size, dtype=torch.float64, generator=generator) + (2 * l - 1)
    t.erfinv_().mul_(std * math.sqrt(2.0)).add_(mean).clamp_(min=low, max=high)
    return t


def make_batch(batch_size, pose_ts, rank=0, seed=1234, pin=False):
    """One host-side training batch + per-ray stratified jitter (CPU tensors; `pin=True`
    page-locks them for the end-to-end benchmark leg).  Random-with-replacement pixels,
    event intervals of 0.2-3 ms, one event per interval, per-rank seed = seed + rank
    (robust_e_nerf/data/datamodule.py:85-89,149-196; robust_e_nerf/utils/datasets.py:20-32)."""
    g = torch.Generator()
    g.manual_seed(seed + rank)
    B = batch_size
    x = torch.randint(0, IMG_W, (B,), generator=g).to(torch.float32)
    y = torch.randint(0, IMG_H, (B,), generator=g).to(torch.float32)
    lo, hi = int(pose_ts[5]), int(pose_ts[-5])
    end_ts = torch.randint(lo + 3_000_001, hi, (B,), generator=g, dtype=torch.int64)
    dur = torch.randint(200_000, 3_000_001, (B,), generator=g, dtype=torch.int64)
    pol = torch.rand((B,), generator=g) < 0.5
    event = {"position": torch.stack([x, y], dim=-1)[None], "start_ts": (end_ts - dur)[None],
             "end_ts": end_ts[None], "num_pos": pol.to(torch.int64)[None],
             "num_neg": (~pol).to(torch.int64)[None]}
    normalized = {"ts_diff": torch.full((1, B), 1.0, dtype=torch.float64),
                  "diff_start_ts": torch.rand((B,), dtype=torch.float64, generator=g)[None],
                  "grad_ts": _trunc_normal(0, 1, (B,), 0.5, 0.25, g)[None]}
    jitter = {n: torch.rand((B,), dtype=torch.float32, generator=g)
              for n in ("grad", "diff_start", "diff_end")}
    batch = {"event": event, "normalized": normalized}
    if pin:
        batch = {k: {kk: vv.contiguous().pin_memory() for kk, vv in v.items()} for k, v in batch.items()}
        jitter = {k: v.pin_memory() for k, v in jitter.items()}
    return batch, jitter


def to_device(batch, device, non_blocking=True):
    return {k: {kk: vv.to(device, non_blocking=non_blocking) for kk, vv in v.items()}
            for k, v in batch.items()}


def batch_nbytes(batch):
    return sum(v.numel() * v.element_size() for d in batch.values() for v in d.values())


def default_config(net_depth=8, net_width=256, n_samples=128, learn_bkgd=False, w_diff=1.0,
                   aabb=(-1.5, -1.5, -1.5, 1.5, 1.5, 1.5), near=0.5, span=4.0):
    """Hyper-parameters of configs/train/synthetic.yaml:24-128 with `arch: mlp`, in the
    fixed-N regime of SURVEY.md section 8(d): exactly `n_samples` per ray."""
    step = span / n_samples
    nerf = dict(
        aabb=list(aabb), contraction_type="sphere",
        occ_grid=dict(resolution=128, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16, enabled=False),
        near_plane=near, far_plane=near + (n_samples + 0.5) * step, render_step_size=step,
        cone_angle=0.0, early_stop_eps=0.0, alpha_thre=0.0, test_chunk_size=16384, arch="mlp",
        load_state_dict=False, freeze=False,
        mlp=dict(net_depth=net_depth, net_width=net_width, skip_layer=4, net_depth_condition=1,
                 net_width_condition=128, hidden_activation="softplus",
                 density_activation="shifted_trunc_exp", radiance_activation="softplus",
                 pos_encoder_max_deg=10, view_encoder_max_deg=4, weight_norm=False))
    return dict(
        git_head_hash="synthetic", eval_target=["novel_view"], num_nodes=1, gpus=[0],
        min_modeled_intensity=1e-3, eval_save_pred_intensity_img=False, checkpoint_filepath=None,
        contrast_threshold=dict(load_state_dict=False, freeze=True),
        refractory_period=dict(load_state_dict=False, freeze=True), nerf=nerf,
        loss=dict(error_fn=dict(log_intensity_grad="mape", log_intensity_diff="mse"),
                  weight=dict(log_intensity_grad=0.0, log_intensity_diff=w_diff,
                              nerf_mlp_weight_decay=1e-6),
                  param_weight=dict(log_intensity_grad=None,
                                    log_intensity_diff="mean_contrast_reciprocal_sq")),
        metric=dict(lpips_net="alex"),
        optimizer=dict(algo="adam", lr=dict(contrast_threshold=0.1, default=0.01),
                       relative_lr=dict(refractory_period=50)),
        lr_scheduler=dict(algo="multi_step_lr", interval="epoch",
                          multi_step_lr=dict(milestones=[20, 30, 36], gamma=0.33)),
        alpha_over_white_bg=learn_bkgd, train_eff_ray_sample_batch_size=1 << 20)
