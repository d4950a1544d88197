"""ORACLE (test infrastructure): synthetic 640x480 event-stream scene + batches.

Writes a dataset directory in the reference's on-disk schema
(robust_e_nerf/data/datasets.py:15-34 `camera_calibration.npz`, :693-724
`camera_poses.npz`, :376-395 `views/transforms_{stage}.json`,
:116-130 `max_refractory_period.pt`) so that the *same directory* feeds the
unmodified reference (oracle.ref_loader), the restatement and the CUDA path, and
draws event batches the way `DataModule` + samplers do
(robust_e_nerf/data/datamodule.py:149-230, robust_e_nerf/data/samplers.py:6-118,
robust_e_nerf/utils/datasets.py:20-32) -- SURVEY.md section 8(d).

Only numpy / torch / json are used (no reference import), so it also runs on the
GPU box.
"""
import json
import math
import os

import numpy as np
import torch

IMG_W, IMG_H = 640, 480
NUM_POSES = 2001
POSE_DT_NS = 1_000_000  # 1 kHz


def _rotmat_to_quat_xyzw(R):
    """Shepperd's method, float64, returns XYZW."""
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        w = 0.25 * s
        x = (R[2, 1] - R[1, 2]) / s
        y = (R[0, 2] - R[2, 0]) / s
        z = (R[1, 0] - R[0, 1]) / s
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        w = (R[2, 1] - R[1, 2]) / s
        x = 0.25 * s
        y = (R[0, 1] + R[1, 0]) / s
        z = (R[0, 2] + R[2, 0]) / s
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        w = (R[0, 2] - R[2, 0]) / s
        x = (R[0, 1] + R[1, 0]) / s
        y = 0.25 * s
        z = (R[1, 2] + R[2, 1]) / s
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        w = (R[1, 0] - R[0, 1]) / s
        x = (R[0, 2] + R[2, 0]) / s
        y = (R[1, 2] + R[2, 1]) / s
        z = 0.25 * s
    return np.array([x, y, z, w])


def make_trajectory(num_poses=NUM_POSES, radius=3.0, center=(0.0, 0.0, 0.0), scale=1.0):
    """Orbit of `radius*scale` about `center`, z = 0.2 sin(3 phi), camera z-axis
    (common convention: x right, y down, z forward) looking at the centre."""
    center = np.asarray(center, dtype=np.float64)
    phi = np.linspace(0.0, 2.0 * math.pi, num_poses)
    pos = np.stack(
        [radius * np.cos(phi), radius * np.sin(phi), 0.2 * np.sin(3.0 * phi)], axis=1
    ) * scale + center
    quats = []
    prev = None
    for p in pos:
        zc = center - p
        zc = zc / np.linalg.norm(zc)
        up = np.array([0.0, 0.0, 1.0])
        xc = np.cross(zc, up)
        xc = xc / np.linalg.norm(xc)
        yc = np.cross(zc, xc)
        R = np.stack([xc, yc, zc], axis=1)
        q = _rotmat_to_quat_xyzw(R)
        q = q / np.linalg.norm(q)
        if prev is not None and np.dot(prev, q) < 0:
            q = -q
        prev = q
        quats.append(q)
    quats = np.stack(quats)
    # negate every 7th quaternion: exercises the shortest-path flip
    # (robust_e_nerf/utils/tensor_ops.py:157-159)
    quats[::7] *= -1.0
    ts = np.arange(num_poses, dtype=np.int64) * POSE_DT_NS
    return pos.astype(np.float32), quats.astype(np.float32), ts


def write_dataset(root, center=(0.0, 0.0, 0.0), scale=1.0, contrast=0.25,
                  refractory_ns=1000, max_refractory_ns=50_000):
    """Create the synthetic dataset directory (idempotent)."""
    os.makedirs(os.path.join(root, "views"), exist_ok=True)
    K = np.array([[500.0, 0.0, 319.5], [0.0, 500.0, 239.5], [0.0, 0.0, 1.0]],
                 dtype=np.float32)  # scripts/preprocess_esim.py:285 (float32)
    np.savez(
        os.path.join(root, "camera_calibration.npz"),
        intrinsics=K,
        distortion_model=np.array("plumb_bob"),
        distortion_params=np.zeros(5),
        img_height=np.array(IMG_H, dtype=np.uint16),
        img_width=np.array(IMG_W, dtype=np.uint16),
        pos_contrast_threshold=np.array(contrast, dtype=np.float32),
        neg_contrast_threshold=np.array(contrast, dtype=np.float32),
        refractory_period=np.array(refractory_ns),  # preprocess_esim.py:153 (int64)
        bayer_pattern=np.array(""),
    )
    pos, quat, ts = make_trajectory(center=center, scale=scale)
    np.savez(
        os.path.join(root, "camera_poses.npz"),
        T_wc_position=pos, T_wc_orientation=quat, T_wc_timestamp=ts,
    )
    torch.save(torch.tensor(max_refractory_ns, dtype=torch.int64),  # datasets.py:185
               os.path.join(root, "max_refractory_period.pt"))
    # one tiny posed view per stage: the reference ctor loads them only to learn
    # H, W, intrinsics (robust_e_nerf/models/robust_e_nerf.py:86-152)
    try:
        import cv2
        img = (np.arange(8 * 8, dtype=np.uint8).reshape(8, 8) * 3 + 10)
        cv2.imwrite(os.path.join(root, "views", "r_0.png"), img)
    except ImportError:  # pragma: no cover - the restatement never reads the views
        pass
    T = np.eye(4)
    T[:3, 3] = pos[0]
    for stage in ("train", "val", "test"):
        with open(os.path.join(root, "views", f"transforms_{stage}.json"), "w") as f:
            json.dump({"intrinsics": K.tolist(),
                       "frames": [{"file_path": "r_0", "transform_matrix": T.tolist()}]}, f)
    return root


def load_scene(root):
    """Reads back what the path needs (same arrays the reference loads)."""
    calib = np.load(os.path.join(root, "camera_calibration.npz"))
    poses = np.load(os.path.join(root, "camera_poses.npz"))
    return {
        "intrinsics": torch.from_numpy(calib["intrinsics"]),                 # f64 (3,3)
        "pos_contrast_threshold": torch.from_numpy(calib["pos_contrast_threshold"]),
        "neg_contrast_threshold": torch.from_numpy(calib["neg_contrast_threshold"]),
        "refractory_period": torch.from_numpy(calib["refractory_period"]),
        "max_refractory_period": torch.load(os.path.join(root, "max_refractory_period.pt")),
        "T_wc_position": torch.tensor(poses["T_wc_position"]),
        "T_wc_orientation": torch.tensor(poses["T_wc_orientation"]),
        "T_wc_timestamp": torch.tensor(poses["T_wc_timestamp"]),
    }


def trunc_normal(low, high, size, mean, std, dtype, generator):
    """Inverse-CDF truncated normal, restating
    robust_e_nerf/data/samplers.py:33-85."""
    def norm_cdf(x):
        return (1.0 + math.erf(x / math.sqrt(2.0))) / 2.0

    l = norm_cdf((low - mean) / std)
    u = norm_cdf((high - mean) / std)
    t = 2 * (u - l) * torch.rand(size, dtype=dtype, generator=generator) + (2 * l - 1)
    t.erfinv_()
    t.mul_(std * math.sqrt(2.0))
    t.add_(mean)
    t.clamp_(min=low, max=high)
    return t


def make_batch(batch_size, scene_ts, rank=0, seed=1234):
    """One training batch exactly as `training_step` receives it (leading dim 1
    from DataLoader(batch_size=1), robust_e_nerf/data/datamodule.py:205-222) plus
    the per-ray stratified jitter `u` that is passed explicitly to both arms."""
    g = torch.Generator()
    g.manual_seed(seed + rank)
    B = batch_size
    x = torch.randint(0, IMG_W, (B,), generator=g).to(torch.float32)
    y = torch.randint(0, IMG_H, (B,), generator=g).to(torch.float32)
    lo, hi = int(scene_ts[5]), int(scene_ts[-5])
    end_ts = torch.randint(lo + 3_000_001, hi, (B,), generator=g, dtype=torch.int64)
    dur = torch.randint(200_000, 3_000_001, (B,), generator=g, dtype=torch.int64)
    start_ts = end_ts - dur
    pol = torch.rand((B,), generator=g) < 0.5
    event = {
        "position": torch.stack([x, y], dim=-1)[None],
        "start_ts": start_ts[None],
        "end_ts": end_ts[None],
        "num_pos": pol.to(torch.int64)[None],
        "num_neg": (~pol).to(torch.int64)[None],
    }
    normalized = {
        "ts_diff": torch.full((1, B), 1.0, dtype=torch.float64),
        "diff_start_ts": torch.rand((B,), dtype=torch.float64, generator=g)[None],
        "grad_ts": trunc_normal(0, 1, (B,), 0.5, 0.25, torch.float64, g)[None],
    }
    jitter = {
        name: torch.rand((B,), dtype=torch.float32, generator=g)
        for name in ("grad", "diff_start", "diff_end")
    }
    return {"event": event, "normalized": normalized}, jitter
