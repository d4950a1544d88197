"""Event dataset preparation on the GPU -- the host-side mirror of the reference's
`robust_e_nerf.data.datasets.Event` class methods for the raw-event conversion
(robust_e_nerf/data/datasets.py:132-328), backed by `ren_event_queue` (one stable radix sort + scan
instead of two per-event Python loops).  SURVEY.md section 8(f) #3.

    events = queue_raw_events(root_directory, camera_calibration)          # == Event.queue_raw_events
    tau_max = extract_max_refractory_period(raw_events, camera_calibration)
    events = colorize_events(events, camera_calibration)
    events = undistort_events(events, camera_calibration)                  # == Event.undistort_events

File names / keys are the reference's (`raw_events.npz` {position (N,2) u16, timestamp (N,) i64 ns,
polarity (N,) bool}, `camera_calibration.npz`, `events.pt`, `max_refractory_period.pt`).
"""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib

RAW_EVENTS_FILENAME = "raw_events.npz"                       # datasets.py:15-27
TF_EVENTS_FILENAME = "events.pt"
MAX_REFRACTORY_PERIOD_FILENAME = "max_refractory_period.pt"
COLOR_CHANNEL_NAME_TO_INDEX = {"R": 0, "G": 1, "B": 2}


def load_raw_events(root_directory):
    """datasets.py:97-102"""
    return dict(np.load(os.path.join(root_directory, RAW_EVENTS_FILENAME)))


def _run(raw_events, camera_calibration, device):
    pos = np.ascontiguousarray(raw_events["position"]).astype(np.uint16, copy=False)
    ts = np.ascontiguousarray(raw_events["timestamp"]).astype(np.int64, copy=False)
    pol = np.ascontiguousarray(raw_events["polarity"]).astype(np.uint8)
    assert len(pos) == len(ts) == len(pol)                   # datasets.py:143-144
    n = len(ts)
    h, w = int(camera_calibration["img_height"]), int(camera_calibration["img_width"])
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("ren_b200 prepares events on a CUDA device only (no CPU fallback)")
    # uint16 has no torch dtype on every version: ship the bytes
    d_pos = torch.from_numpy(pos.view(np.uint8).reshape(-1)).to(dev)
    d_ts = torch.from_numpy(ts).to(dev)
    d_pol = torch.from_numpy(pol).to(dev)
    out = {k: torch.empty((n,), dtype=torch.int64, device=dev) for k in ("start_ts", "end_ts", "num_pos", "num_neg")}
    out_pos = torch.empty((n, 2), dtype=torch.int64, device=dev)
    n_valid = torch.zeros((1,), dtype=torch.int64, device=dev)
    min_int = torch.zeros((1,), dtype=torch.int64, device=dev)
    lib = _lib.load()
    nb = C.c_size_t(0)
    _lib.check(lib.ren_event_queue_workspace_bytes(n, w, h, C.byref(nb)), "ren_event_queue_workspace_bytes")
    ws = torch.empty((int(nb.value),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ren_event_queue(
            n, w, h, _lib.ptr(d_pos), _lib.ptr(d_ts), _lib.ptr(d_pol), _lib.ptr(out_pos), _lib.ptr(out["start_ts"]),
            _lib.ptr(out["end_ts"]), _lib.ptr(out["num_pos"]), _lib.ptr(out["num_neg"]), _lib.ptr(n_valid),
            _lib.ptr(min_int), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "ren_event_queue")
    nv = int(n_valid.item())
    events = {"position": out_pos[:nv]}
    events.update({k: v[:nv] for k, v in out.items()})
    return events, int(min_int.item())


def queue_raw_events(root_directory, camera_calibration, device="cuda"):
    """`Event.queue_raw_events` (datasets.py:189-284): dict of int64 tensors position (M,2), start_ts,
    end_ts, num_pos, num_neg (M,) for the M events that have a same-pixel predecessor at a different
    timestamp, in stream order."""
    events, _ = _run(load_raw_events(root_directory), camera_calibration, device)
    return events


def extract_max_refractory_period(raw_events, camera_calibration, device="cuda"):
    """`Event.extract_max_refractory_period` (datasets.py:132-187): the smallest positive interval
    between consecutive events of any pixel -- an int64 scalar tensor, or float64 inf when no pixel fired
    twice (the reference's `np.array(float("inf"))` initial value)."""
    _, m = _run(raw_events, camera_calibration, device)
    if m == np.iinfo(np.int64).max:
        return torch.tensor(float("inf"), dtype=torch.float64)
    return torch.tensor(m, dtype=torch.int64)


def colorize_events(events, camera_calibration):
    """`Event.colorize_events` (datasets.py:286-328): channel index of every event from the 2x2 Bayer
    pattern string (row-major), nothing for a monochrome sensor."""
    pattern = str(camera_calibration["bayer_pattern"])
    if pattern == "":
        return events
    assert len(pattern) == 4 and set(pattern) == set(COLOR_CHANNEL_NAME_TO_INDEX)
    table = torch.tensor([COLOR_CHANNEL_NAME_TO_INDEX[c] for c in pattern], dtype=torch.uint8,
                         device=events["position"].device)
    x, y = events["position"][:, 0], events["position"][:, 1]
    events = dict(events)
    events["channel_idx"] = table[(y % 2) * 2 + (x % 2)]
    return events


def undistort_events(events, camera_calibration):
    """`Event.undistort_events` (datasets.py:330-364): event pixel positions -> the ideal pinhole image with the
    same intrinsics.  The reference calls `cv2.undistortPoints(..., P=K)` (plumb_bob: k1 k2 p1 p2) or
    `cv2.fisheye.undistortPoints(..., P=K)` (equidistant: k1..k4) on the host; here the same fixed-point /
    Newton iterations (OpenCV's published algorithms, float64 internally, float32 result) run as elementwise
    tensor ops on whatever device holds the events.  Checked against OpenCV itself
    (`tests/test_events_prep.py::test_undistort_events_matches_opencv`)."""
    model = str(camera_calibration["distortion_model"])
    params = np.asarray(camera_calibration["distortion_params"], dtype=np.float64).reshape(-1)
    K = np.asarray(camera_calibration["intrinsics"], dtype=np.float64)
    assert len(params) in (0, 4)
    events = dict(events)
    events["position"] = events["position"].to(torch.get_default_dtype())
    if len(params) == 0:
        return events
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    p = events["position"].to(torch.float64)
    x0, y0 = (p[:, 0] - cx) / fx, (p[:, 1] - cy) / fy
    if model == "plumb_bob":
        k1, k2, p1, p2 = params
        x, y = x0.clone(), y0.clone()
        for _ in range(5):                                   # cv::undistortPoints default: 5 fixed-point sweeps
            r2 = x * x + y * y
            icdist = 1.0 / (1.0 + (k2 * r2 + k1) * r2)
            # a negative inverse distortion factor means the sweep left the model's valid range: keep the
            # undistorted-as-is normalised point (OpenCV does the same)
            bad = icdist < 0
            dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
            dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
            x = torch.where(bad, x0, (x0 - dx) * icdist)
            y = torch.where(bad, y0, (y0 - dy) * icdist)
    elif model == "equidistant":
        k1, k2, k3, k4 = params
        theta_d = torch.sqrt(x0 * x0 + y0 * y0).clamp(-math.pi / 2, math.pi / 2)
        theta = theta_d.clone()
        active = theta_d > 1e-8
        for _ in range(10):                                  # cv::fisheye::undistortPoints: Newton, <= 10 steps
            t2 = theta * theta
            t4, t6, t8 = t2 * t2, t2 * t2 * t2, t2 * t2 * t2 * t2
            fix = (theta * (1 + k1 * t2 + k2 * t4 + k3 * t6 + k4 * t8) - theta_d) / \
                  (1 + 3 * k1 * t2 + 5 * k2 * t4 + 7 * k3 * t6 + 9 * k4 * t8)
            theta = torch.where(active, theta - fix, theta)
            active = active & (fix.abs() >= 1e-8)
        scale = torch.where(theta_d > 1e-8, torch.tan(theta) / theta_d.clamp(min=1e-300), torch.ones_like(theta))
        x, y = x0 * scale, y0 * scale
    else:
        raise NotImplementedError(model)                     # "fov" is a TODO in the reference as well
    events["position"] = torch.stack((x * fx + cx, y * fy + cy), dim=1).to(torch.float32)
    return events


class Event(torch.utils.data.Dataset):
    """`datasets.Event` (datasets.py:14-131,366-373): the transformed events of a sequence -- the cached
    `events.pt` when present, else raw events -> `queue_raw_events` -> `colorize_events` -> `undistort_events`
    (all on `device`) -> cached in the reference's file and schema (CPU tensors: position (M,2) f32, start_ts /
    end_ts / num_pos / num_neg (M,) i64 [, channel_idx (M,) u8]) -- optionally permuted with a seeded
    `torch.randperm` as the reference does to emulate sparser event streams."""
    CAMERA_CALIBRATION_FILENAME = "camera_calibration.npz"

    def __init__(self, root_directory, permutation_seed, device="cuda"):
        super().__init__()
        path = os.path.join(root_directory, TF_EVENTS_FILENAME)
        if os.path.isfile(path):
            self.events = dict(torch.load(path))
        else:
            calib = self.load_camera_calibration(root_directory)
            ev = queue_raw_events(root_directory, calib, device)
            ev = undistort_events(colorize_events(ev, calib), calib)
            self.events = {k: v.cpu() for k, v in ev.items()}
            torch.save(self.events, path)
        if permutation_seed is not None:
            perm = torch.randperm(len(self), generator=torch.Generator().manual_seed(permutation_seed))
            self.events = {k: v[perm] for k, v in self.events.items()}

    @classmethod
    def load_camera_calibration(cls, root_directory):
        return np.load(os.path.join(root_directory, cls.CAMERA_CALIBRATION_FILENAME))

    @staticmethod
    def load_max_refractory_period(root_directory):
        path = os.path.join(root_directory, MAX_REFRACTORY_PERIOD_FILENAME)
        return torch.load(path) if os.path.isfile(path) else None

    @staticmethod
    def save_max_refractory_period(max_refractory_period, root_directory):
        torch.save(max_refractory_period, os.path.join(root_directory, MAX_REFRACTORY_PERIOD_FILENAME))

    def __getitem__(self, index):
        return {k: v[index] for k, v in self.events.items()}

    def __len__(self):
        return len(self.events["position"])
