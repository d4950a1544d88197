"""Event dataset preparation on the GPU -- the host-side mirror of the reference's
`robust_e_nerf.data.datasets.Event` class methods for the raw-event conversion
(robust_e_nerf/data/datasets.py:132-328), backed by `ren_event_queue` (one stable radix sort + scan
instead of two per-event Python loops).  SURVEY.md section 8(f) #3.

    events = queue_raw_events(root_directory, camera_calibration)          # == Event.queue_raw_events
    tau_max = extract_max_refractory_period(raw_events, camera_calibration)
    events = colorize_events(events, camera_calibration)

File names / keys are the reference's (`raw_events.npz` {position (N,2) u16, timestamp (N,) i64 ns,
polarity (N,) bool}, `camera_calibration.npz`, `events.pt`, `max_refractory_period.pt`).
"""
import ctypes as C
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
