"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

numpy restatement of the reference's raw-event conversion
(robust_e_nerf/data/datasets.py:132-284): `extract_max_refractory_period` and `queue_raw_events`,
vectorised with a stable per-pixel sort (the reference walks the time-ordered stream with a 2-deep
sliding window per pixel).  Pinned bit-for-bit to the UNMODIFIED reference functions by
tests/golden/events_prep.npz (minted by oracle/make_golden_events.py in the build container).
"""
import numpy as np


def synth_raw_events(n, width, height, seed=0, dup_rate=0.05):
    """Time-ordered synthetic stream in the `raw_events.npz` schema (scripts/preprocess_esim.py:210-259):
    position (n,2) u16, timestamp (n,) i64 ns (non-decreasing, with repeated timestamps), polarity bool."""
    rng = np.random.default_rng(seed)
    pos = np.stack([rng.integers(0, width, n), rng.integers(0, height, n)], axis=1).astype(np.uint16)
    dt = rng.integers(0, 2000, n).astype(np.int64)
    dt[rng.random(n) < dup_rate] = 0                     # simultaneous events
    ts = np.cumsum(dt) + 1_000_000
    # force some same-pixel duplicates at the same timestamp (the skipped case of both loops)
    idx = rng.choice(n - 1, size=max(1, n // 50), replace=False)
    pos[idx + 1] = pos[idx]
    ts[idx + 1] = ts[idx]
    ts = np.maximum.accumulate(ts)
    pol = rng.random(n) < 0.5
    return {"position": pos, "timestamp": ts.astype(np.int64), "polarity": pol}


def _pairs(raw_events, width):
    pos = raw_events["position"].astype(np.int64)
    ts = raw_events["timestamp"].astype(np.int64)
    key = pos[:, 1] * width + pos[:, 0]
    order = np.argsort(key, kind="stable")               # time order is kept inside a pixel
    same = np.zeros(len(ts), dtype=bool)
    same[1:] = key[order][1:] == key[order][:-1]
    prev = np.zeros(len(ts), dtype=np.int64)
    prev[order[1:]] = ts[order[:-1]]
    has_prev = np.zeros(len(ts), dtype=bool)
    has_prev[order] = same
    return pos, ts, prev, has_prev


def extract_max_refractory_period(raw_events, camera_calibration):
    """datasets.py:132-187: min positive interval between consecutive events of a pixel (events
    repeating the previous timestamp of their pixel are skipped); int64, or float inf if none."""
    pos, ts, prev, has_prev = _pairs(raw_events, int(camera_calibration["img_width"]))
    d = (ts - prev)[has_prev & (ts != prev)]
    return np.int64(d.min()) if d.size else np.float64("inf")


def queue_raw_events(raw_events, camera_calibration):
    """datasets.py:189-284: dict of int64 arrays for the events whose pixel has a previous event at a
    different timestamp: position (M,2), start_ts, end_ts, num_pos, num_neg."""
    pos, ts, prev, has_prev = _pairs(raw_events, int(camera_calibration["img_width"]))
    valid = has_prev & (ts != prev)
    pol = raw_events["polarity"].astype(np.int64)
    return {"position": pos[valid], "start_ts": prev[valid], "end_ts": ts[valid],
            "num_pos": pol[valid], "num_neg": 1 - pol[valid]}
