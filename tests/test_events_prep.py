"""Raw-event conversion: the numpy oracle against the golden vectors minted by the
UNMODIFIED reference (`Event.queue_raw_events`, `Event.extract_max_refractory_period`,
`Event.colorize_events`, robust_e_nerf/data/datasets.py:132-328), and -- on the GPU -- `ren_event_queue`
through the C ABI against both.  Integer work: the bar is bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import events_prep as EP

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "events_prep.npz")
TAGS = ["a", "b", "c"]


def _case(g, tag):
    raw = {k: g[f"{tag}_raw_{k}"] for k in ("position", "timestamp", "polarity")}
    w, h = (int(v) for v in g[f"{tag}_calib"])
    calib = {"img_width": w, "img_height": h, "bayer_pattern": str(g[f"{tag}_bayer"])}
    ref = {k[len(tag) + 3:]: g[k] for k in g.files if k.startswith(f"{tag}_q_")}
    return raw, calib, ref, g[f"{tag}_max_refractory_period"]


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_matches_reference_golden(tag):
    g = np.load(GOLDEN)
    raw, calib, ref, mrp = _case(g, tag)
    q = EP.queue_raw_events(raw, calib)
    for k in ("position", "start_ts", "end_ts", "num_pos", "num_neg"):
        assert q[k].dtype == np.int64 and np.array_equal(q[k], ref[k]), k
    m = EP.extract_max_refractory_period(raw, calib)
    assert np.array_equal(np.asarray(m), mrp) and np.asarray(m).dtype == mrp.dtype


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_event_queue_matches_reference_golden(tag, built_lib, tmp_path):
    from ren_b200 import events as E
    g = np.load(GOLDEN)
    raw, calib, ref, mrp = _case(g, tag)
    np.savez(tmp_path / E.RAW_EVENTS_FILENAME, **raw)
    q = E.colorize_events(E.queue_raw_events(str(tmp_path), calib), calib)
    for k, v in ref.items():
        got = q[k].cpu().numpy()
        assert got.dtype == v.dtype and np.array_equal(got, v), k
    m = E.extract_max_refractory_period(raw, calib)
    assert np.array_equal(m.numpy(), mrp) and m.numpy().dtype == mrp.dtype


@pytest.mark.gpu
def test_event_queue_full_size_vs_oracle(built_lib):
    """4 M events on a 640x480 sensor (the size class of a real sequence) vs the numpy oracle, plus the
    size-independent invariants: stream order kept, every interval positive, num_pos + num_neg == 1."""
    from ren_b200 import events as E
    raw = EP.synth_raw_events(4_000_000, 640, 480, seed=11)
    calib = {"img_width": 640, "img_height": 480, "bayer_pattern": ""}
    q, m = E._run(raw, calib, "cuda")
    ref = EP.queue_raw_events(raw, calib)
    for k, v in ref.items():
        assert np.array_equal(q[k].cpu().numpy(), v), k
    assert m == int(EP.extract_max_refractory_period(raw, calib))
    assert bool((q["end_ts"] > q["start_ts"]).all()) and bool((q["num_pos"] + q["num_neg"] == 1).all())
    assert bool((q["end_ts"][1:] >= q["end_ts"][:-1]).all())
    # empty stream
    empty = {"position": np.zeros((0, 2), np.uint16), "timestamp": np.zeros((0,), np.int64),
             "polarity": np.zeros((0,), bool)}
    q0, m0 = E._run(empty, calib, "cuda")
    assert q0["end_ts"].numel() == 0 and m0 == np.iinfo(np.int64).max


@pytest.mark.gpu
def test_raw_events_to_training_steps(built_lib, dataset_dir):
    """The rows either side of the hot path in one flow, all on the device: `raw_events.npz` -> `ren_event_queue`
    -> `DeviceEventSampler` (randint-with-replacement gather + normalised timestamp samplers) ->
    `RobustENeRF.training_step` + backward + Adam for a few steps."""
    import helpers as H
    from oracle import synth
    from ren_b200 import events as E
    from ren_b200.data import DeviceEventSampler
    from ren_b200.trainer import FlatTrainer
    scene = synth.load_scene(dataset_dir)
    T = scene["T_wc_timestamp"]
    raw = EP.synth_raw_events(300_000, 640, 480, seed=21)
    # spread the stream over the interior of the pose track
    span = int(T[-6] - T[5])
    raw["timestamp"] = (int(T[5]) + (raw["timestamp"] - raw["timestamp"][0]) % span).astype(np.int64)
    raw["timestamp"].sort()
    calib = {"img_width": 640, "img_height": 480, "bayer_pattern": ""}
    q, tau_max = E._run(raw, calib, "cuda")
    assert q["end_ts"].numel() > 1000 and tau_max > 0
    # the sensor model needs refractory period < every interval (event_generation_params.py:117-131 redefines
    # tau from the extracted maximum); this synthetic stream is denser than the calibrated tau, so keep the
    # intervals that respect it
    keep = (q["end_ts"] - q["start_ts"]) > 100_000
    q = {k: v[keep] for k, v in q.items()}
    g = dict(H.load_golden("step_8x256"))
    g["rcfg"] = dict(g["rcfg"], render_step_size=4.0 / 32, far_plane=0.5 + 32.5 * (4.0 / 32))
    m = H.build_product(g, dataset_dir, "cuda")
    tr = FlatTrainer(m)
    sampler = DeviceEventSampler(q, 512, device="cuda", seed=1)
    losses = []
    for i, batch in zip(range(4), sampler):
        losses.append(float(tr.step(batch, i)))
    assert all(np.isfinite(losses)), losses
    assert int(m.logged["train/batch_size"]) == 512
    # the trainer reads the MLP gradients as the one flat buffer the fused backward wrote
    flat = tr._flat_mlp_grad()
    assert flat is not None
    assert torch.equal(flat, torch.cat([p.grad.reshape(-1) for p in m.nerf.mlp_parameters()]))
    assert torch.equal(tr.gbuf[:flat.numel()], flat)


def _undistort_cases():
    K = np.array([[500.0, 0, 319.5], [0, 500.0, 239.5], [0, 0, 1]])
    return K, [("plumb_bob", np.array([-0.3, 0.12, 0.001, -0.002])),
               ("plumb_bob", np.array([0.2, -0.05, -0.003, 0.004])),
               ("equidistant", np.array([0.05, -0.01, 0.002, -0.0003]))]


def _opencv_undistort(model, src, K, D):
    import cv2
    if model == "plumb_bob":
        return cv2.undistortPoints(src, K, D, P=K).squeeze(1)
    return cv2.fisheye.undistortPoints(src[:, None, :], K, D, P=K).squeeze(1)


def test_undistort_events_matches_opencv():
    """`undistort_events` (datasets.py:330-364) against the OpenCV calls the reference makes, on every pixel of a
    640 x 480 sensor: same iterations in float64 -> identical float32 positions on the CPU."""
    from ren_b200 import events as E
    K, cases = _undistort_cases()
    xs, ys = np.meshgrid(np.arange(640), np.arange(480))
    pos = torch.tensor(np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int64))
    for model, D in cases:
        calib = {"distortion_model": np.array(model), "distortion_params": D, "intrinsics": K}
        ev = {"position": pos, "end_ts": torch.zeros(len(pos), dtype=torch.int64)}
        out = E.undistort_events(ev, calib)
        assert out["position"].dtype == torch.float32 and out["end_ts"] is ev["end_ts"]
        ref = _opencv_undistort(model, pos.to(torch.float32).numpy(), K, D)
        assert np.array_equal(out["position"].numpy(), ref), model
    # no distortion parameters: positions are only cast (datasets.py:339-344)
    calib = {"distortion_model": np.array("plumb_bob"), "distortion_params": np.zeros(0), "intrinsics": K}
    out = E.undistort_events({"position": pos}, calib)
    assert out["position"].dtype == torch.float32 and torch.equal(out["position"], pos.float())
    with pytest.raises(NotImplementedError):
        E.undistort_events({"position": pos}, {"distortion_model": np.array("fov"),
                                               "distortion_params": np.zeros(4), "intrinsics": K})


@pytest.mark.gpu
def test_undistort_events_on_device():
    from ren_b200 import events as E
    K, cases = _undistort_cases()
    gen = np.random.default_rng(1)
    pos = torch.tensor(np.stack([gen.integers(0, 640, 200000), gen.integers(0, 480, 200000)], 1).astype(np.int64))
    for model, D in cases:
        calib = {"distortion_model": np.array(model), "distortion_params": D, "intrinsics": K}
        out = E.undistort_events({"position": pos.cuda()}, calib)["position"]
        assert out.is_cuda
        ref = _opencv_undistort(model, pos.to(torch.float32).numpy(), K, D)
        assert float(np.abs(out.cpu().numpy() - ref).max()) < 1e-3          # fused multiply-adds on the device


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_event_dataset_builds_and_caches_like_the_reference(tag, built_lib, tmp_path):
    """`ren_b200.events.Event` from a raw stream: the cached `events.pt` holds what the unmodified reference's
    queue + colorize produced (golden), positions cast to float32 (no distortion parameters), and a second
    construction reads the cache; the seeded permutation is `torch.randperm`'s."""
    from ren_b200 import events as E
    g = np.load(GOLDEN)
    w, h = (int(v) for v in g[f"{tag}_calib"])
    np.savez(tmp_path / "raw_events.npz", **{k: g[f"{tag}_raw_{k}"] for k in ("position", "timestamp", "polarity")})
    np.savez(tmp_path / "camera_calibration.npz", img_width=np.uint16(w), img_height=np.uint16(h),
             bayer_pattern=g[f"{tag}_bayer"], distortion_model=np.array("plumb_bob"),
             distortion_params=np.zeros(0), intrinsics=np.eye(3))
    ds = E.Event(str(tmp_path), None)
    assert os.path.isfile(tmp_path / "events.pt")
    keys = [k[len(tag) + 3:] for k in g.files if k.startswith(f"{tag}_q_")]
    assert sorted(ds.events) == sorted(keys)
    for k in keys:
        want = torch.from_numpy(g[f"{tag}_q_{k}"])
        got = ds.events[k]
        assert not got.is_cuda
        if k == "position":
            assert got.dtype == torch.float32 and torch.equal(got, want.float())
        else:
            assert got.dtype == want.dtype and torch.equal(got, want), k
    again = E.Event(str(tmp_path), 9)                            # from the cache, permuted
    perm = torch.randperm(len(ds), generator=torch.Generator().manual_seed(9))
    assert torch.equal(again.events["end_ts"], ds.events["end_ts"][perm])
    assert set(again[0]) == set(keys) and len(again) == len(ds)


def test_colorize_events_matches_reference_golden_on_cpu():
    """`colorize_events` is plain tensor indexing: on the CPU against the reference's Bayer channel indices."""
    from ren_b200 import events as E
    g = np.load(GOLDEN)
    pos = torch.from_numpy(g["b_q_position"])
    out = E.colorize_events({"position": pos}, {"bayer_pattern": g["b_bayer"]})
    assert out["channel_idx"].dtype == torch.uint8
    assert torch.equal(out["channel_idx"], torch.from_numpy(g["b_q_channel_idx"]))
    mono = E.colorize_events({"position": pos}, {"bayer_pattern": np.array("")})
    assert "channel_idx" not in mono
