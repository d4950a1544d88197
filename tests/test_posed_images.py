"""`ren_b200.posed_images.PosedImage` against the UNMODIFIED reference class (robust_e_nerf/data/datasets.py:
376-690): the golden holds small synthetic view folders (images, transforms json, calibration) and what the
reference made of them (`oracle/make_golden_posed.py`); the folders are re-created here file by file."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.make_golden_posed import CASES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "posed_images.npz")


def _write(root, tag, g):
    import cv2
    dtype, ch, ext, bayer, space, _a, _b, _f, stage, _s = CASES[tag]
    meta = json.loads(str(g[f"{tag}_meta"]))
    os.makedirs(os.path.join(root, "views", stage), exist_ok=True)
    for i, frame in enumerate(meta["frames"]):
        assert cv2.imwrite(os.path.join(root, "views", frame["file_path"] + ext), g[f"{tag}_img{i}"])
    with open(os.path.join(root, "views", f"transforms_{stage}.json"), "w") as f:
        json.dump(meta, f)
    np.savez(os.path.join(root, "camera_calibration.npz"), bayer_pattern=np.array(bayer),
             img_height=np.uint16(6), img_width=np.uint16(8))
    if space is not None:
        np.savez(os.path.join(root, "renderer_params.npz"), interm_color_space=np.array(space),
                 log_eps=np.array(1e-3))


@pytest.mark.parametrize("tag", sorted(CASES))
def test_posed_images_match_reference_golden(tag, tmp_path):
    from ren_b200.posed_images import PosedImage
    g = np.load(GOLDEN)
    _write(str(tmp_path), tag, g)
    spec = CASES[tag]
    ds = PosedImage(str(tmp_path), spec[8], spec[9], alpha_over_white_bg=spec[5])
    for k in ("sample_id", "img", "T_wc_position", "T_wc_orientation", "intrinsics"):
        want = g[f"{tag}_out_{k}"]
        got = ds.posed_imgs[k].numpy()
        assert got.shape == want.shape and got.dtype == want.dtype, k
        assert np.array_equal(got, want), k                      # same arithmetic, same order: bit for bit
    assert float(ds.min_normalized_pixel_value) == float(g[f"{tag}_min"])
    assert float(ds.max_normalized_pixel_value) == float(g[f"{tag}_max"])
    assert len(ds) == 3
    item = ds[1]
    assert set(item) == {"sample_id", "img", "T_wc_position", "T_wc_orientation"}
    assert "".join(map(chr, item["sample_id"].tolist())).strip().startswith("r_")
    # the rotation is still a rotation after the OpenGL -> common camera flip
    R = ds.posed_imgs.T_wc_orientation[0].double()
    assert torch.allclose(R @ R.T, torch.eye(3, dtype=torch.float64), atol=1e-6)


def test_views_folder_may_sit_one_level_up(tmp_path):
    from ren_b200.posed_images import PosedImage
    g = np.load(GOLDEN)
    _write(str(tmp_path), "real_mono8", g)
    seq = tmp_path / "sequence"
    seq.mkdir()
    os.replace(tmp_path / "camera_calibration.npz", seq / "camera_calibration.npz")
    ds = PosedImage(str(seq), "val", None)
    assert len(ds) == 3
    with pytest.raises(FileNotFoundError):
        PosedImage(str(seq), "test", None)                       # no transforms_test.json


def test_module_constructor_registers_eval_stages(tmp_path, built_lib):
    """robust_e_nerf.py:86-152: with a `views` folder in the dataset the constructor sizes the validation / test
    renders from the posed images; the test split may be missing (novel_view), the validation split may not."""
    import shutil
    import helpers as H
    from oracle import synth
    root = synth.write_dataset(str(tmp_path / "seq"))
    g = np.load(GOLDEN)
    extra = tmp_path / "views_src"
    extra.mkdir()
    _write(str(extra), "real_mono8", g)                          # stage "val", intrinsics key
    shutil.rmtree(os.path.join(root, "views"))                   # the oracle's own 8x8 placeholder views
    shutil.copytree(extra / "views", os.path.join(root, "views"))
    gold = H.load_golden("step_plumbing")
    m = H.build_product(gold, root)
    assert m.val_img_pixel_pos.shape == (6, 8, 2)
    assert m.val_min_normalized_pixel_value == 0.5 / 256 and m.val_max_normalized_pixel_value == 1 - 0.5 / 256
    K = torch.tensor([[10.5, 0.0, 3.4], [0.0, 10.25, 2.6], [0.0, 0.0, 1.0]])
    assert torch.allclose(m.val_intrinsics_inv, torch.linalg.inv(K))
    assert not hasattr(m, "test_intrinsics_inv")                 # no transforms_test.json: not an error
    stage = m._eval_stage("val")
    assert stage["min_normalized_pixel_value"] == 0.5 / 256
    with pytest.raises(RuntimeError):
        m._eval_stage("test")
    # a views folder without the validation split is an error, as in the reference
    os.remove(os.path.join(root, "views", "transforms_val.json"))
    with pytest.raises(FileNotFoundError):
        H.build_product(gold, root)
