"""ORACLE (test infrastructure): mint tests/golden/posed_images.npz by running the UNMODIFIED reference
`datasets.PosedImage` (robust_e_nerf/data/datasets.py:376-690) on small synthetic view folders.  The folder
contents (images, transforms json, calibration) are stored next to the reference outputs so that the test can
re-create the files anywhere.

    python -m oracle.make_golden_posed          # build container only (needs /root/reference)
"""
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_loader  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                      "posed_images.npz")

# tag -> (image dtype, channels, extension, bayer pattern, renderer colour space | None, alpha over white,
#         bit_depth key | None, intrinsics from fov?, stage, permutation seed | None)
CASES = {
    "real_mono8": ("uint8", 1, ".png", "", None, False, None, False, "val", 5),
    "synth_bgra8_bayer": ("uint8", 4, ".png", "RGGB", "display", True, None, True, "test", None),
    "synth_bgr16_gray": ("uint16", 3, ".png", "", "display", False, 12, True, "train", 3),
    "synth_bgra8_dropalpha": ("uint8", 4, ".png", "", "display", False, None, False, "val", None),
    "synth_linear_f32": ("float32", 3, ".tiff", "RGGB", "linear", False, None, True, "test", 7),
}


def write_case(root, tag, spec, rng):
    """Writes one dataset folder; returns the dict of arrays needed to write it again."""
    import cv2
    dtype, ch, ext, bayer, space, _alpha, bit_depth, use_fov, stage, _seed = spec
    H, W, N = 6, 8, 3
    os.makedirs(os.path.join(root, "views", stage), exist_ok=True)
    frames, store = [], {}
    for i in range(N):
        shape = (H, W) if ch == 1 else (H, W, ch)
        if dtype == "float32":
            img = rng.random(shape, dtype=np.float32) * 2.0
        else:
            hi = (2 ** bit_depth if bit_depth else np.iinfo(dtype).max + 1)
            img = rng.integers(0, hi, size=shape).astype(dtype)
        rel = f"{stage}/r_{i:03d}"
        assert cv2.imwrite(os.path.join(root, "views", rel + ext), img)
        T = np.eye(4)
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        T[:3, :3] = q
        T[:3, 3] = rng.standard_normal(3)
        frames.append({"file_path": rel, "transform_matrix": T.tolist()})
        store[f"{tag}_img{i}"] = img
    meta = {"frames": frames}
    if use_fov:
        meta["camera_angle_x"] = 0.6911112070083618
    else:
        meta["intrinsics"] = [[10.5, 0.0, 3.4], [0.0, 10.25, 2.6], [0.0, 0.0, 1.0]]
    if bit_depth:
        meta["bit_depth"] = bit_depth
    with open(os.path.join(root, "views", f"transforms_{stage}.json"), "w") as f:
        json.dump(meta, f)
    np.savez(os.path.join(root, "camera_calibration.npz"), bayer_pattern=np.array(bayer),
             img_height=np.uint16(H), img_width=np.uint16(W))
    if space is not None:
        np.savez(os.path.join(root, "renderer_params.npz"), interm_color_space=np.array(space),
                 log_eps=np.array(1e-3))
    store[f"{tag}_meta"] = np.array(json.dumps(meta))
    return store


def main():
    ref = ref_loader.load_reference()
    PosedImage = ref.data.datasets.PosedImage
    rng = np.random.default_rng(11)
    out = {}
    for tag, spec in CASES.items():
        root = tempfile.mkdtemp(prefix="ren_b200_posed_")
        out.update(write_case(root, tag, spec, rng))
        ds = PosedImage(root, spec[8], spec[9], alpha_over_white_bg=spec[5])
        for k, v in ds.posed_imgs.items():
            out[f"{tag}_out_{k}"] = v.numpy()
        out[f"{tag}_min"] = np.asarray(ds.min_normalized_pixel_value, dtype=np.float64)
        out[f"{tag}_max"] = np.asarray(ds.max_normalized_pixel_value, dtype=np.float64)
        print(tag, tuple(ds.posed_imgs.img.shape), float(out[f"{tag}_min"]), float(out[f"{tag}_max"]))
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    main()
