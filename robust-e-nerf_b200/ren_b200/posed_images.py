"""Posed target images of a dataset: `views/transforms_{stage}.json` + image files ->
normalised intensity images, camera poses in the common (x right, y down, z forward) convention, intrinsics and
the target value range -- what `RobustENeRF` needs to size and score its evaluation renders
(robust_e_nerf/models/robust_e_nerf.py:86-152,573-587).  Mirrors `datasets.PosedImage`
(robust_e_nerf/data/datasets.py:376-690): same constructor, attributes (`posed_imgs`, `min_normalized_pixel_value`,
`max_normalized_pixel_value`) and item schema; checked value for value against the unmodified reference class on
synthetic view folders (`tests/test_posed_images.py`, golden minted by `oracle/make_golden_posed.py`).

Host-side, offline code (json + OpenCV decode + a few numpy passes over the images)."""
import glob
import json
import math
import os

import numpy as np
import torch


class _Bag(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class PosedImage(torch.utils.data.Dataset):
    STAGES = ("train", "val", "test")
    NORMALIZED_SAMPLE_ID_CHAR_LEN = 16
    POSED_IMG_FOLDER_NAME = "views"

    def __init__(self, root_directory, stage, permutation_seed, alpha_over_white_bg=False):
        super().__init__()
        if stage not in self.STAGES:
            raise ValueError(f"stage must be one of {self.STAGES}")
        folder = self.posed_img_folder_path(root_directory)
        if folder is None:
            raise FileNotFoundError(f"no `{self.POSED_IMG_FOLDER_NAME}` folder in or next to {root_directory}")
        with open(os.path.join(folder, f"transforms_{stage}.json")) as f:
            meta = json.load(f)
        renderer_path = os.path.join(root_directory, "renderer_params.npz")
        renderer = np.load(renderer_path) if os.path.isfile(renderer_path) else None
        calib = np.load(os.path.join(root_directory, "camera_calibration.npz"))

        ids, imgs, positions, rotations = self._read_frames(folder, meta)
        imgs = self._to_intensity(imgs, alpha_over_white_bg, meta, renderer, str(calib["bayer_pattern"]))
        # OpenGL camera (y up, z backwards) -> common camera (y down, z forwards): flip the y and z axes
        rotations = rotations @ np.diag([1.0, -1.0, -1.0])
        dt = torch.get_default_dtype()
        self.posed_imgs = _Bag(sample_id=torch.tensor(ids), img=torch.tensor(imgs, dtype=dt),
                               T_wc_position=torch.tensor(positions, dtype=dt),
                               T_wc_orientation=torch.tensor(rotations, dtype=dt),
                               intrinsics=torch.tensor(self._intrinsics(meta, imgs.shape[-2:]), dtype=dt))
        if permutation_seed is not None:
            perm = torch.randperm(len(self), generator=torch.Generator().manual_seed(permutation_seed))
            for k in ("sample_id", "img", "T_wc_position", "T_wc_orientation"):
                self.posed_imgs[k] = self.posed_imgs[k][perm]

    @classmethod
    def posed_img_folder_path(cls, root_directory):
        for p in (os.path.join(root_directory, cls.POSED_IMG_FOLDER_NAME),
                  os.path.join(root_directory, "..", cls.POSED_IMG_FOLDER_NAME)):
            if os.path.isdir(p):
                return p
        return None

    @classmethod
    def _read_frames(cls, folder, meta):
        import cv2
        ids, imgs, positions, rotations = [], [], [], []
        for frame in meta["frames"]:
            name = os.path.basename(frame["file_path"]).ljust(cls.NORMALIZED_SAMPLE_ID_CHAR_LEN)
            ids.append([ord(c) for c in name])                  # fixed-length id as code points (collatable)
            path = glob.glob(os.path.join(folder, frame["file_path"] + ".*"))[0]
            imgs.append(cv2.imread(path, cv2.IMREAD_UNCHANGED))  # (H, W[, 3 | 4]) gray / BGR / BGRA
            T = np.array(frame["transform_matrix"])
            positions.append(T[:3, 3])
            rotations.append(T[:3, :3])
        return np.array(ids), np.stack(imgs), np.stack(positions), np.stack(rotations)

    @staticmethod
    def _intrinsics(meta, hw):
        if "camera_angle_x" in meta:                            # horizontal field of view, principal point centred
            H, W = hw
            f = (W / 2) / math.tan(meta["camera_angle_x"] / 2)
            return np.array([[f, 0, W / 2 - 0.5], [0, f, H / 2 - 0.5], [0, 0, 1]])
        return np.array(meta["intrinsics"])

    def _to_intensity(self, imgs, alpha_over_white_bg, meta, renderer, bayer_pattern):
        import cv2
        quantized = np.issubdtype(imgs.dtype, np.unsignedinteger)
        synthetic = renderer is not None
        if not (quantized or np.issubdtype(imgs.dtype, np.floating)) or np.any(imgs < 0):
            raise ValueError("images must be unsigned-integer captures / renders or non-negative float renders")
        channels = 1 if imgs.ndim == 3 else imgs.shape[3]
        if channels not in (1, 3, 4) or (channels == 4 and not synthetic):
            raise ValueError("images must be gray, BGR or (synthetic renders only) BGRA")
        space = str(renderer["interm_color_space"]) if synthetic else None
        if synthetic and space != ("display" if quantized else "linear"):
            raise ValueError("quantized renders are display-space, float renders linear-space")
        if not synthetic and not quantized:
            raise ValueError("real captures are quantized")
        levels = None
        if quantized:
            levels = 2 ** meta["bit_depth"] if "bit_depth" in meta else np.iinfo(imgs.dtype).max + 1

        if alpha_over_white_bg:                                  # composite the render over white
            if space == "display":                               # straight alpha, quantized
                alpha = (imgs[..., 3] / (levels - 1))[..., None]
                imgs = alpha * imgs[..., :3] + (1 - alpha) * (levels - 1)
            elif space == "linear":                              # premultiplied alpha
                imgs = imgs[..., :3] + (1 - imgs[..., 3][..., None])
        elif channels == 4:
            imgs = imgs[..., :3]
        imgs = imgs.astype(np.float32)
        if bayer_pattern != "":                                  # colour sensor: (N, 3, H, W) RGB
            imgs = np.stack([cv2.cvtColor(im, cv2.COLOR_BGR2RGB) for im in imgs]).transpose(0, 3, 1, 2)
        elif channels == 3:                                      # monochrome sensor: luma
            imgs = np.stack([cv2.cvtColor(im, cv2.COLOR_BGR2GRAY) for im in imgs])

        if quantized:
            # an ADC code x stands for the analog range [x, x + 1): its centre, normalised by the number of levels
            self.min_normalized_pixel_value = 0.5 / levels
            imgs = imgs / levels + self.min_normalized_pixel_value
            self.max_normalized_pixel_value = 1 - self.min_normalized_pixel_value
        else:
            self.min_normalized_pixel_value = renderer["log_eps"]
            imgs = imgs + self.min_normalized_pixel_value
            self.max_normalized_pixel_value = imgs.max()
        return imgs

    def __getitem__(self, index):
        return {k: v[index] for k, v in self.posed_imgs.items() if k != "intrinsics"}

    def __len__(self):
        return len(self.posed_imgs.img)
