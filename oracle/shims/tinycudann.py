"""ORACLE SHIM (test infrastructure): torch-CPU stand-in for the one tiny-cuda-nn entry point the reference
calls -- `tcnn.Encoding(n_input_dims, encoding_config, dtype)` (robust_e_nerf/external/ngp.py:166-170), the
multiresolution hash grid of `arch: ngp`.  tiny-cuda-nn is CUDA-only, un-vendored and unpinned
(environment.yml:31): the arithmetic is restated from its published grid.h in `oracle/hashgrid.py`
(PARITY UNPINNED at this seam).  Like the PyTorch binding, the module owns ONE flat fp32 parameter `params`
(state-dict key `...mlp_base.0.params`) and exposes `n_input_dims` / `n_output_dims`."""
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle import hashgrid  # noqa: E402

_OTYPE_TO_GRID = {"HashGrid": "Hash", "DenseGrid": "Dense", "TiledGrid": "Tiled", "Grid": "Hash"}


class Encoding(torch.nn.Module):
    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
        super().__init__()
        cfg = dict(encoding_config)
        if n_input_dims != 3:
            raise NotImplementedError("the shim restates the 3-D grid encoding")
        if cfg.get("otype") not in _OTYPE_TO_GRID:
            raise NotImplementedError(f"otype {cfg.get('otype')!r}")
        if cfg.get("interpolation", "Linear") != "Linear":
            raise NotImplementedError("interpolation: Linear only")
        self.n_input_dims = n_input_dims
        self.dtype = dtype or torch.float32
        self.meta = hashgrid.grid_meta(
            n_levels=cfg.get("n_levels", 16), n_features_per_level=cfg.get("n_features_per_level", 2),
            log2_hashmap_size=cfg.get("log2_hashmap_size", 19), base_resolution=cfg.get("base_resolution", 16),
            per_level_scale=cfg.get("per_level_scale", 2.0), grid_type=_OTYPE_TO_GRID[cfg["otype"]])
        self.n_output_dims = self.meta["n_output_dims"]
        self.params = torch.nn.Parameter(hashgrid.init_params(self.meta))   # torch's global RNG

    def forward(self, x):
        return hashgrid.encode(x, self.params, self.meta).to(self.dtype)
