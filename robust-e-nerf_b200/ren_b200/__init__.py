"""ren_b200 -- B200-native (sm_100a) event-NeRF render/training hot path behind the
reference's `NeRF` / `RobustENeRF` interface (wengflow/robust-e-nerf).  The arithmetic
lives in hand-written CUDA (`csrc/`, built into `lib/libren_b200.so`, C ABI in
`include/ren_b200.h`); this package is the thin Python host side.  There is no CPU or
PyTorch fallback: importing works anywhere, calling needs the built library and a GPU."""
from . import _lib
from .nerf import NeRF, RenderFunction
from .module import (ContrastThreshold, DiffLossFunction, LinearTrajectory, RefractoryPeriod,
                     RobustENeRF)
from .posed_images import PosedImage

__all__ = ["NeRF", "RenderFunction", "RobustENeRF", "ContrastThreshold", "RefractoryPeriod",
           "LinearTrajectory", "DiffLossFunction", "PosedImage", "_lib"]
