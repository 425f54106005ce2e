"""poseprobe_b200 -- B200-native (sm_100a) voxel-grid volume-rendering operators.

A from-scratch implementation of PoseProbe's ``lib/cuda`` hot path behind the reference's own
operator API.  Importing this package loads ``libposeprobe_b200.so`` (hand-written CUDA behind a
C ABI, see ``include/poseprobe_b200.h``); there is no CPU or PyTorch fallback -- a missing library
is an ImportError.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from . import adam_upd_cuda, render_utils_cuda, total_variation_cuda, ub360_utils_cuda  # noqa: F401
from . import camera, dvgo, grid, ops, scene, voxurf  # noqa: F401
from .grid import DenseGrid, MaskGrid  # noqa: F401
from .ops import Alphas2Weights, Raw2Alpha  # noqa: F401

from .dvgo import DirectVoxGO, FusedDVGO  # noqa: F401
from .voxurf import FusedVoxurf, Voxurf  # noqa: F401

__all__ = ["render_utils_cuda", "total_variation_cuda", "adam_upd_cuda", "ub360_utils_cuda", "grid", "ops", "dvgo",
           "voxurf", "camera", "scene", "DenseGrid", "MaskGrid", "Alphas2Weights", "Raw2Alpha", "DirectVoxGO", "FusedDVGO",
           "Voxurf", "FusedVoxurf"]
