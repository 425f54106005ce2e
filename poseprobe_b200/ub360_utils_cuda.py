"""Drop-in for the reference's ``ub360_utils_cuda`` extension (lib/cuda/ub360_utils.cpp:21-23)."""
import torch

from ._lib import check, check_input, dtype_code, lib, ptr, stream


def cumdist_thres(dist, thres):
    """dist [N_rays, N_pts] -> bool mask, running sum with reset (ub360_utils.cpp:15-19)."""
    check_input(dist, "dist")
    dt = dtype_code(dist, "cumdist_thres_cuda")
    n_rays, n_pts = dist.size(0), dist.size(1)
    mask = torch.empty((n_rays, n_pts), dtype=torch.bool, device=dist.device)
    with torch.cuda.device(dist.device):
        check(lib.pp_cumdist_thres(dt, ptr(dist), thres, n_rays, n_pts, ptr(mask), stream()),
              "cumdist_thres")
    return mask
