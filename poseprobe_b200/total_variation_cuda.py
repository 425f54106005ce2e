"""Drop-in for the reference's ``total_variation_cuda`` extension (lib/cuda/total_variation.cpp:29-32).

Both functions mutate ``grad`` in place, exactly like the reference; ``param``/``grad``/``mask`` are
``[1,C,X,Y,Z]`` tensors.  The reference's axis-weight quirks are preserved (see the header).
"""
from ._lib import check, check_input, dtype_code, lib, ptr, stream
import torch


def _dims(param):
    if param.dim() != 5:
        raise RuntimeError("param must be [1,C,X,Y,Z]")
    # the reference treats dim 0 and 1 alike: every [X,Y,Z] volume is an independent plane
    return param.size(0) * param.size(1), param.size(2), param.size(3), param.size(4)


def total_variation_add_grad(param, grad, wx, wy, wz, dense_mode):
    """In-place grad += TV gradient (total_variation.cpp:16-20)."""
    check_input(param, "param")
    check_input(grad, "grad")
    dt = dtype_code(param, "total_variation_add_grad_cuda")
    planes, x, y, z = _dims(param)
    with torch.cuda.device(param.device):
        check(lib.pp_total_variation_add_grad(dt, ptr(param), ptr(grad), None, wx, wy, wz,
                                              int(bool(dense_mode)), planes, x, y, z, stream()),
              "total_variation_add_grad")


def total_variation_add_grad_new(param, grad, mask, wx, wy, wz, dense_mode):
    """Masked variant (total_variation.cpp:22-27); mask is a float tensor shaped like param."""
    check_input(param, "param")
    check_input(grad, "grad")
    check_input(mask, "mask")
    dt = dtype_code(param, "total_variation_add_grad_cuda_new")
    if mask.dtype != param.dtype or mask.numel() != param.numel():
        raise RuntimeError("mask must have param's dtype and shape")
    planes, x, y, z = _dims(param)
    with torch.cuda.device(param.device):
        check(lib.pp_total_variation_add_grad(dt, ptr(param), ptr(grad), ptr(mask), wx, wy, wz,
                                              int(bool(dense_mode)), planes, x, y, z, stream()),
              "total_variation_add_grad_new")
