"""SE(3) pose chain of the reference's pose refinement, for the render path only.

Mirrors the parts of ``lib/camera.py`` that sit between the learnable ``se3_refine`` [n_img,6] and the rays:
``Lie.se3_to_SE3`` (:127-142, Taylor series of A, B, C with nth=10, :165-188), ``Pose.compose_pair`` (:92-99),
``Pose.invert`` (:76-82); and ``get_rays`` / ``get_rays_of_a_view`` (lib/voxurf_coarse.py:1339-1407) restricted
to the selected pixels.  The [n_img,6] -> [n_img,3,4] algebra is a handful of tiny tensor ops and stays in
torch (autograd); ray generation and the N-ray -> n_img reduction of the gradient are one kernel each.
"""
import math

import torch

from ._lib import check, lib, ptr, stream

_NTH = 10


def _series(theta, first_denominator_step, nth=_NTH):
    """sum_i (-1)^i theta^(2i) / d_i with d_0 = 1 (A) or the first factor applied at i = 0 (B, C)."""
    out = torch.zeros_like(theta)
    denom = 1.0
    for i in range(nth + 1):
        a, b = first_denominator_step(i)
        if a:
            denom *= a * b
        out = out + (-1) ** i * theta ** (2 * i) / denom
    return out


def taylor_A(x):   # sin(x)/x, lib/camera.py:165-172
    return _series(x, lambda i: (2 * i, 2 * i + 1) if i > 0 else (0, 0))


def taylor_B(x):   # (1-cos(x))/x^2, :173-180
    return _series(x, lambda i: (2 * i + 1, 2 * i + 2))


def taylor_C(x):   # (x-sin(x))/x^3, :181-188
    return _series(x, lambda i: (2 * i + 2, 2 * i + 3))


def skew(w):
    """[...,3] -> [...,3,3] cross-product matrix (lib/camera.py:157-163)."""
    w0, w1, w2 = w.unbind(-1)
    zero = torch.zeros_like(w0)
    return torch.stack([torch.stack([zero, -w2, w1], -1),
                        torch.stack([w2, zero, -w0], -1),
                        torch.stack([-w1, w0, zero], -1)], -2)


def se3_to_SE3(wu):
    """[...,6] twist (w | u) -> [...,3,4] pose [R | V u]  (lib/camera.py:127-142)."""
    w, u = wu.split([3, 3], dim=-1)
    wx = skew(w)
    theta = w.norm(dim=-1)[..., None, None]
    eye = torch.eye(3, device=wu.device, dtype=torch.float32)
    A, B, C = taylor_A(theta), taylor_B(theta), taylor_C(theta)
    wx2 = wx @ wx
    R = eye + A * wx + B * wx2
    V = eye + B * wx + C * wx2
    return torch.cat([R, V @ u[..., None]], dim=-1)


def compose_pair(pose_a, pose_b):
    """pose_new(x) = pose_b o pose_a(x)  (lib/camera.py:92-99)."""
    R_a, t_a = pose_a[..., :3], pose_a[..., 3:]
    R_b, t_b = pose_b[..., :3], pose_b[..., 3:]
    return torch.cat([R_b @ R_a, R_b @ t_a + t_b], dim=-1)


def invert(pose):
    """[R|t] -> [R^T | -R^T t]  (lib/camera.py:76-82)."""
    R, t = pose[..., :3], pose[..., 3:]
    R_inv = R.transpose(-1, -2)
    return torch.cat([R_inv, -R_inv @ t], dim=-1)


class _RaysFromPixels(torch.autograd.Function):
    @staticmethod
    def forward(ctx, c2w, Ks, img_id, px, py, inverse_y):
        if not c2w.is_cuda:
            raise RuntimeError("c2w must be a CUDA tensor (no CPU fallback)")
        c2w_c, Ks_c = c2w.contiguous().float(), Ks.contiguous().float()
        n = img_id.numel()
        rays_o = torch.empty((n, 3), dtype=torch.float32, device=c2w.device)
        rays_d = torch.empty_like(rays_o)
        with torch.cuda.device(c2w.device):
            check(lib.pp_rays_from_pixels(ptr(Ks_c), ptr(c2w_c), ptr(img_id), ptr(px), ptr(py), int(inverse_y), n,
                                          ptr(rays_o), ptr(rays_d), stream()), "rays_from_pixels")
        ctx.save_for_backward(c2w_c, Ks_c, img_id, px, py)
        ctx.inverse_y = int(inverse_y)
        return rays_o, rays_d

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_o, g_d):
        c2w, Ks, img_id, px, py = ctx.saved_tensors
        n = img_id.numel()
        grad_rays = torch.cat([g_o, g_d], dim=-1).contiguous()
        grad_c2w = torch.empty_like(c2w)
        with torch.cuda.device(c2w.device):
            check(lib.pp_pose_grad_reduce(ptr(Ks), ptr(c2w), ptr(img_id), ptr(px), ptr(py), ctx.inverse_y, n,
                                          c2w.shape[0], ptr(grad_rays), ptr(grad_c2w), stream()), "pose_grad_reduce")
        return grad_c2w, None, None, None, None, None


def rays_from_pixels(c2w, Ks, img_id, px, py, inverse_y=False):
    """c2w [n_img,3,4] (differentiable), Ks [n_img,3,3], img_id [N] int32, px/py [N] pixel coordinates
    (+0.5 for mode='center') -> rays_o, rays_d (unit) [N,3]; viewdirs == rays_d."""
    return _RaysFromPixels.apply(c2w, Ks, img_id.to(torch.int32).contiguous(), px.float().contiguous(),
                                 py.float().contiguous(), inverse_y)
