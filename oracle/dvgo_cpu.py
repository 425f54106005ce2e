"""TEST INFRASTRUCTURE ONLY -- CPU restatements of the DVGO render / train step.

Two formulations of the same path, both pure CPU:

* ``render_composed`` / ``train_step_composed`` -- the chain of reference operators the fused kernels
  replace: sample_pts_on_rays -> ``~mask_outbbox`` compaction (lib/voxurf_coarse.py:686-693) ->
  DenseGrid trilinear (lib/grid.py:47-58, F.grid_sample) -> raw2alpha -> Alphas2Weights
  (lib/voxurf_coarse.py:1316-1332) -> sigmoid(k0) -> per-ray segment sums
  (lib/voxurf_coarse.py:1034-1042 with ``alphainv_last * bg`` as in lib/dvgo_ori.py:363) -> losses ->
  torch autograd -> masked Adam (adam_upd_kernel.cu:25-40).  The native pieces come from the C oracle
  (oracle.pp_oracle.c).  This is the parity oracle for the fused CUDA path.
* ``DenseBaseline`` -- lib/dvgo_ori.py's own dense pure-PyTorch formulation (dense [N,S] sampling,
  F.grid_sample, softplus activation, cumprod compositing, Python Adam of lib/utils.py:176-198).  This
  is the CPU baseline bench.py times.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import oracle


class _Raw2Alpha(torch.autograd.Function):
    @staticmethod
    def forward(ctx, density, shift, interval):
        exp_d, alpha = oracle.raw2alpha(density, shift, interval)
        ctx.save_for_backward(exp_d)
        ctx.interval = interval
        return alpha

    @staticmethod
    def backward(ctx, grad_back):
        (exp_d,) = ctx.saved_tensors
        return oracle.raw2alpha_backward(exp_d, grad_back.contiguous(), ctx.interval), None, None


class _Alphas2Weights(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha, ray_id, n_rays):
        weights, T, last, i_start, i_end = oracle.alpha2weight(alpha, ray_id, n_rays)
        ctx.save_for_backward(alpha, weights, T, last, i_start, i_end)
        ctx.n_rays = n_rays
        return weights, last

    @staticmethod
    def backward(ctx, grad_weights, grad_last):
        alpha, weights, T, last, i_start, i_end = ctx.saved_tensors
        grad = oracle.alpha2weight_backward(alpha, weights, T, last, i_start, i_end, ctx.n_rays,
                                            grad_weights.contiguous(), grad_last.contiguous())
        return grad, None, None


def render_composed(density, k0, rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist, act_shift, interval,
                    bg=0.0, fast_color_thres=0.0):
    """density [1,1,X,Y,Z], k0 [1,3,X,Y,Z] (may require grad) -> dict of per-ray results."""
    n = rays_o.shape[0]
    xyz_min = torch.as_tensor(xyz_min, dtype=torch.float32)
    xyz_max = torch.as_tensor(xyz_max, dtype=torch.float32)
    pts, mask, ray_id, step_id, n_steps, *_ = oracle.sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near,
                                                                        far, stepdist)
    keep = ~mask
    pts, ray_id, step_id = pts[keep], ray_id[keep], step_id[keep]
    dens = oracle.dense_grid_forward(density, pts, xyz_min, xyz_max)
    alpha = _Raw2Alpha.apply(dens, act_shift, interval)
    weights, last = _Alphas2Weights.apply(alpha, ray_id, n)
    _, _, _, i_start, i_end = oracle.alpha2weight(alpha.detach(), ray_id, n)
    col = weights > fast_color_thres
    rgb = torch.sigmoid(oracle.dense_grid_forward(k0, pts[col], xyz_min, xyz_max))
    w_col, id_col = weights[col], ray_id[col]
    rgb_marched = oracle.segment_sum(w_col[:, None] * rgb, id_col, n) + last[:, None] * bg
    depth = oracle.segment_sum((weights * step_id * stepdist)[:, None], ray_id, n)[:, 0]
    return dict(rgb_marched=rgb_marched, alphainv_last=last, depth=depth, weights=weights, raw_rgb=rgb,
                ray_id=ray_id, ray_id_col=id_col, weights_col=w_col, n_samples=int(keep.sum()),
                n_marched=(i_end - i_start).to(torch.int32))


def losses(out, target, w_rgb=1.0, w_entropy_last=0.0, w_rgbper=0.0):
    """Photometric MSE (lib/losses.py:26-29) + DVGO's background-entropy and per-point colour terms."""
    n = target.shape[0]
    mse = F.mse_loss(out["rgb_marched"], target)
    loss = w_rgb * mse
    if w_entropy_last:
        pout = out["alphainv_last"].clamp(1e-6, 1 - 1e-6)
        loss = loss + w_entropy_last * (-(pout * torch.log(pout) + (1 - pout) * torch.log(1 - pout)).mean())
    if w_rgbper:
        rgbper = (out["raw_rgb"] - target[out["ray_id_col"]]).pow(2).sum(-1)
        loss = loss + w_rgbper * (rgbper * out["weights_col"].detach()).sum() / n
    return loss, mse


def masked_adam(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    """In place; adam_upd_kernel.cu:25-40 via the C oracle."""
    oracle.adam_upd(1, param.view(-1), grad.contiguous().view(-1), exp_avg.view(-1), exp_avg_sq.view(-1), None,
                    step, beta1, beta2, lr, eps)


# ---------------------------------------------------------------------------------------------------
class DenseBaseline:
    """lib/dvgo_ori.py restated for the coarse stage (rgbnet_dim=0), CPU, fp32."""

    def __init__(self, density, k0, xyz_min, xyz_max, stepsize, alpha_init=1e-2, near=0.2, far=1e9, bg=0.0,
                 fast_color_thres=0.0, lr=0.1, betas=(0.9, 0.99), eps=1e-8):
        self.density = density.clone().requires_grad_(True)
        self.k0 = k0.clone().requires_grad_(True)
        self.xyz_min = torch.as_tensor(xyz_min, dtype=torch.float32)
        self.xyz_max = torch.as_tensor(xyz_max, dtype=torch.float32)
        self.stepsize, self.near, self.far, self.bg = stepsize, near, far, bg
        self.fast_color_thres = fast_color_thres
        self.act_shift = math.log(1 / (1 - alpha_init) - 1)
        world = torch.tensor(density.shape[2:], dtype=torch.float32)
        self.voxel_size = float(((self.xyz_max - self.xyz_min).prod() / world.prod()).pow(1 / 3))
        self.voxel_size_ratio = 1.0
        self.n_samples = int(np.linalg.norm(np.array(density.shape[2:]) + 1) / stepsize) + 1
        self.hp = (betas[0], betas[1], lr, eps)
        self.state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in (self.density, self.k0)]
        self.step = 0

    def grid_sampler(self, xyz, grid):
        # lib/dvgo_ori.py:246-261
        shape = xyz.shape[:-1]
        q = xyz.reshape(1, 1, 1, -1, 3)
        ind_norm = ((q - self.xyz_min) / (self.xyz_max - self.xyz_min)).flip((-1,)) * 2 - 1
        return F.grid_sample(grid, ind_norm, mode='bilinear', align_corners=True).reshape(
            grid.shape[1], -1).T.reshape(*shape, grid.shape[1]).squeeze(-1)

    def sample_ray(self, rays_o, rays_d):
        # lib/dvgo_ori.py:263-287 (no jitter)
        vec = torch.where(rays_d == 0, torch.full_like(rays_d, 1e-6), rays_d)
        rate_a = (self.xyz_max - rays_o) / vec
        rate_b = (self.xyz_min - rays_o) / vec
        t_min = torch.minimum(rate_a, rate_b).amax(-1).clamp(min=self.near, max=self.far)
        t_max = torch.maximum(rate_a, rate_b).amin(-1).clamp(min=self.near, max=self.far)
        mask_outbbox = (t_max <= t_min)
        rng = torch.arange(self.n_samples)[None].float()
        step = self.stepsize * self.voxel_size * rng
        interpx = (t_min[..., None] + step / rays_d.norm(dim=-1, keepdim=True))
        rays_pts = rays_o[..., None, :] + rays_d[..., None, :] * interpx[..., None]
        mask_outbbox = mask_outbbox[..., None] | ((self.xyz_min > rays_pts) | (rays_pts > self.xyz_max)).any(dim=-1)
        return rays_pts, mask_outbbox

    def forward(self, rays_o, rays_d):
        # lib/dvgo_ori.py:289-379, post-activation branch, rgbnet None
        rays_pts, mask_outbbox = self.sample_ray(rays_o, rays_d)
        interval = self.stepsize * self.voxel_size_ratio
        alpha = torch.zeros_like(rays_pts[..., 0])
        density = self.grid_sampler(rays_pts[~mask_outbbox], self.density)
        alpha[~mask_outbbox] = 1 - torch.exp(-F.softplus(density + self.act_shift) * interval)
        weights, alphainv_cum = oracle.get_ray_marching_ray(alpha)
        mask = weights > self.fast_color_thres
        k0 = torch.zeros(*weights.shape, 3).to(weights)
        k0[mask] = self.grid_sampler(rays_pts[mask], self.k0)
        rgb = torch.sigmoid(k0)
        rgb_marched = (weights[..., None] * rgb).sum(-2) + alphainv_cum[..., [-1]] * self.bg
        return rgb_marched.clamp(0, 1), alphainv_cum[..., -1], int((~mask_outbbox).sum())

    def train_step(self, rays_o, rays_d, target):
        for p in (self.density, self.k0):
            p.grad = None
        rgb, _, m = self.forward(rays_o, rays_d)
        loss = F.mse_loss(rgb, target)
        loss.backward()
        self.step += 1
        b1, b2, lr, eps = self.hp
        with torch.no_grad():
            for p, (m1, m2) in zip((self.density, self.k0), self.state):
                oracle.python_adam_step(p, p.grad, m1, m2, self.step, b1, b2, lr, eps)
        return float(loss), m
