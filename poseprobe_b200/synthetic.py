"""Synthetic scenes / ray batches.  Generated on the CPU under a fixed seed so
that CPU (oracle) and GPU runs see bit-identical inputs; the caller moves tensors to the device."""
import math

import torch


def make_rays(n_rays, seed=777, radius=2.5, target_half=0.8):
    """Origins uniform on the sphere |o| = radius, targets uniform in [-target_half, target_half]^3,
    unit directions.  Returns rays_o, rays_d, viewdirs (all [N,3] fp32, CPU)."""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n_rays, 3, generator=g)
    o = o / o.norm(dim=-1, keepdim=True) * radius
    tgt = (torch.rand(n_rays, 3, generator=g) * 2 - 1) * target_half
    d = tgt - o
    d = d / d.norm(dim=-1, keepdim=True)
    return o.contiguous(), d.contiguous(), d.clone()


def make_dvgo_grids(grid_size, k0_dim=3, seed=777):
    """density [1,1,G,G,G] ~ N(0,1); k0 [1,C,G,G,G] ~ 0.1 N(0,1)  (fp32, CPU)."""
    g = torch.Generator().manual_seed(seed + 1)
    G = grid_size
    density = torch.randn(1, 1, G, G, G, generator=g)
    k0 = torch.randn(1, k0_dim, G, G, G, generator=g) * 0.1
    return density, k0


def make_sdf_grid(grid_size, radius=0.5, noise=0.02, seed=777):
    """distance-to-sphere SDF + noise on [-1,1]^3, [1,1,G,G,G]."""
    g = torch.Generator().manual_seed(seed + 2)
    G = grid_size
    lin = torch.linspace(-1, 1, G)
    x, y, z = torch.meshgrid(lin, lin, lin, indexing="ij")
    sdf = (x * x + y * y + z * z).sqrt() - radius
    sdf = sdf + noise * torch.randn(G, G, G, generator=g)
    return sdf[None, None].contiguous()


def act_shift(alpha_init=1e-2):
    """lib/dvgo_ori.py:43"""
    return math.log(1 / (1 - alpha_init) - 1)


def dense_n_samples(grid_size, stepsize):
    """N_samples of the dense sampler, lib/dvgo_ori.py:266"""
    return int(math.sqrt(3 * (grid_size + 1) ** 2) / stepsize) + 1
