"""Seeded input builders shared by the CPU and GPU test tiers."""
import torch

from poseprobe_b200 import synthetic


def rays_with_edge_cases(n, seed=0):
    """Synthetic camera rays plus the corner cases the sampling kernels branch on."""
    o, d, _ = synthetic.make_rays(n, seed=777 + seed)
    g = torch.Generator().manual_seed(1234 + seed)
    if n >= 16:
        d[0, 0] = 0.0                      # zero direction component -> 1e-6 substitution
        d[1] = torch.tensor([0.0, 0.0, 1.0])
        o[1] = torch.tensor([0.3, -0.2, -3.0])
        o[2] = torch.tensor([5.0, 5.0, 5.0])   # misses the box entirely
        d[2] = torch.tensor([1.0, 0.0, 0.0])
        o[3] = torch.tensor([0.1, 0.2, -0.3])  # origin inside the box
        d[4] = d[4] * 3.7                  # non-unit direction
        d[5] = d[5] * 0.01
        o[6] = torch.tensor([-1.0, 0.0, -2.0])  # grazes a face
        d[6] = torch.tensor([0.0, 0.0, 1.0])
        # a handful of fully random (mostly missing) rays
        o[7:12] = torch.randn(5, 3, generator=g) * 3
        d[7:12] = torch.randn(5, 3, generator=g)
    return o.contiguous(), d.contiguous()


def bbox():
    return torch.tensor([-1.0, -1.0, -1.0]), torch.tensor([1.0, 1.0, 1.0])


def bbox_aniso():
    return torch.tensor([-1.1, -0.7, -0.9]), torch.tensor([0.8, 1.2, 1.05])


def alpha_and_ray_id(n_rays, max_len, seed, p_big=0.0, drop_rays=True):
    """Random per-ray segment lengths (some rays empty), sorted ray_id, alpha in [0,1)."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(0, max_len + 1, (n_rays,), generator=g)
    if drop_rays and n_rays > 4:
        lens[torch.randint(0, n_rays, (max(1, n_rays // 7),), generator=g)] = 0
    ray_id = torch.repeat_interleave(torch.arange(n_rays), lens)
    m = int(lens.sum())
    alpha = torch.rand(m, generator=g) * 0.05
    if p_big > 0:
        big = torch.rand(m, generator=g) < p_big
        alpha = torch.where(big, torch.rand(m, generator=g), alpha)
    if m > 8:
        alpha[3] = 0.0
        alpha[5] = 1.0          # T drops to exactly 0 -> early termination
        alpha[7] = 1e-12        # below the fp64 exactness threshold of 1 - alpha
    return alpha.contiguous(), ray_id.contiguous(), lens
