"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the scene-branch compositing.

Follows the reference's SPARF-derived NeRF: `NeRF.composite` (lib/bg_nerf/source/models/frequency_nerf.py:283-343) and
`Graph.sample_depth_from_pdf` (lib/bg_nerf/source/models/renderer.py:703-738).  Pinned against the reference's own code
by tests/test_oracle_golden.py (tests/golden/bgnerf_composite.npz, made by tests/golden/make_golden_bgnerf.py).
Only tests/ and bench.py's CPU legs may import this module.
"""
import torch


def composite(ray, rgb_samples, density_samples, depth_samples, setbg_opaque=False):
    """ray [B,R,3], rgb_samples [B,R,N,3] (after sigmoid), density_samples [B,R,N] (after softplus), depth_samples
    [B,R,N,1] -> dict(rgb, rgb_var, depth, depth_var, opacity, weights, all_cumulated) (frequency_nerf.py:283-343)."""
    ray_length = ray.norm(dim=-1, keepdim=True)                                                  # :305
    intv = depth_samples[..., 1:, 0] - depth_samples[..., :-1, 0]                                 # :309
    intv = torch.cat([intv, torch.full_like(intv[..., :1], 1e10)], dim=2)                         # :310
    sigma_delta = density_samples * (intv * ray_length)                                           # :311-313
    alpha = 1 - torch.exp(-sigma_delta)                                                           # :315
    T = torch.exp(-torch.cat([torch.zeros_like(sigma_delta[..., :1]), sigma_delta[..., :-1]], dim=2).cumsum(dim=2))
    all_cumulated = T[:, :, -2].clone()                                                           # :320
    weights = (T * alpha)[..., None]                                                              # :322
    depth = (depth_samples * weights).sum(dim=2)                                                  # :330
    depth_var = (weights * (depth_samples - depth.unsqueeze(-1)) ** 2).sum(dim=2)                 # :331
    rgb = (rgb_samples * weights).sum(dim=2)                                                      # :335
    rgb_var = ((rgb_samples - rgb.unsqueeze(-2)).sum(dim=-1, keepdims=True) * weights).sum(dim=2)  # :336
    opacity = weights.sum(dim=2)                                                                  # :338
    if setbg_opaque:
        rgb = rgb + (1. - opacity)                                                                # :340-341
    return dict(rgb=rgb, rgb_var=rgb_var, depth=depth, depth_var=depth_var, opacity=opacity, weights=weights,
                all_cumulated=all_cumulated)


def sample_depth_from_pdf(weights, n_samples_coarse, n_samples_fine, depth_range, grid):
    """weights [B,R,N]; grid [Nf+1] = linspace(0,1,Nf+1) (det) or the reference's ONE shared torch.rand(Nf+1) draw
    (renderer.py:717-721) -> depth samples [B,R,Nf,1] (renderer.py:703-738)."""
    depth_min, depth_max = depth_range
    pdf = weights / (weights.sum(dim=-1, keepdims=True) + 1e-6)
    cdf = pdf.cumsum(dim=-1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
    unif = (0.5 * (grid[:-1] + grid[1:])).repeat(*cdf.shape[:-1], 1)
    idx = torch.searchsorted(cdf, unif, right=True)
    depth_bin = torch.linspace(depth_min, depth_max, n_samples_coarse + 1, device=weights.device).repeat(*cdf.shape[:-1], 1)
    lo, hi = (idx - 1).clamp(min=0), idx.clamp(max=n_samples_coarse)
    depth_low, depth_high = depth_bin.gather(2, lo), depth_bin.gather(2, hi)
    cdf_low, cdf_high = cdf.gather(2, lo), cdf.gather(2, hi)
    t = (unif - cdf_low) / (cdf_high - cdf_low + 1e-8)
    return (depth_low + t * (depth_high - depth_low))[..., None], idx
