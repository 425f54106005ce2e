"""Scene-branch compositing behind the reference's own method signatures.

The reference's scene branch is a SPARF-derived NeRF rendered next to the object grids in every training step
(lib/recon_scene.py:645-649, loss = 0.1 * object + scene).  Its volume rendering is `NeRF.composite`
(lib/bg_nerf/source/models/frequency_nerf.py:283-343) and its hierarchical resampling `Graph.sample_depth` /
`Graph.sample_depth_from_pdf` (lib/bg_nerf/source/models/renderer.py:669-738): ~25 elementwise / cumsum / gather torch
kernels over dense [B, R, N] tensors plus autograd's mirror image of them.  Here each is ONE streaming kernel
(csrc/scene.cu, a warp per ray) behind the same argument names and result keys.  No CPU fallback: CPU tensors raise.
"""
import torch

from ._lib import check, lib, ptr, stream


def _need_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (poseprobe_b200 has no CPU fallback)" % name)
    if t.dtype != torch.float32:
        raise RuntimeError('"nerf_composite" not implemented for %s' % str(t.dtype))


class NerfComposite(torch.autograd.Function):
    """(ray [B,R,3], rgb_samples [B,R,N,3], density_samples [B,R,N], depth_samples [B,R,N,1]) ->
    (rgb [B,R,3], rgb_var [B,R,1], depth [B,R,1], depth_var [B,R,1], opacity [B,R,1], weights [B,R,N,1],
    all_cumulated [B,R]).  Differentiable in rgb, depth, opacity and weights w.r.t. rgb_samples, density_samples and
    ray (what the reference's losses reach, lib/bg_nerf/source/training/core/base_losses.py); rgb_var, depth_var and
    all_cumulated are diagnostics and marked non-differentiable; depth_samples are produced under no_grad in the
    reference (renderer.py:603-611) and get no gradient."""

    @staticmethod
    def forward(ctx, ray, rgb_samples, density_samples, depth_samples, setbg_opaque):
        for t, nme in ((ray, "ray"), (rgb_samples, "rgb_samples"), (density_samples, "density_samples"),
                       (depth_samples, "depth_samples")):
            _need_cuda(t, nme)
        B, R, N = density_samples.shape
        ray_c, rgb_c, den_c = ray.contiguous(), rgb_samples.contiguous(), density_samples.contiguous()
        dep_c = depth_samples.reshape(B, R, N).contiguous()
        dev = ray.device
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        rgb, rgb_var, depth, depth_var, opacity = f(B, R, 3), f(B, R, 1), f(B, R, 1), f(B, R, 1), f(B, R, 1)
        weights, all_cum = f(B, R, N, 1), f(B, R)
        with torch.cuda.device(dev):
            check(lib.pp_nerf_composite_forward(ptr(ray_c), ptr(rgb_c), ptr(den_c), ptr(dep_c), B * R, N,
                                                int(bool(setbg_opaque)), ptr(rgb), ptr(rgb_var), ptr(depth),
                                                ptr(depth_var), ptr(opacity), ptr(weights), ptr(all_cum), stream()),
                  "nerf_composite_forward")
        ctx.save_for_backward(ray_c, rgb_c, den_c, dep_c)
        ctx.opaque = int(bool(setbg_opaque))
        ctx.mark_non_differentiable(rgb_var, depth_var, all_cum)
        return rgb, rgb_var, depth, depth_var, opacity, weights, all_cum

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_rgb, _g_rgb_var, g_depth, _g_depth_var, g_opacity, g_weights, _g_all):
        ray_c, rgb_c, den_c, dep_c = ctx.saved_tensors
        B, R, N = den_c.shape
        need_ray, need_rgb, need_den = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        c = lambda t: None if t is None else t.contiguous()  # noqa: E731
        g_rgb, g_depth, g_opacity, g_weights = c(g_rgb), c(g_depth), c(g_opacity), c(g_weights)
        grad_rgb_s = torch.empty_like(rgb_c) if need_rgb else None
        grad_den = torch.empty_like(den_c) if need_den else None
        grad_ray = torch.empty_like(ray_c) if need_ray else None
        with torch.cuda.device(ray_c.device):
            check(lib.pp_nerf_composite_backward(ptr(ray_c), ptr(rgb_c), ptr(den_c), ptr(dep_c), B * R, N, ctx.opaque,
                                                 ptr(g_rgb), ptr(g_depth), ptr(g_opacity), ptr(g_weights),
                                                 ptr(grad_rgb_s), ptr(grad_den), ptr(grad_ray), stream()),
                  "nerf_composite_backward")
        return grad_ray, grad_rgb_s, grad_den, None, None


def composite(opt, ray, pred_dict, depth_samples):
    """`NeRF.composite(opt, ray, pred_dict, depth_samples)` (frequency_nerf.py:283-343): same arguments, updates and
    returns `pred_dict` with rgb, rgb_var, depth, depth_var, opacity, weights, all_cumulated.  `opt` needs
    `opt.nerf.setbg_opaque` and `opt.mask_img` (frequency_nerf.py:340)."""
    opaque = bool(getattr(getattr(opt, "nerf", None), "setbg_opaque", False) or getattr(opt, "mask_img", False))
    rgb, rgb_var, depth, depth_var, opacity, weights, all_cum = NerfComposite.apply(
        ray, pred_dict["rgb_samples"], pred_dict["density_samples"], depth_samples, opaque)
    pred_dict.update(rgb=rgb, rgb_var=rgb_var, depth=depth, depth_var=depth_var, opacity=opacity, weights=weights,
                     all_cumulated=all_cum)
    return pred_dict


def sample_depth(n_samples, depth_range, batch_size, num_rays, device, stratified=False, param="metric", rand=None):
    """`Graph.sample_depth` (renderer.py:669-701): one depth per interval of the range, mid points or stratified
    (`rand` [B,R,N,1]: the caller's draw, defaults to torch.rand on `device`) -> [B,R,N,1].  Elementwise on N values
    per ray: plain torch, nothing to fuse."""
    depth_min, depth_max = depth_range
    if stratified:
        r = torch.rand(batch_size, num_rays, n_samples, 1, device=device) if rand is None else rand.to(device)
    else:
        r = 0.5 * torch.ones(batch_size, num_rays, n_samples, 1, device=device)
    r = r + torch.arange(n_samples, device=device)[None, None, :, None].float()
    d = r / n_samples * (depth_max - depth_min) + depth_min
    return d if param == "metric" else 1 / (d + 1e-8)


@torch.no_grad()
def sample_depth_from_pdf(weights, n_samples_coarse, n_samples_fine, depth_range, det, grid=None):
    """`Graph.sample_depth_from_pdf` (renderer.py:703-738): weights [B,R,N] -> fine depths [B,R,Nf,1].  `grid`
    ([Nf+1], optional) replaces the reference's `torch.rand(n_samples_fine + 1)` draw (one draw shared by every ray)
    so a caller or a test can fix it; det=True uses linspace(0, 1, Nf + 1) as the reference does."""
    _need_cuda(weights, "weights")
    B, R, N = weights.shape
    assert N == n_samples_coarse
    dev = weights.device
    if grid is None:
        # renderer.py:717-721: linspace on the model's device, the random grid drawn on the CPU and moved over
        grid = torch.linspace(0, 1, n_samples_fine + 1, device=dev) if det else torch.rand(n_samples_fine + 1)
    grid = grid.to(dev, torch.float32)
    unif = (0.5 * (grid[:-1] + grid[1:])).contiguous()
    out = torch.empty(B, R, n_samples_fine, 1, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(lib.pp_nerf_sample_depth_from_pdf(ptr(weights.contiguous()), B * R, N, ptr(unif), n_samples_fine,
                                                float(depth_range[0]), float(depth_range[1]), ptr(out), stream()),
              "nerf_sample_depth_from_pdf")
    return out


@torch.no_grad()
def merge_depth_samples(depth_samples, depth_samples_fine):
    """`torch.cat([depth_samples, depth_samples_fine], dim=2).sort(dim=2).values` (renderer.py:612-614): the coarse and
    the resampled depths of every ray, [B,R,N,1] and [B,R,Nf,1] -> [B,R,N+Nf,1] ascending, in one kernel."""
    _need_cuda(depth_samples, "depth_samples")
    _need_cuda(depth_samples_fine, "depth_samples_fine")
    B, R, N = depth_samples.shape[:3]
    Nf = depth_samples_fine.shape[2]
    a = depth_samples.reshape(B * R, N).contiguous()
    b = depth_samples_fine.reshape(B * R, Nf).contiguous()
    out = torch.empty(B, R, N + Nf, 1, device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        check(lib.pp_nerf_sort_depth(ptr(a), N, ptr(b), Nf, B * R, ptr(out), stream()), "nerf_sort_depth")
    return out
