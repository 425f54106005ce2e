"""Drop-in for the reference's JIT-built ``render_utils_cuda`` extension.

Same 13 function names, positional argument order and return arity as the pybind module in
``lib/cuda/render_utils.cpp:170-184`` -- so ``lib/grid.py:12-17`` / ``lib/voxurf_coarse.py:19-24``
can replace their ``load(name='render_utils_cuda', ...)`` by
``from poseprobe_b200 import render_utils_cuda``.  Each function allocates the outputs with torch
(on the inputs' device, current stream) and calls the C ABI in ``include/poseprobe_b200.h``.

Error behaviour mirrors the reference: non-CUDA / non-contiguous inputs raise ``RuntimeError`` with
the ``CHECK_INPUT`` texts; there is no CPU path.
"""
import torch

from . import _lib
from ._lib import check, check_input, dtype_code, lib, ptr, stream

_pinned_total = None


def _opts(t):
    return dict(dtype=t.dtype, device=t.device)


def infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far):
    """-> [t_min, t_max]   (render_utils.cpp:50-58)"""
    for t, n in ((rays_o, "rays_o"), (rays_d, "rays_d"), (xyz_min, "xyz_min"), (xyz_max, "xyz_max")):
        check_input(t, n)
    dt = dtype_code(rays_o, "infer_t_minmax_cuda")
    n = rays_o.size(0)
    t_min = torch.empty(n, **_opts(rays_o))
    t_max = torch.empty(n, **_opts(rays_o))
    with torch.cuda.device(rays_o.device):
        check(lib.pp_infer_t_minmax(dt, ptr(rays_o), ptr(rays_d), ptr(xyz_min), ptr(xyz_max),
                                    near, far, n, ptr(t_min), ptr(t_max), stream()), "infer_t_minmax")
    return [t_min, t_max]


def infer_n_samples(rays_d, t_min, t_max, stepdist):
    """-> int64 [N]   (render_utils.cpp:60-65)"""
    for t, n in ((rays_d, "rays_d"), (t_min, "t_min"), (t_max, "t_max")):
        check_input(t, n)
    dt = dtype_code(t_min, "infer_n_samples_cuda")
    n = t_min.size(0)
    out = torch.empty(n, dtype=torch.int64, device=t_min.device)
    with torch.cuda.device(t_min.device):
        check(lib.pp_infer_n_samples(dt, ptr(rays_d), ptr(t_min), ptr(t_max), stepdist, n, ptr(out),
                                     stream()), "infer_n_samples")
    return out


def infer_ray_start_dir(rays_o, rays_d, t_min):
    """-> [rays_start, rays_dir]   (render_utils.cpp:67-72)"""
    for t, n in ((rays_o, "rays_o"), (rays_d, "rays_d"), (t_min, "t_min")):
        check_input(t, n)
    dt = dtype_code(rays_o, "infer_ray_start_dir_cuda")
    n = rays_o.size(0)
    start = torch.empty_like(rays_o)
    d = torch.empty_like(rays_o)
    with torch.cuda.device(rays_o.device):
        check(lib.pp_infer_ray_start_dir(dt, ptr(rays_o), ptr(rays_d), ptr(t_min), n, ptr(start),
                                         ptr(d), stream()), "infer_ray_start_dir")
    return [start, d]


def _pinned():
    global _pinned_total
    if _pinned_total is None:
        _pinned_total = torch.zeros(1, dtype=torch.int64).pin_memory()
    return _pinned_total


def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    """-> [rays_pts, mask_outbbox, ray_id, step_id, N_steps, t_min, t_max, rays_start, rays_dir]

    The 9-tuple of this reference (render_utils_kernel.cu:241).  Two kernels and one stream
    synchronisation (to learn M); the reference needs six kernels, two cumsums and a sync.
    """
    for t, n in ((rays_o, "rays_o"), (rays_d, "rays_d"), (xyz_min, "xyz_min"), (xyz_max, "xyz_max")):
        check_input(t, n)
    dt = dtype_code(rays_o, "sample_pts_on_rays_cuda")
    dev = rays_o.device
    n = rays_o.size(0)
    i64 = dict(dtype=torch.int64, device=dev)
    t_min = torch.empty(n, **_opts(rays_o))
    t_max = torch.empty(n, **_opts(rays_o))
    n_steps = torch.empty(n, **i64)
    cumsum = torch.empty(n, **i64)
    start = torch.empty_like(rays_o)
    d = torch.empty_like(rays_o)
    scratch = torch.empty(int(lib.pp_sample_scan_scratch_elems(n)) + 1, **i64)
    total_dev = scratch[-1:]
    pinned = _pinned()
    with torch.cuda.device(dev):
        s = stream()
        check(lib.pp_sample_pts_prepare(dt, ptr(rays_o), ptr(rays_d), ptr(xyz_min), ptr(xyz_max),
                                        near, far, stepdist, n, ptr(t_min), ptr(t_max), ptr(n_steps),
                                        ptr(cumsum), ptr(start), ptr(d), ptr(scratch), ptr(total_dev),
                                        ptr(pinned), s), "sample_pts_on_rays(prepare)")
        torch.cuda.current_stream().synchronize()   # the one unavoidable host read-back (M)
        total = int(pinned[0]) if n > 0 else 0
        pts = torch.empty((total, 3), **_opts(rays_o))
        mask = torch.empty(total, dtype=torch.bool, device=dev)
        ray_id = torch.empty(total, **i64)
        step_id = torch.empty(total, **i64)
        check(lib.pp_sample_pts_fill(dt, ptr(start), ptr(d), ptr(xyz_min), ptr(xyz_max), ptr(n_steps),
                                     ptr(cumsum), stepdist, n, total, ptr(pts), ptr(mask), ptr(ray_id),
                                     ptr(step_id), s), "sample_pts_on_rays(fill)")
    return [pts, mask, ray_id, step_id, n_steps, t_min, t_max, start, d]


def sample_ndc_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, N_samples):
    """-> [rays_pts [N,S,3], mask_outbbox [N,S]]   (render_utils.cpp:87-98)"""
    for t, n in ((rays_o, "rays_o"), (rays_d, "rays_d"), (xyz_min, "xyz_min"), (xyz_max, "xyz_max")):
        check_input(t, n)
    dt = dtype_code(rays_o, "sample_ndc_pts_on_rays_cuda")
    n = rays_o.size(0)
    pts = torch.empty((n, N_samples, 3), **_opts(rays_o))
    mask = torch.empty((n, N_samples), dtype=torch.bool, device=rays_o.device)
    with torch.cuda.device(rays_o.device):
        check(lib.pp_sample_ndc_pts_on_rays(dt, ptr(rays_o), ptr(rays_d), ptr(xyz_min), ptr(xyz_max),
                                            N_samples, n, ptr(pts), ptr(mask), stream()),
              "sample_ndc_pts_on_rays")
    return [pts, mask]


def sample_bg_pts_on_rays(rays_o, rays_d, t_max, bg_preserve, N_samples):
    """-> rays_pts [N,S,3]   (render_utils.cpp:100-107)"""
    for t, n in ((rays_o, "rays_o"), (rays_d, "rays_d"), (t_max, "t_max")):
        check_input(t, n)
    dt = dtype_code(rays_o, "sample_bg_pts_on_rays_cuda")
    n = rays_o.size(0)
    pts = torch.empty((n, N_samples, 3), **_opts(rays_o))
    with torch.cuda.device(rays_o.device):
        check(lib.pp_sample_bg_pts_on_rays(dt, ptr(rays_o), ptr(rays_d), ptr(t_max), bg_preserve,
                                           N_samples, n, ptr(pts), stream()), "sample_bg_pts_on_rays")
    return pts


def maskcache_lookup(world, xyz, xyz2ijk_scale, xyz2ijk_shift):
    """-> bool [n]   (render_utils.cpp:109-118)"""
    for t, n in ((world, "world"), (xyz, "xyz"), (xyz2ijk_scale, "xyz2ijk_scale"),
                 (xyz2ijk_shift, "xyz2ijk_shift")):
        check_input(t, n)
    dt = dtype_code(xyz, "maskcache_lookup_cuda")
    if world.dtype != torch.bool:
        raise RuntimeError("expected scalar type Bool but found %s" % str(world.dtype).replace("torch.", ""))
    n = xyz.size(0)
    out = torch.empty(n, dtype=torch.bool, device=xyz.device)
    with torch.cuda.device(xyz.device):
        check(lib.pp_maskcache_lookup(dt, ptr(world), ptr(xyz), ptr(xyz2ijk_scale), ptr(xyz2ijk_shift),
                                      world.size(0), world.size(1), world.size(2), n, ptr(out),
                                      stream()), "maskcache_lookup")
    return out


def raw2alpha(density, shift, interval):
    """-> [exp_d, alpha]   (render_utils.cpp:120-124)"""
    check_input(density, "density")
    dt = dtype_code(density, "raw2alpha_cuda")
    exp_d = torch.empty_like(density)
    alpha = torch.empty_like(density)
    with torch.cuda.device(density.device):
        check(lib.pp_raw2alpha(dt, ptr(density), shift, interval, None, density.size(0), ptr(exp_d),
                               ptr(alpha), stream()), "raw2alpha")
    return [exp_d, alpha]


def raw2alpha_nonuni(density, shift, interval):
    """-> [exp_d, alpha] with a per-point interval tensor   (render_utils.cpp:125-129)"""
    check_input(density, "density")
    dt = dtype_code(density, "raw2alpha_cuda")
    exp_d = torch.empty_like(density)
    alpha = torch.empty_like(density)
    with torch.cuda.device(density.device):
        check(lib.pp_raw2alpha(dt, ptr(density), shift, 0.0, ptr(interval), density.size(0),
                               ptr(exp_d), ptr(alpha), stream()), "raw2alpha_nonuni")
    return [exp_d, alpha]


def raw2alpha_backward(exp, grad_back, interval):
    """-> grad   (render_utils.cpp:131-135)"""
    check_input(exp, "exp")
    check_input(grad_back, "grad_back")
    dt = dtype_code(exp, "raw2alpha_backward_cuda")
    grad = torch.empty_like(exp)
    with torch.cuda.device(exp.device):
        check(lib.pp_raw2alpha_backward(dt, ptr(exp), ptr(grad_back), interval, None, exp.size(0),
                                        ptr(grad), stream()), "raw2alpha_backward")
    return grad


def raw2alpha_nonuni_backward(exp, grad_back, interval):
    """-> grad   (render_utils.cpp:136-140)"""
    check_input(exp, "exp")
    check_input(grad_back, "grad_back")
    dt = dtype_code(exp, "raw2alpha_backward_cuda")
    grad = torch.empty_like(exp)
    with torch.cuda.device(exp.device):
        check(lib.pp_raw2alpha_backward(dt, ptr(exp), ptr(grad_back), 0.0, ptr(interval), exp.size(0),
                                        ptr(grad), stream()), "raw2alpha_nonuni_backward")
    return grad


def alpha2weight(alpha, ray_id, n_rays):
    """-> [weight, T, alphainv_last, i_start, i_end]   (render_utils.cpp:142-149)"""
    check_input(alpha, "alpha")
    check_input(ray_id, "ray_id")
    dt = dtype_code(alpha, "alpha2weight_cuda")
    if ray_id.dtype != torch.int64:
        raise RuntimeError("expected scalar type Long but found %s" % str(ray_id.dtype).replace("torch.", ""))
    dev = alpha.device
    n_pts = alpha.size(0)
    weight = torch.empty_like(alpha)
    T = torch.empty_like(alpha)
    last = torch.empty(n_rays, **_opts(alpha))
    i_start = torch.empty(n_rays, dtype=torch.int64, device=dev)
    i_end = torch.empty(n_rays, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib.pp_alpha2weight(dt, ptr(alpha), ptr(ray_id), n_pts, n_rays, ptr(weight), ptr(T),
                                  ptr(last), ptr(i_start), ptr(i_end), stream()), "alpha2weight")
    return [weight, T, last, i_start, i_end]


def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights,
                          grad_last):
    """-> grad   (render_utils.cpp:151-167)"""
    for t, n in ((alpha, "alpha"), (weight, "weight"), (T, "T"), (alphainv_last, "alphainv_last"),
                 (i_start, "i_start"), (i_end, "i_end"), (grad_weights, "grad_weights"),
                 (grad_last, "grad_last")):
        check_input(t, n)
    dt = dtype_code(alpha, "alpha2weight_backward_cuda")
    grad = torch.empty_like(alpha)
    with torch.cuda.device(alpha.device):
        check(lib.pp_alpha2weight_backward(dt, ptr(alpha), ptr(weight), ptr(T), ptr(alphainv_last),
                                           ptr(i_start), ptr(i_end), n_rays, alpha.size(0),
                                           ptr(grad_weights), ptr(grad_last), ptr(grad), stream()),
              "alpha2weight_backward")
    return grad
