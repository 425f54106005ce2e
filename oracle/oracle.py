"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle.

Two halves:

* thin ctypes wrappers over ``oracle/libpp_oracle.so`` (``pp_oracle.c``: the plain-C restatement of
  the reference's ``lib/cuda`` kernels), operating on CPU ``torch`` tensors and returning tensors
  with the reference's shapes / dtypes / return arity;
* pure-PyTorch CPU restatements of the operators the reference composes in Python
  (``lib/grid.py``, ``lib/voxurf_coarse.py``, ``lib/dvgo_ori.py``), each citing the lines followed.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product (``poseprobe_b200``) never does.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(HERE, "libpp_oracle.so")
_lib = None


def build():
    """Compile pp_oracle.c with gcc (seconds)."""
    src = os.path.join(HERE, "pp_oracle.c")
    if os.path.exists(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src):
        return _LIB_PATH
    subprocess.check_call(["make", "-s", "-C", HERE, "libpp_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_sample_pts_on_rays.restype = ctypes.c_int64
    return _lib


def _p(t):
    if t is None:
        return ctypes.c_void_p(0)
    assert t.device.type == "cpu" and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


_f = ctypes.c_float
_i64 = ctypes.c_int64
_i32 = ctypes.c_int32


def _f32(t):
    return t.detach().to("cpu", torch.float32).contiguous()


# ---------------------------------------------------------------------------------------------
# render_utils_cuda
# ---------------------------------------------------------------------------------------------
def infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far):
    rays_o, rays_d, xyz_min, xyz_max = map(_f32, (rays_o, rays_d, xyz_min, xyz_max))
    n = rays_o.shape[0]
    t_min, t_max = torch.empty(n), torch.empty(n)
    lib().orc_infer_t_minmax(_p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max), _f(near), _f(far),
                             _i64(n), _p(t_min), _p(t_max))
    return t_min, t_max


def infer_n_samples(rays_d, t_min, t_max, stepdist):
    rays_d, t_min, t_max = map(_f32, (rays_d, t_min, t_max))
    n = t_min.shape[0]
    out = torch.empty(n, dtype=torch.int64)
    lib().orc_infer_n_samples(_p(rays_d), _p(t_min), _p(t_max), _f(stepdist), _i64(n), _p(out))
    return out


def infer_ray_start_dir(rays_o, rays_d, t_min):
    rays_o, rays_d, t_min = map(_f32, (rays_o, rays_d, t_min))
    n = rays_o.shape[0]
    start, d = torch.empty(n, 3), torch.empty(n, 3)
    lib().orc_infer_ray_start_dir(_p(rays_o), _p(rays_d), _p(t_min), _i64(n), _p(start), _p(d))
    return start, d


def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    """9-tuple like the reference (render_utils_kernel.cu:241)."""
    rays_o, rays_d, xyz_min, xyz_max = map(_f32, (rays_o, rays_d, xyz_min, xyz_max))
    n = rays_o.shape[0]
    t_min, t_max = torch.empty(n), torch.empty(n)
    n_steps = torch.empty(n, dtype=torch.int64)
    start, d = torch.empty(n, 3), torch.empty(n, 3)
    args = (_p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max), _f(near), _f(far), _f(stepdist),
            _i64(n), _p(t_min), _p(t_max), _p(n_steps), _p(start), _p(d))
    m = lib().orc_sample_pts_on_rays(*args, _p(None), _p(None), _p(None), _p(None))
    pts = torch.empty(m, 3)
    mask = torch.empty(m, dtype=torch.bool)
    ray_id = torch.empty(m, dtype=torch.int64)
    step_id = torch.empty(m, dtype=torch.int64)
    lib().orc_sample_pts_on_rays(*args, _p(pts), _p(mask), _p(ray_id), _p(step_id))
    return pts, mask, ray_id, step_id, n_steps, t_min, t_max, start, d


def sample_ndc_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, n_samples):
    rays_o, rays_d, xyz_min, xyz_max = map(_f32, (rays_o, rays_d, xyz_min, xyz_max))
    n = rays_o.shape[0]
    pts = torch.empty(n, n_samples, 3)
    mask = torch.empty(n, n_samples, dtype=torch.bool)
    lib().orc_sample_ndc_pts_on_rays(_p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max),
                                     _i32(n_samples), _i64(n), _p(pts), _p(mask))
    return pts, mask


def sample_bg_pts_on_rays(rays_o, rays_d, t_max, bg_preserve, n_samples):
    rays_o, rays_d, t_max = map(_f32, (rays_o, rays_d, t_max))
    n = rays_o.shape[0]
    pts = torch.empty(n, n_samples, 3)
    lib().orc_sample_bg_pts_on_rays(_p(rays_o), _p(rays_d), _p(t_max), _f(bg_preserve),
                                    _i32(n_samples), _i64(n), _p(pts))
    return pts


def maskcache_lookup(world, xyz, scale, shift):
    world = world.detach().cpu().contiguous()
    xyz, scale, shift = map(_f32, (xyz, scale, shift))
    n = xyz.shape[0]
    out = torch.empty(n, dtype=torch.bool)
    lib().orc_maskcache_lookup(_p(world), _p(xyz), _p(scale), _p(shift), _i32(world.shape[0]),
                               _i32(world.shape[1]), _i32(world.shape[2]), _i64(n), _p(out))
    return out


def raw2alpha(density, shift, interval):
    density = _f32(density)
    exp_d, alpha = torch.empty_like(density), torch.empty_like(density)
    lib().orc_raw2alpha(_p(density), _f(shift), _f(interval), _p(None), _i64(density.numel()),
                        _p(exp_d), _p(alpha))
    return exp_d, alpha


def raw2alpha_nonuni(density, shift, interval):
    density, interval = _f32(density), _f32(interval)
    exp_d, alpha = torch.empty_like(density), torch.empty_like(density)
    lib().orc_raw2alpha(_p(density), _f(shift), _f(0.0), _p(interval), _i64(density.numel()),
                        _p(exp_d), _p(alpha))
    return exp_d, alpha


def raw2alpha_backward(exp_d, grad_back, interval):
    exp_d, grad_back = _f32(exp_d), _f32(grad_back)
    grad = torch.empty_like(exp_d)
    lib().orc_raw2alpha_backward(_p(exp_d), _p(grad_back), _f(interval), _p(None),
                                 _i64(exp_d.numel()), _p(grad))
    return grad


def raw2alpha_nonuni_backward(exp_d, grad_back, interval):
    exp_d, grad_back, interval = _f32(exp_d), _f32(grad_back), _f32(interval)
    grad = torch.empty_like(exp_d)
    lib().orc_raw2alpha_backward(_p(exp_d), _p(grad_back), _f(0.0), _p(interval),
                                 _i64(exp_d.numel()), _p(grad))
    return grad


def alpha2weight(alpha, ray_id, n_rays):
    alpha = _f32(alpha)
    ray_id = ray_id.detach().cpu().to(torch.int64).contiguous()
    m = alpha.numel()
    weight, T = torch.empty(m), torch.empty(m)
    last = torch.empty(n_rays)
    i_start = torch.empty(n_rays, dtype=torch.int64)
    i_end = torch.empty(n_rays, dtype=torch.int64)
    lib().orc_alpha2weight(_p(alpha), _p(ray_id), _i64(m), _i64(n_rays), _p(weight), _p(T), _p(last),
                           _p(i_start), _p(i_end))
    return weight, T, last, i_start, i_end


def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights,
                          grad_last):
    alpha, weight, T, alphainv_last, grad_weights, grad_last = map(
        _f32, (alpha, weight, T, alphainv_last, grad_weights, grad_last))
    i_start = i_start.detach().cpu().contiguous()
    i_end = i_end.detach().cpu().contiguous()
    grad = torch.empty_like(alpha)
    lib().orc_alpha2weight_backward(_p(alpha), _p(weight), _p(T), _p(alphainv_last), _p(i_start),
                                    _p(i_end), _i64(n_rays), _i64(alpha.numel()), _p(grad_weights),
                                    _p(grad_last), _p(grad))
    return grad


# ---------------------------------------------------------------------------------------------
# total_variation_cuda / adam_upd_cuda / ub360_utils_cuda  (in place, like the reference)
# ---------------------------------------------------------------------------------------------
def total_variation_add_grad(param, grad, wx, wy, wz, dense_mode, mask=None):
    assert param.dtype == torch.float32 and grad.dtype == torch.float32
    assert param.is_contiguous() and grad.is_contiguous()
    _, c, x, y, z = param.shape
    m = None if mask is None else _f32(mask)
    lib().orc_total_variation_add_grad(_p(param), _p(grad), _p(m), _f(wx), _f(wy), _f(wz),
                                       ctypes.c_int(bool(dense_mode)), _i64(c), _i64(x), _i64(y), _i64(z))


def adam_upd(mode, param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps):
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    lib().orc_adam_upd(ctypes.c_int(mode), _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq),
                       _p(perlr), _i64(param.numel()), _i32(step), _f(beta1), _f(beta2), _f(lr), _f(eps))


def cumdist_thres(dist, thres):
    dist = _f32(dist)
    mask = torch.empty(dist.shape, dtype=torch.bool)
    lib().orc_cumdist_thres(_p(dist), _f(thres), _i64(dist.shape[0]), _i64(dist.shape[1]), _p(mask))
    return mask


# ---------------------------------------------------------------------------------------------
# Python-side operators of the reference, restated with CPU torch ops
# ---------------------------------------------------------------------------------------------
def dense_grid_forward(grid, xyz, xyz_min, xyz_max, padding_mode="zeros"):
    """DenseGrid.forward, lib/grid.py:47-58 (padding 'border' = Voxurf.grid_sampler, :540)."""
    channels = grid.shape[1]
    shape = xyz.shape[:-1]
    q = xyz.reshape(1, 1, 1, -1, 3)
    ind_norm = ((q - xyz_min) / (xyz_max - xyz_min)).flip((-1,)) * 2 - 1
    out = F.grid_sample(grid, ind_norm, mode="bilinear", align_corners=True, padding_mode=padding_mode)
    out = out.reshape(channels, -1).T.reshape(*shape, channels)
    if channels == 1:
        out = out.squeeze(-1)
    return out


def segment_sum(src, ray_id, n_rays):
    """torch_scatter.segment_coo(src, index, out=zeros, reduce='sum') with a sorted index
    (call sites lib/voxurf_coarse.py:1034-1056, :1179-1200)."""
    out = torch.zeros((n_rays,) + tuple(src.shape[1:]), dtype=src.dtype)
    return out.index_add_(0, ray_id, src)


def segment_cumsum(src, ray_id):
    """Inclusive prefix sum restarted at every change of ray_id."""
    out = torch.empty_like(src)
    if src.numel() == 0:
        return out
    rid = ray_id.numpy()
    heads = np.flatnonzero(np.concatenate([[True], rid[1:] != rid[:-1]]))
    bounds = list(heads) + [len(rid)]
    for a, b in zip(bounds[:-1], bounds[1:]):
        out[a:b] = torch.cumsum(src[a:b], 0)
    return out


def neus_alpha(viewdirs, ray_id, dist, sdf, gradients, inv_s):
    """neus_alpha_from_sdf_scatter, lib/voxurf_coarse.py:483-519 (use_mid, cos_anneal_ratio=1)."""
    dirs = viewdirs[ray_id]
    true_cos = (dirs * gradients).sum(-1, keepdim=True)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - 1.0) + F.relu(-true_cos) * 1.0)
    sdf = sdf.unsqueeze(-1)
    dist_t = torch.as_tensor(dist, dtype=sdf.dtype).reshape(-1, 1)
    est_next = sdf + iter_cos * dist_t * 0.5
    est_prev = sdf - iter_cos * dist_t * 0.5
    inv_s_t = torch.as_tensor(inv_s, dtype=sdf.dtype).reshape(-1, 1)
    prev_cdf = torch.sigmoid(est_prev * inv_s_t)
    next_cdf = torch.sigmoid(est_next * inv_s_t)
    p = prev_cdf - next_cdf
    c = prev_cdf
    return ((p + 1e-5) / (c + 1e-5)).clip(0.0, 1.0).squeeze(-1)


def cumprod_exclusive(p):
    """lib/dvgo_ori.py:478-480"""
    return torch.cat([torch.ones_like(p[..., [0]]), p.clamp_min(1e-10).cumprod(-1)], -1)


def get_ray_marching_ray(alpha):
    """lib/dvgo_ori.py:482-485"""
    alphainv_cum = cumprod_exclusive(1 - alpha)
    weights = alpha * alphainv_cum[..., :-1]
    return weights, alphainv_cum


def total_variation_loss(v):
    """lib/dvgo_ori.py:487-495 without mask"""
    tv2 = v.diff(dim=2).abs()
    tv3 = v.diff(dim=3).abs()
    tv4 = v.diff(dim=4).abs()
    return (tv2.mean() + tv3.mean() + tv4.mean()) / 3


def total_variation(v, mask=None):
    """lib/voxurf_coarse.py:1298-1313, restated literally (mask: bool tensor of v's shape)."""
    tv2 = (v[:, :, 1:, :, :] - v[:, :, :-1, :, :]).abs()
    tv3 = (v[:, :, :, 1:, :] - v[:, :, :, :-1, :]).abs()
    tv4 = (v[:, :, :, :, 1:] - v[:, :, :, :, :-1]).abs()
    if mask is not None:
        tv2 = tv2[mask[:, :, :-1] & mask[:, :, 1:]]
        tv3 = tv3[mask[:, :, :, :-1] & mask[:, :, :, 1:]]
        tv4 = tv4[mask[:, :, :, :, :-1] & mask[:, :, :, :, 1:]]
    if mask is None:
        mask = torch.ones_like(v)
    return (tv2.sum() + tv3.sum() + tv4.sum()) / 3 / mask.sum()


def python_adam_step(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps, per_lr=None):
    """The live optimiser's math, lib/utils.py:176-198 (in place)."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    denom = (exp_avg_sq.sqrt() / (bc2 ** 0.5)).add_(eps)
    step_size = lr / bc1
    if per_lr is not None:
        param.addcdiv_(exp_avg * per_lr, denom, value=-step_size)
    else:
        param.addcdiv_(exp_avg, denom, value=-step_size)


def get_rays_of_a_view(H, W, K, c2w, inverse_y=False, flip_x=False, flip_y=False, mode="center"):
    """get_rays + get_rays_of_a_view (no NDC), lib/voxurf_coarse.py:1339-1368, :1402-1407.
    -> rays_o, rays_d (normalised), viewdirs, each [H,W,3]."""
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i, j = i.t().float(), j.t().float()
    if mode == "center":
        i, j = i + 0.5, j + 0.5
    if flip_x:
        i = i.flip((1,))
    if flip_y:
        j = j.flip((0,))
    if inverse_y:
        dirs = torch.stack([(i - K[0][2]) / K[0][0], (j - K[1][2]) / K[1][1], torch.ones_like(i)], -1)
    else:
        dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, 3].expand(rays_d.shape)
    rays_d = rays_d / rays_d.norm(dim=-1, keepdim=True)
    return rays_o, rays_d, rays_d


def se3_to_SE3(wu, nth=10):
    """Lie.se3_to_SE3, lib/camera.py:127-142 with taylor_A/B/C :165-188."""
    w, u = wu.split([3, 3], dim=-1)
    w0, w1, w2 = w.unbind(dim=-1)
    O = torch.zeros_like(w0)
    wx = torch.stack([torch.stack([O, -w2, w1], dim=-1), torch.stack([w2, O, -w0], dim=-1),
                      torch.stack([-w1, w0, O], dim=-1)], dim=-2)
    theta = w.norm(dim=-1)[..., None, None]
    I = torch.eye(3, dtype=torch.float32)

    def taylor(first):
        ans, denom = torch.zeros_like(theta), 1.
        for i in range(nth + 1):
            if first == "A":
                if i > 0:
                    denom *= (2 * i) * (2 * i + 1)
            elif first == "B":
                denom *= (2 * i + 1) * (2 * i + 2)
            else:
                denom *= (2 * i + 2) * (2 * i + 3)
            ans = ans + (-1) ** i * theta ** (2 * i) / denom
        return ans
    A, B, C = taylor("A"), taylor("B"), taylor("C")
    R = I + A * wx + B * wx @ wx
    V = I + B * wx + C * wx @ wx
    return torch.cat([R, (V @ u[..., None])], dim=-1)
