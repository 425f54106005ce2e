"""Host side of the fused DVGO render / train path (C ABI: include/poseprobe_b200_fused.h).

Mirrors the coarse-stage ``DirectVoxGO`` of the reference (lib/dvgo_ori.py:14-435: ``density``
[1,1,X,Y,Z] + ``k0`` [1,3,X,Y,Z], post-activation density, ``sigmoid(k0)`` colour) behind two faces:

* ``DirectVoxGO`` -- an ``nn.Module`` with the reference's constructor arguments, parameter names
  (``density``, ``k0``; checkpoints interchange) and ``forward(rays_o, rays_d, viewdirs, global_step,
  **render_kwargs)`` returning the reference's result dict.  Forward/backward are ONE kernel each.
* ``FusedDVGO`` -- the training engine: voxel records, gradients and Adam moments stay resident in HBM
  as float4 records; ``train_step`` is two kernel launches (march+loss+backward, masked Adam).
"""
import ctypes
import math

import torch
import torch.nn as nn

from ._lib import check, lib, ptr, stream


class _Scene(ctypes.Structure):
    _fields_ = [("sx", ctypes.c_int32), ("sy", ctypes.c_int32), ("sz", ctypes.c_int32),
                ("xyz_min", ctypes.c_float * 3), ("xyz_max", ctypes.c_float * 3),
                ("near", ctypes.c_float), ("far", ctypes.c_float), ("stepdist", ctypes.c_float),
                ("act_shift", ctypes.c_float), ("interval", ctypes.c_float), ("bg", ctypes.c_float),
                ("fast_color_thres", ctypes.c_float)]


class _Loss(ctypes.Structure):
    _fields_ = [("w_rgb", ctypes.c_float), ("w_entropy_last", ctypes.c_float), ("w_rgbper", ctypes.c_float),
                ("mse_scale", ctypes.c_float)]


def make_scene(world_size, xyz_min, xyz_max, near, far, stepdist, act_shift, interval, bg=0.0,
               fast_color_thres=0.0):
    s = _Scene()
    s.sx, s.sy, s.sz = [int(v) for v in world_size]
    for c in range(3):
        s.xyz_min[c] = float(xyz_min[c])
        s.xyz_max[c] = float(xyz_max[c])
    s.near, s.far, s.stepdist = float(near), float(far), float(stepdist)
    s.act_shift, s.interval, s.bg = float(act_shift), float(interval), float(bg)
    s.fast_color_thres = float(fast_color_thres)
    return s


def _need_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32" % name)


def pack_voxels(density, k0):
    """density [1,1,X,Y,Z], k0 [1,3,X,Y,Z] -> float4 records [X*Y*Z, 4]."""
    _need_cuda(density, "density")
    _need_cuda(k0, "k0")
    if k0.shape[1] != 3 or density.shape[1] != 1 or density.shape[2:] != k0.shape[2:]:
        raise RuntimeError("pack_voxels expects density [1,1,X,Y,Z] and k0 [1,3,X,Y,Z]")
    n = density.numel()
    vox = torch.empty((n, 4), dtype=torch.float32, device=density.device)
    with torch.cuda.device(density.device):
        check(lib.pp_voxels_pack(ptr(density), ptr(k0), n, ptr(vox), stream()), "voxels_pack")
    return vox


def unpack_voxels(vox, world_size):
    """float4 records -> (density [1,1,X,Y,Z], k0 [1,3,X,Y,Z])."""
    _need_cuda(vox, "vox")
    n = vox.shape[0]
    x, y, z = [int(v) for v in world_size]
    density = torch.empty((1, 1, x, y, z), dtype=torch.float32, device=vox.device)
    k0 = torch.empty((1, 3, x, y, z), dtype=torch.float32, device=vox.device)
    with torch.cuda.device(vox.device):
        check(lib.pp_voxels_unpack(ptr(vox), n, ptr(density), ptr(k0), stream()), "voxels_unpack")
    return density, k0


_workspaces = {}


def workspace(scene, device):
    """Per-device scratch for the backward / train kernels (grown on demand, reused across calls)."""
    need = int(lib.pp_dvgo_workspace_bytes(ctypes.byref(scene)))
    ws = _workspaces.get(device)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _workspaces[device] = ws
    return ws


def render_forward(scene, vox, rays_o, rays_d, want_depth=True, want_count=False):
    """-> rgb_marched [N,3], alphainv_last [N], depth [N] | None, n_marched [N] | None"""
    for t, nme in ((vox, "vox"), (rays_o, "rays_o"), (rays_d, "rays_d")):
        _need_cuda(t, nme)
    n = rays_o.shape[0]
    dev = rays_o.device
    rgb = torch.empty((n, 3), dtype=torch.float32, device=dev)
    last = torch.empty(n, dtype=torch.float32, device=dev)
    depth = torch.empty(n, dtype=torch.float32, device=dev) if want_depth else None
    cnt = torch.empty(n, dtype=torch.int32, device=dev) if want_count else None
    with torch.cuda.device(dev):
        check(lib.pp_dvgo_render_forward(ctypes.byref(scene), ptr(vox), ptr(rays_o), ptr(rays_d), n, ptr(rgb),
                                         ptr(last), ptr(depth), ptr(cnt), stream()), "dvgo_render_forward")
    return rgb, last, depth, cnt


class _RenderFn(torch.autograd.Function):
    """(density, k0, rays_o, rays_d) -> (rgb_marched, alphainv_last, depth); one kernel each way."""

    @staticmethod
    def forward(ctx, density, k0, rays_o, rays_d, scene):
        vox = pack_voxels(density.contiguous(), k0.contiguous())
        rays_o, rays_d = rays_o.contiguous(), rays_d.contiguous()
        rgb, last, depth, _ = render_forward(scene, vox, rays_o, rays_d)
        ctx.save_for_backward(vox, rays_o, rays_d)
        ctx.scene = scene
        ctx.world = density.shape[2:]
        ctx.mark_non_differentiable(depth)
        return rgb, last, depth

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_rgb, g_last, _g_depth):
        vox, rays_o, rays_d = ctx.saved_tensors
        n = rays_o.shape[0]
        grad_vox = torch.zeros_like(vox)
        need_rays = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        grad_rays = torch.empty((n, 6), dtype=torch.float32, device=vox.device) if need_rays else None
        g_rgb = g_rgb.contiguous() if g_rgb is not None else torch.zeros((n, 3), device=vox.device)
        g_last = g_last.contiguous() if g_last is not None else None
        with torch.cuda.device(vox.device):
            ws = workspace(ctx.scene, vox.device)
            check(lib.pp_dvgo_render_backward(ctypes.byref(ctx.scene), ptr(vox), ptr(rays_o), ptr(rays_d), n,
                                              ptr(g_rgb), ptr(g_last), ptr(grad_vox), ptr(grad_rays), ptr(ws),
                                              ws.numel(), stream()), "dvgo_render_backward")
        g_density, g_k0 = unpack_voxels(grad_vox, ctx.world)
        g_o = grad_rays[:, :3].contiguous() if need_rays else None
        g_d = grad_rays[:, 3:].contiguous() if need_rays else None
        return g_density, g_k0, g_o, g_d, None


class DirectVoxGO(nn.Module):
    """Coarse-stage DirectVoxGO (lib/dvgo_ori.py:14-131) on the fused kernels.

    Supported configuration: ``rgbnet_dim=0`` (colour voxel grid), post-activation density,
    no mask cache -- the configuration BASELINE.json's config B names.
    """

    def __init__(self, xyz_min, xyz_max, num_voxels=0, num_voxels_base=0, alpha_init=None,
                 fast_color_thres=0, rgbnet_dim=0, **kwargs):
        super().__init__()
        if rgbnet_dim > 0:
            raise NotImplementedError("rgbnet (fine stage) is served by the op-level API, not the fused module")
        self.register_buffer('xyz_min', torch.Tensor(xyz_min))
        self.register_buffer('xyz_max', torch.Tensor(xyz_max))
        self.fast_color_thres = fast_color_thres
        self.num_voxels_base = num_voxels_base
        self.voxel_size_base = ((self.xyz_max - self.xyz_min).prod() / self.num_voxels_base).pow(1 / 3)
        self.alpha_init = alpha_init
        self.act_shift = math.log(1 / (1 - alpha_init) - 1)
        self._set_grid_resolution(num_voxels)
        self.density = nn.Parameter(torch.zeros([1, 1, *self.world_size]))
        self.k0_dim = 3
        self.k0 = nn.Parameter(torch.zeros([1, self.k0_dim, *self.world_size]))
        self.rgbnet = None

    def _set_grid_resolution(self, num_voxels):
        # lib/dvgo_ori.py:115-125
        self.num_voxels = num_voxels
        self.voxel_size = ((self.xyz_max - self.xyz_min).prod() / num_voxels).pow(1 / 3)
        self.world_size = ((self.xyz_max - self.xyz_min) / self.voxel_size).long()
        self.voxel_size_ratio = self.voxel_size / self.voxel_size_base

    def scene(self, near, far, stepsize, bg):
        return make_scene(self.world_size.tolist(), self.xyz_min.tolist(), self.xyz_max.tolist(), near, far,
                          stepsize * float(self.voxel_size), self.act_shift,
                          stepsize * float(self.voxel_size_ratio), bg, self.fast_color_thres)

    def forward(self, rays_o, rays_d, viewdirs=None, global_step=None, **render_kwargs):
        """Volume rendering; keys of the reference's result dict that exist per ray are returned."""
        sc = self.scene(render_kwargs['near'], render_kwargs['far'], render_kwargs['stepsize'],
                        float(render_kwargs.get('bg', 0.0)))
        rgb, last, depth = _RenderFn.apply(self.density, self.k0, rays_o, rays_d, sc)
        return {'alphainv_cum': last, 'rgb_marched': rgb, 'depth': depth, 'disp': 1 / depth}


class FusedDVGO:
    """Resident-state DVGO training engine (voxel records + gradient + Adam moments in HBM)."""

    def __init__(self, density, k0, xyz_min, xyz_max, stepsize, voxel_size=None, voxel_size_ratio=1.0,
                 alpha_init=1e-2, near=0.2, far=1e9, bg=0.0, fast_color_thres=0.0,
                 lr_density=0.1, lr_k0=0.1, betas=(0.9, 0.99), eps=1e-8,
                 w_rgb=1.0, w_entropy_last=0.0, w_rgbper=0.0, tv_density=0.0, tv_k0=0.0, tv_dense=False,
                 pose_grad=False, vox_buf=None, grad_buf=None, moment_records=None, mse_scale=1.0):
        self.world_size = tuple(int(v) for v in density.shape[2:])
        if voxel_size is None:
            voxel_size = float((xyz_max[0] - xyz_min[0]) / self.world_size[0])
        self.scene = make_scene(self.world_size, xyz_min, xyz_max, near, far, stepsize * voxel_size,
                                math.log(1 / (1 - alpha_init) - 1), stepsize * voxel_size_ratio, bg,
                                fast_color_thres)
        self.loss = _Loss(float(w_rgb), float(w_entropy_last), float(w_rgbper), float(mse_scale))
        self.vox = pack_voxels(density.contiguous(), k0.contiguous())
        if vox_buf is not None:          # caller-owned (e.g. IPC-shared) parameter / gradient records
            vox_buf.copy_(self.vox)
            self.vox = vox_buf
        self.grad = grad_buf if grad_buf is not None else torch.zeros_like(self.vox)
        self.grad.zero_()
        n_mom = self.vox.shape[0] if moment_records is None else int(moment_records)
        self.exp_avg = torch.zeros((n_mom, 4), dtype=torch.float32, device=self.vox.device)
        self.exp_avg_sq = torch.zeros_like(self.exp_avg)
        self.tv = (float(tv_density), float(tv_k0), bool(tv_dense))
        self.vox_next = torch.empty_like(self.vox) if (tv_density or tv_k0) else None
        self.hp = (float(betas[0]), float(betas[1]), float(lr_density), float(lr_k0), float(eps))
        self.step = 0
        self.pose_grad = pose_grad
        self.device = self.vox.device
        self.loss_buf = torch.zeros(2, dtype=torch.float32, device=self.device)
        self.n_voxels = self.vox.shape[0]
        self.ws = workspace(self.scene, self.device)

    # -- reference-layout views (checkpoint interchange) --
    def export_grids(self):
        return unpack_voxels(self.vox, self.world_size)

    def export_grad(self):
        return unpack_voxels(self.grad, self.world_size)

    def render(self, rays_o, rays_d, **kw):
        return render_forward(self.scene, self.vox, rays_o, rays_d, **kw)

    def backward_only(self, rays_o, rays_d, target, rgb_out=None, grad_rays=None):
        """march + loss + scatter into self.grad (no optimiser step)."""
        n = rays_o.shape[0]
        with torch.cuda.device(self.device):
            check(lib.pp_dvgo_train_step(ctypes.byref(self.scene), ctypes.byref(self.loss), ptr(self.vox),
                                         ptr(rays_o), ptr(rays_d), ptr(target), n, ptr(self.grad), ptr(rgb_out),
                                         ptr(self.loss_buf), ptr(grad_rays), ptr(self.ws), self.ws.numel(),
                                         stream()), "dvgo_train_step")

    def optimizer_step(self):
        b1, b2, lr_d, lr_k, eps = self.hp
        self.step += 1
        tvd, tvk, dense = self.tv
        with torch.cuda.device(self.device):
            if self.vox_next is not None:
                check(lib.pp_voxel_tv_adam(ctypes.byref(self.scene), ptr(self.vox), ptr(self.vox_next),
                                           ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.step, b1,
                                           b2, lr_d, lr_k, eps, tvd, tvk, int(dense), stream()), "voxel_tv_adam")
                self.vox, self.vox_next = self.vox_next, self.vox
            else:
                check(lib.pp_voxel_adam(ptr(self.vox), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                        self.n_voxels, self.step, b1, b2, lr_d, lr_k, eps, 1, stream()),
                      "voxel_adam")

    def train_step(self, rays_o, rays_d, target, rgb_out=None, grad_rays=None):
        """One optimisation step on a ray batch.  Returns the device tensor [loss, mse] (no sync)."""
        self.loss_buf.zero_()
        self.backward_only(rays_o, rays_d, target, rgb_out, grad_rays)
        self.optimizer_step()
        return self.loss_buf


def smoke(dev):
    """Tiny fused train step (used by __graft_entry__.smoke)."""
    from . import synthetic
    G, N = 24, 128
    density, k0 = synthetic.make_dvgo_grids(G)
    o, d, _ = synthetic.make_rays(N)
    eng = FusedDVGO(density.to(dev), k0.to(dev), [-1, -1, -1], [1, 1, 1], stepsize=0.5)
    tgt = torch.rand(N, 3, device=dev)
    loss = eng.train_step(o.to(dev), d.to(dev), tgt)
    torch.cuda.synchronize()
    assert torch.isfinite(loss).all() and float(loss[0]) > 0
    return float(loss[0])
