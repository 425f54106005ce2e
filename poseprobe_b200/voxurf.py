"""Host side of the fused Voxurf fine-stage train / render path (C ABI: include/poseprobe_b200_fused.h).

Mirrors the surface stage of the reference (``Voxurf.forward``, lib/voxurf_coarse.py:922-1091, both branches):
``sdf.grid`` [1,1,X,Y,Z] with the per-step mapping function and the finite-difference gradient grid, ``k0.grid``
[1,12,X,Y,Z], ``rgbnet`` (57 -> 128 x (depth-1) -> 3), NeuS alpha, alpha compositing, the object losses of
lib/losses.py:34-74, and - use_deform=True, the reference's live configuration - the implicit deformation network
with its input Jacobian (forward and the tcgen05 backward of the double backward).
Parameter layouts interchange with the reference's checkpoints (``import_state`` / ``export_state`` use the keys
``sdf.grid``, ``k0.grid``, ``rgbnet.*`` of lib/recon_scene.py:779-791).

``FusedVoxurf.train_step`` is 10 launches: march+features, MLP forward (tcgen05, compositing epilogue), per-ray
loss, MLP forward-recompute+backward (tcgen05), reverse march + scatter, SDF fold + Adam, SDF record rebuild,
feature-record Adam, rgbnet Adam, gradient clears.  With ``pose_grad=True`` the reverse march also emits
d loss / d (rays_o, rays_d, viewdirs) for the SE3 refinement (``ray_gradients`` -> ``camera.rays_from_pixels``'
backward = the fused per-image reduction).  ``FusedVoxurf.train_step(use_deform=True)`` is the live step through the
deformation network; ``Voxurf`` is the nn.Module face with the reference's parameter names and
result dict; ``DistributedVoxurf`` shards rays over the GPUs of a node.
"""
import ctypes
import math

import torch

from ._lib import check, lib, ptr, stream
from .dvgo import _need_cuda, make_scene

FEAT_COLS = 64
DX_COLS = 16


class _Cfg(ctypes.Structure):
    _fields_ = [("inv_s", ctypes.c_float), ("sdf_scale_a", ctypes.c_float), ("sdf_scale_b", ctypes.c_float),
                ("voxel_size", ctypes.c_float), ("pos_pe", ctypes.c_int32), ("view_pe", ctypes.c_int32),
                ("pe_w_pos", ctypes.c_float * 8), ("pe_w_view", ctypes.c_float * 8)]


def barf_weights(progress, barf_c2f, L):
    """positional_encoding_barf's per-frequency weights (lib/voxurf_coarse.py:721-732)."""
    if barf_c2f is None or L == 0:
        return [1.0] * L
    start, end = barf_c2f
    alpha = (progress - start) / (end - start) * L
    return [(1 - math.cos(min(max(alpha - k, 0.0), 1.0) * math.pi)) / 2 for k in range(L)]


def softplus10(x):
    return math.log1p(math.exp(10.0 * x)) / 10.0 if x < 2.0 else x + math.log1p(math.exp(-10.0 * x)) / 10.0


def feature_permutation(pos_pe, view_pe, k0_dim=12):
    """Column of the reference's rgb_feat row (voxurf_coarse.py:1005-1030: k0 | rays_xyz | sin | cos | viewdirs |
    sin | cos | normal) that feeds each of the 64 kernel columns (-1 = zero padding)."""
    if k0_dim != 12 or pos_pe > 5 or view_pe > 1:
        raise RuntimeError("FusedVoxurf: the fused path supports rgbnet_dim=12, posbase_pe<=5, viewbase_pe<=1")
    L, Lv = pos_pe, view_pe
    ref_xyz = k0_dim
    ref_sin = ref_xyz + 3
    ref_cos = ref_sin + 3 * L
    ref_view = ref_cos + 3 * L
    ref_vsin = ref_view + 3
    ref_vcos = ref_vsin + 3 * Lv
    ref_normal = ref_vcos + 3 * Lv
    perm = [-1] * FEAT_COLS
    for c in range(12):
        perm[c] = c
    for c in range(3):
        perm[12 + c] = ref_normal + c
        perm[16 + c] = ref_xyz + c
        perm[49 + c] = ref_view + c
        if Lv:
            perm[52 + c] = ref_vsin + c * Lv
            perm[55 + c] = ref_vcos + c * Lv
        for k in range(L):
            perm[19 + c * 5 + k] = ref_sin + c * L + k
            perm[34 + c * 5 + k] = ref_cos + c * L + k
    return perm, ref_normal + 3


class FusedVoxurf:
    """Training / rendering engine of the SDF fine stage; state stays resident in HBM in the kernels' layouts."""

    def __init__(self, sdf, k0, rgbnet_weights, rgbnet_biases, xyz_min, xyz_max, stepsize=0.5, near=0.2, far=1e9,
                 bg=0.0, s_val=0.05, sdf_alpha=None, sdf_beta=None, posbase_pe=5, viewbase_pe=1, barf_c2f=None,
                 lr_sdf=0.1, lr_k0=0.1, lr_rgbnet=1e-3, betas=(0.9, 0.99), eps=1e-8, w_rgb=1.0,
                 samples_per_ray_cap=None, pose_grad=False, voxel_size=None, adam="reference", lr_warp=1e-3,
                 w_entropy=0.0, w_mask=0.0, w_tv_k0=0.0, w_eikonal=0.0, w_deform=0.0, k0rec_buf=None,
                 grad_flat_buf=None, k0_moment_records=None):
        """adam: "reference" = the optimiser the reference runs (dense Python Adam, lib/utils.py:176-198: every
        element updated each step, eps after the bias correction) or "masked" = masked_adam_upd arithmetic
        (adam_upd_kernel.cu:25-40: elements with a zero gradient are skipped, eps before the bias correction).
        Loss weights of lib/losses.py:34-74 for `train_step`: w_rgb (weight_main), w_entropy (0.01 in the reference),
        w_mask (weight_mask), w_tv_k0, w_eikonal (grad_constraint, 1.0) and w_deform (the dynamic weight of the three
        deformation priors); `set_loss_weights` changes them between steps."""
        _need_cuda(sdf, "sdf")
        _need_cuda(k0, "k0")
        self.device = sdf.device
        self.world = tuple(int(v) for v in sdf.shape[2:])
        if k0.shape[1] != 12 or tuple(k0.shape[2:]) != self.world:
            raise RuntimeError("FusedVoxurf expects sdf [1,1,X,Y,Z] and k0 [1,12,X,Y,Z]")
        self.n_voxels = sdf.numel()
        self.xyz_min = [float(v) for v in xyz_min]
        self.xyz_max = [float(v) for v in xyz_max]
        vol = 1.0
        for a, b in zip(self.xyz_min, self.xyz_max):
            vol *= (b - a)
        # voxurf_coarse.py _set_grid_resolution: (volume / num_voxels)^(1/3); pass it when num_voxels is not the
        # product of the grid extents (the reference floors the extents)
        self.voxel_size = float(voxel_size) if voxel_size is not None else (vol / self.n_voxels) ** (1.0 / 3.0)
        self.stepsize = stepsize
        self.scene = make_scene(self.world, self.xyz_min, self.xyz_max, near, far, stepsize * self.voxel_size, 0.0,
                                stepsize, bg)
        self.posbase_pe, self.viewbase_pe, self.barf_c2f = posbase_pe, viewbase_pe, barf_c2f
        self.sdf_alpha, self.sdf_beta = sdf_alpha, sdf_beta
        self.s_val = s_val
        self.progress = 1.0
        self.cfg = _Cfg()
        self._refresh_cfg()
        self.perm, self.d0_ref = feature_permutation(posbase_pe, viewbase_pe)
        self.d0 = FEAT_COLS          # the kernels see the permuted, zero-padded first layer [128, 64]
        self.depth = len(rgbnet_weights)
        if rgbnet_weights[0].shape[1] != self.d0_ref or rgbnet_weights[-1].shape[0] != 3:
            raise RuntimeError("FusedVoxurf: rgbnet expects %d inputs and 3 outputs" % self.d0_ref)
        dev = self.device
        # parameters in the kernels' layouts
        self.sdf = sdf.detach().clone().reshape(-1).contiguous()
        # (k0rec_buf / grad_flat_buf: caller-owned, e.g. peer-mapped, memory for the feature records and the flat
        # gradient buffer; k0_moment_records: Adam moments for a slice of the feature records only - DistributedVoxurf)
        self.k0rec = (torch.empty((self.n_voxels, 12), dtype=torch.float32, device=dev) if k0rec_buf is None
                      else k0rec_buf.view(-1)[:12 * self.n_voxels].view(self.n_voxels, 12))
        with torch.cuda.device(dev):
            check(lib.pp_voxurf_k0_pack(ptr(k0.contiguous()), self.n_voxels, ptr(self.k0rec), stream()), "k0_pack")
        idx = torch.tensor([max(p, 0) for p in self.perm], device=dev)
        keep = torch.tensor([1.0 if p >= 0 else 0.0 for p in self.perm], device=dev)
        w0 = rgbnet_weights[0].detach().to(dev).float()
        self.weights = [(w0[:, idx] * keep).contiguous()] + [w.detach().to(dev).float().contiguous()
                                                              for w in rgbnet_weights[1:]]
        self.biases = [b.detach().to(dev).float().contiguous() for b in rgbnet_biases]
        self.srec = torch.empty((self.n_voxels, 4), dtype=torch.float32, device=dev)
        # gradients: ONE flat buffer (SDF-record gradients | feature-record gradients | rgbnet gradients), so a
        # multi-GPU step reduces everything with a single all-reduce (DistributedVoxurf)
        n_net = sum(w.numel() for w in self.weights) + sum(b.numel() for b in self.biases)
        if grad_flat_buf is None:
            self.grad_flat = torch.zeros(16 * self.n_voxels + n_net, dtype=torch.float32, device=dev)
        else:
            self.grad_flat = grad_flat_buf.view(-1)[:16 * self.n_voxels + n_net]
            self.grad_flat.zero_()
        self.n_net = n_net
        self.g_srec = self.grad_flat[:4 * self.n_voxels].view(self.n_voxels, 4)
        self.g_k0rec = self.grad_flat[4 * self.n_voxels:16 * self.n_voxels].view(self.n_voxels, 12)
        self.g_w, self.g_b = [], []
        off = 16 * self.n_voxels
        for w in self.weights:
            self.g_w.append(self.grad_flat[off:off + w.numel()].view_as(w))
            off += w.numel()
        for bb in self.biases:
            self.g_b.append(self.grad_flat[off:off + bb.numel()].view_as(bb))
            off += bb.numel()
        self.m_sdf, self.v_sdf = torch.zeros_like(self.sdf), torch.zeros_like(self.sdf)
        if k0_moment_records is None:
            self.m_k0, self.v_k0 = torch.zeros_like(self.k0rec), torch.zeros_like(self.k0rec)
        else:
            self.m_k0 = torch.zeros((int(k0_moment_records), 4), dtype=torch.float32, device=dev)
            self.v_k0 = torch.zeros_like(self.m_k0)
        self.m_w = [torch.zeros_like(w) for w in self.weights]
        self.v_w = [torch.zeros_like(w) for w in self.weights]
        self.m_b = [torch.zeros_like(b) for b in self.biases]
        self.v_b = [torch.zeros_like(b) for b in self.biases]
        self.g_ab = torch.zeros(2, dtype=torch.float32, device=dev)
        self.hp = (betas[0], betas[1], lr_sdf, lr_k0, lr_rgbnet, eps)
        if adam not in ("reference", "masked"):
            raise ValueError("adam must be 'reference' or 'masked'")
        self.adam_dense = adam == "reference"
        self.lr_warp = lr_warp
        self.w_rgb = w_rgb
        self.w_entropy, self.w_mask, self.w_tv_k0, self.w_eikonal, self.w_deform = (w_entropy, w_mask, w_tv_k0,
                                                                                   w_eikonal, w_deform)
        self.step = 0
        # [0] weighted per-ray loss, [1] img_render (mse), [2] entropy, [3] mask BCE, [4] sum |k0 forward differences|
        self.loss_buf = torch.zeros(8, dtype=torch.float32, device=dev)
        self._scratch = torch.zeros(4, dtype=torch.float32, device=dev)
        self._cap_per_ray = samples_per_ray_cap
        self.pose_grad = bool(pose_grad)     # also emit d loss / d (rays_o, rays_d, viewdirs) (SE3 refinement)
        self.dx_cols = FEAT_COLS if self.pose_grad else DX_COLS
        self._bufs = None
        self._arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        self.warp = None                     # (weights, biases, out_range) of the deformation network, see below
        self.rebuild_records()

    # ---- configuration that changes over training --------------------------------------------------------
    def _refresh_cfg(self):
        c = self.cfg
        c.inv_s = 1.0 / float(self.s_val)
        c.sdf_scale_a = softplus10(self.sdf_alpha) if self.sdf_alpha is not None else 0.0
        c.sdf_scale_b = softplus10(self.sdf_beta) if self.sdf_beta is not None else 0.0
        c.voxel_size = self.voxel_size
        c.pos_pe, c.view_pe = self.posbase_pe, self.viewbase_pe
        for k, w in enumerate(barf_weights(self.progress, self.barf_c2f, self.posbase_pe)):
            c.pe_w_pos[k] = w
        for k, w in enumerate(barf_weights(self.progress, self.barf_c2f, self.viewbase_pe)):
            c.pe_w_view[k] = w

    def set_schedule(self, s_val=None, progress=None):
        """s_val of neus_alpha_from_sdf_scatter (:487) and the BARF progress (:928) for the next steps."""
        if s_val is not None:
            self.s_val = s_val
        if progress is not None:
            self.progress = progress
        self._refresh_cfg()

    def rebuild_records(self):
        with torch.cuda.device(self.device):
            check(lib.pp_voxurf_build_records(ctypes.byref(self.scene), ctypes.byref(self.cfg), ptr(self.sdf),
                                              ptr(self.srec), stream()), "voxurf_build_records")

    # ---- buffers ----------------------------------------------------------------------------------------------
    def _buffers(self, n_rays):
        per_ray = self._cap_per_ray
        if per_ray is None:
            diag = math.sqrt(sum((b - a) ** 2 for a, b in zip(self.xyz_min, self.xyz_max)))
            per_ray = int(diag / (self.stepsize * self.voxel_size)) + 4        # every step of every ray
        cap = int(n_rays) * per_ray
        b = self._bufs
        if b is None or b["cap"] < cap or b["n_rays"] < n_rays:
            dev = self.device
            b = {"cap": cap, "n_rays": n_rays,
                 "feat": torch.empty((cap, FEAT_COLS), dtype=torch.bfloat16, device=dev),
                 "w": torch.empty(cap, dtype=torch.float32, device=dev),
                 "rid": torch.empty(cap, dtype=torch.int32, device=dev),
                 "rec": torch.empty((cap, 4), dtype=torch.float32, device=dev),
                 "gw": torch.empty(cap, dtype=torch.float32, device=dev),
                 "dx": torch.empty((cap, self.dx_cols), dtype=torch.float32, device=dev),
                 "grad_rays": torch.zeros((n_rays, 6), dtype=torch.float32, device=dev) if self.pose_grad else None,
                 "grad_view": torch.zeros((n_rays, 3), dtype=torch.float32, device=dev) if self.pose_grad else None,
                 "ray_off": torch.empty(n_rays, dtype=torch.int32, device=dev),
                 "ray_cnt": torch.empty(n_rays, dtype=torch.int32, device=dev),
                 "ray_end": torch.empty(n_rays, dtype=torch.int32, device=dev),
                 "cum_w": torch.empty(n_rays, dtype=torch.float32, device=dev),
                 "last": torch.empty(n_rays, dtype=torch.float32, device=dev),
                 "acc": torch.empty((n_rays, 3), dtype=torch.float32, device=dev),
                 "g_ray": torch.empty((n_rays, 3), dtype=torch.float32, device=dev),
                 "g_cumw": torch.empty(n_rays, dtype=torch.float32, device=dev),
                 "g_last": torch.empty(n_rays, dtype=torch.float32, device=dev),
                 "rgb": torch.empty((n_rays, 3), dtype=torch.float32, device=dev),
                 "counter": torch.zeros(2, dtype=torch.int32, device=dev)}
            self._bufs = b
        return b

    # ---- implicit deformation (render path) --------------------------------------------------------------------
    def set_warp_network(self, weights, biases, out_range):
        """Attach the deformation network (DeformedImplicitField: Linear 3->128, 128->128 x3, 128->4, lib/deformation/
        deform_net.py:12-29; ``out_range`` = range_shape.max()).  `render` then follows Voxurf.inference
        (voxurf_coarse.py:1093-1215): SDF and normal of every sample come from the warped template lookup."""
        from .ops import warpnet_params
        dev = self.device
        same = (self.warp is not None and len(self.warp_w) == len(weights) and
                all(a.shape == b.shape for a, b in zip(self.warp_w, weights)))
        if same:       # keep the optimiser state, refresh the values
            for dst, src in zip(self.warp_w + self.warp_b, list(weights) + list(biases)):
                dst.copy_(src.detach())
        else:
            self.warp_w = [w.detach().to(dev).float().contiguous().clone() for w in weights]      # nn.Linear layouts
            self.warp_b = [b.detach().to(dev).float().contiguous().clone() for b in biases]
            self.m_warp = [torch.zeros_like(t) for t in self.warp_w + self.warp_b]
            self.v_warp = [torch.zeros_like(t) for t in self.warp_w + self.warp_b]
        ws, bs = warpnet_params(self.warp_w, self.warp_b)        # first layer doubled: [128, 6]
        self.warp = (ws, bs, float(out_range))
        if not same:
            self.g_ww = [torch.zeros_like(w) for w in ws]
            self.g_wb = [torch.zeros_like(b) for b in bs]
            self._warp_work = torch.zeros(int(lib.pp_warpnet_backward_workspace_bytes()), dtype=torch.uint8, device=dev)

    def set_loss_weights(self, **kw):
        for k, v in kw.items():
            if k not in ("w_rgb", "w_entropy", "w_mask", "w_tv_k0", "w_eikonal", "w_deform"):
                raise KeyError(k)
            setattr(self, k, float(v))

    def warp_gradients(self):
        """([dW_l], [db_l]) of the accumulated deformation-network gradients, nn.Linear layouts."""
        g0 = self.g_ww[0][:, :3] + self.g_ww[0][:, 3:]
        return [g0] + [g.clone() for g in self.g_ww[1:]], [g.clone() for g in self.g_wb]

    def _inbox_buffers(self, b, points=True):
        dev = self.device
        if "dray_off" not in b:
            b["dray_off"] = torch.empty(b["n_rays"], dtype=torch.int32, device=dev)
            b["dray_cnt"] = torch.zeros(b["n_rays"], dtype=torch.int32, device=dev)
            b["dcounter"] = torch.zeros(4, dtype=torch.int32, device=dev)
        if points and "dpts" not in b:
            b["dpts"] = torch.empty((b["cap"], 4), dtype=torch.float32, device=dev)
        if "reg" not in b:
            b["reg"] = torch.zeros(4, dtype=torch.float32, device=dev)

    def _enumerate_inbox(self, b, rays_o, rays_d, jitter):
        """Runs of ALL in-box samples per ray (the reference's compacted ray_pts, voxurf_coarse.py:940-944) and their
        positions, without the deformation network's rows."""
        self._inbox_buffers(b)
        check(lib.pp_voxurf_deform_rows(ctypes.byref(self.scene), ptr(rays_o), ptr(rays_d), ptr(jitter),
                                        rays_o.shape[0], b["cap"], None, ptr(b["dray_off"]), ptr(b["dcounter"]),
                                        ptr(b["dpts"]), ptr(b["dray_cnt"]), stream()), "voxurf_deform_rows")

    def inbox_order(self, n):
        """-> (idx [M], ray_id [M]): kernel slot and ray of every in-box sample in the reference's ray-major order
        (runs are handed out in completion order on the device).  Host sync; raises on a capacity overflow."""
        b = self._bufs
        c = b["dcounter"].tolist()
        if c[1]:
            raise RuntimeError("FusedVoxurf: sample capacity exceeded (raise samples_per_ray_cap)")
        m = c[0]
        dev = self.device
        cnt, off = b["dray_cnt"][:n].long(), b["dray_off"][:n].long()
        ray_id = torch.repeat_interleave(torch.arange(n, device=dev), cnt, output_size=m)
        starts = torch.cumsum(cnt, 0) - cnt
        idx = off[ray_id] + (torch.arange(m, device=dev) - starts[ray_id])
        return idx, ray_id

    def _march_deform(self, b, rays_o, rays_d, viewdirs, jitter, depth=None, normal=None, keep_points=False):
        n, cap = rays_o.shape[0], b["cap"]
        self._inbox_buffers(b, keep_points)
        if "drows" not in b:
            dev = self.device
            b["drows"] = torch.empty((4 * cap, 16), dtype=torch.float16, device=dev)
            b["dout"] = torch.empty((4 * cap, 4), dtype=torch.float32, device=dev)
        ws, bs, out_range = self.warp
        check(lib.pp_voxurf_deform_rows(ctypes.byref(self.scene), ptr(rays_o), ptr(rays_d), ptr(jitter), n, cap,
                                        ptr(b["drows"]), ptr(b["dray_off"]), ptr(b["dcounter"]),
                                        ptr(b["dpts"]) if keep_points else None, ptr(b["dray_cnt"]), stream()),
              "voxurf_deform_rows")
        check(lib.pp_warpnet_forward_packed(ptr(b["drows"]), 4 * cap, len(ws), 128, self._arr(ws), self._arr(bs),
                                            ptr(b["dout"]), ptr(b["dcounter"][2:]), stream()), "warpnet_forward")
        check(lib.pp_voxurf_march_features_deform(
            ctypes.byref(self.scene), ctypes.byref(self.cfg), ptr(self.srec), ptr(self.k0rec), ptr(rays_o),
            ptr(rays_d), ptr(viewdirs), ptr(jitter), n, cap, ptr(b["feat"]), ptr(b["w"]), ptr(b["rid"]),
            ptr(b["rec"]), ptr(b["ray_off"]), ptr(b["ray_cnt"]), ptr(b["ray_end"]), ptr(b["cum_w"]), ptr(b["last"]),
            ptr(b["counter"]), ptr(depth), ptr(normal), ptr(b["dout"]), ptr(b["dray_off"]), out_range, stream()),
            "voxurf_march_features_deform")

    # ---- forward pieces -----------------------------------------------------------------------------------------
    def _march(self, b, rays_o, rays_d, viewdirs, jitter, depth=None, normal=None):
        n = rays_o.shape[0]
        check(lib.pp_voxurf_march_features(
            ctypes.byref(self.scene), ctypes.byref(self.cfg), ptr(self.srec), ptr(self.k0rec), ptr(rays_o),
            ptr(rays_d), ptr(viewdirs), ptr(jitter), n, b["cap"], ptr(b["feat"]), ptr(b["w"]), ptr(b["rid"]),
            ptr(b["rec"]), ptr(b["ray_off"]), ptr(b["ray_cnt"]), ptr(b["ray_end"]), ptr(b["cum_w"]), ptr(b["last"]),
            ptr(b["counter"]), ptr(depth), ptr(normal), stream()), "voxurf_march_features")

    def _mlp_forward(self, b, n):
        b["acc"][:n].zero_()
        check(lib.pp_rgbnet_forward_packed(ptr(b["feat"]), b["cap"], self.d0, self.depth, 128, self._arr(self.weights),
                                           self._arr(self.biases), 1, None, ptr(b["w"]), ptr(b["rid"]), ptr(b["acc"]),
                                           ptr(b["counter"]), stream()), "rgbnet_forward")

    def render(self, rays_o, rays_d, viewdirs, want_depth=False, want_normal=False, use_deform=None, jitter=None):
        """-> rgb_marched [N,3], alphainv_last [N], cum_weights [N] (+ depth [N], normal_marched [N,3] on request):
        the per-ray outputs of Voxurf.forward / inference that lib/nvs_fun.py:74-88 reads.  use_deform (default: True
        once set_warp_network was called) selects the implicit-deformation lookup of Voxurf.inference."""
        if use_deform is None:
            use_deform = self.warp is not None
        if use_deform and self.warp is None:
            raise RuntimeError("FusedVoxurf.render: use_deform=True needs set_warp_network first")
        for t, nme in ((rays_o, "rays_o"), (rays_d, "rays_d"), (viewdirs, "viewdirs")):
            _need_cuda(t, nme)
        n = rays_o.shape[0]
        b = self._buffers(n)
        depth = torch.empty(n, dtype=torch.float32, device=self.device) if want_depth else None
        normal = torch.empty((n, 3), dtype=torch.float32, device=self.device) if want_normal else None
        with torch.cuda.device(self.device):
            (self._march_deform if use_deform else self._march)(b, rays_o, rays_d, viewdirs, jitter, depth, normal)
            self._mlp_forward(b, n)
            check(lib.pp_voxurf_ray_loss(n, ptr(b["acc"]), ptr(b["cum_w"]), ptr(b["acc"]), self.scene.bg, 0.0,
                                         ptr(b["g_ray"]), ptr(b["rgb"]), None, stream()), "voxurf_ray_loss")
        out = (b["rgb"][:n].clone(), b["last"][:n].clone(), b["cum_w"][:n].clone())
        if want_depth:
            out = out + (depth,)
        if want_normal:
            out = out + (normal,)
        return out

    def n_samples(self):
        """Composited samples of the last march (host sync); raises if the sample capacity overflowed."""
        c = self._bufs["counter"].tolist()
        dc = self._bufs["dcounter"].tolist() if "dcounter" in self._bufs else [0, 0]
        if c[1] or dc[1]:
            raise RuntimeError("FusedVoxurf: sample capacity exceeded (raise samples_per_ray_cap)")
        return c[0]

    # ---- training -----------------------------------------------------------------------------------------------
    def backward_only(self, rays_o, rays_d, viewdirs, target, jitter=None, mask=None):
        """march -> MLP -> per-ray object losses (lib/losses.py:34-74: photometric, entropy, mask BCE) -> MLP backward
        -> reverse march.  Accumulates into the gradient buffers and loss_buf; no parameter changes."""
        for t, nme in ((rays_o, "rays_o"), (rays_d, "rays_d"), (viewdirs, "viewdirs"), (target, "target")):
            _need_cuda(t, nme)
        n = rays_o.shape[0]
        b = self._buffers(n)
        with torch.cuda.device(self.device):
            self._march(b, rays_o, rays_d, viewdirs, jitter)
            self._mlp_forward(b, n)
            self._object_loss(b, n, target, mask)
            check(lib.pp_rgbnet_backward_fused(ptr(b["feat"]), b["cap"], ptr(b["counter"]), self.d0, self.depth, 128,
                                               self._arr(self.weights), self._arr(self.biases), ptr(b["w"]),
                                               ptr(b["rid"]), ptr(b["g_ray"]), ptr(b["gw"]), ptr(b["dx"]),
                                               self.dx_cols, self._arr(self.g_w), self._arr(self.g_b), stream()),
                  "rgbnet_backward_fused")
            check(lib.pp_voxurf_backward(ctypes.byref(self.scene), ctypes.byref(self.cfg), ptr(rays_o), ptr(rays_d),
                                         ptr(viewdirs), ptr(jitter), n, ptr(b["rec"]), ptr(b["gw"]), ptr(b["dx"]),
                                         self.dx_cols, ptr(b["ray_off"]), ptr(b["ray_cnt"]), ptr(b["ray_end"]),
                                         ptr(b["g_ray"]), ptr(self.g_srec), ptr(self.g_k0rec), ptr(self.srec),
                                         ptr(self.k0rec), ptr(b["grad_rays"]), ptr(b["grad_view"]), ptr(b["g_cumw"]),
                                         ptr(b["g_last"]), stream()), "voxurf_backward")
            if self.w_eikonal > 0:
                # grad_constraint of lib/losses.py:7-10 on ALL in-box samples (also past the early termination)
                self._enumerate_inbox(b, rays_o, rays_d, jitter)
                b["reg"].zero_()
                check(lib.pp_voxurf_sample_records_backward(ctypes.byref(self.scene), ptr(self.srec), ptr(b["dpts"]),
                                                            ptr(b["dcounter"]), b["cap"], None, self.w_eikonal,
                                                            ptr(self.g_srec), ptr(b["reg"]), stream()),
                      "voxurf_sample_records_backward")
        return b["rgb"][:n]

    def _object_loss(self, b, n, target, mask):
        check(lib.pp_voxurf_object_loss(n, ptr(b["acc"]), ptr(b["cum_w"]), ptr(b["last"]), ptr(target), ptr(mask),
                                        self.scene.bg, self.w_rgb, self.w_entropy, self.w_mask, ptr(b["g_ray"]),
                                        ptr(b["g_cumw"]), ptr(b["g_last"]), ptr(b["rgb"]), ptr(self.loss_buf),
                                        ptr(self._scratch), stream()), "voxurf_object_loss")

    def forward_sums_deform(self, rays_o, rays_d, viewdirs, jitter=None, want_samples=False, depth=None, normal=None):
        """Deform branch of `forward_sums`: -> (rgb_acc, cum_weights, alphainv_last); with want_samples also the
        per-sample entries of ALL in-box samples (b["sout"] fp32 [cap,16], see pp_voxurf_deform_sample_forward)."""
        if self.warp is None:
            raise RuntimeError("the deformation branch needs set_warp_network first")
        n = rays_o.shape[0]
        b = self._buffers(n)
        with torch.cuda.device(self.device):
            self._march_deform(b, rays_o, rays_d, viewdirs, jitter, depth, normal, keep_points=True)
            self._mlp_forward(b, n)
            if want_samples:
                if "sout" not in b:
                    b["sout"] = torch.empty((b["cap"], 16), dtype=torch.float32, device=self.device)
                check(lib.pp_voxurf_deform_sample_forward(ctypes.byref(self.scene), ptr(self.srec), ptr(b["dpts"]),
                                                          ptr(b["dout"]), ptr(b["dcounter"]), b["cap"], self.warp[2],
                                                          ptr(b["sout"]), None, stream()),
                      "voxurf_deform_sample_forward")
        return b["acc"][:n], b["cum_w"][:n], b["last"][:n]

    def backward_deform_from_ray_grads(self, rays_o, rays_d, viewdirs, g_acc, g_cumw, g_last, jitter=None,
                                       reg_weights=(0.0, 0.0, 0.0, 0.0), ext_grad=None):
        """Back-propagate d loss / d (rgb_acc, cum_weights, alphainv_last) of the last `forward_sums_deform` (+ the
        regularisers of deform_implicit_loss with `reg_weights` = weights of grad_constraint, grad_deform_constraint,
        sdf_correct_constraint, sdf_deform_constraint, lib/losses.py:6-24; + ext_grad fp32 [cap,14], the gradient of
        an external loss w.r.t. the per-sample entries) through the whole deformation branch
        (voxurf_coarse.py:958-985): rgbnet backward -> reverse march writing {dL/d sdf_final, dL/d gradient} per
        sample -> per-sample backward of the warped lookup -> tcgen05 backward of the deformation network.
        Accumulates into the rgbnet / feature-grid / deformation-network gradient buffers; the SDF grid is frozen in
        this mode (:136-138).  Leaves ``g_dout`` (gradient w.r.t. the network's outputs), ``g_pts`` [samples,3] and
        ``reg`` (the four unweighted terms) in the buffers; with pose_grad the ray gradients are complete
        (`ray_gradients`)."""
        n = rays_o.shape[0]
        b = self._bufs
        cap = b["cap"]
        g_acc = g_acc.contiguous()
        ws, bs, out_range = self.warp
        with torch.cuda.device(self.device):
            check(lib.pp_rgbnet_backward_fused(ptr(b["feat"]), cap, ptr(b["counter"]), self.d0, self.depth, 128,
                                               self._arr(self.weights), self._arr(self.biases), ptr(b["w"]),
                                               ptr(b["rid"]), ptr(g_acc), ptr(b["gw"]), ptr(b["dx"]),
                                               self.dx_cols, self._arr(self.g_w), self._arr(self.g_b), stream()),
                  "rgbnet_backward_fused")
            if "g_val" not in b:
                dev = self.device
                b["g_val"] = torch.empty((cap, 4), dtype=torch.float32, device=dev)
                b["g_dout"] = torch.empty((4 * cap, 4), dtype=torch.float32, device=dev)
                b["g_pts"] = torch.empty((cap, 3), dtype=torch.float32, device=dev)
            b["g_val"].zero_()
            b["reg"].zero_()
            check(lib.pp_voxurf_backward_deform(ctypes.byref(self.scene), ctypes.byref(self.cfg), ptr(rays_o),
                                                ptr(rays_d), ptr(viewdirs), ptr(jitter), n, ptr(b["rec"]),
                                                ptr(b["gw"]), ptr(b["dx"]), self.dx_cols, ptr(b["ray_off"]),
                                                ptr(b["ray_cnt"]), ptr(b["ray_end"]), ptr(g_acc),
                                                ptr(self.g_k0rec), ptr(g_cumw.contiguous()) if g_cumw is not None else None,
                                                ptr(g_last.contiguous()) if g_last is not None else None,
                                                ptr(b["g_val"]), ptr(b["dray_off"]), ptr(self.srec), ptr(self.k0rec),
                                                ptr(b["grad_rays"]), ptr(b["grad_view"]), stream()),
                  "voxurf_backward_deform")
            w = [float(x) for x in reg_weights]
            check(lib.pp_voxurf_deform_sample_backward(ctypes.byref(self.scene), ptr(self.srec), ptr(b["dpts"]),
                                                       ptr(b["dout"]), ptr(b["g_val"]), ptr(b["dcounter"]), cap,
                                                       out_range, w[0], w[1], w[2], w[3], ptr(b["g_dout"]),
                                                       ptr(b["g_pts"]), ptr(b["reg"]), ptr(ext_grad), stream()),
                  "voxurf_deform_sample_backward")
            if self.pose_grad:
                check(lib.pp_voxurf_deform_ray_grad(ctypes.byref(self.scene), ptr(rays_o), ptr(rays_d), ptr(jitter), n,
                                                    ptr(b["dray_off"]), ptr(b["g_pts"]), ptr(b["grad_rays"]),
                                                    stream()), "voxurf_deform_ray_grad")
            check(lib.pp_warpnet_backward_packed(ptr(b["drows"]), ptr(b["g_dout"]), 4 * cap, ptr(b["dcounter"][2:]),
                                                 len(ws), 128, self._arr(ws), self._arr(bs), self._arr(self.g_ww),
                                                 self._arr(self.g_wb), ptr(self._warp_work), stream()),
                  "warpnet_backward")

    def backward_deform(self, rays_o, rays_d, viewdirs, target, jitter=None, mask=None, reg_weights=None):
        """The reference's live train step up to the optimiser (lib/recon_scene.py:603-606, :649): forward through the
        implicit deformation, the per-ray object losses (w_rgb / w_entropy / w_mask), the per-sample regularisers
        (reg_weights, default (w_eikonal, w_deform, w_deform, w_deform)), full backward.  Returns rgb_marched."""
        for t, nme in ((rays_o, "rays_o"), (rays_d, "rays_d"), (viewdirs, "viewdirs"), (target, "target")):
            _need_cuda(t, nme)
        if reg_weights is None:
            reg_weights = (self.w_eikonal, self.w_deform, self.w_deform, self.w_deform)
        n = rays_o.shape[0]
        self.forward_sums_deform(rays_o, rays_d, viewdirs, jitter)
        b = self._bufs
        with torch.cuda.device(self.device):
            self._object_loss(b, n, target, mask)
        self.backward_deform_from_ray_grads(rays_o, rays_d, viewdirs, b["g_ray"], b["g_cumw"], b["g_last"], jitter,
                                            reg_weights)
        return b["rgb"][:n]

    backward_deform_pieces = backward_deform

    # ---- autograd-facing pieces (used by the Voxurf module below) ---------------------------------------------
    def load_params(self, sdf, k0, weights, biases, sdf_alpha=None, sdf_beta=None):
        """Copy parameters given in the reference's layouts into the engine's layouts and rebuild the records."""
        self.sdf.copy_(sdf.detach().reshape(-1))
        with torch.cuda.device(self.device):
            check(lib.pp_voxurf_k0_pack(ptr(k0.detach().contiguous()), self.n_voxels, ptr(self.k0rec), stream()),
                  "k0_pack")
        idx = torch.tensor([max(p, 0) for p in self.perm], device=self.device)
        keep = torch.tensor([1.0 if p >= 0 else 0.0 for p in self.perm], device=self.device)
        self.weights[0].copy_(weights[0].detach()[:, idx] * keep)
        for dst, src in zip(self.weights[1:], weights[1:]):
            dst.copy_(src.detach())
        for dst, src in zip(self.biases, biases):
            dst.copy_(src.detach())
        if sdf_alpha is not None:
            self.sdf_alpha, self.sdf_beta = float(sdf_alpha), float(sdf_beta)
        self._refresh_cfg()
        self.rebuild_records()

    def forward_sums(self, rays_o, rays_d, viewdirs, jitter=None, want_samples=False, depth=None, normal=None):
        """-> (rgb_acc [N,3] = sum_i w_i rgb_i, cum_weights [N], alphainv_last [N]); state is kept for
        backward_from_ray_grads."""
        n = rays_o.shape[0]
        b = self._buffers(n)
        with torch.cuda.device(self.device):
            self._march(b, rays_o, rays_d, viewdirs, jitter, depth, normal)
            self._mlp_forward(b, n)
            if want_samples:       # {sdf, gradient} of ALL in-box samples (the dict's `gradient`, :1082)
                self._enumerate_inbox(b, rays_o, rays_d, jitter)
                if "srec_s" not in b:
                    b["srec_s"] = torch.empty((b["cap"], 4), dtype=torch.float32, device=self.device)
                check(lib.pp_voxurf_sample_records(ctypes.byref(self.scene), ptr(self.srec), ptr(b["dpts"]),
                                                   ptr(b["dcounter"]), b["cap"], ptr(b["srec_s"]), stream()),
                      "voxurf_sample_records")
        return b["acc"][:n], b["cum_w"][:n], b["last"][:n]

    def backward_from_ray_grads(self, rays_o, rays_d, viewdirs, g_acc, g_cumw, g_last, jitter=None, g_samples=None):
        """Back-propagate d loss / d (rgb_acc, cum_weights, alphainv_last) of the last forward_sums into the
        gradient buffers (accumulating).  g_samples (nullable) fp32 [cap,4]: d loss / d {sdf, gradient} of the in-box
        samples of forward_sums(want_samples=True), kernel slot order."""
        n = rays_o.shape[0]
        b = self._bufs
        g_acc = g_acc.contiguous()
        with torch.cuda.device(self.device):
            check(lib.pp_rgbnet_backward_fused(ptr(b["feat"]), b["cap"], ptr(b["counter"]), self.d0, self.depth, 128,
                                               self._arr(self.weights), self._arr(self.biases), ptr(b["w"]),
                                               ptr(b["rid"]), ptr(g_acc), ptr(b["gw"]), ptr(b["dx"]),
                                               self.dx_cols, self._arr(self.g_w), self._arr(self.g_b), stream()),
                  "rgbnet_backward_fused")
            check(lib.pp_voxurf_backward(ctypes.byref(self.scene), ctypes.byref(self.cfg), ptr(rays_o), ptr(rays_d),
                                         ptr(viewdirs), ptr(jitter), n, ptr(b["rec"]), ptr(b["gw"]), ptr(b["dx"]),
                                         self.dx_cols, ptr(b["ray_off"]), ptr(b["ray_cnt"]), ptr(b["ray_end"]),
                                         ptr(g_acc), ptr(self.g_srec), ptr(self.g_k0rec), ptr(self.srec),
                                         ptr(self.k0rec), ptr(b["grad_rays"]), ptr(b["grad_view"]),
                                         ptr(g_cumw.contiguous()), ptr(g_last.contiguous()) if g_last is not None else None,
                                         stream()), "voxurf_backward")
            if g_samples is not None:
                check(lib.pp_voxurf_sample_records_backward(ctypes.byref(self.scene), ptr(self.srec), ptr(b["dpts"]),
                                                            ptr(b["dcounter"]), b["cap"], ptr(g_samples), 0.0,
                                                            ptr(self.g_srec), None, stream()),
                      "voxurf_sample_records_backward")

    def ray_gradients(self, rays_d=None):
        """d loss / d (rays_o | rays_d) [N,6] and d loss / d viewdirs [N,3] of the last backward (pose_grad=True).
        With rays_d given, the viewdirs gradient is folded through viewdirs = rays_d / |rays_d|
        (voxurf_coarse.py:1365) into the rays_d half, ready for pp_pose_grad_reduce / camera.pose_grad."""
        if not self.pose_grad:
            raise RuntimeError("FusedVoxurf was built with pose_grad=False")
        b = self._bufs
        n = b["n_rays"]
        g_rays, g_view = b["grad_rays"][:n].clone(), b["grad_view"][:n].clone()
        if rays_d is None:
            return g_rays, g_view
        nrm = rays_d.norm(dim=-1, keepdim=True)
        v = rays_d / nrm
        g_rays[:, 3:] += (g_view - v * (v * g_view).sum(-1, keepdim=True)) / nrm
        return g_rays

    def grid_gradients(self):
        """(d loss / d sdf [1,1,X,Y,Z], d loss / d k0 [1,12,X,Y,Z], [dW], [db]) of the accumulated gradients, in
        the reference's layouts (first-layer columns un-permuted)."""
        g_sdf = torch.empty_like(self.sdf)
        with torch.cuda.device(self.device):
            check(lib.pp_voxurf_sdf_adam(ctypes.byref(self.scene), ctypes.byref(self.cfg), ptr(self.sdf),
                                         ptr(self.g_srec), None, None, 1, 0.9, 0.99, 0.0, 1e-8, ptr(g_sdf),
                                         ptr(self.g_ab), 0, stream()), "voxurf_sdf_adam")
            g_k0 = torch.empty((1, 12) + self.world, dtype=torch.float32, device=self.device)
            check(lib.pp_voxurf_k0_unpack(ptr(self.g_k0rec), self.n_voxels, ptr(g_k0), stream()), "k0_unpack")
        gw0 = torch.zeros((128, self.d0_ref), dtype=torch.float32, device=self.device)
        cols = [(c, p) for c, p in enumerate(self.perm) if p >= 0]
        gw0[:, [p for _, p in cols]] = self.g_w[0][:, [c for c, _ in cols]]
        return g_sdf.reshape((1, 1) + self.world), g_k0, [gw0] + [g.clone() for g in self.g_w[1:]], \
            [g.clone() for g in self.g_b]

    def _k0_tv(self):
        if self.w_tv_k0 > 0:
            check(lib.pp_voxurf_k0_tv(ctypes.byref(self.scene), ptr(self.k0rec), ptr(self.g_k0rec), self.w_tv_k0,
                                      ptr(self.loss_buf[4:]), stream()), "voxurf_k0_tv")

    def optimizer_step(self, deform=False, skip_k0=False):
        """Adam on everything that has a gradient.  deform=True: the live configuration of the reference (SDF grid
        frozen, voxurf_coarse.py:136-138; feature grid, rgbnet and deformation network trained).  skip_k0: the feature
        records were already updated (by the multi-GPU exchange kernel, which owns their Adam moments)."""
        b1, b2, lr_sdf, lr_k0, lr_net, eps = self.hp
        self.step += 1
        dense = 1 if self.adam_dense else 0
        mode = 3 if self.adam_dense else 0
        with torch.cuda.device(self.device):
            s = stream()
            if not deform:
                check(lib.pp_voxurf_sdf_adam(ctypes.byref(self.scene), ctypes.byref(self.cfg), ptr(self.sdf),
                                             ptr(self.g_srec), ptr(self.m_sdf), ptr(self.v_sdf), self.step, b1, b2,
                                             lr_sdf, eps, None, ptr(self.g_ab), dense, s), "voxurf_sdf_adam")
                self.g_srec.zero_()
                check(lib.pp_voxurf_build_records(ctypes.byref(self.scene), ctypes.byref(self.cfg), ptr(self.sdf),
                                                  ptr(self.srec), s), "voxurf_build_records")
            # feature records: Adam over 3 V float4 records, gradient reset fused
            if not skip_k0:
                check(lib.pp_voxel_adam(ptr(self.k0rec), ptr(self.g_k0rec), ptr(self.m_k0), ptr(self.v_k0),
                                        3 * self.n_voxels, self.step, b1, b2, lr_k0, lr_k0, eps, 1 | (2 * dense), s),
                      "voxel_adam")
            for p, g, m, v in zip(self.weights + self.biases, self.g_w + self.g_b, self.m_w + self.m_b,
                                  self.v_w + self.v_b):
                check(lib.pp_adam_upd(0, mode, ptr(p), ptr(g), ptr(m), ptr(v), None, p.numel(), self.step, b1, b2,
                                      lr_net, eps, s), "adam_upd")
                g.zero_()
            if deform:
                gws, gbs = self.warp_gradients()
                for p, g, m, v in zip(self.warp_w + self.warp_b, gws + gbs, self.m_warp, self.v_warp):
                    g = g.contiguous()
                    check(lib.pp_adam_upd(0, mode, ptr(p), ptr(g), ptr(m), ptr(v), None, p.numel(), self.step, b1,
                                          b2, self.lr_warp, eps, s), "adam_upd")
                self.warp[0][0].copy_(torch.cat([self.warp_w[0], self.warp_w[0]], dim=1))
                for dst, src in zip(self.warp[0][1:] + self.warp[1], self.warp_w[1:] + self.warp_b):
                    if dst.data_ptr() != src.data_ptr():
                        dst.copy_(src)
                for g in self.g_ww + self.g_wb:
                    g.zero_()
            self.g_ab.zero_()

    def train_step(self, rays_o, rays_d, viewdirs, target, jitter=None, use_deform=False, mask=None):
        """One optimisation step; returns loss_buf ([0] weighted per-ray loss, [1] photometric MSE, [2] entropy,
        [3] mask BCE, [4] k0 TV sum).  use_deform=True is the reference's live step (recon_scene.py:603)."""
        self.loss_buf.zero_()
        if use_deform:
            self.backward_deform(rays_o, rays_d, viewdirs, target, jitter, mask)
        else:
            self.backward_only(rays_o, rays_d, viewdirs, target, jitter, mask)
        with torch.cuda.device(self.device):
            self._k0_tv()
        self.optimizer_step(deform=use_deform)
        return self.loss_buf

    # ---- checkpoints ----------------------------------------------------------------------------------------------
    def export_state(self):
        """{'sdf.grid', 'k0.grid', 'rgbnet.<i>.weight/bias'} in the reference's layouts."""
        k0 = torch.empty((1, 12) + self.world, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.pp_voxurf_k0_unpack(ptr(self.k0rec), self.n_voxels, ptr(k0), stream()), "k0_unpack")
        out = {"sdf.grid": self.sdf.reshape((1, 1) + self.world).clone(), "k0.grid": k0}
        w0 = torch.zeros((128, self.d0_ref), dtype=torch.float32, device=self.device)
        cols = [(c, p) for c, p in enumerate(self.perm) if p >= 0]
        w0[:, [p for _, p in cols]] = self.weights[0][:, [c for c, _ in cols]]
        ws = [w0] + [w.clone() for w in self.weights[1:]]
        # nn.Sequential indices of the reference's rgbnet (voxurf_coarse.py:208-216): 0, 2.0, 3.0, ..., last
        names = ["0"] + ["%d.0" % (i + 1) for i in range(1, self.depth - 1)] + [str(self.depth)]
        for nme, w, bb in zip(names, ws, self.biases):
            out["rgbnet.%s.weight" % nme] = w
            out["rgbnet.%s.bias" % nme] = bb.clone()
        return out


class _VoxurfRenderFn(torch.autograd.Function):
    """(sdf.grid, k0.grid, sdf_alpha, sdf_beta, rgbnet parameters, rays) -> (rgb_acc, cum_weights, alphainv_last,
    {sdf, gradient} of every in-box sample) with both directions on the fused kernels; any loss written on those
    outputs (lib/losses.py) trains the reference's parameters and, through the rays, the poses."""

    @staticmethod
    def forward(ctx, eng, rays_o, rays_d, viewdirs, jitter, depth, sdf, k0, ab, *params):
        eng.load_params(sdf, k0, params[:depth], params[depth:], float(ab[0]), float(ab[1]))
        rays_o, rays_d, viewdirs = rays_o.contiguous(), rays_d.contiguous(), viewdirs.contiguous()
        n = rays_o.shape[0]
        eng.aux = {"depth": torch.empty(n, device=eng.device), "normal": torch.empty((n, 3), device=eng.device)}
        acc, cumw, last = eng.forward_sums(rays_o, rays_d, viewdirs, jitter, want_samples=True,
                                           depth=eng.aux["depth"], normal=eng.aux["normal"])
        idx, ray_id = eng.inbox_order(n)
        eng.aux.update(idx=idx, ray_id=ray_id)
        ctx.eng, ctx.depth = eng, depth
        ctx.save_for_backward(rays_o, rays_d, viewdirs, jitter if jitter is not None else torch.empty(0), ab, idx)
        return acc.clone(), cumw.clone(), last.clone(), eng._bufs["srec_s"][idx]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_acc, g_cumw, g_last, g_samp):
        eng, depth = ctx.eng, ctx.depth
        rays_o, rays_d, viewdirs, jitter, ab, idx = ctx.saved_tensors
        jitter = jitter if jitter.numel() else None
        eng.grad_flat.zero_()
        eng.g_ab.zero_()
        g_samples = None
        if g_samp is not None and idx.numel():
            g_samples = torch.zeros((eng._bufs["cap"], 4), dtype=torch.float32, device=eng.device)
            g_samples[idx] = g_samp
        eng.backward_from_ray_grads(rays_o, rays_d, viewdirs, g_acc, g_cumw, g_last, jitter, g_samples)
        g_sdf, g_k0, g_w, g_b = eng.grid_gradients()
        # (a, b) = softplus10(sdf_alpha, sdf_beta): d softplus_beta(x) / dx = sigmoid(beta x)
        g_param_ab = eng.g_ab * torch.sigmoid(10.0 * ab)
        g_o = g_d = g_v = None
        if eng.pose_grad:
            nr = rays_o.shape[0]
            g_rays, g_v = eng.ray_gradients()
            g_o, g_d, g_v = g_rays[:nr, :3], g_rays[:nr, 3:], g_v[:nr]
        return (None, g_o, g_d, g_v, None, None, g_sdf, g_k0, g_param_ab) + tuple(g_w) + tuple(g_b)


class _VoxurfDeformFn(torch.autograd.Function):
    """The implicit-deformation branch (voxurf_coarse.py:958-985) with autograd on both sides: (k0.grid, rgbnet and
    warp_network parameters, rays) -> (rgb_acc, cum_weights, alphainv_last, per-sample [M,14] = gradient | grad_deform |
    sdf_correct | sdf_deform of every in-box sample).  The backward is the reference's double backward in closed form:
    rgbnet backward (tcgen05) -> reverse march -> warped-lookup backward (cell Hessian) -> deformation-network backward
    (tcgen05, csrc/warp_bwd.cu).  The SDF grid is frozen in this branch (:136-138): no gradient is returned for it."""

    @staticmethod
    def forward(ctx, eng, rays_o, rays_d, viewdirs, jitter, depth, out_range, sdf, k0, ab, *params):
        rgb, warp = params[:2 * depth], params[2 * depth:]
        nw = len(warp) // 2
        eng.load_params(sdf, k0, rgb[:depth], rgb[depth:], float(ab[0]), float(ab[1]))
        eng.set_warp_network([w.detach() for w in warp[:nw]], [b.detach() for b in warp[nw:]], out_range)
        rays_o, rays_d, viewdirs = rays_o.contiguous(), rays_d.contiguous(), viewdirs.contiguous()
        n = rays_o.shape[0]
        eng.aux = {"depth": torch.empty(n, device=eng.device), "normal": torch.empty((n, 3), device=eng.device)}
        acc, cumw, last = eng.forward_sums_deform(rays_o, rays_d, viewdirs, jitter, want_samples=True,
                                                  depth=eng.aux["depth"], normal=eng.aux["normal"])
        idx, ray_id = eng.inbox_order(n)
        sout = eng._bufs["sout"][idx]
        eng.aux.update(idx=idx, ray_id=ray_id, sdf=sout[:, 0])
        ctx.eng, ctx.depth, ctx.nw = eng, depth, nw
        ctx.save_for_backward(rays_o, rays_d, viewdirs, jitter if jitter is not None else torch.empty(0), idx)
        return acc.clone(), cumw.clone(), last.clone(), sout[:, 1:15].contiguous()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_acc, g_cumw, g_last, g_samp):
        eng, depth, nw = ctx.eng, ctx.depth, ctx.nw
        rays_o, rays_d, viewdirs, jitter, idx = ctx.saved_tensors
        jitter = jitter if jitter.numel() else None
        eng.grad_flat.zero_()
        for g in eng.g_ww + eng.g_wb:
            g.zero_()
        ext = None
        if g_samp is not None and idx.numel():
            ext = torch.zeros((eng._bufs["cap"], 14), dtype=torch.float32, device=eng.device)
            ext[idx] = g_samp
        eng.backward_deform_from_ray_grads(rays_o, rays_d, viewdirs, g_acc, g_cumw, g_last, jitter, ext_grad=ext)
        _, g_k0, g_w, g_b = eng.grid_gradients()
        g_ww, g_wb = eng.warp_gradients()
        g_o = g_d = g_v = None
        if eng.pose_grad:
            nr = rays_o.shape[0]
            g_rays, g_v = eng.ray_gradients()
            g_o, g_d, g_v = g_rays[:nr, :3], g_rays[:nr, 3:], g_v[:nr]
        return (None, g_o, g_d, g_v, None, None, None, None, g_k0, None) + tuple(g_w) + tuple(g_b) + tuple(g_ww) + \
            tuple(g_wb)


class _LazyResult(dict):
    """The reference's result dict (voxurf_coarse.py:1069-1091).  Entries the training losses read are tensors with
    autograd; the per-sample / diagnostic entries the trainer only logs (`weights`, `raw_alpha`, `mask`, `depth`, ...:
    lib/recon_scene.py:725-736) are computed on first access."""

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if callable(v) and not torch.is_tensor(v):
            v = v()
            dict.__setitem__(self, key, v)
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default

    def items(self):
        return [(k, self[k]) for k in dict.keys(self)]

    def values(self):
        return [self[k] for k in dict.keys(self)]


class Voxurf(torch.nn.Module):
    """Surface-stage ``Voxurf`` (lib/voxurf_coarse.py:44-1215) on the fused kernels, both branches of ``forward`` with
    autograd: the plain SDF lookup (use_deform=False) and the implicit deformation (use_deform=True, the reference's
    default and live train step, lib/recon_scene.py:603).

    Same parameter names as the reference (``sdf.grid`` [1,1,X,Y,Z], ``k0.grid`` [1,12,X,Y,Z], ``rgbnet.*``,
    ``warp_network.*``, ``sdf_alpha``, ``sdf_beta``, ``s_val``), so ``last_ckpt.tar`` state dicts load with
    ``strict=False``, and the same ``forward(rays_o, rays_d, viewdirs, use_deform=True, global_step=None,
    **render_kwargs)`` returning the reference's result dict: ``rgb_marched``, ``alphainv_cum``, ``cum_weights``,
    ``gradient`` [M,3] of every in-box sample, ``k0_tv`` and (deform) ``grad_deform``, ``sdf_correct``, ``sdf_deform``
    carry autograd, so lib/losses.py:34-74 runs unchanged on it; ``weights``, ``raw_alpha``, ``raw_rgb``, ``mask``,
    ``mask_outbbox``, ``depth``, ``disp``, ``normal_marched``, ``s_val`` are detached (the trainer only logs them) and
    materialised on first access.  Rays that require grad receive the pose gradient.  ``inference`` renders through the
    deformation network without building a graph."""

    def __init__(self, xyz_min, xyz_max, num_voxels=0, num_voxels_base=0, rgbnet_dim=12, rgbnet_depth=3,
                 rgbnet_width=128, posbase_pe=5, viewbase_pe=1, geo_rgb_dim=3, s_ratio=2000, s_start=0.2,
                 step_start=0, barf_c2f=None, N_iters=20000, **kwargs):
        super().__init__()
        from .grid import DenseGrid
        if rgbnet_dim != 12 or rgbnet_width != 128 or geo_rgb_dim != 3:
            raise NotImplementedError("fused Voxurf: rgbnet_dim=12, rgbnet_width=128, geo_rgb_dim=3")
        self.register_buffer("xyz_min", torch.Tensor(xyz_min))
        self.register_buffer("xyz_max", torch.Tensor(xyz_max))
        self.num_voxels_base = num_voxels_base
        self.voxel_size = ((self.xyz_max - self.xyz_min).prod() / num_voxels).pow(1 / 3)          # :_set_grid_resolution
        self.world_size = ((self.xyz_max - self.xyz_min) / self.voxel_size).long()
        ws = self.world_size.tolist()
        self.sdf = DenseGrid(channels=1, world_size=ws, xyz_min=self.xyz_min, xyz_max=self.xyz_max)
        self.k0 = DenseGrid(channels=rgbnet_dim, world_size=ws, xyz_min=self.xyz_min, xyz_max=self.xyz_max)
        self.posbase_pe, self.viewbase_pe, self.barf_c2f, self.N_iters = posbase_pe, viewbase_pe, barf_c2f, N_iters
        dim0 = (3 + 3 * posbase_pe * 2) + (3 + 3 * viewbase_pe * 2) + rgbnet_dim + geo_rgb_dim             # :198-206
        nn = torch.nn
        self.rgbnet = nn.Sequential(
            nn.Linear(dim0, rgbnet_width), nn.ReLU(inplace=True),
            *[nn.Sequential(nn.Linear(rgbnet_width, rgbnet_width), nn.ReLU(inplace=True)) for _ in range(rgbnet_depth - 2)],
            nn.Linear(rgbnet_width, 3))
        nn.init.constant_(self.rgbnet[-1].bias, 0)
        self.sdf_alpha = nn.Parameter(torch.Tensor([10.0]))          # :118-119, use_mapping_fun
        self.sdf_beta = nn.Parameter(torch.Tensor([2.0]))
        self.s_val = nn.Parameter(torch.ones(1), requires_grad=False)
        self.s_ratio, self.s_start, self.step_start = s_ratio, s_start, step_start
        from .deform import DeformedImplicitField
        range_shape = kwargs.get("range_shape")
        self.range_shape = (self.xyz_max - self.xyz_min) if range_shape is None else torch.as_tensor(range_shape)
        self.warp_network = DeformedImplicitField(range_shape=self.range_shape)                       # :84-85
        self._eng = None

    def _engine(self, render_kwargs, pose_grad=False):
        from .ops import rgbnet_from_sequential
        ws, bs = rgbnet_from_sequential(self.rgbnet)
        key = (float(render_kwargs["near"]), float(render_kwargs["far"]), float(render_kwargs["stepsize"]),
               float(render_kwargs.get("bg", 0.0)), bool(pose_grad))
        if self._eng is None or self._eng_key != key:
            self._eng = FusedVoxurf(self.sdf.grid.detach(), self.k0.grid.detach(), [w.detach() for w in ws],
                                    [b.detach() for b in bs], self.xyz_min.tolist(), self.xyz_max.tolist(),
                                    stepsize=key[2], near=key[0], far=key[1], bg=key[3], s_val=float(self.s_val),
                                    sdf_alpha=float(self.sdf_alpha), sdf_beta=float(self.sdf_beta),
                                    posbase_pe=self.posbase_pe, viewbase_pe=self.viewbase_pe, barf_c2f=self.barf_c2f,
                                    voxel_size=float(self.voxel_size), pose_grad=pose_grad)
            self._eng_key = key
        return self._eng, ws, bs

    def grid_sampler(self, xyz, *grids, mode=None, align_corners=True, smooth=False, displace=0., use_custom=False):
        """Voxurf.grid_sampler (lib/voxurf_coarse.py:522-543): trilinear lookup with border padding, or with
        use_custom=True the twice-differentiable grid_sample_3d (:545-659).  xyz [...,3] -> [...,C] (squeezed)."""
        from .ops import trilinear
        if (mode not in (None, "bilinear")) or not align_corners or smooth or displace != 0:
            raise NotImplementedError("grid_sampler: bilinear, align_corners=True, no smoothing / displacement")
        shape = xyz.shape[:-1]
        grid = grids[0]
        vals = trilinear(grid, xyz.reshape(-1, 3), self.xyz_min, self.xyz_max, "custom" if use_custom else "border")
        return vals.reshape(*shape, grid.shape[1]).squeeze()

    def grid_sample_3d(self, image, optical):
        from .ops import grid_sample_3d
        return grid_sample_3d(image, optical)

    def k0_total_variation(self, k0_tv=1., k0_grad_tv=0.):
        """voxurf_coarse.py:443-455 (no nonempty mask in this module)."""
        from .ops import total_variation
        if k0_grad_tv > 0:
            raise NotImplementedError
        return total_variation(self.k0.grid) if k0_tv > 0 else 0

    def forward(self, rays_o, rays_d, viewdirs, use_deform=True, global_step=None, **render_kwargs):
        """Same signature, default and result dict as the reference (lib/voxurf_coarse.py:922-1091)."""
        if not torch.is_grad_enabled() and use_deform:
            return self.inference(rays_o, rays_d, viewdirs, global_step=global_step, **render_kwargs)
        pose = bool(rays_o.requires_grad or rays_d.requires_grad or viewdirs.requires_grad)
        eng, ws, bs = self._engine(render_kwargs, pose_grad=pose)
        if global_step is not None:                      # neus_alpha_from_sdf_scatter, :485-488, and the BARF progress
            s_val = 1.0 / (global_step + self.s_ratio / self.s_start - self.step_start) * self.s_ratio
            self.s_val.data.fill_(s_val)
            eng.set_schedule(progress=global_step / self.N_iters)
        else:
            eng.set_schedule(progress=1.0)
        eng.s_val = float(self.s_val)
        eng._refresh_cfg()
        n = rays_o.shape[0]
        # sample_ray_ori's per-ray training jitter (:712-713); render_kwargs['jitter'] pins it (tests, replays)
        jitter = render_kwargs.get("jitter")
        if jitter is None and global_step is not None:
            jitter = torch.rand(n, device=rays_o.device)
        ab = torch.cat([self.sdf_alpha, self.sdf_beta])
        bg = render_kwargs.get("bg", 0.0)
        out = _LazyResult()
        if use_deform:
            wws, wbs = self.warp_network.parameters_for_kernel()
            acc, cumw, last, samp = _VoxurfDeformFn.apply(
                eng, rays_o, rays_d, viewdirs, jitter, len(ws), self.warp_network.output_range, self.sdf.grid.detach(),
                self.k0.grid, ab.detach(), *ws, *bs, *wws, *wbs)
            gradient = samp[:, 0:3]
            out.update({"sdf_deform": samp[:, 13], "grad_deform": samp[:, 3:12].reshape(-1, 3, 3),
                        "sdf_correct": samp[:, 12:13]})
            sdf_s = eng.aux["sdf"]
        else:
            acc, cumw, last, samp = _VoxurfRenderFn.apply(eng, rays_o, rays_d, viewdirs, jitter, len(ws), self.sdf.grid,
                                                          self.k0.grid, ab, *ws, *bs)
            gradient = samp[:, 1:4]
            sdf_s = samp[:, 0].detach()
        rgb_marched = (acc + (1 - cumw[:, None]) * bg).clamp(0, 1)                                     # :1040-1042
        aux = eng.aux
        ray_id = aux["ray_id"]
        dist = float(render_kwargs["stepsize"]) * float(self.voxel_size)
        memo = {}

        def sample_stats():
            # alpha / weights of ALL in-box samples from the operators (NeuS alpha, alpha2weight), detached
            if not memo:
                from .ops import Alphas2Weights, neus_alpha
                with torch.no_grad():
                    alpha = neus_alpha(viewdirs.detach().contiguous(), ray_id, dist, sdf_s.detach().contiguous(),
                                       gradient.detach().contiguous(), 1.0 / float(self.s_val))
                    weights, _ = Alphas2Weights.apply(alpha, ray_id, n)
                memo.update(alpha=alpha, weights=weights)
            return memo

        def depth():
            # train-mode definition (:1052-1058): t_min / |d| + sum_i w_i step_i (step in world units)
            with torch.no_grad():
                vec = torch.where(rays_d == 0, torch.full_like(rays_d, 1e-6), rays_d)
                ra, rb = (self.xyz_max - rays_o) / vec, (self.xyz_min - rays_o) / vec
                t_min = torch.minimum(ra, rb).amax(-1).clamp(min=render_kwargs["near"], max=render_kwargs["far"])
                return t_min / rays_d.norm(dim=-1) + aux["depth"]

        def dense_mask():
            S = int(float((self.world_size.float() + 1).norm()) / float(render_kwargs["stepsize"])) + 1      # :700
            cnt = eng._bufs["dray_cnt"][:n].long()
            return (torch.arange(S, device=cnt.device)[None, :] < cnt[:, None]).flatten()

        def raw_rgb():
            # colour of EVERY in-box sample (the fused path only evaluates the composited ones): op-level path
            from .ops import rgbnet as rgbnet_op
            with torch.no_grad():
                pts = eng._bufs["dpts"][aux["idx"], :3]
                xyz = (pts - self.xyz_min) / (self.xyz_max - self.xyz_min)
                L, Lv = self.posbase_pe, self.viewbase_pe
                pf = torch.tensor([2.0 ** i for i in range(L)], device=pts.device)
                vf = torch.tensor([2.0 ** i for i in range(Lv)], device=pts.device)
                wp = torch.tensor(barf_weights(eng.progress, self.barf_c2f, L), device=pts.device)
                wv = torch.tensor(barf_weights(eng.progress, self.barf_c2f, Lv), device=pts.device)
                xe = (xyz.unsqueeze(-1) * pf).flatten(-2)
                xe = torch.cat([(xe.sin().view(-1, 3, L) * wp).flatten(-2), (xe.cos().view(-1, 3, L) * wp).flatten(-2)], -1)
                vd = viewdirs.detach()
                ve = (vd.unsqueeze(-1) * vf).flatten(-2)
                ve = torch.cat([(ve.sin().view(-1, 3, Lv) * wv).flatten(-2), (ve.cos().view(-1, 3, Lv) * wv).flatten(-2)], -1)
                g = gradient.detach()
                feat = torch.cat([self.k0(pts), xyz, xe, torch.cat([vd, ve], -1)[ray_id],
                                  g / (g.norm(dim=-1, keepdim=True) + 1e-5)], -1)
                return rgbnet_op(feat.contiguous(), [w.detach() for w in ws], [b.detach() for b in bs], True)

        out.update({
            "alphainv_cum": last, "cum_weights": cumw[:, None], "rgb_marched": rgb_marched, "gradient": gradient,
            "s_val": float(self.s_val), "k0_tv": self.k0_total_variation(),
            "weights": lambda: sample_stats()["weights"], "raw_alpha": lambda: sample_stats()["alpha"],
            "raw_rgb": raw_rgb, "ray_id": ray_id,
            "normal_marched": aux["normal"] if render_kwargs.get("render_grad", False) else None,
            "depth": depth, "disp": lambda: 1 / out["depth"],
            # `mask` = ~mask_outbbox over the DENSE [N * S] sample grid before the compaction (:937-944; the trainer
            # logs its mean): the per-ray in-box counts are exact, the flags are packed at the start of each ray's
            # row.  After the compaction every remaining sample is inside the box.
            "mask": dense_mask, "mask_outbbox": lambda: torch.zeros_like(ray_id, dtype=torch.bool),
        })
        return out

    def _rgbnet_params(self):
        from .ops import rgbnet_from_sequential
        ws, bs = rgbnet_from_sequential(self.rgbnet)
        return [w.detach() for w in ws], [b.detach() for b in bs]

    @torch.no_grad()
    def inference(self, rays_o, rays_d, viewdirs, global_step=None, **render_kwargs):
        """Voxurf.inference (lib/voxurf_coarse.py:1093-1215): render with the implicit deformation - SDF and normal
        of every in-box sample from the warped template lookup - forward only.  Returns the per-ray entries that
        lib/nvs_fun.py:74-88 reads: rgb_marched, alphainv_cum, cum_weights, depth, disp, normal_marched."""
        render_kwargs.pop("training", None)
        eng, _, _ = self._engine(render_kwargs, pose_grad=self._eng_key[4] if self._eng is not None else False)
        ws, bs = self.warp_network.parameters_for_kernel()
        eng.load_params(self.sdf.grid, self.k0.grid, *self._rgbnet_params(), float(self.sdf_alpha),
                        float(self.sdf_beta))
        eng.set_warp_network([w.detach() for w in ws], [b.detach() for b in bs], self.warp_network.output_range)
        eng.set_schedule(s_val=float(self.s_val), progress=1.0 if global_step is None else global_step / self.N_iters)
        rgb, last, cumw, depth, normal = eng.render(rays_o, rays_d, viewdirs, want_depth=True, want_normal=True,
                                                    use_deform=True)
        return {"rgb_marched": rgb, "alphainv_cum": last, "cum_weights": cumw[:, None], "depth": depth,
                "disp": 1 / depth, "normal_marched": normal, "s_val": float(self.s_val)}


    @torch.no_grad()
    def _query_surface(self, rays_o, rays_d, global_step, keep_dim, return_depth, use_deform, render_kwargs):
        """Depth-only march of `query_sdf_point_wocuda_render` (lib/voxurf_coarse.py:839-920): NeuS alpha +
        alpha2weight along every ray, n_step = sum_i w_i step_i, depth = t_min + n_step / |d|, surface point
        rays_o + rays_d * depth, valid where n_step > 0.  One march launch (plus the deformation network when
        use_deform), no colour network.  Forward only: the reference lets the projection loss of
        lib/recon_scene.py:622-637 reach the model through this depth; here the points are detached."""
        render_kwargs = dict(render_kwargs)
        render_kwargs.pop("training", None)
        jitter = render_kwargs.pop("jitter", None)
        eng, _, _ = self._engine(render_kwargs, pose_grad=self._eng_key[4] if self._eng is not None else False)
        eng.load_params(self.sdf.grid, self.k0.grid, *self._rgbnet_params(), float(self.sdf_alpha), float(self.sdf_beta))
        if use_deform:
            ws, bs = self.warp_network.parameters_for_kernel()
            eng.set_warp_network([w.detach() for w in ws], [b.detach() for b in bs], self.warp_network.output_range)
        if global_step is not None:          # neus_alpha_from_sdf_scatter(is_train=True): the s_val schedule (:485-488)
            self.s_val.data.fill_(1.0 / (global_step + self.s_ratio / self.s_start - self.step_start) * self.s_ratio)
        eng.set_schedule(s_val=float(self.s_val), progress=1.0 if global_step is None else global_step / self.N_iters)
        rays_o, rays_d = rays_o.detach().contiguous(), rays_d.detach().contiguous()
        n = rays_o.shape[0]
        if jitter is None and global_step is not None:
            jitter = torch.rand(n, device=rays_o.device)                                     # sample_ray_ori, :712-713
        viewdirs = (rays_d / rays_d.norm(dim=-1, keepdim=True)).contiguous()
        b = eng._buffers(n)
        n_step = torch.empty(n, dtype=torch.float32, device=eng.device)
        with torch.cuda.device(eng.device):
            (eng._march_deform if use_deform else eng._march)(b, rays_o, rays_d, viewdirs, jitter, n_step, None)
        vec = torch.where(rays_d == 0, torch.full_like(rays_d, 1e-6), rays_d)
        ra, rb = (self.xyz_max - rays_o) / vec, (self.xyz_min - rays_o) / vec
        t_min = torch.minimum(ra, rb).amax(-1).clamp(min=render_kwargs["near"], max=render_kwargs["far"])
        depth = t_min[..., None] + n_step[..., None] / rays_d.norm(dim=-1, keepdim=True)
        mask = n_step > 0.
        if keep_dim:
            return rays_o + rays_d * depth, mask, depth
        pts_sdf0 = rays_o[mask] + rays_d[mask] * depth[mask]
        return (pts_sdf0, depth[mask]) if return_depth else pts_sdf0

    def query_sdf_point_wocuda_render(self, rays_o, rays_d, global_step=None, keep_dim=False, return_depth=False,
                                      use_deform=True, **render_kwargs):
        """lib/voxurf_coarse.py:839-920 (the second, small render of every train step: lib/recon_scene.py:622-637)."""
        return self._query_surface(rays_o, rays_d, global_step, keep_dim, return_depth, use_deform, render_kwargs)


def smoke(dev):
    """Tiny fused fine-stage train step (used by __graft_entry__.smoke): runs the march, both tcgen05 rgbnet
    kernels, the reverse march and the optimiser tail once."""
    import torch.nn as nn
    from . import ops, synthetic
    G, N = 24, 256
    torch.manual_seed(0)
    sdf = synthetic.make_sdf_grid(G, 0.5, 0.01)
    k0 = 0.3 * torch.randn(1, 12, G, G, G)
    net = nn.Sequential(nn.Linear(57, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(), nn.Linear(128, 3))
    ws, bs = ops.rgbnet_from_sequential(net)
    eng = FusedVoxurf(sdf.to(dev), k0.to(dev), [w.detach() for w in ws], [b.detach() for b in bs], [-1, -1, -1],
                      [1, 1, 1], stepsize=0.5, s_val=0.05, bg=0.5, sdf_alpha=10.0, sdf_beta=0.35)
    o, d, v = synthetic.make_rays(N)
    tgt = torch.rand(N, 3, device=dev)
    loss = eng.train_step(o.to(dev), d.to(dev), v.to(dev), tgt)
    torch.cuda.synchronize()
    assert torch.isfinite(loss).all() and float(loss[0]) > 0 and eng.n_samples() > 0
    return float(loss[0]), eng, (o, d, v)


class DistributedVoxurf:
    """FusedVoxurf replicated over the ranks of a process group: every rank marches its own ray batch; the loss is
    the global mean over all ranks' rays, so N ranks with batches b_0..b_{N-1} equal one rank stepping on their
    concatenation (up to fp32 summation order).

    mode="nvls" (default when the system offers a multicast address): the feature records - 3/4 of the gradient
    bytes and of the optimiser work - go through ONE kernel per rank that reduces the gradient slice in the NVSwitch
    (multimem.ld_reduce), applies Adam to 1/world of the records (moments exist for the slice only) and multicasts
    the result into every replica (pp_records_adam_exchange_mc); the SDF gradient records (coupled by the
    finite-difference fold, so every rank needs all of them) and the rgbnet gradients are summed in place by the same
    switch path without an optimiser (pp_allreduce_mc) and applied identically on every rank.  The cross-rank
    barriers are flag words in peer memory (the first is fused into the exchange kernel and carries the global loss):
    no NCCL call on the step.
    mode="nccl": the flat gradient buffer (SDF-record, feature-record and rgbnet gradients) is summed with ONE NCCL
    all-reduce and every rank applies the full optimiser step (the comparison baseline)."""

    def __init__(self, sdf, k0, rgbnet_weights, rgbnet_biases, *args, group=None, mode="auto", **kw):
        import torch.distributed as dist
        from . import dist as ppdist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        kw = dict(kw)
        kw["w_rgb"] = kw.get("w_rgb", 1.0) / self.world
        auto = mode == "auto"
        if auto:
            mode = "nvls" if (self.world in (2, 4, 8) and ppdist.nvls_available() and not kw.get("w_tv_k0")) else "nccl"
        if mode not in ("nvls", "nccl"):
            raise ValueError("DistributedVoxurf: mode must be 'auto', 'nvls' or 'nccl'")
        if self.world == 1:
            mode = "nccl"
        dev = sdf.device
        if mode == "nvls":
            if kw.get("w_tv_k0"):
                raise RuntimeError("DistributedVoxurf(mode='nvls'): the k0 total-variation term needs mode='nccl'")
            V = sdf.numel()
            # (the engine holds the first layer permuted and zero-padded to FEAT_COLS input columns)
            n_net = (rgbnet_weights[0].shape[0] * FEAT_COLS + sum(w.numel() for w in rgbnet_weights[1:]) +
                     sum(b.numel() for b in rgbnet_biases))
            n_grad_rec = (16 * V + n_net + 3) // 4
            ok, err = True, None
            try:
                self.k0_mem = ppdist.SymmRecords(3 * V, dev, group, self.world)
                self.grad_mem = ppdist.SymmRecords(n_grad_rec, dev, group, self.world, zero=True)
                n_sync = int(lib.pp_peer_sync_bytes()) // 16
                self.sync_mem = ppdist.SymmRecords(n_sync, dev, group, self.world, zero=True)
                ok = bool(self.k0_mem.multicast_ptr and self.grad_mem.multicast_ptr)
            except Exception as e:      # no symmetric memory / multicast on this system
                ok, err = False, e
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if not int(flag):
                if not auto:
                    raise RuntimeError("DistributedVoxurf(mode='nvls'): no multicast address (%s)" % (err,))
                mode = "nccl"
        self.mode = mode
        if mode == "nvls":
            b, e = ppdist.shard_bounds(3 * V, self.world, self.rank)
            self.eng = FusedVoxurf(sdf, k0, rgbnet_weights, rgbnet_biases, *args, k0rec_buf=self.k0_mem.tensor,
                                   grad_flat_buf=self.grad_mem.tensor, k0_moment_records=max(e - b, 1), **kw)
            self.sync_ptrs = self.sync_mem.table
            self.epoch = 0
            self.loss_global = torch.zeros(2, dtype=torch.float32, device=dev)
            assert self.eng.n_net == n_net
            self.n_net_rec = (n_net + 3) // 4
            torch.cuda.synchronize(dev)
            dist.barrier(group=group)
        else:
            self.eng = FusedVoxurf(sdf, k0, rgbnet_weights, rgbnet_biases, *args, **kw)

    def train_step(self, rays_o, rays_d, viewdirs, target, jitter=None):
        eng = self.eng
        eng.loss_buf.zero_()
        eng.backward_only(rays_o, rays_d, viewdirs, target, jitter)
        if self.mode == "nccl":
            if self.world > 1:
                self.dist.all_reduce(eng.grad_flat, group=self.group)
                self.dist.all_reduce(eng.loss_buf, group=self.group)
            eng.optimizer_step()
            eng.loss_buf[1] /= self.world          # [global w_rgb-weighted loss, global mean-squared error]
            return eng.loss_buf
        b1, b2, _, lr_k0, _, eps = eng.hp
        V = eng.n_voxels
        g_mc = self.grad_mem.multicast_ptr
        with torch.cuda.device(eng.device):
            s = stream()
            self.epoch += 1
            # feature records: in-switch reduce-scatter + Adam on this rank's slice + multicast store; the prologue is
            # the "all gradients complete" barrier and sums the ranks' [loss, mse]
            check(lib.pp_records_adam_exchange_mc(ptr(eng.k0rec), self.k0_mem.multicast_ptr, g_mc + 16 * V,
                                                  ptr(eng.m_k0), ptr(eng.v_k0), 3 * V, self.world, self.rank,
                                                  eng.step + 1, b1, b2, lr_k0, eps, 1 if eng.adam_dense else 0,
                                                  self.sync_ptrs, self.epoch, ptr(eng.loss_buf), ptr(self.loss_global),
                                                  s), "records_adam_exchange_mc")
            # SDF gradient records and rgbnet gradients: summed in place, applied identically by every rank below
            check(lib.pp_allreduce_mc(g_mc, V, self.world, self.rank, self.sync_ptrs, s), "allreduce_mc")
            check(lib.pp_allreduce_mc(g_mc + 64 * V, self.n_net_rec, self.world, self.rank, self.sync_ptrs, s),
                  "allreduce_mc")
            self.epoch += 1
            check(lib.pp_peer_sync(self.sync_ptrs, self.world, self.rank, self.epoch, None, None, s), "peer_sync")
        eng.optimizer_step(skip_k0=True)
        eng.g_k0rec.zero_()
        self.loss_global[1] /= self.world
        return self.loss_global

    def check_health(self):
        """Raises if a cross-rank barrier of this rank ever timed out (see DistributedDVGO.check_health)."""
        if self.mode == "nvls":
            with torch.cuda.device(self.eng.device):
                if lib.pp_peer_sync_timed_out(self.sync_ptrs[self.rank]) != 0:
                    raise RuntimeError("DistributedVoxurf: a cross-rank barrier timed out on rank %d" % self.rank)

    def close(self):
        if self.mode == "nvls":
            torch.cuda.synchronize(self.eng.device)
            self.dist.barrier(group=self.group)
