"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the Voxurf fine-stage render / train step.

Follows ``Voxurf.forward`` of the reference, non-deform branch (lib/voxurf_coarse.py:922-1091), op by op in
plain PyTorch so that autograd yields the reference's gradients:

    sample (CUDA sampler semantics, render_utils_kernel.cu:196-242 via the C oracle; optional per-ray jitter of
    sample_ray_ori :712-713) -> ``~mask_outbbox`` compaction (:686-693 / :940-945) -> sdf grid remap (:946-949)
    -> neus_sdf_gradient 'interpolate' (:463-467) -> grid_sampler (border, :522-543) of sdf and gradient ->
    neus_alpha_from_sdf_scatter (:494-517) -> Alphas2Weights (:1316-1332, C oracle) -> k0 DenseGrid (lib/grid.py:
    47-58) -> rgb_feat assembly with BARF weights (:1005-1030, :721-732) -> rgbnet + sigmoid (:1032-1033) ->
    segment_coo sums restated as index_add_ (:1034-1042) -> photometric MSE (lib/losses.py:26-29).

With ``warp=`` the implicit-deformation branch (:958-985) runs instead of the plain SDF lookup: the warp MLP
(``DeformedImplicitField``, lib/deformation/deform_net.py:12-29: 3 -> 128 x4 -> 4, ReLU, output times the box extent),
the template SDF at the warped point through the twice-differentiable 8-gather sampler ``grid_sample_3d`` (:545-659),
the Jacobian of the warped point (``grad_deform``), the scalar correction, and the normal as d sdf_final / d point by
``autograd.grad(create_graph=True)``; ``deform_implicit_loss`` restates lib/losses.py:6-24.  Pinned by
tests/golden/voxurf_deform.npz (reference code run by tests/golden/make_golden_deform.py).

Pinned: tests/golden/voxurf_forward.npz holds the outputs of the reference's OWN Voxurf.forward (executed unmodified on
the CPU by tests/golden/make_golden_voxurf.py); tests/test_oracle_golden.py replays them against this module (2e-5).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
import math

import torch
import torch.nn.functional as F

from . import oracle
from .dvgo_cpu import _Alphas2Weights


def softplus10(x):
    return F.softplus(torch.as_tensor(float(x)), beta=10).item()


def sdf_gradient_grid(sdf, voxel_size):
    """neus_sdf_gradient(mode='interpolate'), voxurf_coarse.py:463-467."""
    g = torch.zeros([1, 3] + list(sdf.shape[-3:]), dtype=sdf.dtype)
    gx = (sdf[:, 0, 2:, :, :] - sdf[:, 0, :-2, :, :]) / 2 / voxel_size
    gy = (sdf[:, 0, :, 2:, :] - sdf[:, 0, :, :-2, :]) / 2 / voxel_size
    gz = (sdf[:, 0, :, :, 2:] - sdf[:, 0, :, :, :-2]) / 2 / voxel_size
    g = torch.cat([F.pad(gx, (0, 0, 0, 0, 1, 1))[:, None], F.pad(gy, (0, 0, 1, 1, 0, 0))[:, None],
                   F.pad(gz, (1, 1, 0, 0, 0, 0))[:, None]], 1)
    return g


def grid_sampler(xyz, grid, xyz_min, xyz_max, padding_mode):
    ind = ((xyz - xyz_min) / (xyz_max - xyz_min)).flip((-1,)) * 2 - 1
    out = F.grid_sample(grid, ind.reshape(1, 1, 1, -1, 3), mode="bilinear", align_corners=True,
                        padding_mode=padding_mode)
    return out.reshape(grid.shape[1], -1).T


def grid_sample_custom(xyz, grid, xyz_min, xyz_max):
    """Voxurf.grid_sampler(use_custom=True) -> grid_sample_3d, voxurf_coarse.py:522-538, :545-659: trilinear weights from
    the UNCLAMPED floor of the continuous voxel coordinate, corner indices clamped into the grid afterwards; written
    with differentiable torch ops only so that it can be differentiated twice with respect to the points.
    xyz [M,3], grid [1,C,X,Y,Z] -> [M,C]."""
    _, C, X, Y, Z = grid.shape
    dims = torch.tensor([X, Y, Z], dtype=xyz.dtype)
    ind_norm = ((xyz - xyz_min) / (xyz_max - xyz_min)) * 2 - 1
    pos = ((ind_norm + 1) / 2) * (dims - 1)                      # continuous voxel coordinate per axis
    with torch.no_grad():
        base = torch.floor(pos)
    hi_w = pos - base                                            # weight of the +1 corner along each axis
    lo_w = (base + 1) - pos
    flat = grid.reshape(C, -1)
    out = 0
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                with torch.no_grad():
                    ix = (base[:, 0] + dx).clamp(0, X - 1)
                    iy = (base[:, 1] + dy).clamp(0, Y - 1)
                    iz = (base[:, 2] + dz).clamp(0, Z - 1)
                    lin = ((ix * Y + iy) * Z + iz).long()
                w = (hi_w[:, 0] if dx else lo_w[:, 0]) * (hi_w[:, 1] if dy else lo_w[:, 1]) * \
                    (hi_w[:, 2] if dz else lo_w[:, 2])
                out = out + flat[:, lin].T * w[:, None]
    return out


def warp_mlp(x, weights, biases, out_range):
    """DeformedImplicitField.forward (deform_net.py:23-29) on top of SingleBVPNet / FCBlock (modules.py:43-124):
    Linear+ReLU x (len-1), Linear; output scaled by ``out_range`` (= range_shape.max()).
    -> (deformation [M,3], correction [M,1])."""
    h = x
    for i, (w, b) in enumerate(zip(weights, biases)):
        h = F.linear(h, w, b)
        if i < len(weights) - 1:
            h = torch.relu(h)
    h = h * out_range
    return h[:, :3], h[:, 3:]


def deform_implicit_loss(out, use_deform=True):
    """lib/losses.py:6-24 -> dict of the mean constraints (Eikonal, Jacobian norm, correction, deformation)."""
    d = {"grad_constraint": (out["gradient"].norm(dim=-1) - 1).abs().mean()}
    if use_deform:
        d["grad_deform_constraint"] = out["grad_deform"].norm(dim=-1).mean()
        d["sdf_correct_constraint"] = out["sdf_correct"].abs().mean()
        d["sdf_deform_constraint"] = out["sdf_deform"].abs().mean()
    return d


def barf_weights(progress, barf_c2f, L):
    if barf_c2f is None:
        return torch.ones(L)
    start, end = barf_c2f
    alpha = (progress - start) / (end - start) * L
    k = torch.arange(L, dtype=torch.float32)
    return (1 - (alpha - k).clamp(min=0, max=1).mul(math.pi).cos()) / 2


def sample(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist, jitter=None):
    """-> pts [M,3], ray_id [M], step_id [M] of the in-box samples (reference order)."""
    pts, mask, ray_id, step_id, n_steps, t_min, t_max = oracle.sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max,
                                                                                   near, far, stepdist)[:7]
    if jitter is not None:
        # same enumeration, points shifted by jitter steps along the normalised direction
        start, dirn = oracle.infer_ray_start_dir(rays_o, rays_d, t_min)
        dist = stepdist * (step_id.float() + jitter[ray_id])
        pts = start[ray_id] + dirn[ray_id] * dist[:, None]
        mask = ((xyz_min > pts) | (pts > xyz_max)).any(-1)
    keep = ~mask
    return pts[keep], ray_id[keep], step_id[keep]


def forward(sdf_grid, k0_grid, rgbnet, rays_o, rays_d, viewdirs, xyz_min, xyz_max, stepsize, voxel_size, s_val,
            near=0.2, far=1e9, bg=0.0, sdf_alpha=None, sdf_beta=None, posbase_pe=5, viewbase_pe=1, barf_c2f=None,
            progress=1.0, jitter=None, differentiable_points=False, warp=None):
    """sdf_grid [1,1,X,Y,Z], k0_grid [1,12,X,Y,Z] (may require grad), rgbnet: callable rows -> logits.
    warp: None, or (weights, biases, out_range) of the deformation network -> implicit-deformation branch.
    differentiable_points: rebuild the sample positions from (rays_o, rays_d) with torch ops (the formulas of
    sample_ray_ori, voxurf_coarse.py:702-716) so that autograd reaches the rays (pose gradient)."""
    xyz_min = torch.as_tensor(xyz_min, dtype=torch.float32)
    xyz_max = torch.as_tensor(xyz_max, dtype=torch.float32)
    n = rays_o.shape[0]
    stepdist = stepsize * voxel_size
    with torch.no_grad():
        pts, ray_id, step_id = sample(rays_o.detach(), rays_d.detach(), xyz_min, xyz_max, near, far, stepdist, jitter)
    if differentiable_points:
        vec = torch.where(rays_d == 0, torch.full_like(rays_d, 1e-6), rays_d)
        t_min = torch.minimum((xyz_max - rays_o) / vec, (xyz_min - rays_o) / vec).amax(-1).clamp(min=near, max=far)
        steps = step_id.float() + (jitter[ray_id] if jitter is not None else 0.0)
        interpx = t_min[ray_id] + stepdist * steps / rays_d.norm(dim=-1)[ray_id]
        pts = rays_o[ray_id] + rays_d[ray_id] * interpx[:, None]
    grid = sdf_grid
    if sdf_alpha is not None:
        grid = softplus10(sdf_alpha) * (torch.sigmoid(softplus10(sdf_beta) * sdf_grid) - 0.5)
    extra = {}
    if warp is None:
        sdf = grid_sampler(pts, grid, xyz_min, xyz_max, "border")[:, 0]
        gradient = grid_sampler(pts, sdf_gradient_grid(grid, voxel_size), xyz_min, xyz_max, "border")
    else:
        # implicit deformation, voxurf_coarse.py:958-985 (bool_mapping_deformspace = False, :104)
        if not pts.requires_grad:
            pts = pts.detach().requires_grad_(True)
        deform, correction = warp_mlp(pts, *warp)
        new_coords = deform + pts
        sdf_template = grid_sample_custom(new_coords, grid, xyz_min, xyz_max)[:, 0]
        ones = torch.ones_like(new_coords[:, 0])
        grad_deform = torch.stack([torch.autograd.grad(new_coords[:, k], [pts], grad_outputs=ones, create_graph=True)[0]
                                   for k in range(3)], dim=2)
        sdf = sdf_template + correction.squeeze(-1)
        sdf_deform = sdf - grid_sample_custom(pts, grid, xyz_min, xyz_max)[:, 0]
        gradient = torch.autograd.grad(sdf, [pts], grad_outputs=torch.ones_like(sdf), create_graph=True)[0]
        extra = {"sdf_deform": sdf_deform, "grad_deform": grad_deform, "sdf_correct": correction}
    # neus_alpha_from_sdf_scatter, :494-517
    dirs = viewdirs[ray_id]
    inv_s = 1.0 / s_val
    true_cos = (dirs * gradient).sum(-1, keepdim=True)
    iter_cos = -F.relu(-true_cos)
    s = sdf.unsqueeze(-1)
    nxt = s + iter_cos * stepdist * 0.5
    prv = s - iter_cos * stepdist * 0.5
    prev_cdf = torch.sigmoid(prv * inv_s)
    next_cdf = torch.sigmoid(nxt * inv_s)
    alpha = ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).clip(0.0, 1.0).squeeze(-1)
    weights, alphainv_last = _Alphas2Weights.apply(alpha, ray_id, n)
    # colour features, :1005-1030
    k0 = grid_sampler(pts, k0_grid, xyz_min, xyz_max, "zeros")
    rays_xyz = (pts - xyz_min) / (xyz_max - xyz_min)
    posfreq = torch.tensor([2.0 ** i for i in range(posbase_pe)])
    viewfreq = torch.tensor([2.0 ** i for i in range(viewbase_pe)])
    xyz_emb = (rays_xyz.unsqueeze(-1) * posfreq).flatten(-2)
    xyz_emb = torch.cat([xyz_emb.sin(), xyz_emb.cos()], -1)
    if posbase_pe:
        xyz_emb = (xyz_emb.reshape(-1, posbase_pe) * barf_weights(progress, barf_c2f, posbase_pe)).reshape(xyz_emb.shape)
    xyz_emb = torch.cat([rays_xyz, xyz_emb], -1)
    view_emb = (viewdirs.unsqueeze(-1) * viewfreq).flatten(-2)
    view_emb = torch.cat([view_emb.sin(), view_emb.cos()], -1)
    if viewbase_pe:
        view_emb = (view_emb.reshape(-1, viewbase_pe) * barf_weights(progress, barf_c2f, viewbase_pe)).reshape(view_emb.shape)
    view_emb = torch.cat([viewdirs, view_emb], -1)
    normal = gradient / (gradient.norm(dim=-1, keepdim=True) + 1e-5)
    rgb_feat = torch.cat([k0, xyz_emb, view_emb[ray_id], normal], -1)
    rgb = torch.sigmoid(rgbnet(rgb_feat))
    rgb_acc = torch.zeros(n, 3).index_add_(0, ray_id, weights.unsqueeze(-1) * rgb)
    cum_weights = torch.zeros(n, 1).index_add_(0, ray_id, weights.unsqueeze(-1))
    rgb_marched = (rgb_acc + (1 - cum_weights) * bg).clamp(0, 1)
    return {"rgb_marched": rgb_marched, "alphainv_cum": alphainv_last, "weights": weights, "cum_weights": cum_weights,
            "raw_alpha": alpha, "raw_rgb": rgb, "ray_id": ray_id, "step_id": step_id, "gradient": gradient,
            "sdf": sdf, "rgb_feat": rgb_feat, "normal": normal, "k0": k0, "pts": pts, **extra}


def mse_loss(out, target, w_rgb=1.0):
    return w_rgb * F.mse_loss(out["rgb_marched"], target)
