"""TEST INFRASTRUCTURE ONLY -- closed-form (no autograd) restatement of the implicit-deformation branch, forward AND
backward, in the four-rows-per-sample formulation the tcgen05 kernels use.

The reference obtains the warp Jacobian, the normal and every gradient of the deform branch from torch.autograd with
create_graph=True (lib/voxurf_coarse.py:958-985).  The kernels cannot: they need the derivatives as formulas.  This
module states them in plain fp32 torch ops and tests/test_voxurf_cpu.py checks them against
oracle/voxurf_cpu.forward(warp=...) + autograd (which tests/test_oracle_golden.py pins against the reference's own code):

* ``warp_rows_forward``   value row + three tangent rows through a ReLU MLP: t_l = (z_l > 0) * (W_l t_{l-1})
* ``warp_rows_backward``  the exact weight / bias / input gradients from gradients w.r.t. net(x) AND its Jacobian: the
                          same layer-by-layer backward applied to the four rows, gates taken from the value row
* ``lookup_forward``      x' = x + R d, trilinear value s and world-space gradient g of the template SDF with the custom
                          sampler's arithmetic (unclamped floor, clamped corners, :545-659); sdf = s + R c;
                          gradient_i = sum_k (delta_ki + R dd_k/dx_i) g_k + R dc/dx_i
* ``lookup_backward``     from G_s = dL/d sdf, G_g = dL/d gradient: gradients w.r.t. the 16 network outputs of the
                          sample and w.r.t. x' (through the cell's off-diagonal Hessian)
* ``regulariser_grads``   deform_implicit_loss (lib/losses.py:6-24) terms as direct gradients

Only tests/ may import this module.
"""
import torch


# ---- warp network on four rows per sample ------------------------------------------------------------------------
def warp_rows_forward(x, weights, biases):
    """-> y [M,4] = net(x), Jt [M,3,4] with Jt[m,i,k] = d y_k / d x_i, cache for the backward."""
    h = x
    T = torch.eye(3, dtype=x.dtype).expand(x.shape[0], 3, 3)          # tangent rows e_0, e_1, e_2
    acts = []
    for w, b in zip(weights[:-1], biases[:-1]):
        z = h @ w.T + b
        gate = (z > 0).to(x.dtype)
        acts.append((h, T, gate))
        h = torch.relu(z)
        T = gate[:, None, :] * (T @ w.T)
    acts.append((h, T, None))
    y = h @ weights[-1].T + biases[-1]
    Jt = T @ weights[-1].T
    return y, Jt, acts


def warp_rows_backward(weights, acts, G_y, G_Jt):
    """G_y [M,4] = dL/d net(x), G_Jt [M,3,4] = dL/d Jt -> ([dW_l], [db_l], dL/dx through the value row)."""
    gw, gb = [None] * len(weights), [None] * len(weights)
    h, T, _ = acts[-1]
    gw[-1] = G_y.T @ h + torch.einsum("mik,mij->kj", G_Jt, T)
    gb[-1] = G_y.sum(0)
    dh = G_y @ weights[-1]
    dT = G_Jt @ weights[-1]
    for l in range(len(weights) - 2, -1, -1):
        h_prev, T_prev, gate = acts[l]
        dz = dh * gate
        dTz = dT * gate[:, None, :]
        gw[l] = dz.T @ h_prev + torch.einsum("mik,mij->kj", dTz, T_prev)
        gb[l] = dz.sum(0)
        dh = dz @ weights[l]
        dT = dTz @ weights[l]
    return gw, gb, dh


# ---- warped template lookup ---------------------------------------------------------------------------------------
_SX = [(-1.0 if not (q & 4) else 1.0) for q in range(8)]
_SY = [(-1.0 if not (q & 2) else 1.0) for q in range(8)]
_SZ = [(-1.0 if not (q & 1) else 1.0) for q in range(8)]


def _cell(grid, pts, xyz_min, xyz_max):
    """corner values u [M,8] (corner q: bit 0 = +Z, bit 1 = +Y, bit 2 = +X), fractions f [M,3], world scale [3]"""
    _, _, X, Y, Z = grid.shape
    dims = torch.tensor([X, Y, Z], dtype=pts.dtype)
    pos = ((((pts - xyz_min) / (xyz_max - xyz_min)) * 2 - 1) + 1) / 2 * (dims - 1)
    base = torch.floor(pos)
    f = pos - base
    flat = grid.reshape(-1)
    u = []
    for q in range(8):
        ix = (base[:, 0] + ((q >> 2) & 1)).clamp(0, X - 1)
        iy = (base[:, 1] + ((q >> 1) & 1)).clamp(0, Y - 1)
        iz = (base[:, 2] + (q & 1)).clamp(0, Z - 1)
        u.append(flat[((ix * Y + iy) * Z + iz).long()])
    return torch.stack(u, 1), f, (dims - 1) / (xyz_max - xyz_min)


def _cell_value_grad_hess(u, f, scale):
    """trilinear value s [M], world gradient g [M,3], world Hessian H [M,3,3] (zero diagonal)"""
    fx, fy, fz = f[:, 0], f[:, 1], f[:, 2]
    wx = [(fx if q & 4 else 1 - fx) for q in range(8)]
    wy = [(fy if q & 2 else 1 - fy) for q in range(8)]
    wz = [(fz if q & 1 else 1 - fz) for q in range(8)]
    s = sum(u[:, q] * wx[q] * wy[q] * wz[q] for q in range(8))
    gx = sum(u[:, q] * _SX[q] * wy[q] * wz[q] for q in range(8))
    gy = sum(u[:, q] * wx[q] * _SY[q] * wz[q] for q in range(8))
    gz = sum(u[:, q] * wx[q] * wy[q] * _SZ[q] for q in range(8))
    hxy = sum(u[:, q] * _SX[q] * _SY[q] * wz[q] for q in range(8))
    hxz = sum(u[:, q] * _SX[q] * wy[q] * _SZ[q] for q in range(8))
    hyz = sum(u[:, q] * wx[q] * _SY[q] * _SZ[q] for q in range(8))
    g = torch.stack([gx, gy, gz], 1) * scale
    zero = torch.zeros_like(hxy)
    H = torch.stack([torch.stack([zero, hxy, hxz], 1), torch.stack([hxy, zero, hyz], 1),
                     torch.stack([hxz, hyz, zero], 1)], 1) * scale[:, None] * scale[None, :]
    return s, g, H


def lookup_forward(grid, pts, y, Jt, R, xyz_min, xyz_max):
    """-> sdf_final [M], gradient [M,3], grad_deform [M,3,3] (reference layout: [m, i, k] = d x'_k / d x_i),
    sdf_deform [M], correction [M,1], cache"""
    xyz_min = torch.as_tensor(xyz_min, dtype=pts.dtype)
    xyz_max = torch.as_tensor(xyz_max, dtype=pts.dtype)
    new = pts + R * y[:, :3]
    u, f, scale = _cell(grid, new, xyz_min, xyz_max)
    s, g, H = _cell_value_grad_hess(u, f, scale)
    A = torch.eye(3, dtype=pts.dtype)[None] + R * Jt[:, :, :3]          # A[m,i,k] = d x'_k / d x_i
    gradient = torch.einsum("mik,mk->mi", A, g) + R * Jt[:, :, 3]
    sdf = s + R * y[:, 3]
    u0, f0, _ = _cell(grid, pts, xyz_min, xyz_max)
    s0, g0, _ = _cell_value_grad_hess(u0, f0, scale)
    return sdf, gradient, A, sdf - s0, R * y[:, 3:], dict(g=g, H=H, A=A, g0=g0, R=R)


def lookup_backward(cache, G_s, G_g, G_A=None, G_corr=None, G_sdf_deform=None):
    """Gradients w.r.t. sdf_final [M], gradient [M,3] (+ optionally grad_deform [M,3,3], correction [M,1],
    sdf_deform [M]) -> G_y [M,4], G_Jt [M,3,4], G_x [M,3] (through x' and the unwarped lookup of sdf_deform; the
    caller adds the value row's input gradient of warp_rows_backward)."""
    g, H, A, g0, R = cache["g"], cache["H"], cache["A"], cache["g0"], cache["R"]
    if G_sdf_deform is not None:
        G_s = G_s + G_sdf_deform
    h = torch.einsum("mik,mi->mk", A, G_g)                              # dL/d g
    G_new = G_s[:, None] * g + torch.einsum("mjk,mj->mk", H, h)         # dL/d x'
    G_y = torch.cat([R * G_new, R * G_s[:, None]], 1)
    if G_corr is not None:
        G_y = G_y + torch.cat([torch.zeros_like(G_new), R * G_corr], 1)
    G_A = G_g[:, :, None] * g[:, None, :] + (0 if G_A is None else G_A)  # dL/d A[m,i,k]
    G_Jt = torch.cat([R * G_A, R * G_g[:, :, None]], 2)
    G_x = G_new if G_sdf_deform is None else G_new - G_sdf_deform[:, None] * g0
    return G_y, G_Jt, G_x


def regulariser_grads(gradient, A, correction, sdf_deform, w_eik, w_jac, w_corr, w_def):
    """deform_implicit_loss (lib/losses.py:6-24), weighted sum -> (value, dL/d gradient, dL/d A, dL/d correction,
    dL/d sdf_deform)."""
    m = gradient.shape[0]
    gn = gradient.norm(dim=-1)
    an = A.norm(dim=-1)                                                  # [M,3]: norm over k
    val = w_eik * (gn - 1).abs().mean() + w_jac * an.mean() + w_corr * correction.abs().mean() + \
        w_def * sdf_deform.abs().mean()
    G_g = w_eik * torch.sign(gn - 1)[:, None] * gradient / gn.clamp_min(1e-30)[:, None] / m
    G_A = w_jac * A / an.clamp_min(1e-30)[:, :, None] / (3 * m)
    G_c = w_corr * torch.sign(correction) / m
    G_d = w_def * torch.sign(sdf_deform) / m
    return val, G_g, G_A, G_c, G_d
