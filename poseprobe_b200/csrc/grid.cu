// Dense-grid trilinear gather (forward) and scatter (backward) on reference-layout grids.
//
// The reference has no native source for these: DenseGrid.forward (lib/grid.py:47-58) and
// Voxurf.grid_sampler (lib/voxurf_coarse.py:522-543) call ATen's grid_sampler_3d through
// F.grid_sample(mode='bilinear', align_corners=True) and rely on autograd for the scatter and for
// d/dxyz (voxurf_coarse.py:972-984 even carries an 8-gather re-implementation to get a second
// derivative).  These kernels compute the same thing in one pass each:
//   forward : value [n,C] and, optionally, the analytic spatial gradient d value / d xyz [n,C,3]
//   backward: grad_grid += scatter(grad_out) with red.global.add, and optionally grad_xyz [n,3]
// Arithmetic follows ATen's order (normalise -> unnormalise -> floor -> 8 weights -> 8 FMAs in
// tnw,tne,tsw,tse,bnw,bne,bsw,bse order) so fp32 results agree with F.grid_sample to the last bits.
//
// Layout: grid is [1,C,X,Y,Z] (channels first, Z contiguous).  One thread owns one query point and
// loops over channels; neighbouring threads are neighbouring samples of a ray, so the 8 corner
// reads of a warp fall into a handful of 128-byte lines per channel plane.
#include "common.cuh"

namespace pp {

struct GridGeom {
  int sx, sy, sz;        // X, Y, Z extents
  float mn[3], mx[3];    // xyz_min / xyz_max
  int border;            // 0 zeros padding, 1 border padding, 2 the reference's custom sampler grid_sample_3d
                         // (voxurf_coarse.py:545-659): weights from the UNCLAMPED coordinate, corner indices clamped
};

struct Corner8 {
  int off[8];      // linear offsets inside one channel plane (valid only where ok; planes are < 2^31 elements)
  float w[8];      // trilinear weights
  bool ok[8];      // corner inside the grid
  float fx, fy, fz, gx, gy, gz;  // fractional parts along X,Y,Z and their complements
  float mx_, my_, mz_;           // d(clipped coord)/d(coord): 0 where border padding clipped the axis
};

// Continuous voxel coordinate along one axis, replicating
//   t = (x - min) / (max - min);  c = t*2 - 1        (lib/grid.py:53, separate torch kernels: no FMA)
//   i = ((c + 1) / 2) * (size - 1)                   (ATen grid_sampler_unnormalize, align_corners)
__device__ __forceinline__ float voxel_coord(float x, float mn, float mx, int size, int border,
                                             float& dclip) {
  const float t = (x - mn) / (mx - mn);
  const float c = add_rn(mul_rn(t, 2.f), -1.f);
  float i = mul_rn(add_rn(c, 1.f) / 2.f, static_cast<float>(size - 1));
  dclip = 1.f;
  if (border == 1) {
    // clip_coordinates_set_grad: the derivative vanishes on and beyond the border
    const float hi = static_cast<float>(size - 1);
    if (!(i > 0.f && i < hi)) dclip = 0.f;
    i = fminf(hi, fmaxf(i, 0.f));
  }
  return i;
}

__device__ __forceinline__ void make_corners(const GridGeom& g, float x, float y, float z, Corner8& c) {
  const float ix = voxel_coord(x, g.mn[0], g.mx[0], g.sx, g.border, c.mx_);
  const float iy = voxel_coord(y, g.mn[1], g.mx[1], g.sy, g.border, c.my_);
  const float iz = voxel_coord(z, g.mn[2], g.mx[2], g.sz, g.border, c.mz_);
  const float x0 = floorf(ix), y0 = floorf(iy), z0 = floorf(iz);
  c.fx = ix - x0; c.gx = (x0 + 1.f) - ix;
  c.fy = iy - y0; c.gy = (y0 + 1.f) - iy;
  c.fz = iz - z0; c.gz = (z0 + 1.f) - iz;
  // ATen's (x,y,z) = (W,H,D) = our (Z,Y,X).  Corner order tnw,tne,tsw,tse,bnw,bne,bsw,bse:
  //   e/w toggles Z, s/n toggles Y, b/t toggles X.  Weight product order is (W term)*(H term)*(D term).
  const float lim = 1e9f;   // (keeps the float -> int conversion defined for far-away points)
  const int xi = static_cast<int>(fminf(fmaxf(x0, -lim), lim)), yi = static_cast<int>(fminf(fmaxf(y0, -lim), lim)),
            zi = static_cast<int>(fminf(fmaxf(z0, -lim), lim));
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int dz = q & 1, dy = (q >> 1) & 1, dx = (q >> 2) & 1;
    int cx = xi + dx, cy = yi + dy, cz = zi + dz;
    if (g.border == 2) {   // torch.clamp of the corner indices (:575-598); the weights stay those of the free cell
      cx = min(max(cx, 0), g.sx - 1);
      cy = min(max(cy, 0), g.sy - 1);
      cz = min(max(cz, 0), g.sz - 1);
    }
    c.ok[q] = (cx >= 0) & (cx < g.sx) & (cy >= 0) & (cy < g.sy) & (cz >= 0) & (cz < g.sz);
    c.off[q] = (cx * g.sy + cy) * g.sz + cz;
    c.w[q] = mul_rn(mul_rn(dz ? c.fz : c.gz, dy ? c.fy : c.gy), dx ? c.fx : c.gx);
  }
}

// (<= 64 registers: the kernel is gather-latency bound and wants all 2048 threads of an SM resident)
template <bool kWithDxyz>
__global__ void __launch_bounds__(256, 4)
trilinear_fwd_kernel(const float* __restrict__ grid, int n_ch, GridGeom g,
                     const float* __restrict__ xyz, int64_t n_pts, float* __restrict__ out,
                     float* __restrict__ dout_dxyz) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n_pts) return;
  Corner8 c;
  make_corners(g, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], c);
  const int64_t plane = (int64_t)g.sx * g.sy * g.sz;
  // d(voxel coord)/d(world coord) per axis
  const float kx = c.mx_ * (g.sx - 1) / (g.mx[0] - g.mn[0]);
  const float ky = c.my_ * (g.sy - 1) / (g.mx[1] - g.mn[1]);
  const float kz = c.mz_ * (g.sz - 1) / (g.mx[2] - g.mn[2]);
  for (int ch = 0; ch < n_ch; ++ch) {
    const float* gp = grid + ch * plane;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = c.ok[q] ? __ldg(gp + c.off[q]) : 0.f;
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc = fma_rn(v[q], c.w[q], acc);
    out[p * n_ch + ch] = acc;
    if (kWithDxyz) {
      // q bit0 -> Z, bit1 -> Y, bit2 -> X
      const float dX = ((v[4] - v[0]) * c.gz + (v[5] - v[1]) * c.fz) * c.gy +
                       ((v[6] - v[2]) * c.gz + (v[7] - v[3]) * c.fz) * c.fy;
      const float dY = ((v[2] - v[0]) * c.gz + (v[3] - v[1]) * c.fz) * c.gx +
                       ((v[6] - v[4]) * c.gz + (v[7] - v[5]) * c.fz) * c.fx;
      const float dZ = ((v[1] - v[0]) * c.gy + (v[3] - v[2]) * c.fy) * c.gx +
                       ((v[5] - v[4]) * c.gy + (v[7] - v[6]) * c.fy) * c.fx;
      float* d = dout_dxyz + (p * n_ch + ch) * 3;
      d[0] = dX * kx;
      d[1] = dY * ky;
      d[2] = dZ * kz;
    }
  }
}

__device__ __forceinline__ void red_add(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__global__ void __launch_bounds__(256, 4)
trilinear_bwd_kernel(const float* __restrict__ grid, int n_ch, GridGeom g,
                     const float* __restrict__ xyz, int64_t n_pts,
                     const float* __restrict__ grad_out, float* __restrict__ grad_grid,
                     float* __restrict__ grad_xyz) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n_pts) return;
  Corner8 c;
  make_corners(g, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], c);
  const int64_t plane = (int64_t)g.sx * g.sy * g.sz;
  float gX = 0.f, gY = 0.f, gZ = 0.f;
  for (int ch = 0; ch < n_ch; ++ch) {
    const float go = grad_out[p * n_ch + ch];
    if (grad_grid) {
      float* gp = grad_grid + ch * plane;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (c.ok[q]) red_add(gp + c.off[q], mul_rn(c.w[q], go));
    }
    if (grad_xyz) {
      const float* vp = grid + ch * plane;
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = c.ok[q] ? __ldg(vp + c.off[q]) : 0.f;
      gX += go * (((v[4] - v[0]) * c.gz + (v[5] - v[1]) * c.fz) * c.gy +
                  ((v[6] - v[2]) * c.gz + (v[7] - v[3]) * c.fz) * c.fy);
      gY += go * (((v[2] - v[0]) * c.gz + (v[3] - v[1]) * c.fz) * c.gx +
                  ((v[6] - v[4]) * c.gz + (v[7] - v[5]) * c.fz) * c.fx);
      gZ += go * (((v[1] - v[0]) * c.gy + (v[3] - v[2]) * c.fy) * c.gx +
                  ((v[5] - v[4]) * c.gy + (v[7] - v[6]) * c.fy) * c.fx);
    }
  }
  if (grad_xyz) {
    grad_xyz[3 * p] = gX * (c.mx_ * (g.sx - 1) / (g.mx[0] - g.mn[0]));
    grad_xyz[3 * p + 1] = gY * (c.my_ * (g.sy - 1) / (g.mx[1] - g.mn[1]));
    grad_xyz[3 * p + 2] = gZ * (c.mz_ * (g.sz - 1) / (g.mx[2] - g.mn[2]));
  }
}

// Backward of the backward (the reference reaches it with torch.autograd.grad(..., create_graph=True) through its
// 8-gather sampler, voxurf_coarse.py:966-985): the first backward's grad_xyz[a] = k_a sum_c go_c D_a(v_c) is a function
// of the grid values, of the point (through the weights) and of go.  Given h = d loss / d grad_xyz [n,3]:
//   d loss / d go_c      = sum_a h_a k_a D_a(v_c)                          (directional derivative of the value)
//   d loss / d grid[q,c] = go_c sum_a h_a k_a d w_q / d i_a                (scatter with derivative weights)
//   d loss / d xyz_b     = k_b sum_c go_c sum_a h_a k_a H_ab(v_c)          (H: off-diagonal Hessian of the cell)
// with i = voxel coordinate, k_a = d i_a / d xyz_a, D = d value / d i.  All three outputs are optional.
__global__ void __launch_bounds__(256)
trilinear_bwd2_kernel(const float* __restrict__ grid, int n_ch, GridGeom g, const float* __restrict__ xyz,
                      int64_t n_pts, const float* __restrict__ grad_out, const float* __restrict__ h_in,
                      float* __restrict__ gg_out, float* __restrict__ grad_grid, float* __restrict__ grad_xyz) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n_pts) return;
  Corner8 c;
  make_corners(g, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], c);
  const int64_t plane = (int64_t)g.sx * g.sy * g.sz;
  const float kx = c.mx_ * (g.sx - 1) / (g.mx[0] - g.mn[0]);
  const float ky = c.my_ * (g.sy - 1) / (g.mx[1] - g.mn[1]);
  const float kz = c.mz_ * (g.sz - 1) / (g.mx[2] - g.mn[2]);
  const float hx = h_in[3 * p] * kx, hy = h_in[3 * p + 1] * ky, hz = h_in[3 * p + 2] * kz;
  // d w_q / d i_a contracted with h: q bit0 -> Z, bit1 -> Y, bit2 -> X
  float dw[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float wz = (q & 1) ? c.fz : c.gz, wy = (q & 2) ? c.fy : c.gy, wx = (q & 4) ? c.fx : c.gx;
    const float sz = (q & 1) ? 1.f : -1.f, sy = (q & 2) ? 1.f : -1.f, sx = (q & 4) ? 1.f : -1.f;
    dw[q] = hx * sx * wy * wz + hy * wx * sy * wz + hz * wx * wy * sz;
  }
  float gX = 0.f, gY = 0.f, gZ = 0.f;
  for (int ch = 0; ch < n_ch; ++ch) {
    const float go = grad_out[p * n_ch + ch];
    const float* vp = grid + ch * plane;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = c.ok[q] ? __ldg(vp + c.off[q]) : 0.f;
    if (gg_out) {
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) acc = fma_rn(v[q], dw[q], acc);
      gg_out[p * n_ch + ch] = acc;
    }
    if (grad_grid) {
      float* gp = grad_grid + ch * plane;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (c.ok[q]) red_add(gp + c.off[q], dw[q] * go);
    }
    if (grad_xyz) {
      const float hxy = ((v[6] - v[4]) - (v[2] - v[0])) * c.gz + ((v[7] - v[5]) - (v[3] - v[1])) * c.fz;
      const float hxz = ((v[5] - v[4]) - (v[1] - v[0])) * c.gy + ((v[7] - v[6]) - (v[3] - v[2])) * c.fy;
      const float hyz = ((v[3] - v[2]) - (v[1] - v[0])) * c.gx + ((v[7] - v[6]) - (v[5] - v[4])) * c.fx;
      gX += go * (hxy * hy + hxz * hz);
      gY += go * (hxy * hx + hyz * hz);
      gZ += go * (hxz * hx + hyz * hy);
    }
  }
  if (grad_xyz) {
    grad_xyz[3 * p] = gX * kx;
    grad_xyz[3 * p + 1] = gY * ky;
    grad_xyz[3 * p + 2] = gZ * kz;
  }
}

static int make_geom(GridGeom& g, int32_t sx, int32_t sy, int32_t sz, const float* mn,
                     const float* mx, int padding) {
  PP_REQUIRE(sx > 0 && sy > 0 && sz > 0, PP_ERR_INVALID_ARG, "trilinear: empty grid");
  PP_REQUIRE(mn && mx, PP_ERR_INVALID_ARG, "trilinear: xyz_min/xyz_max (host) are null");
  PP_REQUIRE(padding >= 0 && padding <= 2, PP_ERR_INVALID_ARG, "trilinear: padding must be 0|1|2");
  PP_REQUIRE((int64_t)sx * sy * sz <= INT32_MAX, PP_ERR_TOO_LARGE, "trilinear: more than 2^31 voxels per channel");
  g.sx = sx; g.sy = sy; g.sz = sz; g.border = padding;
  for (int c = 0; c < 3; ++c) { g.mn[c] = mn[c]; g.mx[c] = mx[c]; }
  return PP_OK;
}

}  // namespace pp

using namespace pp;

extern "C" int pp_trilinear_forward(const float* grid, int32_t n_ch, int32_t sz_x, int32_t sz_y,
                                    int32_t sz_z, const float* xyz, int64_t n_pts,
                                    const float* xyz_min_host, const float* xyz_max_host,
                                    int padding, float* out, float* dout_dxyz, pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0 && n_ch >= 0, PP_ERR_INVALID_ARG, "trilinear_forward: negative extent");
  if (n_pts == 0 || n_ch == 0) return PP_OK;
  PP_REQUIRE(grid && xyz && out, PP_ERR_INVALID_ARG, "trilinear_forward: null pointer");
  GridGeom g;
  const int rc = make_geom(g, sz_x, sz_y, sz_z, xyz_min_host, xyz_max_host, padding);
  if (rc != PP_OK) return rc;
  const int blocks = ceil_div(n_pts, 256);
  if (dout_dxyz)
    trilinear_fwd_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(grid, n_ch, g, xyz, n_pts, out, dout_dxyz);
  else
    trilinear_fwd_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(grid, n_ch, g, xyz, n_pts, out, nullptr);
  PP_CHECK_LAUNCH("trilinear_forward");
  return PP_OK;
}

extern "C" int pp_trilinear_backward(const float* grid, int32_t n_ch, int32_t sz_x, int32_t sz_y,
                                     int32_t sz_z, const float* xyz, int64_t n_pts,
                                     const float* xyz_min_host, const float* xyz_max_host,
                                     int padding, const float* grad_out, float* grad_grid,
                                     float* grad_xyz, pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0 && n_ch >= 0, PP_ERR_INVALID_ARG, "trilinear_backward: negative extent");
  if (n_pts == 0 || n_ch == 0 || (!grad_grid && !grad_xyz)) return PP_OK;
  PP_REQUIRE(xyz && grad_out, PP_ERR_INVALID_ARG, "trilinear_backward: null pointer");
  PP_REQUIRE(!grad_xyz || grid, PP_ERR_INVALID_ARG, "trilinear_backward: grad_xyz needs grid");
  GridGeom g;
  const int rc = make_geom(g, sz_x, sz_y, sz_z, xyz_min_host, xyz_max_host, padding);
  if (rc != PP_OK) return rc;
  trilinear_bwd_kernel<<<ceil_div(n_pts, 256), 256, 0, as_stream(stream)>>>(
      grid, n_ch, g, xyz, n_pts, grad_out, grad_grid, grad_xyz);
  PP_CHECK_LAUNCH("trilinear_backward");
  return PP_OK;
}

extern "C" int pp_trilinear_backward2(const float* grid, int32_t n_ch, int32_t sz_x, int32_t sz_y, int32_t sz_z,
                                      const float* xyz, int64_t n_pts, const float* xyz_min_host,
                                      const float* xyz_max_host, int padding, const float* grad_out,
                                      const float* grad_grad_xyz, float* grad_grad_out, float* grad_grid,
                                      float* grad_xyz, pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0 && n_ch >= 0, PP_ERR_INVALID_ARG, "trilinear_backward2: negative extent");
  if (n_pts == 0 || n_ch == 0 || (!grad_grad_out && !grad_grid && !grad_xyz)) return PP_OK;
  PP_REQUIRE(grid && xyz && grad_out && grad_grad_xyz, PP_ERR_INVALID_ARG, "trilinear_backward2: null pointer");
  GridGeom g;
  const int rc = make_geom(g, sz_x, sz_y, sz_z, xyz_min_host, xyz_max_host, padding);
  if (rc != PP_OK) return rc;
  trilinear_bwd2_kernel<<<ceil_div(n_pts, 256), 256, 0, as_stream(stream)>>>(
      grid, n_ch, g, xyz, n_pts, grad_out, grad_grad_xyz, grad_grad_out, grad_grid, grad_xyz);
  PP_CHECK_LAUNCH("trilinear_backward2");
  return PP_OK;
}
