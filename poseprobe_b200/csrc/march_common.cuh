// Pieces shared by the fused march kernels (dvgo_fused.cu, voxurf_fused.cu): voxel-record cells, per-ray
// setup with the stand-alone sampler's arithmetic, the lane product scan and the device-side ray queue.
#pragma once
#include "../../include/poseprobe_b200_fused.h"
#include "common.cuh"
#include "ray_math.cuh"

namespace pp {

// ---- voxel record helpers -----------------------------------------------------------------------
struct Cell {
  int base;           // record index of the (x0,y0,z0) corner
  int ex, ey, ez;     // record strides towards the +X/+Y/+Z corner (0 on the max face: weight is 0)
  float fx, fy, fz;   // fractional position inside the cell
  // complements: ATen computes (floor+1) - i, whose exact value is 1 - f, so 1.f - f rounds identically
  __device__ __forceinline__ float gx() const { return 1.f - fx; }
  __device__ __forceinline__ float gy() const { return 1.f - fy; }
  __device__ __forceinline__ float gz() const { return 1.f - fz; }
};

// ATen-order continuous voxel index (see grid.cu voxel_coord) for a point known to be inside the box.
__device__ __forceinline__ float vcoord(float x, float mn, float mx, int size) {
  const float t = (x - mn) / (mx - mn);
  const float c = add_rn(mul_rn(t, 2.f), -1.f);
  return mul_rn(add_rn(c, 1.f) / 2.f, static_cast<float>(size - 1));
}

__device__ __forceinline__ Cell make_cell(const pp_dvgo_scene& sc, float px, float py, float pz) {
  Cell c;
  float ix = vcoord(px, sc.xyz_min[0], sc.xyz_max[0], sc.sx);
  float iy = vcoord(py, sc.xyz_min[1], sc.xyz_max[1], sc.sy);
  float iz = vcoord(pz, sc.xyz_min[2], sc.xyz_max[2], sc.sz);
  // rounding can push an in-box point a hair outside [0, size-1]; zero padding would give it weight
  // ~1e-7 less, clamping keeps the record index legal
  ix = fminf(fmaxf(ix, 0.f), static_cast<float>(sc.sx - 1));
  iy = fminf(fmaxf(iy, 0.f), static_cast<float>(sc.sy - 1));
  iz = fminf(fmaxf(iz, 0.f), static_cast<float>(sc.sz - 1));
  const float x0 = floorf(ix), y0 = floorf(iy), z0 = floorf(iz);
  c.fx = ix - x0;
  c.fy = iy - y0;
  c.fz = iz - z0;
  const int xi = static_cast<int>(x0), yi = static_cast<int>(y0), zi = static_cast<int>(z0);
  c.base = (xi * sc.sy + yi) * sc.sz + zi;
  c.ez = (zi + 1 < sc.sz) ? 1 : 0;
  c.ey = (yi + 1 < sc.sy) ? sc.sz : 0;
  c.ex = (xi + 1 < sc.sx) ? sc.sy * sc.sz : 0;
  return c;
}

__device__ __forceinline__ int corner_off(const Cell& c, int q) {
  return c.base + ((q & 1) ? c.ez : 0) + ((q & 2) ? c.ey : 0) + ((q & 4) ? c.ex : 0);
}
__device__ __forceinline__ float corner_w(const Cell& c, int q) {
  // (Z term * Y term) * X term : ATen's tnw..bse product order
  return mul_rn(mul_rn((q & 1) ? c.fz : c.gz(), (q & 2) ? c.fy : c.gy()), (q & 4) ? c.fx : c.gx());
}

__device__ __forceinline__ void gather8(const float4* __restrict__ vox, const Cell& c, float4 (&v)[8]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] = __ldg(vox + corner_off(c, q));
}

// pull the 8 corner records of a future sample towards L1 (no registers are tied up)
__device__ __forceinline__ void prefetch8(const float4* __restrict__ vox, const Cell& c) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(vox + corner_off(c, q)));
}

__device__ __forceinline__ float4 interp8(const Cell& c, const float4 (&v)[8]) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float w = corner_w(c, q);
    a.x = fma_rn(v[q].x, w, a.x);
    a.y = fma_rn(v[q].y, w, a.y);
    a.z = fma_rn(v[q].z, w, a.z);
    a.w = fma_rn(v[q].w, w, a.w);
  }
  return a;
}

__device__ __forceinline__ void red_add_v4(float4* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

// sigmoid with the SFU exponential and reciprocal (<= ~4e-7 relative to torch.sigmoid)
__device__ __forceinline__ float sigmoid_ref(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// ---- per-ray state ---------------------------------------------------------------------------------
struct Ray {
  float start[3], dir[3];
  float o[3], d[3];
  float t_min, rnorm;
  int n_steps;
  int enter_axis;   // axis whose slab decided t_min, -1 when t_min was clamped to near/far
};

__device__ __forceinline__ Ray setup_ray(const pp_dvgo_scene& sc, const float* __restrict__ rays_o,
                                         const float* __restrict__ rays_d, int64_t r) {
  Ray ray;
  load3(rays_o, r, ray.o);
  load3(rays_d, r, ray.d);
  float t_max;
  ray_bbox_range(ray.o, ray.d, sc.xyz_min, sc.xyz_max, sc.near, sc.far, ray.t_min, t_max);
  ray.rnorm = ray_norm(ray.d);
  const int64_t n = ray_step_count<float>(ray.t_min, t_max, ray.rnorm, sc.stepdist);
  ray.n_steps = static_cast<int>(min64(n, 1 << 30));
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    ray.start[c] = fma_rn(ray.d[c], ray.t_min, ray.o[c]);
    ray.dir[c] = ray.d[c] / ray.rnorm;
  }
  ray.enter_axis = -1;
  return ray;
}

__device__ __forceinline__ bool sample_point(const pp_dvgo_scene& sc, const Ray& ray, int i, float (&p)[3]) {
  const float dist = sc.stepdist * static_cast<float>(i);
#pragma unroll
  for (int c = 0; c < 3; ++c) p[c] = fma_rn(ray.dir[c], dist, ray.start[c]);
  const bool out = (sc.xyz_min[0] > p[0]) | (sc.xyz_min[1] > p[1]) | (sc.xyz_min[2] > p[2]) |
                   (sc.xyz_max[0] < p[0]) | (sc.xyz_max[1] < p[1]) | (sc.xyz_max[2] < p[2]);
  return !out;
}

// ---- the march ------------------------------------------------------------------------------------
struct LaneSample {
  bool live;
  Cell c;
};

__device__ __forceinline__ LaneSample lane_sample(const pp_dvgo_scene& sc, const Ray& ray, int i, int limit) {
  LaneSample s;
  float p[3];
  s.live = (i >= 0) && (i < limit) && sample_point(sc, ray, i, p);
  if (s.live) s.c = make_cell(sc, p[0], p[1], p[2]);
  return s;
}

// inclusive product scan over the lanes; returns inclusive P and exclusive Pex
__device__ __forceinline__ void lane_product_scan(float q, int lane, float& P, float& Pex) {
  P = q;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float up = __shfl_up_sync(kFull, P, o);
    if (lane >= o) P *= up;
  }
  Pex = __shfl_up_sync(kFull, P, 1);
  if (lane == 0) Pex = 1.f;
}

// rays are handed out through a device-side queue so that long and short rays balance across warps
constexpr int kQueueSlots = 64;

}  // namespace pp
