// Per-ray geometry shared by the sampling operators and the fused render kernels.
// Every expression keeps the reference's operand types, evaluation order and FMA contraction
// (see sampling.cu) because integer outputs depend on the roundings.
#pragma once
#include "common.cuh"

namespace pp {

// ---- per-ray geometry --------------------------------------------------------------------------
// Slab test, reference render_utils_kernel.cu:20-34.  Arithmetic in scalar_t, results kept in float.
template <typename scalar_t>
__device__ __forceinline__ void ray_bbox_range(const scalar_t o[3], const scalar_t d[3],
                                               const scalar_t mn[3], const scalar_t mx[3],
                                               float near, float far, float& t_min, float& t_max) {
  float a[3], b[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // (d==0 ? 1e-6 : d) is a double expression narrowed to float.
    const float v = static_cast<float>((d[c] == scalar_t(0)) ? 1e-6 : static_cast<double>(d[c]));
    a[c] = static_cast<float>((mx[c] - o[c]) / v);
    b[c] = static_cast<float>((mn[c] - o[c]) / v);
  }
  const float enter = fmaxf(fmaxf(fminf(a[0], b[0]), fminf(a[1], b[1])), fminf(a[2], b[2]));
  const float leave = fminf(fminf(fmaxf(a[0], b[0]), fmaxf(a[1], b[1])), fmaxf(a[2], b[2]));
  t_min = fmaxf(fminf(enter, far), near);
  t_max = fmaxf(fminf(leave, far), near);
}

// |d|, reference :48-51 (and :68-71).  nvcc -fmad=true turns dx*dx + dy*dy + dz*dz into
// fma(dz,dz, fma(dx,dx, dy*dy)) -- i.e. the *y* product is the one that is rounded on its own
// (checked in the SASS of the reference build: FMUL y*y, FFMA x*x+., FFMA z*z+.).
template <typename scalar_t>
__device__ __forceinline__ float ray_norm(const scalar_t d[3]) {
  const scalar_t s = fma_rn(d[2], d[2], fma_rn(d[0], d[0], mul_rn(d[1], d[1])));
  return static_cast<float>(sqrt(s));
}

// Step count, reference :53: max(ceil((t_max-t_min)*rnorm/stepdist), 1.) -> int64.
template <typename scalar_t>
__device__ __forceinline__ int64_t ray_step_count(float t_min, float t_max, float rnorm,
                                                  float stepdist) {
  const scalar_t span = static_cast<scalar_t>(t_max) - static_cast<scalar_t>(t_min);
  const scalar_t steps = ceil(span * rnorm / stepdist);
  // fmax semantics: a NaN span still yields 1.
  const double n = fmax(static_cast<double>(steps), 1.0);
  return static_cast<int64_t>(n);
}

template <typename scalar_t>
__device__ __forceinline__ void load3(const scalar_t* p, int64_t i, scalar_t out[3]) {
  out[0] = p[3 * i];
  out[1] = p[3 * i + 1];
  out[2] = p[3 * i + 2];
}

}  // namespace pp
