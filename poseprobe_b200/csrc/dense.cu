// Dense whole-grid operators: total-variation gradient, Adam variants, cumulative-distance mask.
//
// Replaces total_variation_kernel.cu:13-133, adam_upd_kernel.cu:8-132, ub360_utils_kernel.cu:12-47.
// All three are HBM-bound streaming passes over [C,X,Y,Z] tensors; the reference launches one
// thread per element with 64-bit div/mod index recovery and scalar 4-byte accesses.  Here:
//   * TV uses a 3-D launch (k fastest, j, plane*X+i) so no division is needed and every one of the
//     7 stencil reads is a coalesced row access that hits L1/L2 after its first touch;
//   * Adam is a grid-stride float4 pass (128-bit loads/stores, 5 streams) sized to a multiple of
//     the SM count;
//   * cumdist_thres keeps the reference's strictly sequential running sum (its float roundings
//     decide a bool output) on one lane per ray but moves the data through a shared-memory tile so
//     global accesses are coalesced.
#include <math.h>

#include "common.cuh"

namespace pp {

// ---- total variation -------------------------------------------------------------------------------
__device__ __forceinline__ float clamp1(float v) { return fminf(fmaxf(v, -1.f), 1.f); }
__device__ __forceinline__ double clamp1(double v) { return fmin(fmax(v, -1.), 1.); }

// One thread per element.  blockDim = (32, 8): x runs along k (contiguous), y along j.
// grid = (ceil(Z/32), ceil(Y/8), planes*X).
// kMasked == false : reference kernel :13-35, weights (k, j, i) = (wz, wy, wz)   [sic: wx unused]
// kMasked == true  : reference kernel :38-66, weights (k, j, i) = (wx, wy, wz), each term times
//                    mask[idx]*mask[nbr]
template <typename scalar_t, bool kDense, bool kMasked>
__global__ void __launch_bounds__(256)
tv_add_grad_kernel(const scalar_t* __restrict__ param, scalar_t* __restrict__ grad,
                   const scalar_t* __restrict__ mask, float wk, float wj, float wi, int sz_i,
                   int sz_j, int sz_k, int chunk, int chunks_per_plane) {
  const int k = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (k >= sz_k || j >= sz_j) return;
  // blockIdx.z = plane * chunks_per_plane + chunk index: the thread walks `chunk` consecutive slices of X with a
  // sliding window (previous, current, next) of its own column, so every parameter is loaded once for the three
  // X terms it takes part in, and one block does `chunk` slices of work per launch slot
  const int plane = blockIdx.z / chunks_per_plane;
  const int i_begin = (blockIdx.z - plane * chunks_per_plane) * chunk;
  const int i_end = min(i_begin + chunk, sz_i);
  const int64_t sk = 1, sj = sz_k, si = (int64_t)sz_k * sz_j;
  int64_t idx = (((int64_t)plane * sz_i + i_begin) * sz_j + j) * sz_k + k;
  scalar_t p_prev = (i_begin > 0) ? param[idx - si] : scalar_t(0);
  scalar_t p = param[idx];
  for (int i = i_begin; i < i_end; ++i, idx += si) {
    const scalar_t p_next = (i + 1 < sz_i) ? param[idx + si] : scalar_t(0);
    const scalar_t g = grad[idx];
    if (kDense || g != scalar_t(0)) {
      const scalar_t m0 = kMasked ? mask[idx] : scalar_t(1);
      float acc = 0.f;  // the reference accumulates in a float local
      auto term = [&](bool edge, scalar_t nbr, int64_t off, float w) {
        if (edge) return;
        // w * clamp(p - p_nbr, -1, 1) [* mask * mask_nbr], added to the float accumulator
        // (reference SASS: rounded product, then plain adds -- no FMA contraction here)
        scalar_t t = mul_rn(static_cast<scalar_t>(w), clamp1(p - nbr));
        if (kMasked) t = mul_rn(mul_rn(t, m0), mask[idx + off]);
        acc = static_cast<float>(add_rn(static_cast<scalar_t>(acc), t));
      };
      if (k != 0) term(false, param[idx - sk], -sk, wk);
      if (k != sz_k - 1) term(false, param[idx + sk], sk, wk);
      if (j != 0) term(false, param[idx - sj], -sj, wj);
      if (j != sz_j - 1) term(false, param[idx + sj], sj, wj);
      term(i == 0, p_prev, -si, wi);
      term(i == sz_i - 1, p_next, si, wi);
      grad[idx] = add_rn(g, static_cast<scalar_t>(acc));
    }
    p_prev = p;
    p = p_next;
  }
}

// ---- Python-side total variation loss actually used by the reference (voxurf_coarse.py:1298-1313, called from
//      k0_total_variation :443-456 and density_total_variation :407-411): (sum|d2| + sum|d3| + sum|d4|) / 3 / count.
// The reference builds >= 9 grid-sized temporaries per call (forward and autograd).  Here: one pass each way, the
// same sliding window along X as the add-grad kernel.  kBackward: grad += scale * d(sum)/dv.
template <bool kBackward>
__global__ void __launch_bounds__(256)
tv_loss_kernel(const float* __restrict__ v, const uint8_t* __restrict__ mask, int sz_i, int sz_j, int sz_k,
               int chunk, int chunks_per_plane, float scale, float* __restrict__ grad, double* __restrict__ sum_out) {
  const int k = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y * 8 + threadIdx.y;
  float acc = 0.f;
  if (k < sz_k && j < sz_j) {
    const int plane = blockIdx.z / chunks_per_plane;
    const int i_begin = (blockIdx.z - plane * chunks_per_plane) * chunk;
    const int i_end = min(i_begin + chunk, sz_i);
    const int64_t sj = sz_k, si = (int64_t)sz_k * sz_j;
    int64_t cell = ((int64_t)i_begin * sz_j + j) * sz_k + k;           // index inside one plane (mask index)
    int64_t idx = (int64_t)plane * sz_i * si + cell;
    auto ok = [&](int64_t c) { return mask == nullptr || mask[c] != 0; };
    float p_prev = (i_begin > 0) ? v[idx - si] : 0.f;
    float p = v[idx];
    for (int i = i_begin; i < i_end; ++i, idx += si, cell += si) {
      const float p_next = (i + 1 < sz_i) ? v[idx + si] : 0.f;
      const bool m0 = ok(cell);
      if (!kBackward) {
        // forward differences owned by this voxel: towards +X, +Y, +Z
        if (m0) {
          if (i + 1 < sz_i && ok(cell + si)) acc += fabsf(p_next - p);
          if (j + 1 < sz_j && ok(cell + sj)) acc += fabsf(v[idx + sj] - p);
          if (k + 1 < sz_k && ok(cell + 1)) acc += fabsf(v[idx + 1] - p);
        }
      } else if (m0) {
        // d/dp of |p - p_prev| + |p_next - p| per axis (torch.abs backward: sign, 0 at 0)
        auto sgn = [](float x) { return (x > 0.f) - (x < 0.f); };
        int g = 0;
        if (i > 0 && ok(cell - si)) g += sgn(p - p_prev);
        if (i + 1 < sz_i && ok(cell + si)) g -= sgn(p_next - p);
        if (j > 0 && ok(cell - sj)) g += sgn(p - v[idx - sj]);
        if (j + 1 < sz_j && ok(cell + sj)) g -= sgn(v[idx + sj] - p);
        if (k > 0 && ok(cell - 1)) g += sgn(p - v[idx - 1]);
        if (k + 1 < sz_k && ok(cell + 1)) g -= sgn(v[idx + 1] - p);
        if (g != 0) grad[idx] += scale * static_cast<float>(g);
      }
      p_prev = p;
      p = p_next;
    }
  }
  if (!kBackward) {
    __shared__ float s_part[8];
    acc = warp_sum(acc);
    if (threadIdx.x == 0) s_part[threadIdx.y] = acc;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      double tot = 0.;
      for (int w = 0; w < 8; ++w) tot += static_cast<double>(s_part[w]);
      if (tot != 0.) atomicAdd(sum_out, tot);
    }
  }
}

// ---- Adam ------------------------------------------------------------------------------------------
// mode 0 plain, 1 masked (skip grad == 0), 2 per-element lr.  Reference :18-22 / :35-39 / :53-57.
template <typename scalar_t, int kMode>
__device__ __forceinline__ void adam_one(scalar_t& p, const scalar_t g, scalar_t& m, scalar_t& v,
                                         const scalar_t plr, float step_size, float beta1,
                                         float beta2, float eps) {
  if (kMode == 1 && g == scalar_t(0)) return;
  // contraction pattern of the reference build (SASS): the beta*state product is the fused one
  const scalar_t b1 = beta1, b2 = beta2, c1 = 1 - beta1, c2 = 1 - beta2;
  m = fma_rn(b1, m, mul_rn(c1, g));
  v = fma_rn(b2, v, mul_rn(mul_rn(c2, g), g));
  const scalar_t num = (kMode == 2) ? mul_rn(mul_rn(static_cast<scalar_t>(step_size), plr), m)
                                    : mul_rn(static_cast<scalar_t>(step_size), m);
  p = p - num / (sqrt(v) + eps);
}

template <typename scalar_t, int kMode>
__global__ void __launch_bounds__(256)
adam_scalar_kernel(scalar_t* __restrict__ param, const scalar_t* __restrict__ grad,
                   scalar_t* __restrict__ exp_avg, scalar_t* __restrict__ exp_avg_sq,
                   const scalar_t* __restrict__ perlr, int64_t n, float step_size, float beta1,
                   float beta2, float eps) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const scalar_t g = grad[i];
    if (kMode == 1 && g == scalar_t(0)) continue;
    scalar_t p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
    adam_one<scalar_t, kMode>(p, g, m, v, kMode == 2 ? perlr[i] : scalar_t(1), step_size, beta1,
                              beta2, eps);
    param[i] = p;
    exp_avg[i] = m;
    exp_avg_sq[i] = v;
  }
}

// fp32, 16-byte aligned, n % 4 == 0
template <int kMode>
__global__ void __launch_bounds__(256)
adam_vec4_kernel(float4* __restrict__ param, const float4* __restrict__ grad,
                 float4* __restrict__ exp_avg, float4* __restrict__ exp_avg_sq,
                 const float4* __restrict__ perlr, int64_t n4, float step_size, float beta1,
                 float beta2, float eps) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 g = grad[i];
    if (kMode == 1 && g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
    float4 p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
    float4 l = make_float4(1.f, 1.f, 1.f, 1.f);
    if (kMode == 2) l = perlr[i];
    adam_one<float, kMode>(p.x, g.x, m.x, v.x, l.x, step_size, beta1, beta2, eps);
    adam_one<float, kMode>(p.y, g.y, m.y, v.y, l.y, step_size, beta1, beta2, eps);
    adam_one<float, kMode>(p.z, g.z, m.z, v.z, l.z, step_size, beta1, beta2, eps);
    adam_one<float, kMode>(p.w, g.w, m.w, v.w, l.w, step_size, beta1, beta2, eps);
    param[i] = p;
    exp_avg[i] = m;
    exp_avg_sq[i] = v;
  }
}

// ---- cumdist_thres ---------------------------------------------------------------------------------
// reference :20-31.  One lane per ray for the recurrence, 32x32 shared tiles for the traffic.
constexpr int kCdWarps = 4;
template <typename scalar_t>
__global__ void __launch_bounds__(kCdWarps* kWarp)
cumdist_kernel(const scalar_t* __restrict__ dist, float thres, int64_t n_rays, int64_t n_pts,
               uint8_t* __restrict__ mask) {
  __shared__ scalar_t s_d[kCdWarps][32][33];
  __shared__ uint8_t s_m[kCdWarps][32][36];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t ray0 = (blockIdx.x * (int64_t)kCdWarps + w) * 32;
  if (ray0 >= n_rays) return;
  const int rows = static_cast<int>(min64(32, n_rays - ray0));
  float cum = 0.f;
  for (int64_t p0 = 0; p0 < n_pts; p0 += 32) {
    const int cols = static_cast<int>(min64(32, n_pts - p0));
    for (int r = 0; r < rows; ++r)
      if (lane < cols) s_d[w][r][lane] = dist[(ray0 + r) * n_pts + p0 + lane];
    __syncwarp();
    if (lane < rows) {
      for (int c = 0; c < cols; ++c) {
        cum += s_d[w][lane][c];              // scalar_t widened add, narrowed to the float local
        const bool over = cum > thres;
        cum *= float(!over);
        s_m[w][lane][c] = over;
      }
    }
    __syncwarp();
    for (int r = 0; r < rows; ++r)
      if (lane < cols) mask[(ray0 + r) * n_pts + p0 + lane] = s_m[w][r][lane];
    __syncwarp();
  }
}

}  // namespace pp

using namespace pp;

#define PP_DISPATCH(dtype, ...)                                 \
  do {                                                          \
    if ((dtype) == PP_F32) {                                    \
      using scalar_t = float;                                   \
      __VA_ARGS__                                               \
    } else if ((dtype) == PP_F64) {                             \
      using scalar_t = double;                                  \
      __VA_ARGS__                                               \
    } else {                                                    \
      set_error("unknown dtype %d", (int)(dtype));              \
      return PP_ERR_INVALID_ARG;                                \
    }                                                           \
  } while (0)

extern "C" int pp_total_variation_add_grad(pp_dtype dtype, const void* param, void* grad,
                                           const void* mask, float wx, float wy, float wz,
                                           int dense_mode, int64_t n_planes, int64_t sz_i,
                                           int64_t sz_j, int64_t sz_k, pp_stream_t stream) {
  PP_REQUIRE(n_planes >= 0 && sz_i >= 0 && sz_j >= 0 && sz_k >= 0, PP_ERR_INVALID_ARG,
             "total_variation_add_grad: negative extent");
  if (n_planes * sz_i * sz_j * sz_k == 0) return PP_OK;
  PP_REQUIRE(param && grad, PP_ERR_INVALID_ARG, "total_variation_add_grad: null pointer");
  PP_REQUIRE((sz_j + 7) / 8 <= 65535, PP_ERR_TOO_LARGE, "total_variation_add_grad: sz_j too large");
  // the reference host code divides the three weights by 6 (total_variation_kernel.cu:76-78)
  wx /= 6;
  wy /= 6;
  wz /= 6;
  const dim3 block(32, 8, 1);
  const int chunk = 16;
  const int chunks_per_plane = (int)((sz_i + chunk - 1) / chunk);
  PP_REQUIRE(n_planes * chunks_per_plane <= 65535, PP_ERR_TOO_LARGE, "total_variation_add_grad: too many planes");
  const dim3 grid((unsigned)((sz_k + 31) / 32), (unsigned)((sz_j + 7) / 8), (unsigned)(n_planes * chunks_per_plane));
  cudaStream_t s = as_stream(stream);
  PP_DISPATCH(dtype, {
    const scalar_t* p = (const scalar_t*)param;
    scalar_t* g = (scalar_t*)grad;
    const scalar_t* m = (const scalar_t*)mask;
    if (!mask) {
      // old kernel: k -> wz, j -> wy, i -> wz
      if (dense_mode)
        tv_add_grad_kernel<scalar_t, true, false><<<grid, block, 0, s>>>(p, g, nullptr, wz, wy, wz, (int)sz_i, (int)sz_j, (int)sz_k, chunk, chunks_per_plane);
      else
        tv_add_grad_kernel<scalar_t, false, false><<<grid, block, 0, s>>>(p, g, nullptr, wz, wy, wz, (int)sz_i, (int)sz_j, (int)sz_k, chunk, chunks_per_plane);
    } else {
      // new kernel: k -> wx, j -> wy, i -> wz
      if (dense_mode)
        tv_add_grad_kernel<scalar_t, true, true><<<grid, block, 0, s>>>(p, g, m, wx, wy, wz, (int)sz_i, (int)sz_j, (int)sz_k, chunk, chunks_per_plane);
      else
        tv_add_grad_kernel<scalar_t, false, true><<<grid, block, 0, s>>>(p, g, m, wx, wy, wz, (int)sz_i, (int)sz_j, (int)sz_k, chunk, chunks_per_plane);
    }
  });
  PP_CHECK_LAUNCH("total_variation_add_grad");
  return PP_OK;
}

template <typename scalar_t>
static int launch_adam(int mode, void* param, const void* grad, void* exp_avg, void* exp_avg_sq,
                       const void* perlr, int64_t n, float step_size, float beta1, float beta2,
                       float eps, cudaStream_t s) {
  const int64_t max_blocks = (int64_t)sm_count() * 8;
  const bool vec = sizeof(scalar_t) == 4 && (n % 4 == 0) &&
                   (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg |
                     (uintptr_t)exp_avg_sq | (uintptr_t)perlr) % 16 == 0);
  if (vec) {
    const int64_t n4 = n / 4;
    const int blocks = (int)min64(max_blocks, (n4 + 255) / 256);
#define PP_ADAM_V(M)                                                                            \
  adam_vec4_kernel<M><<<blocks, 256, 0, s>>>((float4*)param, (const float4*)grad,               \
                                             (float4*)exp_avg, (float4*)exp_avg_sq,             \
                                             (const float4*)perlr, n4, step_size, beta1, beta2, eps)
    if (mode == 0) PP_ADAM_V(0); else if (mode == 1) PP_ADAM_V(1); else PP_ADAM_V(2);
#undef PP_ADAM_V
  } else {
    const int blocks = (int)min64(max_blocks, (n + 255) / 256);
#define PP_ADAM_S(M)                                                                              \
  adam_scalar_kernel<scalar_t, M><<<blocks, 256, 0, s>>>(                                        \
      (scalar_t*)param, (const scalar_t*)grad, (scalar_t*)exp_avg, (scalar_t*)exp_avg_sq,        \
      (const scalar_t*)perlr, n, step_size, beta1, beta2, eps)
    if (mode == 0) PP_ADAM_S(0); else if (mode == 1) PP_ADAM_S(1); else PP_ADAM_S(2);
#undef PP_ADAM_S
  }
  PP_CHECK_LAUNCH("adam_upd");
  return PP_OK;
}

extern "C" int pp_adam_upd(pp_dtype dtype, int mode, void* param, const void* grad, void* exp_avg,
                           void* exp_avg_sq, const void* perlr, int64_t numel, int32_t step,
                           float beta1, float beta2, float lr, float eps, pp_stream_t stream) {
  PP_REQUIRE(numel >= 0, PP_ERR_INVALID_ARG, "adam_upd: numel < 0");
  PP_REQUIRE(mode >= 0 && mode <= 3, PP_ERR_INVALID_ARG, "adam_upd: unknown mode %d", mode);
  if (numel == 0) return PP_OK;
  PP_REQUIRE(param && grad && exp_avg && exp_avg_sq, PP_ERR_INVALID_ARG, "adam_upd: null pointer");
  PP_REQUIRE(mode != 2 || perlr, PP_ERR_INVALID_ARG, "adam_upd: perlr is null");
  // reference host code (adam_upd_kernel.cu:72): all-float bias-corrected step size
  const float step_size =
      lr * sqrtf(1.f - powf(beta2, (float)step)) / (1.f - powf(beta1, (float)step));
  if (mode == 3) {
    // the reference's LIVE optimiser (lib/utils.py:176-198): denom = sqrt(v) / sqrt(bc2) + eps, step lr / bc1, i.e.
    // m / (sqrt(v)/s + eps) = s m / (sqrt(v) + eps s) with s = sqrt(bc2): the plain kernel with eps scaled by s
    eps *= sqrtf(1.f - powf(beta2, (float)step));
    mode = 0;
  }
  if (dtype == PP_F32)
    return launch_adam<float>(mode, param, grad, exp_avg, exp_avg_sq, perlr, numel, step_size,
                              beta1, beta2, eps, as_stream(stream));
  if (dtype == PP_F64)
    return launch_adam<double>(mode, param, grad, exp_avg, exp_avg_sq, perlr, numel, step_size,
                               beta1, beta2, eps, as_stream(stream));
  set_error("unknown dtype %d", (int)dtype);
  return PP_ERR_INVALID_ARG;
}

extern "C" int pp_cumdist_thres(pp_dtype dtype, const void* dist, float thres, int64_t n_rays,
                                int64_t n_pts, uint8_t* mask, pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0 && n_pts >= 0, PP_ERR_INVALID_ARG, "cumdist_thres: negative extent");
  if (n_rays * n_pts == 0) return PP_OK;
  PP_REQUIRE(dist && mask, PP_ERR_INVALID_ARG, "cumdist_thres: null pointer");
  PP_REQUIRE(n_rays * n_pts <= INT32_MAX, PP_ERR_TOO_LARGE, "cumdist_thres: too many samples");
  const int rays_per_block = kCdWarps * 32;
  PP_DISPATCH(dtype, {
    cumdist_kernel<scalar_t><<<ceil_div(n_rays, rays_per_block), kCdWarps * kWarp, 0, as_stream(stream)>>>(
        (const scalar_t*)dist, thres, n_rays, n_pts, mask);
  });
  PP_CHECK_LAUNCH("cumdist_thres");
  return PP_OK;
}

extern "C" int pp_total_variation_loss(const float* v, const uint8_t* mask, int64_t n_planes, int64_t sz_i,
                                       int64_t sz_j, int64_t sz_k, double* sum_out, pp_stream_t stream) {
  PP_REQUIRE(n_planes >= 0 && sz_i >= 0 && sz_j >= 0 && sz_k >= 0, PP_ERR_INVALID_ARG,
             "total_variation_loss: negative extent");
  PP_REQUIRE(sum_out, PP_ERR_INVALID_ARG, "total_variation_loss: sum_out is null");
  cudaStream_t s = as_stream(stream);
  PP_CHECK_CUDA(cudaMemsetAsync(sum_out, 0, sizeof(double), s));
  if (n_planes * sz_i * sz_j * sz_k == 0) return PP_OK;
  PP_REQUIRE(v, PP_ERR_INVALID_ARG, "total_variation_loss: null pointer");
  const int chunk = 16;
  const int cpp = (int)((sz_i + chunk - 1) / chunk);
  PP_REQUIRE(n_planes * cpp <= 65535 && (sz_j + 7) / 8 <= 65535, PP_ERR_TOO_LARGE, "total_variation_loss: too large");
  const dim3 block(32, 8, 1);
  const dim3 grid((unsigned)((sz_k + 31) / 32), (unsigned)((sz_j + 7) / 8), (unsigned)(n_planes * cpp));
  tv_loss_kernel<false><<<grid, block, 0, s>>>(v, mask, (int)sz_i, (int)sz_j, (int)sz_k, chunk, cpp, 0.f, nullptr,
                                               sum_out);
  PP_CHECK_LAUNCH("total_variation_loss");
  return PP_OK;
}

extern "C" int pp_total_variation_loss_backward(const float* v, const uint8_t* mask, int64_t n_planes, int64_t sz_i,
                                                int64_t sz_j, int64_t sz_k, float scale, float* grad,
                                                pp_stream_t stream) {
  PP_REQUIRE(n_planes >= 0 && sz_i >= 0 && sz_j >= 0 && sz_k >= 0, PP_ERR_INVALID_ARG,
             "total_variation_loss_backward: negative extent");
  if (n_planes * sz_i * sz_j * sz_k == 0) return PP_OK;
  PP_REQUIRE(v && grad, PP_ERR_INVALID_ARG, "total_variation_loss_backward: null pointer");
  const int chunk = 16;
  const int cpp = (int)((sz_i + chunk - 1) / chunk);
  PP_REQUIRE(n_planes * cpp <= 65535 && (sz_j + 7) / 8 <= 65535, PP_ERR_TOO_LARGE,
             "total_variation_loss_backward: too large");
  const dim3 block(32, 8, 1);
  const dim3 grid((unsigned)((sz_k + 31) / 32), (unsigned)((sz_j + 7) / 8), (unsigned)(n_planes * cpp));
  tv_loss_kernel<true><<<grid, block, 0, as_stream(stream)>>>(v, mask, (int)sz_i, (int)sz_j, (int)sz_k, chunk, cpp,
                                                             scale, grad, nullptr);
  PP_CHECK_LAUNCH("total_variation_loss_backward");
  return PP_OK;
}
