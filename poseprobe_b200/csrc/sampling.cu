// Ray / bounding-box sampling operators.
//
// Replaces the sampling half of the reference's render_utils extension
// (lib/cuda/render_utils_kernel.cu:12-424).  The reference enumerates samples with six small
// kernels, two ATen cumsums and a host sync; here phase 1 is ONE kernel (per-ray geometry + a
// single-pass decoupled-look-back scan of N_steps) and phase 2 is ONE kernel (a warp writes the
// contiguous segment of its ray, so ray_id/step_id/pts/mask stores are fully coalesced and no
// per-sample search or scatter is needed).
//
// Integer / bool outputs must be bit-exact with the reference, so every float expression keeps the
// reference's operand types, evaluation order and (nvcc -fmad=true) FMA contraction; the
// contractions are spelled out with fma_rn().
#include "common.cuh"
#include "ray_math.cuh"

namespace pp {

// ---- the three small exported helpers ----------------------------------------------------------
template <typename scalar_t>
__global__ void t_minmax_kernel(const scalar_t* __restrict__ rays_o,
                                const scalar_t* __restrict__ rays_d,
                                const scalar_t* __restrict__ xyz_min,
                                const scalar_t* __restrict__ xyz_max, float near, float far,
                                int64_t n_rays, scalar_t* __restrict__ t_min,
                                scalar_t* __restrict__ t_max) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  scalar_t o[3], d[3], mn[3], mx[3];
  load3(rays_o, r, o);
  load3(rays_d, r, d);
  load3(xyz_min, 0, mn);
  load3(xyz_max, 0, mx);
  float lo, hi;
  ray_bbox_range(o, d, mn, mx, near, far, lo, hi);
  t_min[r] = lo;
  t_max[r] = hi;
}

template <typename scalar_t>
__global__ void n_samples_kernel(const scalar_t* __restrict__ rays_d,
                                 const scalar_t* __restrict__ t_min,
                                 const scalar_t* __restrict__ t_max, float stepdist,
                                 int64_t n_rays, int64_t* __restrict__ n_samples) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  scalar_t d[3];
  load3(rays_d, r, d);
  const float rnorm = ray_norm(d);
  // The reference subtracts the stored scalar_t values (:53); they may be arbitrary user tensors
  // here, so do not round them through float.
  const scalar_t steps = ceil((t_max[r] - t_min[r]) * rnorm / stepdist);
  n_samples[r] = static_cast<int64_t>(fmax(static_cast<double>(steps), 1.0));
}

template <typename scalar_t>
__global__ void ray_start_dir_kernel(const scalar_t* __restrict__ rays_o,
                                     const scalar_t* __restrict__ rays_d,
                                     const scalar_t* __restrict__ t_min, int64_t n_rays,
                                     scalar_t* __restrict__ rays_start,
                                     scalar_t* __restrict__ rays_dir) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  scalar_t o[3], d[3];
  load3(rays_o, r, o);
  load3(rays_d, r, d);
  const float rnorm = ray_norm(d);
  const scalar_t t0 = t_min[r];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    rays_start[3 * r + c] = fma_rn(d[c], t0, o[c]);  // o + d*t_min  (:72-74)
    rays_dir[3 * r + c] = d[c] / rnorm;              // true division (:75-77)
  }
}

// ---- phase 1: fused per-ray geometry + single-pass scan ---------------------------------------
constexpr int kPrepBlock = 256;
// tile status word: [63:62] flag (0 = not ready, 1 = aggregate, 2 = inclusive prefix), [61:0] value
constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

template <typename scalar_t>
__global__ void __launch_bounds__(kPrepBlock)
ray_prepare_kernel(const scalar_t* __restrict__ rays_o, const scalar_t* __restrict__ rays_d,
                   const scalar_t* __restrict__ xyz_min, const scalar_t* __restrict__ xyz_max,
                   float near, float far, float stepdist, int64_t n_rays,
                   scalar_t* __restrict__ t_min_out, scalar_t* __restrict__ t_max_out,
                   int64_t* __restrict__ n_steps_out, int64_t* __restrict__ cumsum_out,
                   scalar_t* __restrict__ rays_start, scalar_t* __restrict__ rays_dir,
                   unsigned long long* __restrict__ scratch,  // [0] ticket, [1..] tile status
                   int64_t* __restrict__ total_dev, volatile int64_t* total_host) {
  __shared__ unsigned int s_tile;
  __shared__ long long s_warp_tot[kPrepBlock / kWarp];
  __shared__ long long s_prefix;
  if (threadIdx.x == 0) s_tile = static_cast<unsigned int>(atomicAdd(&scratch[0], 1ull));
  __syncthreads();
  const unsigned int tile = s_tile;  // logical tile id in launch order (deadlock-free look-back)
  const int64_t r = tile * (int64_t)kPrepBlock + threadIdx.x;

  long long n = 0;
  if (r < n_rays) {
    scalar_t o[3], d[3], mn[3], mx[3];
    load3(rays_o, r, o);
    load3(rays_d, r, d);
    load3(xyz_min, 0, mn);
    load3(xyz_max, 0, mx);
    float lo, hi;
    ray_bbox_range(o, d, mn, mx, near, far, lo, hi);
    const float rnorm = ray_norm(d);
    n = ray_step_count<scalar_t>(lo, hi, rnorm, stepdist);
    t_min_out[r] = lo;
    t_max_out[r] = hi;
    n_steps_out[r] = n;
    const scalar_t t0 = lo;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      rays_start[3 * r + c] = fma_rn(d[c], t0, o[c]);
      rays_dir[3 * r + c] = d[c] / rnorm;
    }
  }

  // block-wide inclusive scan of n
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long inc = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long up = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += up;
  }
  if (lane == 31) s_warp_tot[warp] = inc;
  __syncthreads();
  long long warp_off = 0, block_total = 0;
#pragma unroll
  for (int w = 0; w < kPrepBlock / kWarp; ++w) {
    const long long t = s_warp_tot[w];
    if (w < warp) warp_off += t;
    block_total += t;
  }
  inc += warp_off;

  // decoupled look-back across tiles
  if (threadIdx.x == 0) {
    unsigned long long* status = scratch + 1;
    long long prefix = 0;
    if (tile == 0) {
      atomicExch(&status[0], kFlagPrefix | static_cast<unsigned long long>(block_total));
    } else {
      atomicExch(&status[tile], kFlagAgg | static_cast<unsigned long long>(block_total));
      for (long long p = (long long)tile - 1; p >= 0; --p) {
        unsigned long long w;
        do {
          w = atomicAdd(&status[p], 0ull);
        } while ((w >> 62) == 0);
        prefix += static_cast<long long>(w & kValueMask);
        if ((w >> 62) == 2) break;
      }
      atomicExch(&status[tile], kFlagPrefix | static_cast<unsigned long long>(prefix + block_total));
    }
    s_prefix = prefix;
    if (tile == gridDim.x - 1) {
      const long long total = prefix + block_total;
      *total_dev = total;
      if (total_host) {
        *total_host = total;
        __threadfence_system();
      }
    }
  }
  __syncthreads();
  if (r < n_rays) cumsum_out[r] = s_prefix + inc;
}

// ---- phase 2: a warp writes its ray's segment --------------------------------------------------
constexpr int kFillBlock = 256;

template <typename scalar_t>
__global__ void __launch_bounds__(kFillBlock)
ray_fill_kernel(const scalar_t* __restrict__ rays_start, const scalar_t* __restrict__ rays_dir,
                const scalar_t* __restrict__ xyz_min, const scalar_t* __restrict__ xyz_max,
                const int64_t* __restrict__ n_steps, const int64_t* __restrict__ cumsum,
                float stepdist, int64_t n_rays, int64_t total, scalar_t* __restrict__ rays_pts,
                uint8_t* __restrict__ mask_outbbox, int64_t* __restrict__ ray_id,
                int64_t* __restrict__ step_id) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (kFillBlock / kWarp);
  scalar_t mn[3], mx[3];
  load3(xyz_min, 0, mn);
  load3(xyz_max, 0, mx);
  for (int64_t r = blockIdx.x * (int64_t)(kFillBlock / kWarp) + (threadIdx.x >> 5); r < n_rays;
       r += warps_total) {
    const int64_t n = n_steps[r];
    const int64_t base = cumsum[r] - n;
    scalar_t s[3], d[3];
    load3(rays_start, r, s);
    load3(rays_dir, r, d);
    for (int64_t i = lane; i < n; i += kWarp) {
      const int64_t idx = base + i;
      if (idx >= total) break;
      const float dist = stepdist * static_cast<float>(static_cast<int>(i));  // :184
      // start + dir*dist is an FMA in scalar_t, narrowed to float (:185-187)
      const float px = static_cast<float>(fma_rn(d[0], static_cast<scalar_t>(dist), s[0]));
      const float py = static_cast<float>(fma_rn(d[1], static_cast<scalar_t>(dist), s[1]));
      const float pz = static_cast<float>(fma_rn(d[2], static_cast<scalar_t>(dist), s[2]));
      if (rays_pts) {
        rays_pts[3 * idx] = px;
        rays_pts[3 * idx + 1] = py;
        rays_pts[3 * idx + 2] = pz;
      }
      if (mask_outbbox)
        mask_outbbox[idx] = (mn[0] > px) | (mn[1] > py) | (mn[2] > pz) | (mx[0] < px) |
                            (mx[1] < py) | (mx[2] < pz);
      if (ray_id) ray_id[idx] = r;
      if (step_id) step_id[idx] = i;
    }
  }
}

// ---- fixed-count samplers ------------------------------------------------------------------------
// reference :245-270
template <typename scalar_t>
__global__ void ndc_sample_kernel(const scalar_t* __restrict__ rays_o,
                                  const scalar_t* __restrict__ rays_d,
                                  const scalar_t* __restrict__ xyz_min,
                                  const scalar_t* __restrict__ xyz_max, int n_samples,
                                  int64_t n_total, scalar_t* __restrict__ rays_pts,
                                  uint8_t* __restrict__ mask_outbbox) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n_total) return;
  const int64_t r = idx / n_samples;
  const int i = static_cast<int>(idx - r * n_samples);
  const float dist = static_cast<float>(i) / (n_samples - 1);
  scalar_t o[3], d[3];
  load3(rays_o, r, o);
  load3(rays_d, r, d);
  const float px = static_cast<float>(fma_rn(d[0], static_cast<scalar_t>(dist), o[0]));
  const float py = static_cast<float>(fma_rn(d[1], static_cast<scalar_t>(dist), o[1]));
  const float pz = static_cast<float>(fma_rn(d[2], static_cast<scalar_t>(dist), o[2]));
  rays_pts[3 * idx] = px;
  rays_pts[3 * idx + 1] = py;
  rays_pts[3 * idx + 2] = pz;
  mask_outbbox[idx] = (xyz_min[0] > px) | (xyz_min[1] > py) | (xyz_min[2] > pz) |
                      (xyz_max[0] < px) | (xyz_max[1] < py) | (xyz_max[2] < pz);
}

// reference :301-340 (inverse-sphere background warp).  Float locals, double literals.
template <typename scalar_t>
__global__ void bg_sample_kernel(const scalar_t* __restrict__ rays_o,
                                 const scalar_t* __restrict__ rays_d,
                                 const scalar_t* __restrict__ t_max, float bg_preserve,
                                 int n_samples, int64_t n_total, scalar_t* __restrict__ rays_pts) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n_total) return;
  const int64_t r = idx / n_samples;
  const int i = static_cast<int>(idx - r * n_samples);
  const float t_inner = static_cast<float>(t_max[r]);
  // t_inner - 1. + 1. / (1. - (float)i / N): the float quotient is widened, the rest is double
  const float frac = static_cast<float>(i) / n_samples;
  const float t = static_cast<float>(static_cast<double>(t_inner) - 1. + 1. / (1. - static_cast<double>(frac)));
  scalar_t o[3], d[3];
  load3(rays_o, r, o);
  load3(rays_d, r, d);
  const float qx = static_cast<float>(fma_rn(d[0], static_cast<scalar_t>(t), o[0]));
  const float qy = static_cast<float>(fma_rn(d[1], static_cast<scalar_t>(t), o[1]));
  const float qz = static_cast<float>(fma_rn(d[2], static_cast<scalar_t>(t), o[2]));
  const float t_outer = sqrtf(fma_rn(qz, qz, fma_rn(qx, qx, mul_rn(qy, qy))));  // same contraction as ray_norm
  const float m = fmaxf(fabsf(qx), fmaxf(fabsf(qy), fabsf(qz)));
  const float R = t_outer / m;
  // R*R/(t*t)*(1.-bg) + R/t*bg : products/quotients in float, the (1.-bg) factor and the sum in
  // double (FMA-contracted there), narrowed to float.
  const double lhs = static_cast<double>(mul_rn(R, R) / mul_rn(t_outer, t_outer));
  const double keep = 1. - static_cast<double>(bg_preserve);
  const float rhs = mul_rn(R / t_outer, bg_preserve);
  const float o2i = static_cast<float>(fma_rn(lhs, keep, static_cast<double>(rhs)));
  rays_pts[3 * idx] = mul_rn(qx, o2i);
  rays_pts[3 * idx + 1] = mul_rn(qy, o2i);
  rays_pts[3 * idx + 2] = mul_rn(qz, o2i);
}

// reference :374-392.  Writes every output (false when the voxel index is out of range).
template <typename scalar_t>
__global__ void maskcache_kernel(const uint8_t* __restrict__ world,
                                 const scalar_t* __restrict__ xyz, const scalar_t* __restrict__ scale,
                                 const scalar_t* __restrict__ shift, int sz_i, int sz_j, int sz_k,
                                 int64_t n_pts, uint8_t* __restrict__ out) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n_pts) return;
  // round(x*scale + shift): FMA in scalar_t, round-half-away, float->int conversion saturates
  const int i = static_cast<int>(round(fma_rn(xyz[3 * p], scale[0], shift[0])));
  const int j = static_cast<int>(round(fma_rn(xyz[3 * p + 1], scale[1], shift[1])));
  const int k = static_cast<int>(round(fma_rn(xyz[3 * p + 2], scale[2], shift[2])));
  uint8_t v = 0;
  if (0 <= i && i < sz_i && 0 <= j && j < sz_j && 0 <= k && k < sz_k)
    v = world[((int64_t)i * sz_j + j) * sz_k + k];
  out[p] = v;
}

}  // namespace pp

// ================================================================================================
// C ABI
// ================================================================================================
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

extern "C" int pp_infer_t_minmax(pp_dtype dtype, const void* rays_o, const void* rays_d,
                                 const void* xyz_min, const void* xyz_max, float near, float far,
                                 int64_t n_rays, void* t_min, void* t_max, pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0, PP_ERR_INVALID_ARG, "infer_t_minmax: n_rays < 0");
  if (n_rays == 0) return PP_OK;
  PP_REQUIRE(rays_o && rays_d && xyz_min && xyz_max && t_min && t_max, PP_ERR_INVALID_ARG,
             "infer_t_minmax: null pointer");
  PP_DISPATCH(dtype, {
    t_minmax_kernel<scalar_t><<<ceil_div(n_rays, 256), 256, 0, as_stream(stream)>>>(
        (const scalar_t*)rays_o, (const scalar_t*)rays_d, (const scalar_t*)xyz_min,
        (const scalar_t*)xyz_max, near, far, n_rays, (scalar_t*)t_min, (scalar_t*)t_max);
  });
  PP_CHECK_LAUNCH("infer_t_minmax");
  return PP_OK;
}

extern "C" int pp_infer_n_samples(pp_dtype dtype, const void* rays_d, const void* t_min,
                                  const void* t_max, float stepdist, int64_t n_rays,
                                  int64_t* n_samples, pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0, PP_ERR_INVALID_ARG, "infer_n_samples: n_rays < 0");
  if (n_rays == 0) return PP_OK;
  PP_REQUIRE(rays_d && t_min && t_max && n_samples, PP_ERR_INVALID_ARG,
             "infer_n_samples: null pointer");
  PP_DISPATCH(dtype, {
    n_samples_kernel<scalar_t><<<ceil_div(n_rays, 256), 256, 0, as_stream(stream)>>>(
        (const scalar_t*)rays_d, (const scalar_t*)t_min, (const scalar_t*)t_max, stepdist, n_rays,
        n_samples);
  });
  PP_CHECK_LAUNCH("infer_n_samples");
  return PP_OK;
}

extern "C" int pp_infer_ray_start_dir(pp_dtype dtype, const void* rays_o, const void* rays_d,
                                      const void* t_min, int64_t n_rays, void* rays_start,
                                      void* rays_dir, pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0, PP_ERR_INVALID_ARG, "infer_ray_start_dir: n_rays < 0");
  if (n_rays == 0) return PP_OK;
  PP_REQUIRE(rays_o && rays_d && t_min && rays_start && rays_dir, PP_ERR_INVALID_ARG,
             "infer_ray_start_dir: null pointer");
  PP_DISPATCH(dtype, {
    ray_start_dir_kernel<scalar_t><<<ceil_div(n_rays, 256), 256, 0, as_stream(stream)>>>(
        (const scalar_t*)rays_o, (const scalar_t*)rays_d, (const scalar_t*)t_min, n_rays,
        (scalar_t*)rays_start, (scalar_t*)rays_dir);
  });
  PP_CHECK_LAUNCH("infer_ray_start_dir");
  return PP_OK;
}

extern "C" int64_t pp_sample_scan_scratch_elems(int64_t n_rays) {
  return 1 + (n_rays + kPrepBlock - 1) / kPrepBlock;
}

extern "C" int pp_sample_pts_prepare(pp_dtype dtype, const void* rays_o, const void* rays_d,
                                     const void* xyz_min, const void* xyz_max, float near,
                                     float far, float stepdist, int64_t n_rays, void* t_min,
                                     void* t_max, int64_t* n_steps, int64_t* n_steps_cumsum,
                                     void* rays_start, void* rays_dir, int64_t* scan_scratch,
                                     int64_t* total_dev, int64_t* total_host_pinned,
                                     pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0, PP_ERR_INVALID_ARG, "sample_pts_prepare: n_rays < 0");
  PP_REQUIRE(total_dev, PP_ERR_INVALID_ARG, "sample_pts_prepare: total_dev is null");
  cudaStream_t s = as_stream(stream);
  if (n_rays == 0) {
    PP_CHECK_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(int64_t), s));
    if (total_host_pinned) *total_host_pinned = 0;
    return PP_OK;
  }
  PP_REQUIRE(rays_o && rays_d && xyz_min && xyz_max && t_min && t_max && n_steps &&
                 n_steps_cumsum && rays_start && rays_dir && scan_scratch,
             PP_ERR_INVALID_ARG, "sample_pts_prepare: null pointer");
  const int64_t tiles = (n_rays + kPrepBlock - 1) / kPrepBlock;
  PP_CHECK_CUDA(cudaMemsetAsync(scan_scratch, 0, sizeof(int64_t) * (1 + tiles), s));
  PP_DISPATCH(dtype, {
    ray_prepare_kernel<scalar_t><<<(unsigned)tiles, kPrepBlock, 0, s>>>(
        (const scalar_t*)rays_o, (const scalar_t*)rays_d, (const scalar_t*)xyz_min,
        (const scalar_t*)xyz_max, near, far, stepdist, n_rays, (scalar_t*)t_min, (scalar_t*)t_max,
        n_steps, n_steps_cumsum, (scalar_t*)rays_start, (scalar_t*)rays_dir,
        (unsigned long long*)scan_scratch, total_dev, total_host_pinned);
  });
  PP_CHECK_LAUNCH("sample_pts_prepare");
  return PP_OK;
}

extern "C" int pp_sample_pts_fill(pp_dtype dtype, const void* rays_start, const void* rays_dir,
                                  const void* xyz_min, const void* xyz_max, const int64_t* n_steps,
                                  const int64_t* n_steps_cumsum, float stepdist, int64_t n_rays,
                                  int64_t total, void* rays_pts, uint8_t* mask_outbbox,
                                  int64_t* ray_id, int64_t* step_id, pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0 && total >= 0, PP_ERR_INVALID_ARG, "sample_pts_fill: negative extent");
  // the reference indexes samples with 32-bit ints (render_utils_kernel.cu:177-183, item<int>() :212)
  PP_REQUIRE(total <= INT32_MAX / 3, PP_ERR_TOO_LARGE,
             "sample_pts_fill: %lld samples exceed the 32-bit sample index space", (long long)total);
  if (n_rays == 0 || total == 0) return PP_OK;
  PP_REQUIRE(rays_start && rays_dir && xyz_min && xyz_max && n_steps && n_steps_cumsum,
             PP_ERR_INVALID_ARG, "sample_pts_fill: null pointer");
  const int warps_per_block = kFillBlock / kWarp;
  int64_t blocks = (n_rays + warps_per_block - 1) / warps_per_block;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  PP_DISPATCH(dtype, {
    ray_fill_kernel<scalar_t><<<(unsigned)blocks, kFillBlock, 0, as_stream(stream)>>>(
        (const scalar_t*)rays_start, (const scalar_t*)rays_dir, (const scalar_t*)xyz_min,
        (const scalar_t*)xyz_max, n_steps, n_steps_cumsum, stepdist, n_rays, total,
        (scalar_t*)rays_pts, mask_outbbox, ray_id, step_id);
  });
  PP_CHECK_LAUNCH("sample_pts_fill");
  return PP_OK;
}

extern "C" int pp_sample_ndc_pts_on_rays(pp_dtype dtype, const void* rays_o, const void* rays_d,
                                         const void* xyz_min, const void* xyz_max,
                                         int32_t n_samples, int64_t n_rays, void* rays_pts,
                                         uint8_t* mask_outbbox, pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0 && n_samples >= 0, PP_ERR_INVALID_ARG, "sample_ndc: negative extent");
  const int64_t n_total = n_rays * n_samples;
  PP_REQUIRE(n_total <= INT32_MAX / 3, PP_ERR_TOO_LARGE, "sample_ndc: too many samples");
  if (n_total == 0) return PP_OK;
  PP_REQUIRE(rays_o && rays_d && xyz_min && xyz_max && rays_pts && mask_outbbox,
             PP_ERR_INVALID_ARG, "sample_ndc: null pointer");
  PP_DISPATCH(dtype, {
    ndc_sample_kernel<scalar_t><<<ceil_div(n_total, 256), 256, 0, as_stream(stream)>>>(
        (const scalar_t*)rays_o, (const scalar_t*)rays_d, (const scalar_t*)xyz_min,
        (const scalar_t*)xyz_max, n_samples, n_total, (scalar_t*)rays_pts, mask_outbbox);
  });
  PP_CHECK_LAUNCH("sample_ndc_pts_on_rays");
  return PP_OK;
}

extern "C" int pp_sample_bg_pts_on_rays(pp_dtype dtype, const void* rays_o, const void* rays_d,
                                        const void* t_max, float bg_preserve, int32_t n_samples,
                                        int64_t n_rays, void* rays_pts, pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0 && n_samples >= 0, PP_ERR_INVALID_ARG, "sample_bg: negative extent");
  const int64_t n_total = n_rays * n_samples;
  PP_REQUIRE(n_total <= INT32_MAX / 3, PP_ERR_TOO_LARGE, "sample_bg: too many samples");
  if (n_total == 0) return PP_OK;
  PP_REQUIRE(rays_o && rays_d && t_max && rays_pts, PP_ERR_INVALID_ARG, "sample_bg: null pointer");
  PP_DISPATCH(dtype, {
    bg_sample_kernel<scalar_t><<<ceil_div(n_total, 256), 256, 0, as_stream(stream)>>>(
        (const scalar_t*)rays_o, (const scalar_t*)rays_d, (const scalar_t*)t_max, bg_preserve,
        n_samples, n_total, (scalar_t*)rays_pts);
  });
  PP_CHECK_LAUNCH("sample_bg_pts_on_rays");
  return PP_OK;
}

extern "C" int pp_maskcache_lookup(pp_dtype dtype, const uint8_t* world, const void* xyz,
                                   const void* xyz2ijk_scale, const void* xyz2ijk_shift,
                                   int32_t sz_i, int32_t sz_j, int32_t sz_k, int64_t n_pts,
                                   uint8_t* out, pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0, PP_ERR_INVALID_ARG, "maskcache_lookup: n_pts < 0");
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(world && xyz && xyz2ijk_scale && xyz2ijk_shift && out, PP_ERR_INVALID_ARG,
             "maskcache_lookup: null pointer");
  PP_DISPATCH(dtype, {
    maskcache_kernel<scalar_t><<<ceil_div(n_pts, 256), 256, 0, as_stream(stream)>>>(
        world, (const scalar_t*)xyz, (const scalar_t*)xyz2ijk_scale,
        (const scalar_t*)xyz2ijk_shift, sz_i, sz_j, sz_k, n_pts, out);
  });
  PP_CHECK_LAUNCH("maskcache_lookup");
  return PP_OK;
}
