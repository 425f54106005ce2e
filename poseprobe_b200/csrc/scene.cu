// Scene-branch compositing: the dense exp-cumsum volume rendering of the reference's
// SPARF-derived NeRF and its hierarchical resampling,
//   NeRF.composite                lib/bg_nerf/source/models/frequency_nerf.py:283-343
//   Graph.sample_depth_from_pdf   lib/bg_nerf/source/models/renderer.py:703-738
// which the reference runs as ~25 elementwise / cumsum / gather torch kernels over [B, R, N] tensors (plus autograd's
// mirror image of them) in every training step (lib/recon_scene.py:645-649: loss = 0.1 * object + scene).
// One warp per ray, lanes = 32 consecutive samples: the interval, sigma * delta, the exclusive prefix sum (shuffle
// scan carried across chunks), transmittance, weight and all seven reductions in ONE pass over the ray's samples
// (second short pass for the two variances, which need the means); the backward is one reverse pass with a suffix
// scan.  HBM-bound streaming kernels: 20 B read + 4 B written per sample forward, 28 B + 16 B backward.
#include <math.h>

#include "../../include/poseprobe_b200_fused.h"
#include "common.cuh"

namespace pp {

__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float up = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += up;
  }
  return v;
}

// per-sample quantities of sample n of a ray (frequency_nerf.py:309-315)
struct NerfSample {
  float depth, intv, sd, ema;   // depth, interval to the next sample (1e10 for the last), sigma*delta, exp(-sigma*delta)
};

__device__ __forceinline__ NerfSample nerf_sample(const float* __restrict__ depth, const float* __restrict__ dens, int n,
                                                  int N, float ray_len) {
  NerfSample s;
  s.depth = depth[n];
  s.intv = (n + 1 < N) ? depth[n + 1] - s.depth : 1e10f;
  s.sd = dens[n] * (s.intv * ray_len);
  s.ema = expf(-s.sd);
  return s;
}

__global__ void __launch_bounds__(256)
nerf_composite_fwd_kernel(const float* __restrict__ ray, const float* __restrict__ rgb_s, const float* __restrict__ dens,
                          const float* __restrict__ depth, int64_t n_rays, int N, int opaque, float* __restrict__ rgb,
                          float* __restrict__ rgb_var, float* __restrict__ depth_out, float* __restrict__ depth_var,
                          float* __restrict__ opacity, float* __restrict__ weights, float* __restrict__ all_cum) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rays) return;
  const float rx = ray[3 * r], ry = ray[3 * r + 1], rz = ray[3 * r + 2];
  const float ray_len = sqrtf(rx * rx + ry * ry + rz * rz);
  const float* d_r = depth + r * N;
  const float* s_r = dens + r * N;
  const float* c_r = rgb_s + r * (int64_t)N * 3;
  float* w_r = weights + r * N;
  float carry = 0.f;            // sum of sigma*delta of all samples before this chunk
  float a_r = 0.f, a_g = 0.f, a_b = 0.f, a_d = 0.f, a_o = 0.f;
  for (int base = 0; base < N; base += 32) {
    const int n = base + lane;
    float sd = 0.f, ema = 1.f, dn = 0.f;
    if (n < N) {
      const NerfSample s = nerf_sample(d_r, s_r, n, N, ray_len);
      sd = s.sd;
      ema = s.ema;
      dn = s.depth;
    }
    const float incl = warp_incl_scan(sd, lane);
    // exclusive prefix sum (frequency_nerf.py:316) taken from the left neighbour, never as incl - sd: the last
    // sample's sigma * delta carries the 1e10 interval and would absorb the prefix
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 0.f;
    const float T = expf(-(carry + excl));
    carry += __shfl_sync(kFull, incl, 31);
    if (n < N) {
      const float w = T * (1.f - ema);
      w_r[n] = w;
      a_r = fmaf(w, c_r[3 * n], a_r);
      a_g = fmaf(w, c_r[3 * n + 1], a_g);
      a_b = fmaf(w, c_r[3 * n + 2], a_b);
      a_d = fmaf(w, dn, a_d);
      a_o += w;
      if (n == N - 2 && all_cum) all_cum[r] = T;           // T[:, :, -2] (:320)
    }
  }
  a_r = warp_sum(a_r);
  a_g = warp_sum(a_g);
  a_b = warp_sum(a_b);
  a_d = warp_sum(a_d);
  a_o = warp_sum(a_o);
  // variances around the means (:331, :336; rgb_var uses rgb before the background is added)
  float v_d = 0.f, v_c = 0.f;
  if (depth_var || rgb_var) {
    __syncwarp();
    for (int n = lane; n < N; n += 32) {
      const float w = w_r[n];
      const float dd = d_r[n] - a_d;
      v_d = fmaf(w, dd * dd, v_d);
      v_c = fmaf(w, (c_r[3 * n] - a_r) + (c_r[3 * n + 1] - a_g) + (c_r[3 * n + 2] - a_b), v_c);
    }
    v_d = warp_sum(v_d);
    v_c = warp_sum(v_c);
  }
  if (lane == 0) {
    const float bgc = opaque ? 1.f - a_o : 0.f;            // rgb + (1 - opacity) (:340-341)
    rgb[3 * r] = a_r + bgc;
    rgb[3 * r + 1] = a_g + bgc;
    rgb[3 * r + 2] = a_b + bgc;
    depth_out[r] = a_d;
    opacity[r] = a_o;
    if (depth_var) depth_var[r] = v_d;
    if (rgb_var) rgb_var[r] = v_c;
    if (N < 2 && all_cum) all_cum[r] = 1.f;
  }
}

// d loss / d (rgb_samples, density_samples, ray) from d loss / d (rgb, depth, opacity, weights).
//   w_n = T_n (1 - e^{-sd_n}),  T_n = exp(-sum_{j<n} sd_j),  sd_n = dens_n * intv_n * |ray|
//   dL/dsd_n = gw_n T_n e^{-sd_n} - sum_{m>n} gw_m w_m,   gw_n = dL/dw_n
// Pass 1 walks the ray forward and leaves the prefix sum entering chunk k in lane k (N <= 1024); pass 2 walks it in
// reverse with a suffix scan of gw_m w_m.
__global__ void __launch_bounds__(256)
nerf_composite_bwd_kernel(const float* __restrict__ ray, const float* __restrict__ rgb_s, const float* __restrict__ dens,
                          const float* __restrict__ depth, int64_t n_rays, int N, int opaque,
                          const float* __restrict__ g_rgb, const float* __restrict__ g_depth,
                          const float* __restrict__ g_opacity, const float* __restrict__ g_weights,
                          float* __restrict__ grad_rgb_s, float* __restrict__ grad_dens, float* __restrict__ grad_ray) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rays) return;
  const float rx = ray[3 * r], ry = ray[3 * r + 1], rz = ray[3 * r + 2];
  const float ray_len = sqrtf(rx * rx + ry * ry + rz * rz);
  const float* d_r = depth + r * N;
  const float* s_r = dens + r * N;
  const float* c_r = rgb_s + r * (int64_t)N * 3;
  const float gr = g_rgb ? g_rgb[3 * r] : 0.f, gg = g_rgb ? g_rgb[3 * r + 1] : 0.f, gb = g_rgb ? g_rgb[3 * r + 2] : 0.f;
  const float gd = g_depth ? g_depth[r] : 0.f;
  const float go = (g_opacity ? g_opacity[r] : 0.f) - (opaque ? (gr + gg + gb) : 0.f);
  // pass 1: prefix sum of sigma*delta entering every chunk
  float carry_of_my_chunk = 0.f, carry = 0.f;
  for (int base = 0, k = 0; base < N; base += 32, ++k) {
    const int n = base + lane;
    const float sd = (n < N) ? nerf_sample(d_r, s_r, n, N, ray_len).sd : 0.f;
    if (lane == k) carry_of_my_chunk = carry;
    carry += warp_sum(sd);
  }
  // pass 2: reverse
  float suffix = 0.f, g_len = 0.f;
  const int last_base = ((N - 1) / 32) * 32;
  for (int base = last_base, k = last_base / 32; base >= 0; base -= 32, --k) {
    const int n = base + lane;
    const float carry_in = __shfl_sync(kFull, carry_of_my_chunk, k);
    NerfSample s;
    s.depth = 0.f; s.intv = 0.f; s.sd = 0.f; s.ema = 1.f;
    float dn = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    if (n < N) {
      s = nerf_sample(d_r, s_r, n, N, ray_len);
      dn = s_r[n];
      cr = c_r[3 * n];
      cg = c_r[3 * n + 1];
      cb = c_r[3 * n + 2];
    }
    const float incl = warp_incl_scan(s.sd, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 0.f;
    const float T = expf(-(carry_in + excl));
    const float w = (n < N) ? T * (1.f - s.ema) : 0.f;
    const float gw = (n < N) ? gr * cr + gg * cg + gb * cb + gd * s.depth + go + (g_weights ? g_weights[r * N + n] : 0.f)
                             : 0.f;
    const float c = gw * w;
    float S = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float dnv = __shfl_down_sync(kFull, S, o);
      if (lane + o < 32) S += dnv;
    }
    float right = __shfl_down_sync(kFull, S, 1);
    if (lane == 31) right = 0.f;
    const float after = suffix + right;                     // sum over the samples strictly after n
    suffix += __shfl_sync(kFull, S, 0);
    if (n < N) {
      if (grad_rgb_s) {
        float* o = grad_rgb_s + (r * N + n) * 3;
        o[0] = gr * w;
        o[1] = gg * w;
        o[2] = gb * w;
      }
      const float g_sd = gw * T * s.ema - after;
      if (grad_dens) grad_dens[r * N + n] = g_sd * (s.intv * ray_len);
      g_len += g_sd * dn * s.intv;
    }
  }
  if (grad_ray) {
    g_len = warp_sum(g_len);
    if (lane == 0) {
      const float k = ray_len > 0.f ? g_len / ray_len : 0.f;
      grad_ray[3 * r] = k * rx;
      grad_ray[3 * r + 1] = k * ry;
      grad_ray[3 * r + 2] = k * rz;
    }
  }
}

// Inverse-CDF resampling of the coarse weights (renderer.py:703-738).  One warp per ray: pdf = w / (sum w + 1e-6),
// cdf = [0, cumsum(pdf)] staged in shared memory (accumulated in double, as ATen's CPU cumsum does for float), then
// every lane takes fine samples lane, lane + 32, ...: idx = searchsorted(cdf, u, right=True) by bisection, depth bins
// linspace(depth_min, depth_max, N + 1) with ATen's two-sided formula, linear interpolation with the + 1e-8 guard.
__global__ void __launch_bounds__(256)
nerf_sample_pdf_kernel(const float* __restrict__ weights, int64_t n_rays, int N, const float* __restrict__ unif, int Nf,
                       float depth_min, float depth_max, float* __restrict__ out) {
  extern __shared__ float s_cdf_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* cdf = s_cdf_all + (size_t)warp * (N + 1);
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rays) return;
  const float* w_r = weights + r * N;
  float tot = 0.f;
  for (int n = lane; n < N; n += 32) tot += w_r[n];
  tot = warp_sum(tot);
  const float denom = tot + 1e-6f;
  double carry = 0.;
  if (lane == 0) cdf[0] = 0.f;
  for (int base = 0; base < N; base += 32) {
    const int n = base + lane;
    const double p = (n < N) ? static_cast<double>(w_r[n] / denom) : 0.;
    double incl = p;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += up;
    }
    if (n < N) cdf[n + 1] = static_cast<float>(carry + incl);
    carry += __shfl_sync(kFull, incl, 31);
  }
  __syncwarp();
  const float step = (depth_max - depth_min) / static_cast<float>(N);
  auto bin = [&](int i) {   // at::linspace: start + step * i below the middle, end - step * (steps - 1 - i) above
    return (i < (N + 1) / 2) ? depth_min + step * static_cast<float>(i) : depth_max - step * static_cast<float>(N - i);
  };
  for (int f = lane; f < Nf; f += 32) {
    const float u = unif[f];
    int lo = 0, hi = N + 1;                 // first index with cdf[idx] > u
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] > u) hi = mid; else lo = mid + 1;
    }
    const int idx = lo;
    const int il = max(idx - 1, 0), ih = min(idx, N);
    const float cl = cdf[il], ch = cdf[ih];
    const float dl = bin(il), dh = bin(ih);
    const float t = (u - cl) / (ch - cl + 1e-8f);
    out[r * Nf + f] = dl + t * (dh - dl);
  }
}


// cat([depth_samples, depth_samples_fine], dim=2).sort(dim=2).values (renderer.py:612-614): the union of the coarse and
// the resampled depths of a ray in ascending order.  One warp per ray: the na + nb values (<= 1024) are padded with +inf
// to a power of two in shared memory and sorted with a bitonic network (compare-exchange pairs strided over the lanes).
__global__ void __launch_bounds__(256)
nerf_sort_depth_kernel(const float* __restrict__ a, int na, const float* __restrict__ b, int nb, int64_t n_rays,
                       int n_pad, float* __restrict__ out) {
  extern __shared__ float s_sort_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* v = s_sort_all + (size_t)warp * n_pad;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rays) return;
  const int n = na + nb;
  for (int i = lane; i < n_pad; i += 32)
    v[i] = i < na ? a[r * na + i] : (i < n ? b[r * nb + (i - na)] : __int_as_float(0x7f800000));
  __syncwarp();
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n_pad >> 1); t += 32) {
        // t-th compare-exchange of this stage: partner indices i < l = i ^ j
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const bool up = (i & k) == 0;
        const float x = v[i], y = v[l];
        if ((x > y) == up) {
          v[i] = y;
          v[l] = x;
        }
      }
      __syncwarp();
    }
  }
  for (int i = lane; i < n; i += 32) out[r * n + i] = v[i];
}

}  // namespace pp

using namespace pp;

extern "C" int pp_nerf_composite_forward(const float* ray, const float* rgb_samples, const float* density_samples,
                                         const float* depth_samples, int64_t n_rays, int32_t n_samples,
                                         int setbg_opaque, float* rgb, float* rgb_var, float* depth, float* depth_var,
                                         float* opacity, float* weights, float* all_cumulated, pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0 && n_samples >= 0, PP_ERR_INVALID_ARG, "nerf_composite_forward: negative extent");
  if (n_rays == 0 || n_samples == 0) return PP_OK;
  PP_REQUIRE(ray && rgb_samples && density_samples && depth_samples && rgb && depth && opacity && weights,
             PP_ERR_INVALID_ARG, "nerf_composite_forward: null pointer");
  nerf_composite_fwd_kernel<<<ceil_div(n_rays * 32, 256), 256, 0, as_stream(stream)>>>(
      ray, rgb_samples, density_samples, depth_samples, n_rays, n_samples, setbg_opaque, rgb, rgb_var, depth, depth_var,
      opacity, weights, all_cumulated);
  PP_CHECK_LAUNCH("nerf_composite_forward");
  return PP_OK;
}

extern "C" int pp_nerf_composite_backward(const float* ray, const float* rgb_samples, const float* density_samples,
                                          const float* depth_samples, int64_t n_rays, int32_t n_samples,
                                          int setbg_opaque, const float* grad_rgb, const float* grad_depth,
                                          const float* grad_opacity, const float* grad_weights, float* grad_rgb_samples,
                                          float* grad_density_samples, float* grad_ray, pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0 && n_samples >= 0, PP_ERR_INVALID_ARG, "nerf_composite_backward: negative extent");
  PP_REQUIRE(n_samples <= 1024, PP_ERR_UNSUPPORTED, "nerf_composite_backward: at most 1024 samples per ray");
  if (n_rays == 0 || n_samples == 0) return PP_OK;
  PP_REQUIRE(ray && rgb_samples && density_samples && depth_samples, PP_ERR_INVALID_ARG,
             "nerf_composite_backward: null pointer");
  nerf_composite_bwd_kernel<<<ceil_div(n_rays * 32, 256), 256, 0, as_stream(stream)>>>(
      ray, rgb_samples, density_samples, depth_samples, n_rays, n_samples, setbg_opaque, grad_rgb, grad_depth,
      grad_opacity, grad_weights, grad_rgb_samples, grad_density_samples, grad_ray);
  PP_CHECK_LAUNCH("nerf_composite_backward");
  return PP_OK;
}

extern "C" int pp_nerf_sample_depth_from_pdf(const float* weights, int64_t n_rays, int32_t n_samples_coarse,
                                             const float* unif, int32_t n_samples_fine, float depth_min,
                                             float depth_max, float* depth_samples_fine, pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0 && n_samples_coarse >= 1 && n_samples_fine >= 0, PP_ERR_INVALID_ARG,
             "nerf_sample_depth_from_pdf: bad extent");
  PP_REQUIRE(n_samples_coarse <= 4095, PP_ERR_UNSUPPORTED, "nerf_sample_depth_from_pdf: at most 4095 coarse samples");
  if (n_rays == 0 || n_samples_fine == 0) return PP_OK;
  PP_REQUIRE(weights && unif && depth_samples_fine, PP_ERR_INVALID_ARG, "nerf_sample_depth_from_pdf: null pointer");
  const size_t smem = (size_t)8 * (n_samples_coarse + 1) * sizeof(float);
  if (smem > 48 * 1024)
    PP_CHECK_CUDA(cudaFuncSetAttribute(nerf_sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nerf_sample_pdf_kernel<<<ceil_div(n_rays * 32, 256), 256, smem, as_stream(stream)>>>(
      weights, n_rays, n_samples_coarse, unif, n_samples_fine, depth_min, depth_max, depth_samples_fine);
  PP_CHECK_LAUNCH("nerf_sample_depth_from_pdf");
  return PP_OK;
}

extern "C" int pp_nerf_sort_depth(const float* depth_a, int32_t n_a, const float* depth_b, int32_t n_b, int64_t n_rays,
                                  float* out, pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0 && n_a >= 0 && n_b >= 0, PP_ERR_INVALID_ARG, "nerf_sort_depth: negative extent");
  PP_REQUIRE(n_a + n_b <= 1024, PP_ERR_UNSUPPORTED, "nerf_sort_depth: at most 1024 samples per ray");
  if (n_rays == 0 || n_a + n_b == 0) return PP_OK;
  PP_REQUIRE((depth_a || n_a == 0) && (depth_b || n_b == 0) && out, PP_ERR_INVALID_ARG, "nerf_sort_depth: null pointer");
  int n_pad = 2;
  while (n_pad < n_a + n_b) n_pad <<= 1;
  const size_t smem = (size_t)8 * n_pad * sizeof(float);
  nerf_sort_depth_kernel<<<ceil_div(n_rays * 32, 256), 256, smem, as_stream(stream)>>>(depth_a, n_a, depth_b, n_b,
                                                                                         n_rays, n_pad, out);
  PP_CHECK_LAUNCH("nerf_sort_depth");
  return PP_OK;
}
