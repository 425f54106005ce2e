// Density -> alpha (DVGO raw2alpha family) and SDF -> alpha (NeuS mid-point rule).
//
// raw2alpha* replace render_utils_kernel.cu:431-574.  They are pure streaming element-wise maps
// (4 B in, 8 B out forward; 8 B in, 4 B out backward), so the only B200 concerns are coalescing and
// keeping enough bytes in flight: each thread handles one element of a 256-thread block and the
// grid covers the array once.  The arithmetic keeps the reference's operand types (float/double
// promotion of the literals) so results are bit-identical to the reference kernels for fp32.
#include "common.cuh"
#include "neus.cuh"

namespace pp {

// reference :437-442 / :452-457
template <typename scalar_t, bool kNonUni>
__global__ void raw2alpha_kernel(const scalar_t* __restrict__ density, float shift, float interval,
                                 const scalar_t* __restrict__ interval_pts, int64_t n,
                                 scalar_t* __restrict__ exp_d, scalar_t* __restrict__ alpha) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const scalar_t e = exp(density[i] + shift);  // may be inf
  exp_d[i] = e;
  if (kNonUni)
    alpha[i] = scalar_t(1) - pow(scalar_t(1) + e, -interval_pts[i]);
  else
    alpha[i] = scalar_t(1) - pow(scalar_t(1) + e, static_cast<scalar_t>(-interval));
}

// reference :513-516 / :526-529: min(e, 1e10) is a double, so the whole product is evaluated in
// double and narrowed once.
template <typename scalar_t, bool kNonUni>
__global__ void raw2alpha_bwd_kernel(const scalar_t* __restrict__ exp_d,
                                     const scalar_t* __restrict__ grad_back, float interval,
                                     const scalar_t* __restrict__ interval_pts, int64_t n,
                                     scalar_t* __restrict__ grad) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const scalar_t e = exp_d[i];
  const double capped = fmin(static_cast<double>(e), 1e10);
  double g;
  if (kNonUni) {
    const scalar_t iv = interval_pts[i];
    const scalar_t p = pow(scalar_t(1) + e, -iv - scalar_t(1));
    g = capped * static_cast<double>(p) * static_cast<double>(iv) * static_cast<double>(grad_back[i]);
  } else {
    const scalar_t p = pow(scalar_t(1) + e, static_cast<scalar_t>(-interval - 1.f));
    g = capped * static_cast<double>(p) * static_cast<double>(interval) * static_cast<double>(grad_back[i]);
  }
  grad[i] = static_cast<scalar_t>(g);
}

__global__ void neus_alpha_fwd_kernel(const float* __restrict__ sdf, const float* __restrict__ gradient,
                                      const float* __restrict__ viewdirs,
                                      const int64_t* __restrict__ ray_id, float dist, float inv_s,
                                      int64_t n, float* __restrict__ alpha) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t r = ray_id[i];
  const float g[3] = {gradient[3 * i], gradient[3 * i + 1], gradient[3 * i + 2]};
  const float v[3] = {viewdirs[3 * r], viewdirs[3 * r + 1], viewdirs[3 * r + 2]};
  const NeusTerms t = neus_terms(sdf[i], g, v, dist, inv_s);
  alpha[i] = fminf(fmaxf(t.num / t.den, 0.f), 1.f);
}

__global__ void neus_alpha_bwd_kernel(const float* __restrict__ sdf, const float* __restrict__ gradient,
                                      const float* __restrict__ viewdirs,
                                      const int64_t* __restrict__ ray_id, float dist, float inv_s,
                                      int64_t n, const float* __restrict__ grad_alpha,
                                      float* __restrict__ grad_sdf, float* __restrict__ grad_gradient) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t r = ray_id[i];
  const float g[3] = {gradient[3 * i], gradient[3 * i + 1], gradient[3 * i + 2]};
  const float v[3] = {viewdirs[3 * r], viewdirs[3 * r + 1], viewdirs[3 * r + 2]};
  const NeusTerms t = neus_terms(sdf[i], g, v, dist, inv_s);
  const float ratio = t.num / t.den;
  // clamp passes gradient on the closed interval [0, 1] (torch.clamp backward)
  const float ga = (ratio >= 0.f && ratio <= 1.f) ? grad_alpha[i] : 0.f;
  const float gP = ga * (t.den - t.num) / (t.den * t.den);   // d(num/den)/dP with dnum/dP = dden/dP = 1
  const float gQ = -ga / t.den;
  const float gu = gP * t.P * (1.f - t.P);
  const float gv = gQ * t.Q * (1.f - t.Q);
  if (grad_sdf) grad_sdf[i] = (gu + gv) * inv_s;
  if (grad_gradient) {
    // prev = sdf - ic*dist/2, next = sdf + ic*dist/2, ic = -relu(-tc): d ic / d tc = [tc < 0]
    const float g_tc = (t.tc < 0.f) ? (gv - gu) * inv_s * dist * 0.5f : 0.f;
    grad_gradient[3 * i] = g_tc * v[0];
    grad_gradient[3 * i + 1] = g_tc * v[1];
    grad_gradient[3 * i + 2] = g_tc * v[2];
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

extern "C" int pp_raw2alpha(pp_dtype dtype, const void* density, float shift, float interval,
                            const void* interval_pts, int64_t n_pts, void* exp_d, void* alpha,
                            pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0, PP_ERR_INVALID_ARG, "raw2alpha: n_pts < 0");
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(density && exp_d && alpha, PP_ERR_INVALID_ARG, "raw2alpha: null pointer");
  const int blocks = ceil_div(n_pts, 256);
  PP_DISPATCH(dtype, {
    if (interval_pts)
      raw2alpha_kernel<scalar_t, true><<<blocks, 256, 0, as_stream(stream)>>>(
          (const scalar_t*)density, shift, interval, (const scalar_t*)interval_pts, n_pts,
          (scalar_t*)exp_d, (scalar_t*)alpha);
    else
      raw2alpha_kernel<scalar_t, false><<<blocks, 256, 0, as_stream(stream)>>>(
          (const scalar_t*)density, shift, interval, nullptr, n_pts, (scalar_t*)exp_d,
          (scalar_t*)alpha);
  });
  PP_CHECK_LAUNCH("raw2alpha");
  return PP_OK;
}

extern "C" int pp_raw2alpha_backward(pp_dtype dtype, const void* exp_d, const void* grad_back,
                                     float interval, const void* interval_pts, int64_t n_pts,
                                     void* grad, pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0, PP_ERR_INVALID_ARG, "raw2alpha_backward: n_pts < 0");
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(exp_d && grad_back && grad, PP_ERR_INVALID_ARG, "raw2alpha_backward: null pointer");
  const int blocks = ceil_div(n_pts, 256);
  PP_DISPATCH(dtype, {
    if (interval_pts)
      raw2alpha_bwd_kernel<scalar_t, true><<<blocks, 256, 0, as_stream(stream)>>>(
          (const scalar_t*)exp_d, (const scalar_t*)grad_back, interval,
          (const scalar_t*)interval_pts, n_pts, (scalar_t*)grad);
    else
      raw2alpha_bwd_kernel<scalar_t, false><<<blocks, 256, 0, as_stream(stream)>>>(
          (const scalar_t*)exp_d, (const scalar_t*)grad_back, interval, nullptr, n_pts,
          (scalar_t*)grad);
  });
  PP_CHECK_LAUNCH("raw2alpha_backward");
  return PP_OK;
}

extern "C" int pp_neus_alpha_forward(const float* sdf, const float* gradient, const float* viewdirs,
                                     const int64_t* ray_id, float dist, float inv_s, int64_t n_pts,
                                     float* alpha, pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0, PP_ERR_INVALID_ARG, "neus_alpha_forward: n_pts < 0");
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(sdf && gradient && viewdirs && ray_id && alpha, PP_ERR_INVALID_ARG,
             "neus_alpha_forward: null pointer");
  neus_alpha_fwd_kernel<<<ceil_div(n_pts, 256), 256, 0, as_stream(stream)>>>(
      sdf, gradient, viewdirs, ray_id, dist, inv_s, n_pts, alpha);
  PP_CHECK_LAUNCH("neus_alpha_forward");
  return PP_OK;
}

extern "C" int pp_neus_alpha_backward(const float* sdf, const float* gradient, const float* viewdirs,
                                      const int64_t* ray_id, float dist, float inv_s, int64_t n_pts,
                                      const float* grad_alpha, float* grad_sdf,
                                      float* grad_gradient, pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0, PP_ERR_INVALID_ARG, "neus_alpha_backward: n_pts < 0");
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(sdf && gradient && viewdirs && ray_id && grad_alpha, PP_ERR_INVALID_ARG,
             "neus_alpha_backward: null pointer");
  neus_alpha_bwd_kernel<<<ceil_div(n_pts, 256), 256, 0, as_stream(stream)>>>(
      sdf, gradient, viewdirs, ray_id, dist, inv_s, n_pts, grad_alpha, grad_sdf, grad_gradient);
  PP_CHECK_LAUNCH("neus_alpha_backward");
  return PP_OK;
}
