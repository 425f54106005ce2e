// NeuS mid-point SDF -> alpha terms shared by the stand-alone operator (alpha.cu) and the fused fine-stage
// kernels (voxurf_fused.cu).
#pragma once
#include "common.cuh"

namespace pp {

// ---- NeuS SDF -> alpha (voxurf_coarse.py:483-519, use_mid branch, cos_anneal_ratio = 1) ------------
__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.f / (1.f + expf(-x)); }

struct NeusTerms {
  float tc;          // viewdir . gradient
  float P, Q;        // prev_cdf, next_cdf
  float num, den;    // (P - Q + 1e-5), (P + 1e-5)
};

__device__ __forceinline__ NeusTerms neus_terms(float sdf, const float g[3], const float v[3],
                                                float dist, float inv_s) {
  NeusTerms t;
  // torch: (dirs * gradients).sum(-1) - three rounded products, summed left to right
  t.tc = add_rn(add_rn(mul_rn(v[0], g[0]), mul_rn(v[1], g[1])), mul_rn(v[2], g[2]));
  const float iter_cos = -fmaxf(-t.tc, 0.f);
  const float half = mul_rn(mul_rn(iter_cos, dist), 0.5f);
  const float next = add_rn(sdf, half);
  const float prev = add_rn(sdf, -half);
  t.P = sigmoidf_ref(mul_rn(prev, inv_s));
  t.Q = sigmoidf_ref(mul_rn(next, inv_s));
  t.num = add_rn(add_rn(t.P, -t.Q), 1e-5f);
  t.den = add_rn(t.P, 1e-5f);
  return t;
}

// gradient of alpha = clamp(num/den, 0, 1) w.r.t. (sdf, tc): returns d alpha/d sdf in g_sdf and
// d alpha / d tc in g_tc, both already multiplied by the incoming ga
__device__ __forceinline__ void neus_backward(const NeusTerms& t, float ga, float dist, float inv_s, float& g_sdf,
                                              float& g_tc) {
  const float ratio = t.num / t.den;
  ga = (ratio >= 0.f && ratio <= 1.f) ? ga : 0.f;   // torch.clamp passes gradient on the closed interval
  const float gP = ga * (t.den - t.num) / (t.den * t.den);
  const float gQ = -ga / t.den;
  const float gu = gP * t.P * (1.f - t.P);
  const float gv = gQ * t.Q * (1.f - t.Q);
  g_sdf = (gu + gv) * inv_s;
  g_tc = (t.tc < 0.f) ? (gv - gu) * inv_s * dist * 0.5f : 0.f;
}

}  // namespace pp
