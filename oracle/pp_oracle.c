/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the poseprobe_b200 hot path.
 *
 * A plain-C (single thread, scalar loops) restatement of what the reference's CUDA kernels compute,
 * one function per exported operator, each citing the reference lines it follows
 * (paths relative to /root/reference/lib/cuda).  It exists so that tests can check the B200 kernels
 * against an independent implementation on a machine without the reference, and so that bench.py
 * can time a CPU baseline.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py (-m gpu) runs this file against the
 * reference's own kernels (oracle/_ref, compiled unmodified from /root/reference) on the same
 * inputs, and tests/golden/ holds outputs of those reference kernels generated on a B200
 * (tests/golden/make_golden.py) that the CPU suite replays against this file.
 *
 * Floating-point contract: fp32 storage.  The reference is built with nvcc's default -fmad=true;
 * where its SASS shows a fused multiply-add this file calls fmaf()/fma() explicitly and the file
 * is compiled with -ffp-contract=off so nothing else is fused.  exp/pow come from libm here and
 * from libdevice there (<= 2 ulp apart), so alpha values are compared with a tolerance while every
 * integer / bool output is compared exactly.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* float -> int32 the way the GPU's cvt.rzi.s32.f32 does it: truncate, saturate, NaN -> 0 */
static int32_t f2i_sat(float v) {
  if (v != v) return 0;
  if (v >= 2147483648.0f) return INT32_MAX;
  if (v <= -2147483648.0f) return INT32_MIN;
  return (int32_t)v;
}

static int64_t d2l_sat(double v) {
  if (v != v) return 0;
  if (v >= 9223372036854775808.0) return INT64_MAX;
  if (v <= -9223372036854775808.0) return INT64_MIN;
  return (int64_t)v;
}

/* |d| with the reference build's contraction: fma(dz,dz, fma(dx,dx, dy*dy))
 * (render_utils_kernel.cu:48-51, :68-71; SASS of oracle/_ref: FMUL y*y, FFMA x, FFMA z). */
static float ray_norm(const float* d) { return sqrtf(fmaf(d[2], d[2], fmaf(d[0], d[0], d[1] * d[1]))); }

/* render_utils_kernel.cu:12-35 */
ORC_API void orc_infer_t_minmax(const float* rays_o, const float* rays_d, const float* xyz_min,
                                const float* xyz_max, float near, float far, int64_t n_rays,
                                float* t_min, float* t_max) {
  for (int64_t r = 0; r < n_rays; ++r) {
    float a[3], b[3];
    for (int c = 0; c < 3; ++c) {
      const float d = rays_d[3 * r + c];
      const float v = (d == 0) ? (float)1e-6 : d;
      a[c] = (xyz_max[c] - rays_o[3 * r + c]) / v;
      b[c] = (xyz_min[c] - rays_o[3 * r + c]) / v;
    }
    const float enter = fmaxf(fmaxf(fminf(a[0], b[0]), fminf(a[1], b[1])), fminf(a[2], b[2]));
    const float leave = fminf(fminf(fmaxf(a[0], b[0]), fmaxf(a[1], b[1])), fmaxf(a[2], b[2]));
    t_min[r] = fmaxf(fminf(enter, far), near);
    t_max[r] = fmaxf(fminf(leave, far), near);
  }
}

/* render_utils_kernel.cu:38-55 */
ORC_API void orc_infer_n_samples(const float* rays_d, const float* t_min, const float* t_max,
                                 float stepdist, int64_t n_rays, int64_t* n_samples) {
  for (int64_t r = 0; r < n_rays; ++r) {
    const float rnorm = ray_norm(rays_d + 3 * r);
    const float steps = ceilf((t_max[r] - t_min[r]) * rnorm / stepdist);
    n_samples[r] = d2l_sat(fmax((double)steps, 1.));
  }
}

/* render_utils_kernel.cu:58-79 */
ORC_API void orc_infer_ray_start_dir(const float* rays_o, const float* rays_d, const float* t_min,
                                     int64_t n_rays, float* rays_start, float* rays_dir) {
  for (int64_t r = 0; r < n_rays; ++r) {
    const float rnorm = ray_norm(rays_d + 3 * r);
    for (int c = 0; c < 3; ++c) {
      rays_start[3 * r + c] = fmaf(rays_d[3 * r + c], t_min[r], rays_o[3 * r + c]);
      rays_dir[3 * r + c] = rays_d[3 * r + c] / rnorm;
    }
  }
}

/* render_utils_kernel.cu:144-242, sample enumeration.  Call once with ray_id == NULL to obtain the
 * per-ray outputs and the return value M, then again with buffers of M entries. */
ORC_API int64_t orc_sample_pts_on_rays(const float* rays_o, const float* rays_d, const float* xyz_min,
                                       const float* xyz_max, float near, float far, float stepdist,
                                       int64_t n_rays, float* t_min, float* t_max, int64_t* n_steps,
                                       float* rays_start, float* rays_dir, float* rays_pts,
                                       uint8_t* mask_outbbox, int64_t* ray_id, int64_t* step_id) {
  orc_infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far, n_rays, t_min, t_max);
  orc_infer_n_samples(rays_d, t_min, t_max, stepdist, n_rays, n_steps);
  orc_infer_ray_start_dir(rays_o, rays_d, t_min, n_rays, rays_start, rays_dir);
  int64_t total = 0;
  for (int64_t r = 0; r < n_rays; ++r) total += n_steps[r];
  if (!ray_id) return total;
  int64_t idx = 0;
  for (int64_t r = 0; r < n_rays; ++r) {
    for (int64_t i = 0; i < n_steps[r]; ++i, ++idx) {
      ray_id[idx] = r;   /* scatter-1 + cumsum of :144-152, :216 */
      step_id[idx] = i;  /* :154-164 */
      const float dist = stepdist * (float)(int32_t)i;  /* :184 */
      float p[3];
      for (int c = 0; c < 3; ++c) p[c] = fmaf(rays_dir[3 * r + c], dist, rays_start[3 * r + c]);
      memcpy(rays_pts + 3 * idx, p, sizeof p);
      mask_outbbox[idx] = (xyz_min[0] > p[0]) | (xyz_min[1] > p[1]) | (xyz_min[2] > p[2]) |
                          (xyz_max[0] < p[0]) | (xyz_max[1] < p[1]) | (xyz_max[2] < p[2]);
    }
  }
  return total;
}

/* render_utils_kernel.cu:245-270 */
ORC_API void orc_sample_ndc_pts_on_rays(const float* rays_o, const float* rays_d,
                                        const float* xyz_min, const float* xyz_max, int32_t n_samples,
                                        int64_t n_rays, float* rays_pts, uint8_t* mask_outbbox) {
  for (int64_t r = 0; r < n_rays; ++r)
    for (int32_t i = 0; i < n_samples; ++i) {
      const int64_t idx = r * n_samples + i;
      const float dist = (float)i / (float)(n_samples - 1);
      float p[3];
      for (int c = 0; c < 3; ++c) p[c] = fmaf(rays_d[3 * r + c], dist, rays_o[3 * r + c]);
      memcpy(rays_pts + 3 * idx, p, sizeof p);
      mask_outbbox[idx] = (xyz_min[0] > p[0]) | (xyz_min[1] > p[1]) | (xyz_min[2] > p[2]) |
                          (xyz_max[0] < p[0]) | (xyz_max[1] < p[1]) | (xyz_max[2] < p[2]);
    }
}

/* render_utils_kernel.cu:301-340 */
ORC_API void orc_sample_bg_pts_on_rays(const float* rays_o, const float* rays_d, const float* t_max,
                                       float bg_preserve, int32_t n_samples, int64_t n_rays,
                                       float* rays_pts) {
  for (int64_t r = 0; r < n_rays; ++r)
    for (int32_t i = 0; i < n_samples; ++i) {
      const int64_t idx = r * n_samples + i;
      const float frac = (float)i / (float)n_samples;
      const float t = (float)((double)t_max[r] - 1. + 1. / (1. - (double)frac));
      float q[3];
      for (int c = 0; c < 3; ++c) q[c] = fmaf(rays_d[3 * r + c], t, rays_o[3 * r + c]);
      const float t_outer = ray_norm(q);
      const float m = fmaxf(fabsf(q[0]), fmaxf(fabsf(q[1]), fabsf(q[2])));
      const float R = t_outer / m;
      const double lhs = (double)((R * R) / (t_outer * t_outer));
      const float rhs = (R / t_outer) * bg_preserve;
      const float o2i = (float)fma(lhs, 1. - (double)bg_preserve, (double)rhs);
      for (int c = 0; c < 3; ++c) rays_pts[3 * idx + c] = q[c] * o2i;
    }
}

/* render_utils_kernel.cu:367-424 */
ORC_API void orc_maskcache_lookup(const uint8_t* world, const float* xyz, const float* scale,
                                  const float* shift, int32_t sz_i, int32_t sz_j, int32_t sz_k,
                                  int64_t n_pts, uint8_t* out) {
  for (int64_t p = 0; p < n_pts; ++p) {
    const int32_t i = f2i_sat(roundf(fmaf(xyz[3 * p], scale[0], shift[0])));
    const int32_t j = f2i_sat(roundf(fmaf(xyz[3 * p + 1], scale[1], shift[1])));
    const int32_t k = f2i_sat(roundf(fmaf(xyz[3 * p + 2], scale[2], shift[2])));
    out[p] = 0; /* torch::zeros, :405 */
    if (0 <= i && i < sz_i && 0 <= j && j < sz_j && 0 <= k && k < sz_k)
      out[p] = world[((int64_t)i * sz_j + j) * sz_k + k];
  }
}

/* render_utils_kernel.cu:431-458; interval_pts != NULL is the nonuni variant */
ORC_API void orc_raw2alpha(const float* density, float shift, float interval,
                           const float* interval_pts, int64_t n, float* exp_d, float* alpha) {
  for (int64_t i = 0; i < n; ++i) {
    const float e = expf(density[i] + shift);
    exp_d[i] = e;
    alpha[i] = 1.f - powf(1.f + e, interval_pts ? -interval_pts[i] : -interval);
  }
}

/* render_utils_kernel.cu:507-530 */
ORC_API void orc_raw2alpha_backward(const float* exp_d, const float* grad_back, float interval,
                                    const float* interval_pts, int64_t n, float* grad) {
  for (int64_t i = 0; i < n; ++i) {
    const float iv = interval_pts ? interval_pts[i] : interval;
    const float p = powf(1.f + exp_d[i], -iv - 1.f);
    grad[i] = (float)(fmin((double)exp_d[i], 1e10) * (double)p * (double)iv * (double)grad_back[i]);
  }
}

/* render_utils_kernel.cu:577-651 */
ORC_API void orc_alpha2weight(const float* alpha, const int64_t* ray_id, int64_t n_pts, int64_t n_rays,
                              float* weight, float* T, float* alphainv_last, int64_t* i_start,
                              int64_t* i_end) {
  for (int64_t i = 0; i < n_pts; ++i) { weight[i] = 0.f; T[i] = 1.f; }           /* :624-625 */
  for (int64_t r = 0; r < n_rays; ++r) { alphainv_last[r] = 1.f; i_start[r] = i_end[r] = 0; }
  if (n_pts == 0) return;
  for (int64_t i = 1; i < n_pts; ++i)                                             /* :607-617 */
    if (ray_id[i] != ray_id[i - 1]) { i_start[ray_id[i]] = i; i_end[ray_id[i - 1]] = i; }
  i_end[ray_id[n_pts - 1]] = n_pts;                                               /* :635 */
  for (int64_t r = 0; r < n_rays; ++r) {                                          /* :586-604 */
    float T_cum = 1.f;
    int64_t i;
    for (i = i_start[r]; i < i_end[r]; ++i) {
      T[i] = T_cum;
      weight[i] = T_cum * alpha[i];
      T_cum = (float)((double)T_cum * (1. - (double)alpha[i]));
      if ((double)T_cum < 1e-3) { i += 1; break; }
    }
    i_end[r] = i;
    alphainv_last[r] = T_cum;
  }
}

/* render_utils_kernel.cu:654-707 */
ORC_API void orc_alpha2weight_backward(const float* alpha, const float* weight, const float* T,
                                       const float* alphainv_last, const int64_t* i_start,
                                       const int64_t* i_end, int64_t n_rays, int64_t n_pts,
                                       const float* grad_weights, const float* grad_last,
                                       float* grad) {
  for (int64_t i = 0; i < n_pts; ++i) grad[i] = 0.f;
  for (int64_t r = 0; r < n_rays; ++r) {
    float back = grad_last[r] * alphainv_last[r];
    for (int64_t i = i_end[r] - 1; i >= i_start[r]; --i) {
      const float gT = grad_weights[i] * T[i];
      const float om = 1.f - alpha[i];
      grad[i] = (float)((double)gT - (double)back / ((double)om + 1e-10));
      back = fmaf(grad_weights[i], weight[i], back);
    }
  }
}

/* total_variation_kernel.cu:13-133.  mask == NULL: old kernel (weights k,j,i = wz,wy,wz);
 * otherwise the _new kernel (wx,wy,wz and the mask products).  w* are the caller's values. */
static float clamp1(float v) { return fminf(fmaxf(v, -1.f), 1.f); }
ORC_API void orc_total_variation_add_grad(const float* param, float* grad, const float* mask,
                                          float wx, float wy, float wz, int dense_mode,
                                          int64_t n_planes, int64_t sz_i, int64_t sz_j, int64_t sz_k) {
  wx /= 6; wy /= 6; wz /= 6;                                                      /* :76-78 */
  const float wk = mask ? wx : wz, wj = wy, wi = wz;
  const int64_t sj = sz_k, si = sz_k * sz_j;
  for (int64_t pl = 0; pl < n_planes; ++pl)
    for (int64_t i = 0; i < sz_i; ++i)
      for (int64_t j = 0; j < sz_j; ++j)
        for (int64_t k = 0; k < sz_k; ++k) {
          const int64_t idx = ((pl * sz_i + i) * sz_j + j) * sz_k + k;
          if (!dense_mode && grad[idx] == 0) continue;
          const int edge[6] = {k == 0, k == sz_k - 1, j == 0, j == sz_j - 1, i == 0, i == sz_i - 1};
          const int64_t off[6] = {-1, 1, -sj, sj, -si, si};
          const float w[6] = {wk, wk, wj, wj, wi, wi};
          float acc = 0.f;
          for (int q = 0; q < 6; ++q) {
            if (edge[q]) continue;
            float t = w[q] * clamp1(param[idx] - param[idx + off[q]]);
            if (mask) t = t * mask[idx] * mask[idx + off[q]];
            acc += t;
          }
          grad[idx] += acc;
        }
}

/* adam_upd_kernel.cu:8-132; mode 0 adam_upd, 1 masked_adam_upd, 2 adam_upd_with_perlr */
ORC_API void orc_adam_upd(int mode, float* param, const float* grad, float* exp_avg,
                          float* exp_avg_sq, const float* perlr, int64_t n, int32_t step,
                          float beta1, float beta2, float lr, float eps) {
  const float step_size = lr * sqrtf(1.f - powf(beta2, (float)step)) / (1.f - powf(beta1, (float)step));
  for (int64_t i = 0; i < n; ++i) {
    const float g = grad[i];
    if (mode == 1 && g == 0) continue;
    exp_avg[i] = fmaf(beta1, exp_avg[i], (1.f - beta1) * g);
    exp_avg_sq[i] = fmaf(beta2, exp_avg_sq[i], ((1.f - beta2) * g) * g);
    const float num = (mode == 2) ? (step_size * perlr[i]) * exp_avg[i] : step_size * exp_avg[i];
    param[i] = param[i] - num / (sqrtf(exp_avg_sq[i]) + eps);
  }
}

/* ub360_utils_kernel.cu:12-47 */
ORC_API void orc_cumdist_thres(const float* dist, float thres, int64_t n_rays, int64_t n_pts,
                               uint8_t* mask) {
  for (int64_t r = 0; r < n_rays; ++r) {
    float cum = 0.f;
    for (int64_t i = r * n_pts; i < (r + 1) * n_pts; ++i) {
      cum += dist[i];
      const int over = cum > thres;
      cum *= (float)(!over);
      mask[i] = (uint8_t)over;
    }
  }
}
