/* A host program in plain C on top of the C ABI: no Python, no torch, no C++.  It is what a maintainer binding the
 * library from another host language would write first: device buffers from the CUDA runtime, raw pointers and
 * extents into pp_* entry points, status codes out.  Checks two operators against values computed on the host:
 *   pp_infer_t_minmax  (render_utils_kernel.cu:20-34),  pp_alpha2weight (:577-651),  pp_dense_composite_forward
 *   (lib/dvgo_ori.py:478-485)  and  pp_nerf_composite_forward (lib/bg_nerf/source/models/frequency_nerf.py:283-343).
 * Build: gcc -O2 -I include -I /usr/local/cuda/include abi_example.c -L poseprobe_b200 -lposeprobe_b200 \
 *            -L /usr/local/cuda/lib64 -lcudart -lm -o abi_example
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "poseprobe_b200.h"
#include "poseprobe_b200_fused.h"

#define CK(x)                                                                 \
  do {                                                                        \
    cudaError_t e_ = (x);                                                     \
    if (e_ != cudaSuccess) {                                                  \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                \
      return 2;                                                               \
    }                                                                         \
  } while (0)
#define PP(x)                                                                 \
  do {                                                                        \
    int rc_ = (x);                                                            \
    if (rc_ != PP_OK) {                                                       \
      fprintf(stderr, "%s -> %d: %s\n", #x, rc_, pp_last_error());            \
      return 3;                                                               \
    }                                                                         \
  } while (0)

int main(void) {
  if (pp_abi_version() != 1) return 1;
  /* ---- ray / box intersection of 4 rays ---- */
  const float o[12] = {0, 0, -3, 0, 0, 3, -3, 0.5f, 0.25f, 5, 5, 5};
  const float d[12] = {0, 0, 1, 0, 0, -1, 1, 0, 0, 1, 0, 0};
  const float mn[3] = {-1, -1, -1}, mx[3] = {1, 1, 1};
  float *d_o, *d_d, *d_mn, *d_mx, *d_tmin, *d_tmax;
  CK(cudaMalloc((void**)&d_o, sizeof o));
  CK(cudaMalloc((void**)&d_d, sizeof d));
  CK(cudaMalloc((void**)&d_mn, sizeof mn));
  CK(cudaMalloc((void**)&d_mx, sizeof mx));
  CK(cudaMalloc((void**)&d_tmin, 4 * sizeof(float)));
  CK(cudaMalloc((void**)&d_tmax, 4 * sizeof(float)));
  CK(cudaMemcpy(d_o, o, sizeof o, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_d, d, sizeof d, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_mn, mn, sizeof mn, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_mx, mx, sizeof mx, cudaMemcpyHostToDevice));
  PP(pp_infer_t_minmax(PP_F32, d_o, d_d, d_mn, d_mx, 0.2f, 1e9f, 4, d_tmin, d_tmax, NULL));
  float tmin[4], tmax[4];
  CK(cudaMemcpy(tmin, d_tmin, sizeof tmin, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(tmax, d_tmax, sizeof tmax, cudaMemcpyDeviceToHost));
  const float want_min[3] = {2, 2, 2}, want_max[3] = {4, 4, 4};
  for (int r = 0; r < 3; ++r)
    if (tmin[r] != want_min[r] || tmax[r] != want_max[r]) {
      fprintf(stderr, "ray %d: t = [%g, %g]\n", r, tmin[r], tmax[r]);
      return 4;
    }
  if (!(tmax[3] <= tmin[3])) return 5; /* the fourth ray misses the box */

  /* ---- alpha compositing of 3 rays x 12 samples, alpha = 0.5: T_i = 0.5^i, stops once T < 1e-3 ---- */
  enum { R = 3, S = 12, M = R * S };
  float alpha[M];
  int64_t ray_id[M];
  for (int i = 0; i < M; ++i) {
    alpha[i] = 0.5f;
    ray_id[i] = i / S;
  }
  float *d_alpha, *d_w, *d_T, *d_last;
  int64_t *d_rid, *d_is, *d_ie;
  CK(cudaMalloc((void**)&d_alpha, sizeof alpha));
  CK(cudaMalloc((void**)&d_w, sizeof alpha));
  CK(cudaMalloc((void**)&d_T, sizeof alpha));
  CK(cudaMalloc((void**)&d_last, R * sizeof(float)));
  CK(cudaMalloc((void**)&d_rid, sizeof ray_id));
  CK(cudaMalloc((void**)&d_is, R * sizeof(int64_t)));
  CK(cudaMalloc((void**)&d_ie, R * sizeof(int64_t)));
  CK(cudaMemcpy(d_alpha, alpha, sizeof alpha, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_rid, ray_id, sizeof ray_id, cudaMemcpyHostToDevice));
  PP(pp_alpha2weight(PP_F32, d_alpha, d_rid, M, R, d_w, d_T, d_last, d_is, d_ie, NULL));
  float w[M], last[R];
  int64_t ie[R], is[R];
  CK(cudaMemcpy(w, d_w, sizeof w, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(last, d_last, sizeof last, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(is, d_is, sizeof is, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ie, d_ie, sizeof ie, cudaMemcpyDeviceToHost));
  for (int r = 0; r < R; ++r) {
    /* 0.5^10 = 9.77e-4 < 1e-3 after the 10th sample: 10 samples composited, 2 keep weight 0 */
    if (is[r] != r * S || ie[r] != r * S + 10 || last[r] != ldexpf(1.f, -10)) {
      fprintf(stderr, "ray %d: i_start %lld i_end %lld last %g\n", r, (long long)is[r], (long long)ie[r], last[r]);
      return 6;
    }
    for (int i = 0; i < S; ++i) {
      const float want = i < 10 ? ldexpf(1.f, -(i + 1)) : 0.f;
      if (w[r * S + i] != want) {
        fprintf(stderr, "weight[%d,%d] = %g, want %g\n", r, i, w[r * S + i], want);
        return 7;
      }
    }
  }
  /* ---- dense compositing (no early termination): the same alphas as a [3, 12] array -> weights 0.5^(i+1) for ALL
   *      12 samples, alphainv_cum [3, 13] = 0.5^i ---- */
  float *d_dw, *d_cum;
  CK(cudaMalloc((void**)&d_dw, sizeof alpha));
  CK(cudaMalloc((void**)&d_cum, R * (S + 1) * sizeof(float)));
  PP(pp_dense_composite_forward(d_alpha, R, S, d_dw, d_cum, NULL));
  float dw[M], cum[R * (S + 1)];
  CK(cudaMemcpy(dw, d_dw, sizeof dw, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cum, d_cum, sizeof cum, cudaMemcpyDeviceToHost));
  for (int r = 0; r < R; ++r)
    for (int i = 0; i <= S; ++i) {
      if (cum[r * (S + 1) + i] != ldexpf(1.f, -i)) return 10;
      if (i < S && dw[r * S + i] != ldexpf(1.f, -(i + 1))) return 11;
    }

  /* ---- scene branch: one ray of unit length, 4 samples at depths 1, 2, 3, 4 with density ln 2 and colour
   *      (1, 0.5, 0.25): every unit interval halves the transmittance, the last interval (1e10) absorbs the rest:
   *      weights 1/2, 1/4, 1/8, 1/8; opacity 1; rgb = colour; depth = 1/2 + 2/4 + 3/8 + 4/8 = 1.875 ---- */
  enum { NS = 4 };
  const float ray1[3] = {0, 0, 1};
  const float depth_s[NS] = {1, 2, 3, 4};
  float dens_s[NS], rgb_s[NS * 3];
  for (int i = 0; i < NS; ++i) {
    dens_s[i] = 0.69314718056f;
    rgb_s[3 * i] = 1.f;
    rgb_s[3 * i + 1] = 0.5f;
    rgb_s[3 * i + 2] = 0.25f;
  }
  float *d_ray, *d_rgbs, *d_dens, *d_dep, *d_out;   /* d_out: rgb 3 | depth 1 | opacity 1 | weights 4 */
  CK(cudaMalloc((void**)&d_ray, sizeof ray1));
  CK(cudaMalloc((void**)&d_rgbs, sizeof rgb_s));
  CK(cudaMalloc((void**)&d_dens, sizeof dens_s));
  CK(cudaMalloc((void**)&d_dep, sizeof depth_s));
  CK(cudaMalloc((void**)&d_out, 9 * sizeof(float)));
  CK(cudaMemcpy(d_ray, ray1, sizeof ray1, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_rgbs, rgb_s, sizeof rgb_s, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_dens, dens_s, sizeof dens_s, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_dep, depth_s, sizeof depth_s, cudaMemcpyHostToDevice));
  PP(pp_nerf_composite_forward(d_ray, d_rgbs, d_dens, d_dep, 1, NS, 0, d_out, NULL, d_out + 3, NULL, d_out + 4,
                               d_out + 5, NULL, NULL));
  float res[9];
  CK(cudaMemcpy(res, d_out, sizeof res, cudaMemcpyDeviceToHost));
  const float want_s[9] = {1.f, 0.5f, 0.25f, 1.875f, 1.f, 0.5f, 0.25f, 0.125f, 0.125f};
  for (int i = 0; i < 9; ++i)
    if (fabsf(res[i] - want_s[i]) > 2e-6f) {
      fprintf(stderr, "nerf composite [%d] = %.8g, want %.8g\n", i, res[i], want_s[i]);
      return 12;
    }

  /* ---- error convention: bad arguments return a status and a message, nothing is thrown ---- */
  if (pp_alpha2weight(PP_F32, NULL, d_rid, M, R, d_w, d_T, d_last, d_is, d_ie, NULL) == PP_OK) return 8;
  if (!pp_last_error() || !pp_last_error()[0]) return 9;
  CK(cudaDeviceSynchronize());
  printf("c abi example ok (launches: %lld)\n", (long long)pp_launch_count());
  return 0;
}
