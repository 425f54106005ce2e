// Fused DVGO render / train-step kernels (see include/poseprobe_b200_fused.h).
//
// One warp marches one ray; the 32 lanes are 32 consecutive steps of the march.
//
//   gather    : every in-box sample reads its 8 corner voxel RECORDS (float4 = density + 3 colour
//               channels) with 8 x LDG.128, so one request returns everything both interpolations need;
//               consecutive lanes are consecutive steps of the same ray, i.e. neighbouring voxels.
//   composite : transmittance is an inclusive product scan over the lanes (5 shuffle steps) carried
//               from chunk to chunk; early termination (T < 1e-3 after a sample, alpha2weight's rule,
//               render_utils_kernel.cu:597) is a ballot.
//   backward  : alpha and T of the march are staged in shared memory (8 B per step), then the ray is
//               walked in reverse with a suffix scan for alpha2weight_backward's recurrence
//               (render_utils_kernel.cu:671-675); the voxel gradient is scattered with one 128-bit
//               red.global.add.v4.f32 per corner instead of 4 scalar atomics.
//   loss      : in train mode the photometric MSE (lib/losses.py:26-29), the background entropy and the
//               per-point colour term are evaluated in the same kernel, so the forward never leaves
//               the SM.
//
// Sample positions, step counts and the in-box test are computed with exactly the arithmetic of the
// stand-alone sampler (ray_math.cuh), so the set of samples is identical to sample_pts_on_rays + the
// `~mask_outbbox` compaction of the reference's Python wrapper (voxurf_coarse.py:686-693).
#include <stdlib.h>

#include <atomic>

#include "march_common.cuh"

namespace pp {

enum { kModeForward = 0, kModeBackward = 1, kModeTrain = 2 };

// raw2alpha (reference :437-442): alpha = 1 - (1+e)^(-interval), e = exp(sigma + shift).  The power is
// evaluated as exp2(-interval * log2(1+e)) with the accurate log2f/exp2f (a third of powf's
// instructions); it agrees with powf to ~3e-7 in the power, the same distance libm's powf has.
__device__ __forceinline__ float density_to_alpha(float sigma, float shift, float interval, float& e) {
  e = expf(sigma + shift);
  return 1.f - exp2f(-interval * log2f(1.f + e));
}

// rays are handed out through a device-side queue so that long and short rays balance across warps
__device__ unsigned int g_ray_queue[kQueueSlots];

// dynamic shared memory: per warp, s_alpha[cap] (alpha of every step, 0 where nothing was composited)
// followed by s_tin[cap/32] (transmittance entering each 32-step chunk)
template <int kMode, bool kPoseGrad, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks)
dvgo_march_kernel(pp_dvgo_scene sc, pp_dvgo_loss ls, const float4* __restrict__ vox,
                  const float* __restrict__ rays_o, const float* __restrict__ rays_d, int64_t n_rays,
                  const float* __restrict__ target, const float* __restrict__ grad_rgb_in,
                  const float* __restrict__ grad_last_in, float4* __restrict__ grad_vox,
                  float* __restrict__ rgb_marched, float* __restrict__ alphainv_last,
                  float* __restrict__ depth, int32_t* __restrict__ n_marched,
                  float* __restrict__ loss_out, float* __restrict__ grad_rays, int cap,
                  unsigned int* __restrict__ queue, float4* __restrict__ stash_base, int flags) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_warp = cap + cap / kWarp;
  float* s_alpha = smem + (size_t)warp * per_warp;
  float* s_tin = s_alpha + cap;
  // per-warp slice of the caller's workspace: (d sigma / d alpha, r, g, b) of every step of the ray in
  // flight, written coalesced in pass 1 and read back coalesced in pass 2 (it stays L2 resident), so the
  // reverse walk does not gather the 8 corner records again
  float4* stash = stash_base ? stash_base + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * cap : nullptr;

  for (;;) {
    unsigned int ticket = 0;
    if (lane == 0) ticket = atomicAdd(queue, 1u);
    const int64_t r = __shfl_sync(kFull, ticket, 0);
    if (r >= n_rays) break;

    const Ray ray = setup_ray(sc, rays_o, rays_d, r);
    const int n_steps = (kMode == kModeForward) ? ray.n_steps : min(ray.n_steps, cap);

    // ---------------- pass 1: forward march ----------------
    float T_in = 1.f;
    float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_depth = 0.f, acc_per = 0.f;
    int n_live = 0;
    int end = n_steps;          // exclusive index of the last composited step (+1)
    float tgt[3] = {0.f, 0.f, 0.f};
    if (kMode == kModeTrain) load3(target, r, tgt);
    LaneSample cur = lane_sample(sc, ray, lane, n_steps);
    for (int base = 0; base < n_steps; base += kWarp) {
      const int i = base + lane;
      // software pipeline: the next chunk's cell is computed now and its records start moving towards L1
      LaneSample nxt = lane_sample(sc, ray, i + kWarp, n_steps);
      if ((flags & 1) && nxt.live) prefetch8(vox, nxt.c);
      float alpha = 0.f, e = 0.f;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cur.live) {
        float4 v[8];
        gather8(vox, cur.c, v);
        val = interp8(cur.c, v);
        alpha = density_to_alpha(val.x, sc.act_shift, sc.interval, e);
      }
      float T_before, T_after;
      if (flags & 2) {
        // exact mode (parity tests): raw2alpha's own expression and alpha2weight's recurrence in the reference's
        // ORDER and precision (render_utils_kernel.cu:437-442, :586-604: T *= (1. - alpha) in double, narrowed to
        // float every step), so the early-termination index is the operator chain's, bit for bit
        if (cur.live) alpha = 1.f - powf(1.f + e, -sc.interval);
        float T = T_in;
        T_before = T_in;
        T_after = T_in;
        for (int j = 0; j < 32; ++j) {
          const float aj = __shfl_sync(kFull, alpha, j);
          if (j == lane) T_before = T;
          T = static_cast<float>(static_cast<double>(T) * (1. - static_cast<double>(aj)));
          if (j == lane) T_after = T;
        }
      } else {
        float P, Pex;
        lane_product_scan(1.f - alpha, lane, P, Pex);
        T_before = T_in * Pex;
        T_after = T_in * P;
      }
      const unsigned term = __ballot_sync(kFull, cur.live && (T_after < 1e-3f));
      bool counted = cur.live;
      int last_lane = 31;
      if (term) {
        last_lane = __ffs(term) - 1;
        counted = cur.live && (lane <= last_lane);
      }
      if (!counted) alpha = 0.f;
      if (kMode != kModeForward) {
        if (i < n_steps) s_alpha[i] = alpha;
        if (lane == 0) s_tin[base / kWarp] = T_in;
      }
      const float w = T_before * alpha;
      float4 keep = make_float4(0.f, 0.f, 0.f, 0.f);
      if (counted)   // raw2alpha_backward's factor (reference :513-516), (1+e)^(-iv-1) = (1-alpha)/(1+e)
        keep.x = fminf(e, 1e10f) * ((1.f - alpha) / (1.f + e)) * sc.interval;
      if (counted && w > sc.fast_color_thres) {
        const float cr = sigmoid_ref(val.y), cg = sigmoid_ref(val.z), cb = sigmoid_ref(val.w);
        keep.y = cr;
        keep.z = cg;
        keep.w = cb;
        acc_r = fma_rn(w, cr, acc_r);
        acc_g = fma_rn(w, cg, acc_g);
        acc_b = fma_rn(w, cb, acc_b);
        if (kMode == kModeTrain && ls.w_rgbper != 0.f) {
          const float dr = cr - tgt[0], dg = cg - tgt[1], db = cb - tgt[2];
          acc_per += w * (dr * dr + dg * dg + db * db);
        }
      }
      if (counted) {
        acc_depth = fma_rn(w, sc.stepdist * static_cast<float>(i), acc_depth);
        n_live += 1;
      }
      if (kMode != kModeForward && i < n_steps) stash[i] = keep;
      T_in = __shfl_sync(kFull, T_after, last_lane);
      if (term) {
        end = base + last_lane + 1;
        break;
      }
      cur = nxt;
    }
    const float T_last = T_in;
    acc_r = warp_sum(acc_r);
    acc_g = warp_sum(acc_g);
    acc_b = warp_sum(acc_b);
    const float rgbm[3] = {fma_rn(T_last, sc.bg, acc_r), fma_rn(T_last, sc.bg, acc_g),
                           fma_rn(T_last, sc.bg, acc_b)};
    if (kMode != kModeBackward) {
      if (depth) acc_depth = warp_sum(acc_depth);
      if (n_marched) n_live = __reduce_add_sync(kFull, n_live);
      if (lane == 0) {
        if (rgb_marched) {
          rgb_marched[3 * r] = rgbm[0];
          rgb_marched[3 * r + 1] = rgbm[1];
          rgb_marched[3 * r + 2] = rgbm[2];
        }
        if (alphainv_last) alphainv_last[r] = T_last;
        if (depth) depth[r] = acc_depth;
        if (n_marched) n_marched[r] = n_live;
      }
    }
    if (kMode == kModeForward) continue;

    // ---------------- loss / incoming gradients ----------------
    float g[3];          // d loss / d rgb_marched
    float g_last;        // d loss / d alphainv_last (total, including the bg term of rgb_marched)
    float per_scale = 0.f;
    if (kMode == kModeTrain) {
      const float inv_n = 1.f / static_cast<float>(n_rays);
      float mse = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float diff = rgbm[c] - tgt[c];
        mse += diff * diff;
        g[c] = ls.w_rgb * 2.f * diff * inv_n * (1.f / 3.f);
      }
      g_last = (g[0] + g[1] + g[2]) * sc.bg;
      float loss = ls.w_rgb * mse * inv_n * (1.f / 3.f);
      if (ls.w_entropy_last != 0.f) {
        // -(p log p + (1-p) log(1-p)) on clamp(alphainv_last, 1e-6, 1-1e-6), mean over rays
        const float pc = fminf(fmaxf(T_last, 1e-6f), 1.f - 1e-6f);
        loss += ls.w_entropy_last * inv_n * -(pc * logf(pc) + (1.f - pc) * logf(1.f - pc));
        if (T_last >= 1e-6f && T_last <= 1.f - 1e-6f)
          g_last += ls.w_entropy_last * inv_n * -(logf(pc) - logf(1.f - pc));
      }
      if (ls.w_rgbper != 0.f) {
        acc_per = warp_sum(acc_per);
        loss += ls.w_rgbper * inv_n * acc_per;
        per_scale = ls.w_rgbper * inv_n * 2.f;
      }
      if (lane == 0 && loss_out) {
        atomicAdd(loss_out, loss);
        atomicAdd(loss_out + 1, ls.mse_scale * mse * inv_n * (1.f / 3.f));
      }
    } else {
      load3(grad_rgb_in, r, g);
      g_last = (g[0] + g[1] + g[2]) * sc.bg;
      if (grad_last_in) g_last += grad_last_in[r];
    }
    __syncwarp();

    // ---------------- pass 2: reverse walk ----------------
    float back = g_last * T_last;   // back_cum of alpha2weight_backward (reference :671)
    float G0[3] = {0.f, 0.f, 0.f}, G1[3] = {0.f, 0.f, 0.f};
    const int last_base = ((end - 1) / kWarp) * kWarp;
    cur = lane_sample(sc, ray, last_base + lane, end);
    float4 kept = (last_base + lane < end) ? stash[last_base + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = last_base; base >= 0; base -= kWarp) {
      const int i = base + lane;
      const LaneSample nxt = lane_sample(sc, ray, i - kWarp, end);
      const float4 kept_nxt = (i - kWarp >= 0) ? stash[i - kWarp] : make_float4(0.f, 0.f, 0.f, 0.f);
      // transmittance of this chunk, recomputed exactly as in pass 1 from the staged alphas
      const float alpha = (i < end) ? s_alpha[i] : 0.f;
      float P, Pex;
      lane_product_scan(1.f - alpha, lane, P, Pex);
      const float T = s_tin[base / kWarp] * Pex;
      const float w = cur.live ? T * alpha : 0.f;
      const bool col = w > sc.fast_color_thres;
      const float rgb[3] = {kept.y, kept.z, kept.w};
      const float gw = col ? g[0] * rgb[0] + g[1] * rgb[1] + g[2] * rgb[2] : 0.f;
      // exclusive suffix sum of gw*w over the lanes above me
      const float contrib = gw * w;
      float S = contrib;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float dn = __shfl_down_sync(kFull, S, o);
        if (lane + o < 32) S += dn;
      }
      const float back_i = back + (S - contrib);
      back += __shfl_sync(kFull, S, 0);
      // per-sample gradient wrt the interpolated record (density, k0_r, k0_g, k0_b)
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cur.live) {
        const float d_alpha = gw * T - back_i / (1.f - alpha + 1e-10f);
        d.x = kept.x * d_alpha;
        if (col) {
          float gc[3];
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            gc[ch] = w * g[ch];
            if (kMode == kModeTrain) gc[ch] += per_scale * w * (rgb[ch] - tgt[ch]);
            gc[ch] *= rgb[ch] * (1.f - rgb[ch]);
          }
          d.y = gc[0];
          d.z = gc[1];
          d.w = gc[2];
        }
      }
      // scatter.  Neighbouring lanes are neighbouring steps (half a voxel apart) and often sit in the
      // same cell: such pairs are merged with one shuffle so the pair issues 8 reductions instead of 16.
      const Cell& c = cur.c;
      const int key = cur.live ? c.base : (-1 - lane);
      const int key_up = __shfl_down_sync(kFull, key, 1);
      const int key_dn = __shfl_up_sync(kFull, key, 1);
      const unsigned heads = __ballot_sync(kFull, lane == 0 || key != key_dn);
      const int run_start = 31 - __clz(heads & ((2u << lane) - 1u));
      const bool even_pos = ((lane - run_start) & 1) == 0;
      const bool absorb = (flags & 4) == 0 && cur.live && lane < 31 && key_up == key && even_pos;
      const bool absorbed = (flags & 4) == 0 && cur.live && lane > 0 && key_dn == key && !even_pos;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float wq = cur.live ? corner_w(c, q) : 0.f;
        float a0 = wq * d.x, a1 = wq * d.y, a2 = wq * d.z, a3 = wq * d.w;
        const float b0 = __shfl_down_sync(kFull, a0, 1), b1 = __shfl_down_sync(kFull, a1, 1),
                    b2 = __shfl_down_sync(kFull, a2, 1), b3 = __shfl_down_sync(kFull, a3, 1);
        if (absorb) {
          a0 += b0;
          a1 += b1;
          a2 += b2;
          a3 += b3;
        }
        if (cur.live && !absorbed && !(flags & 2) && (a0 != 0.f || a1 != 0.f || a2 != 0.f || a3 != 0.f))
          red_add_v4(grad_vox + corner_off(c, q), a0, a1, a2, a3);
      }
      if (kPoseGrad && cur.live) {
        // d loss / d point = sum_ch d_ch * grad_xyz(interp_ch); voxel units -> world units below
        float4 v[8];
        gather8(vox, c, v);
        float gp[3] = {0.f, 0.f, 0.f};
        const float dch[4] = {d.x, d.y, d.z, d.w};
        const float fx = c.fx, fy = c.fy, fz = c.fz, gx = c.gx(), gy = c.gy(), gz = c.gz();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float u[8];
#pragma unroll
          for (int q = 0; q < 8; ++q)
            u[q] = ch == 0 ? v[q].x : ch == 1 ? v[q].y : ch == 2 ? v[q].z : v[q].w;
          const float dX = ((u[4] - u[0]) * gz + (u[5] - u[1]) * fz) * gy +
                           ((u[6] - u[2]) * gz + (u[7] - u[3]) * fz) * fy;
          const float dY = ((u[2] - u[0]) * gz + (u[3] - u[1]) * fz) * gx +
                           ((u[6] - u[4]) * gz + (u[7] - u[5]) * fz) * fx;
          const float dZ = ((u[1] - u[0]) * gy + (u[3] - u[2]) * fy) * gx +
                           ((u[5] - u[4]) * gy + (u[7] - u[6]) * fy) * fx;
          gp[0] += dch[ch] * dX;
          gp[1] += dch[ch] * dY;
          gp[2] += dch[ch] * dZ;
        }
        gp[0] *= (sc.sx - 1) / (sc.xyz_max[0] - sc.xyz_min[0]);
        gp[1] *= (sc.sy - 1) / (sc.xyz_max[1] - sc.xyz_min[1]);
        gp[2] *= (sc.sz - 1) / (sc.xyz_max[2] - sc.xyz_min[2]);
        const float sdist = sc.stepdist * static_cast<float>(i);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          G0[k] += gp[k];
          G1[k] += gp[k] * sdist;
        }
      }
      cur = nxt;
      kept = kept_nxt;
    }
    if (kPoseGrad) {
      // point_i = o + d*t_min + (d/|d|) * s_i.  t_min is treated as locally constant in (o, d) except
      // through the entering slab: t_min = (face - o_a)/d_a  =>  dt/do_a = -1/d_a, dt/dd_a = -t_min/d_a.
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        G0[k] = warp_sum(G0[k]);
        G1[k] = warp_sum(G1[k]);
      }
      if (lane == 0) {
        float go[3] = {G0[0], G0[1], G0[2]};
        float gd[3] = {G0[0] * ray.t_min, G0[1] * ray.t_min, G0[2] * ray.t_min};
        int axis = -1;
        float best = -INFINITY;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float vv = (ray.d[k] == 0.f) ? 1e-6f : ray.d[k];
          const float a = (sc.xyz_max[k] - ray.o[k]) / vv;
          const float b = (sc.xyz_min[k] - ray.o[k]) / vv;
          const float m = fminf(a, b);
          if (m > best) { best = m; axis = k; }
        }
        if (best == ray.t_min && ray.d[axis] != 0.f) {   // not clamped by near / far
          const float gdot = G0[0] * ray.d[0] + G0[1] * ray.d[1] + G0[2] * ray.d[2];
          go[axis] += gdot * (-1.f / ray.d[axis]);
          gd[axis] += gdot * (-ray.t_min / ray.d[axis]);
        }
        const float inv = 1.f / ray.rnorm;
        const float dotg = (G1[0] * ray.d[0] + G1[1] * ray.d[1] + G1[2] * ray.d[2]) * inv * inv * inv;
#pragma unroll
        for (int k = 0; k < 3; ++k) gd[k] += G1[k] * inv - ray.d[k] * dotg;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          grad_rays[6 * r + k] = go[k];
          grad_rays[6 * r + 3 + k] = gd[k];
        }
      }
    }
    __syncwarp();
  }
}

// ---- pack / unpack ---------------------------------------------------------------------------------
__global__ void voxels_pack_kernel(const float* __restrict__ density, const float* __restrict__ k0,
                                   int64_t n, float4* __restrict__ vox) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    vox[i] = make_float4(density[i], k0[i], k0[n + i], k0[2 * n + i]);
}

__global__ void voxels_unpack_kernel(const float4* __restrict__ vox, int64_t n,
                                     float* __restrict__ density, float* __restrict__ k0) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = vox[i];
    if (density) density[i] = v.x;
    if (k0) {
      k0[i] = v.y;
      k0[n + i] = v.z;
      k0[2 * n + i] = v.w;
    }
  }
}

// ---- optimiser -------------------------------------------------------------------------------------
// masked_adam_upd per element (adam_upd_kernel.cu:35-39) with the reference build's contraction
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float step_size,
                                          float b1, float b2, float eps, bool dense = false) {
  if (!dense && g == 0.f) return;
  m = fma_rn(b1, m, mul_rn(1.f - b1, g));
  v = fma_rn(b2, v, mul_rn(mul_rn(1.f - b2, g), g));
  p = p - mul_rn(step_size, m) / (sqrtf(v) + eps);
}

// Streaming pass over 4 arrays of float4 records.  All four loads of a record are issued up front
// (independent, 64 B in flight per thread per record, two records per iteration) instead of making the
// state loads wait for the gradient test; stores happen only for records whose gradient is non-zero.
__global__ void __launch_bounds__(256)
voxel_adam_kernel(float4* __restrict__ vox, float4* __restrict__ grad, float4* __restrict__ m4,
                  float4* __restrict__ v4, int64_t n, float ss_density, float ss_k0, float b1, float b2,
                  float eps, int flags) {
  const bool zero_grad = flags & 1, dense = flags & 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < n; i0 += 2 * stride) {
    const int64_t idx[2] = {i0, i0 + stride};
    float4 g[2], p[2], m[2], v[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (idx[u] < n) {
        g[u] = grad[idx[u]];
        p[u] = vox[idx[u]];
        m[u] = m4[idx[u]];
        v[u] = v4[idx[u]];
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (idx[u] >= n) continue;
      if (!dense && g[u].x == 0.f && g[u].y == 0.f && g[u].z == 0.f && g[u].w == 0.f) continue;
      adam_elem(p[u].x, g[u].x, m[u].x, v[u].x, ss_density, b1, b2, eps, dense);
      adam_elem(p[u].y, g[u].y, m[u].y, v[u].y, ss_k0, b1, b2, eps, dense);
      adam_elem(p[u].z, g[u].z, m[u].z, v[u].z, ss_k0, b1, b2, eps, dense);
      adam_elem(p[u].w, g[u].w, m[u].w, v[u].w, ss_k0, b1, b2, eps, dense);
      vox[idx[u]] = p[u];
      m4[idx[u]] = m[u];
      v4[idx[u]] = v[u];
      if (zero_grad) grad[idx[u]] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

__device__ __forceinline__ float clamp_pm1(float x) { return fminf(fmaxf(x, -1.f), 1.f); }

// TV (old kernel, weights k,j,i = wz,wy,wz with wx=wy=wz=w here) + masked Adam, ping-pong.
// blockDim (32, 8): x along Z, y along Y; gridDim.z strides X.
__global__ void __launch_bounds__(256)
voxel_tv_adam_kernel(const float4* __restrict__ vin, float4* __restrict__ vout,
                     float4* __restrict__ grad, float4* __restrict__ m4, float4* __restrict__ v4, int sx,
                     int sy, int sz, float ss_density, float ss_k0, float b1, float b2, float eps,
                     float tvw_density, float tvw_k0, int dense_mode) {
  const int k = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (k >= sz || j >= sy) return;
  for (int i = blockIdx.z; i < sx; i += gridDim.z) {
    const int64_t idx = ((int64_t)i * sy + j) * sz + k;
    float4 g = grad[idx];
    const float4 p = vin[idx];
    // the reference applies TV per element where that element's gradient is non-zero (sparse mode)
    const bool tx = dense_mode || g.x != 0.f, ty = dense_mode || g.y != 0.f,
               tz = dense_mode || g.z != 0.f, tw = dense_mode || g.w != 0.f;
    if ((tvw_density != 0.f || tvw_k0 != 0.f) && (tx || ty || tz || tw)) {
      float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;
      auto term = [&](bool edge, int64_t off) {
        if (edge) return;
        const float4 q = vin[idx + off];
        ax = add_rn(ax, mul_rn(tvw_density, clamp_pm1(p.x - q.x)));
        ay = add_rn(ay, mul_rn(tvw_k0, clamp_pm1(p.y - q.y)));
        az = add_rn(az, mul_rn(tvw_k0, clamp_pm1(p.z - q.z)));
        aw = add_rn(aw, mul_rn(tvw_k0, clamp_pm1(p.w - q.w)));
      };
      term(k == 0, -1);
      term(k == sz - 1, 1);
      term(j == 0, -sz);
      term(j == sy - 1, sz);
      term(i == 0, -(int64_t)sz * sy);
      term(i == sx - 1, (int64_t)sz * sy);
      if (tx && tvw_density != 0.f) g.x = add_rn(g.x, ax);
      if (ty && tvw_k0 != 0.f) g.y = add_rn(g.y, ay);
      if (tz && tvw_k0 != 0.f) g.z = add_rn(g.z, az);
      if (tw && tvw_k0 != 0.f) g.w = add_rn(g.w, aw);
    }
    float4 pn = p;
    if (g.x != 0.f || g.y != 0.f || g.z != 0.f || g.w != 0.f) {
      float4 m = m4[idx], v = v4[idx];
      adam_elem(pn.x, g.x, m.x, v.x, ss_density, b1, b2, eps);
      adam_elem(pn.y, g.y, m.y, v.y, ss_k0, b1, b2, eps);
      adam_elem(pn.z, g.z, m.z, v.z, ss_k0, b1, b2, eps);
      adam_elem(pn.w, g.w, m.w, v.w, ss_k0, b1, b2, eps);
      m4[idx] = m;
      v4[idx] = v;
      grad[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    vout[idx] = pn;
  }
}

// ---- host side -------------------------------------------------------------------------------------
static int check_scene(const pp_dvgo_scene* sc, const char* who) {
  PP_REQUIRE(sc, PP_ERR_INVALID_ARG, "%s: scene is null", who);
  PP_REQUIRE(sc->sx > 0 && sc->sy > 0 && sc->sz > 0, PP_ERR_INVALID_ARG, "%s: empty grid", who);
  PP_REQUIRE((int64_t)sc->sx * sc->sy * sc->sz < (1LL << 31), PP_ERR_TOO_LARGE, "%s: grid too large", who);
  PP_REQUIRE(sc->stepdist > 0.f, PP_ERR_INVALID_ARG, "%s: stepdist must be positive", who);
  for (int c = 0; c < 3; ++c)
    PP_REQUIRE(sc->xyz_max[c] > sc->xyz_min[c], PP_ERR_INVALID_ARG, "%s: empty bounding box", who);
  return PP_OK;
}

// upper bound of N_steps for any ray: the in-box chord is at most the box diagonal
static int step_capacity(const pp_dvgo_scene* sc) {
  double d2 = 0;
  for (int c = 0; c < 3; ++c) {
    const double e = (double)sc->xyz_max[c] - (double)sc->xyz_min[c];
    d2 += e * e;
  }
  const int n = (int)ceil(sqrt(d2) / (double)sc->stepdist) + 3;
  return (n + 31) / 32 * 32;
}

static int tune_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// a __device__ symbol has a different address on every device: resolved per device (like sm_count)
static int ray_queue_base(unsigned int** out, const char* who) {
  static unsigned int* queue_bases[64] = {nullptr};
  int dev = 0;
  PP_CHECK_CUDA(cudaGetDevice(&dev));
  PP_REQUIRE(dev >= 0 && dev < 64, PP_ERR_UNSUPPORTED, "%s: device ordinal %d out of range", who, dev);
  if (!queue_bases[dev]) PP_CHECK_CUDA(cudaGetSymbolAddress((void**)&queue_bases[dev], g_ray_queue));
  *out = queue_bases[dev];
  return PP_OK;
}

template <int kMode, bool kPoseGrad, int kMinBlocks>
static int launch_march_occ(const pp_dvgo_scene* sc, const pp_dvgo_loss* ls, const void* vox,
                        const float* rays_o, const float* rays_d, int64_t n_rays, const float* target,
                        const float* grad_rgb, const float* grad_last, void* grad_vox,
                        float* rgb_marched, float* alphainv_last, float* depth, int32_t* n_marched,
                        float* loss_out, float* grad_rays, void* workspace, int64_t workspace_bytes,
                        cudaStream_t s, const char* who) {
  int warps = 8;
  int cap = 0;
  size_t smem = 0;
  if (kMode != kModeForward) {
    cap = step_capacity(sc);
    const size_t per_warp = (size_t)(cap + cap / kWarp) * sizeof(float);
    smem = warps * per_warp;
    while (smem > 200 * 1024 && warps > 1) {
      warps /= 2;
      smem = warps * per_warp;
    }
    PP_REQUIRE(smem <= 200 * 1024, PP_ERR_UNSUPPORTED,
               "%s: %d steps per ray exceed the shared-memory staging capacity", who, cap);
  }
  auto kernel = dvgo_march_kernel<kMode, kPoseGrad, kMinBlocks>;
  if (smem > 48 * 1024)
    PP_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // persistent grid: every resident warp pulls rays from a device-side queue
  static std::atomic<unsigned> next_slot{0};
  unsigned int* queue_base = nullptr;
  const int qrc = ray_queue_base(&queue_base, who);
  if (qrc != PP_OK) return qrc;
  unsigned int* queue = queue_base + (next_slot.fetch_add(1) % kQueueSlots);
  PP_CHECK_CUDA(cudaMemsetAsync(queue, 0, sizeof(unsigned int), s));
  int per_sm = 0;
  PP_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t blocks_needed = (n_rays + warps - 1) / warps;
  const int64_t blocks = min64(blocks_needed, (int64_t)sm_count() * per_sm);
  if (kMode != kModeForward) {
    const int64_t need = blocks * warps * (int64_t)cap * (int64_t)sizeof(float4);
    PP_REQUIRE(workspace && workspace_bytes >= need, PP_ERR_INVALID_ARG,
               "%s: workspace of %lld bytes required (pp_dvgo_workspace_bytes), got %lld", who,
               (long long)need, (long long)workspace_bytes);
  }
  pp_dvgo_loss l = {1.f, 0.f, 0.f, 1.f};
  if (ls) l = *ls;
  kernel<<<(unsigned)blocks, warps * 32, smem, s>>>(
      *sc, l, (const float4*)vox, rays_o, rays_d, n_rays, target, grad_rgb, grad_last,
      (float4*)grad_vox, rgb_marched, alphainv_last, depth, n_marched, loss_out, grad_rays, cap, queue,
      (float4*)workspace, tune_env("PP_MARCH_FLAGS", 0));
  PP_CHECK_LAUNCH(who);
  return PP_OK;
}

template <int kMode, bool kPoseGrad>
static int launch_march(const pp_dvgo_scene* sc, const pp_dvgo_loss* ls, const void* vox,
                        const float* rays_o, const float* rays_d, int64_t n_rays, const float* target,
                        const float* grad_rgb, const float* grad_last, void* grad_vox,
                        float* rgb_marched, float* alphainv_last, float* depth, int32_t* n_marched,
                        float* loss_out, float* grad_rays, void* workspace, int64_t workspace_bytes,
                        cudaStream_t s, const char* who) {
  // resident blocks per SM the kernel is compiled for (register budget 64 / 80 regs per thread)
  const int occ = kPoseGrad ? 2 : tune_env("PP_MARCH_OCC", 3);
#define PP_MARCH_ARGS sc, ls, vox, rays_o, rays_d, n_rays, target, grad_rgb, grad_last, grad_vox, rgb_marched, \
                      alphainv_last, depth, n_marched, loss_out, grad_rays, workspace, workspace_bytes, s, who
  if (kPoseGrad) return launch_march_occ<kMode, kPoseGrad, 2>(PP_MARCH_ARGS);
  if (occ >= 4) return launch_march_occ<kMode, false, 4>(PP_MARCH_ARGS);
  return launch_march_occ<kMode, false, 3>(PP_MARCH_ARGS);
#undef PP_MARCH_ARGS
}

}  // namespace pp

using namespace pp;

extern "C" int pp_voxels_pack(const float* density, const float* k0, int64_t n_voxels, void* vox,
                              pp_stream_t stream) {
  PP_REQUIRE(n_voxels >= 0, PP_ERR_INVALID_ARG, "voxels_pack: n_voxels < 0");
  if (n_voxels == 0) return PP_OK;
  PP_REQUIRE(density && k0 && vox, PP_ERR_INVALID_ARG, "voxels_pack: null pointer");
  const int blocks = (int)min64((n_voxels + 255) / 256, (int64_t)sm_count() * 16);
  voxels_pack_kernel<<<blocks, 256, 0, as_stream(stream)>>>(density, k0, n_voxels, (float4*)vox);
  PP_CHECK_LAUNCH("voxels_pack");
  return PP_OK;
}

extern "C" int pp_voxels_unpack(const void* vox, int64_t n_voxels, float* density, float* k0,
                                pp_stream_t stream) {
  PP_REQUIRE(n_voxels >= 0, PP_ERR_INVALID_ARG, "voxels_unpack: n_voxels < 0");
  if (n_voxels == 0 || (!density && !k0)) return PP_OK;
  PP_REQUIRE(vox, PP_ERR_INVALID_ARG, "voxels_unpack: null pointer");
  const int blocks = (int)min64((n_voxels + 255) / 256, (int64_t)sm_count() * 16);
  voxels_unpack_kernel<<<blocks, 256, 0, as_stream(stream)>>>((const float4*)vox, n_voxels, density, k0);
  PP_CHECK_LAUNCH("voxels_unpack");
  return PP_OK;
}

extern "C" int pp_dvgo_render_forward(const pp_dvgo_scene* scene, const void* vox, const float* rays_o,
                                      const float* rays_d, int64_t n_rays, float* rgb_marched,
                                      float* alphainv_last, float* depth, int32_t* n_marched,
                                      pp_stream_t stream) {
  const int rc = check_scene(scene, "dvgo_render_forward");
  if (rc != PP_OK) return rc;
  PP_REQUIRE(n_rays >= 0, PP_ERR_INVALID_ARG, "dvgo_render_forward: n_rays < 0");
  if (n_rays == 0) return PP_OK;
  PP_REQUIRE(vox && rays_o && rays_d, PP_ERR_INVALID_ARG, "dvgo_render_forward: null pointer");
  return launch_march<kModeForward, false>(scene, nullptr, vox, rays_o, rays_d, n_rays, nullptr, nullptr,
                                           nullptr, nullptr, rgb_marched, alphainv_last, depth,
                                           n_marched, nullptr, nullptr, nullptr, 0, as_stream(stream),
                                           "dvgo_render_forward");
}

extern "C" int64_t pp_dvgo_workspace_bytes(const pp_dvgo_scene* scene) {
  if (!scene || scene->stepdist <= 0.f) return 0;
  // upper bound over every kernel variant: 8 warps per block, at most 4 resident blocks per SM
  return (int64_t)sm_count() * 4 * 8 * (int64_t)step_capacity(scene) * (int64_t)sizeof(float4);
}

extern "C" int pp_dvgo_render_backward(const pp_dvgo_scene* scene, const void* vox, const float* rays_o,
                                       const float* rays_d, int64_t n_rays,
                                       const float* grad_rgb_marched, const float* grad_alphainv_last,
                                       void* grad_vox, float* grad_rays, void* workspace,
                                       int64_t workspace_bytes, pp_stream_t stream) {
  const int rc = check_scene(scene, "dvgo_render_backward");
  if (rc != PP_OK) return rc;
  PP_REQUIRE(n_rays >= 0, PP_ERR_INVALID_ARG, "dvgo_render_backward: n_rays < 0");
  if (n_rays == 0) return PP_OK;
  PP_REQUIRE(vox && rays_o && rays_d && grad_rgb_marched && grad_vox, PP_ERR_INVALID_ARG,
             "dvgo_render_backward: null pointer");
  if (grad_rays)
    return launch_march<kModeBackward, true>(scene, nullptr, vox, rays_o, rays_d, n_rays, nullptr,
                                             grad_rgb_marched, grad_alphainv_last, grad_vox, nullptr,
                                             nullptr, nullptr, nullptr, nullptr, grad_rays, workspace,
                                             workspace_bytes, as_stream(stream), "dvgo_render_backward");
  return launch_march<kModeBackward, false>(scene, nullptr, vox, rays_o, rays_d, n_rays, nullptr,
                                            grad_rgb_marched, grad_alphainv_last, grad_vox, nullptr,
                                            nullptr, nullptr, nullptr, nullptr, nullptr, workspace,
                                            workspace_bytes, as_stream(stream), "dvgo_render_backward");
}

extern "C" int pp_dvgo_train_step(const pp_dvgo_scene* scene, const pp_dvgo_loss* loss, const void* vox,
                                  const float* rays_o, const float* rays_d, const float* target,
                                  int64_t n_rays, void* grad_vox, float* rgb_marched, float* loss_out,
                                  float* grad_rays, void* workspace, int64_t workspace_bytes,
                                  pp_stream_t stream) {
  const int rc = check_scene(scene, "dvgo_train_step");
  if (rc != PP_OK) return rc;
  PP_REQUIRE(n_rays >= 0, PP_ERR_INVALID_ARG, "dvgo_train_step: n_rays < 0");
  if (n_rays == 0) return PP_OK;
  PP_REQUIRE(vox && rays_o && rays_d && target && grad_vox, PP_ERR_INVALID_ARG,
             "dvgo_train_step: null pointer");
  if (grad_rays)
    return launch_march<kModeTrain, true>(scene, loss, vox, rays_o, rays_d, n_rays, target, nullptr,
                                          nullptr, grad_vox, rgb_marched, nullptr, nullptr, nullptr,
                                          loss_out, grad_rays, workspace, workspace_bytes, as_stream(stream),
                                          "dvgo_train_step");
  return launch_march<kModeTrain, false>(scene, loss, vox, rays_o, rays_d, n_rays, target, nullptr,
                                         nullptr, grad_vox, rgb_marched, nullptr, nullptr, nullptr,
                                         loss_out, nullptr, workspace, workspace_bytes, as_stream(stream),
                                         "dvgo_train_step");
}

static float adam_step_size(float lr, float beta1, float beta2, int step) {
  // adam_upd_kernel.cu:72
  return lr * sqrtf(1.f - powf(beta2, (float)step)) / (1.f - powf(beta1, (float)step));
}

extern "C" int pp_voxel_adam(void* vox, void* grad_vox, void* exp_avg, void* exp_avg_sq,
                             int64_t n_voxels, int32_t step, float beta1, float beta2,
                             float lr_density, float lr_k0, float eps, int flags,
                             pp_stream_t stream) {
  PP_REQUIRE(n_voxels >= 0, PP_ERR_INVALID_ARG, "voxel_adam: n_voxels < 0");
  if (n_voxels == 0) return PP_OK;
  PP_REQUIRE(vox && grad_vox && exp_avg && exp_avg_sq, PP_ERR_INVALID_ARG, "voxel_adam: null pointer");
  const int blocks = (int)min64((n_voxels + 255) / 256, (int64_t)sm_count() * 8);
  voxel_adam_kernel<<<blocks, 256, 0, as_stream(stream)>>>(
      (float4*)vox, (float4*)grad_vox, (float4*)exp_avg, (float4*)exp_avg_sq, n_voxels,
      adam_step_size(lr_density, beta1, beta2, step), adam_step_size(lr_k0, beta1, beta2, step), beta1,
      beta2, (flags & 2) ? eps * sqrtf(1.f - powf(beta2, (float)step)) : eps, flags);
  PP_CHECK_LAUNCH("voxel_adam");
  return PP_OK;
}

extern "C" int pp_voxel_tv_adam(const pp_dvgo_scene* scene, const void* vox_in, void* vox_out,
                                void* grad_vox, void* exp_avg, void* exp_avg_sq, int32_t step,
                                float beta1, float beta2, float lr_density, float lr_k0, float eps,
                                float tv_density, float tv_k0, int dense_mode, pp_stream_t stream) {
  const int rc = check_scene(scene, "voxel_tv_adam");
  if (rc != PP_OK) return rc;
  PP_REQUIRE(vox_in && vox_out && grad_vox && exp_avg && exp_avg_sq, PP_ERR_INVALID_ARG,
             "voxel_tv_adam: null pointer");
  PP_REQUIRE(vox_in != vox_out, PP_ERR_INVALID_ARG, "voxel_tv_adam: vox_in and vox_out must differ");
  const dim3 block(32, 8, 1);
  const dim3 grid((scene->sz + 31) / 32, (scene->sy + 7) / 8, (unsigned)min64(scene->sx, 65535));
  voxel_tv_adam_kernel<<<grid, block, 0, as_stream(stream)>>>(
      (const float4*)vox_in, (float4*)vox_out, (float4*)grad_vox, (float4*)exp_avg, (float4*)exp_avg_sq,
      scene->sx, scene->sy, scene->sz, adam_step_size(lr_density, beta1, beta2, step),
      adam_step_size(lr_k0, beta1, beta2, step), beta1, beta2, eps, tv_density / 6, tv_k0 / 6, dense_mode);
  PP_CHECK_LAUNCH("voxel_tv_adam");
  return PP_OK;
}

// ---- one training step as one CUDA-graph launch ----------------------------------------------------------
// The step of the coarse stage is five stream operations (ray batch host -> device, loss reset, ray-queue reset,
// march, Adam) plus the loss read-back, issued one by one from the host whenever the caller reads the loss of every
// step (as the reference's trainer does).  Captured once and replayed, the step is ONE launch - for hosts that are
// slow to issue (measured on an idle host: no gain, 0.301 vs 0.300 ms; the gap to the device-resident 0.269 ms is
// the PCIe copy and the host round trip, which dvgo.py's HostPipeline hides instead).  What changes from step to
// step - the pinned host buffer the rays come from
// and Adam's bias-corrected step sizes (host-computed in float exactly as adam_upd_kernel.cu:72, so the graph
// replay and the eager path give the same bits) - is patched into the instantiated graph before each launch.
namespace pp {

struct AdamArgs {
  float4 *vox, *grad, *m4, *v4;
  int64_t n;
  float ss_density, ss_k0, b1, b2, eps;
  int flags;
};

struct StepGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaGraphNode_t h2d_node = nullptr, adam_node = nullptr;
  cudaKernelNodeParams adam_params{};
  AdamArgs adam{};
  void* adam_argv[11];
  float* rays_dev = nullptr;
  const float* rays_host = nullptr;
  size_t ray_bytes = 0;
  float lr_density = 0.f, lr_k0 = 0.f, eps = 0.f;
  int device = 0;
  int kernels = 0;
};

static void step_graph_free(StepGraph* g) {
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
}

}  // namespace pp

extern "C" int pp_dvgo_step_graph_create(const pp_dvgo_scene* scene, const pp_dvgo_loss* loss, void* vox,
                                         void* grad_vox, void* exp_avg, void* exp_avg_sq,
                                         const float* rays_host, float* rays_dev, int64_t n_rays,
                                         float* loss_dev, float* loss_host, void* workspace,
                                         int64_t workspace_bytes, float beta1, float beta2, float lr_density,
                                         float lr_k0, float eps, int adam_flags, void** graph_out) {
  const int rc = check_scene(scene, "dvgo_step_graph_create");
  if (rc != PP_OK) return rc;
  PP_REQUIRE(graph_out, PP_ERR_INVALID_ARG, "dvgo_step_graph_create: graph_out is null");
  *graph_out = nullptr;
  PP_REQUIRE(n_rays > 0, PP_ERR_INVALID_ARG, "dvgo_step_graph_create: n_rays must be positive");
  PP_REQUIRE(vox && grad_vox && exp_avg && exp_avg_sq && rays_host && rays_dev && loss_dev && loss_host,
             PP_ERR_INVALID_ARG, "dvgo_step_graph_create: null pointer");
  cudaPointerAttributes attr;
  PP_CHECK_CUDA(cudaPointerGetAttributes(&attr, rays_host));
  PP_REQUIRE(attr.type == cudaMemoryTypeHost, PP_ERR_INVALID_ARG,
             "dvgo_step_graph_create: rays_host must be pinned host memory");
  PP_CHECK_CUDA(cudaPointerGetAttributes(&attr, loss_host));
  PP_REQUIRE(attr.type == cudaMemoryTypeHost, PP_ERR_INVALID_ARG,
             "dvgo_step_graph_create: loss_host must be pinned host memory");
  const int64_t n_voxels = (int64_t)scene->sx * scene->sy * scene->sz;
  // everything with process-wide lazy state is touched before the capture starts
  unsigned int* qb = nullptr;
  const int qrc = ray_queue_base(&qb, "dvgo_step_graph_create");
  if (qrc != PP_OK) return qrc;
  (void)sm_count();

  StepGraph* g = new StepGraph();
  PP_CHECK_CUDA(cudaGetDevice(&g->device));
  g->rays_dev = rays_dev;
  g->rays_host = rays_host;
  g->ray_bytes = (size_t)n_rays * 9 * sizeof(float);
  g->lr_density = lr_density;
  g->lr_k0 = lr_k0;
  g->eps = eps;
  cudaStream_t cap = nullptr;
  cudaError_t e = cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    step_graph_free(g);
    set_error("dvgo_step_graph_create: stream: %s", cudaGetErrorString(e));
    return PP_ERR_CUDA;
  }
  const uint64_t launches0 = pp_launch_count();
  int step_rc = PP_OK;
  e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed);
  if (e == cudaSuccess) {
    e = cudaMemcpyAsync(rays_dev, rays_host, g->ray_bytes, cudaMemcpyHostToDevice, cap);
    if (e == cudaSuccess) e = cudaMemsetAsync(loss_dev, 0, 2 * sizeof(float), cap);
    if (e == cudaSuccess) {
      step_rc = pp_dvgo_train_step(scene, loss, vox, rays_dev, rays_dev + 3 * n_rays, rays_dev + 6 * n_rays, n_rays,
                                   grad_vox, nullptr, loss_dev, nullptr, workspace, workspace_bytes,
                                   (pp_stream_t)cap);
      if (step_rc == PP_OK)
        step_rc = pp_voxel_adam(vox, grad_vox, exp_avg, exp_avg_sq, n_voxels, 1, beta1, beta2, lr_density, lr_k0, eps,
                                adam_flags, (pp_stream_t)cap);
    }
    if (e == cudaSuccess && step_rc == PP_OK)
      e = cudaMemcpyAsync(loss_host, loss_dev, 2 * sizeof(float), cudaMemcpyDeviceToHost, cap);
    const cudaError_t e_end = cudaStreamEndCapture(cap, &g->graph);
    if (e == cudaSuccess) e = e_end;
  }
  g->kernels = (int)(pp_launch_count() - launches0);
  count_launch(-g->kernels);   // recorded, not launched: replays are counted in pp_dvgo_step_graph_launch
  cudaStreamDestroy(cap);
  if (step_rc != PP_OK) {
    step_graph_free(g);
    return step_rc;
  }
  if (e != cudaSuccess || !g->graph) {
    step_graph_free(g);
    set_error("dvgo_step_graph_create: capture failed: %s", cudaGetErrorString(e));
    return PP_ERR_CUDA;
  }
  // locate the two nodes that are patched per launch
  size_t n_nodes = 0;
  cudaGraphNode_t nodes[32];
  e = cudaGraphGetNodes(g->graph, nullptr, &n_nodes);
  if (e == cudaSuccess && n_nodes <= 32) e = cudaGraphGetNodes(g->graph, nodes, &n_nodes);
  if (e != cudaSuccess || n_nodes > 32) {
    step_graph_free(g);
    set_error("dvgo_step_graph_create: unexpected graph (%zu nodes): %s", n_nodes, cudaGetErrorString(e));
    return PP_ERR_CUDA;
  }
  for (size_t i = 0; i < n_nodes; ++i) {
    cudaGraphNodeType type;
    if (cudaGraphNodeGetType(nodes[i], &type) != cudaSuccess) continue;
    if (type == cudaGraphNodeTypeMemcpy) {
      cudaMemcpy3DParms mp;
      if (cudaGraphMemcpyNodeGetParams(nodes[i], &mp) == cudaSuccess && mp.dstPtr.ptr == (void*)rays_dev)
        g->h2d_node = nodes[i];
    } else if (type == cudaGraphNodeTypeKernel) {
      cudaKernelNodeParams kp;
      // the Adam pass is the kernel whose first two arguments are the parameter and gradient record arrays (the
      // march takes the scene struct first)
      if (cudaGraphKernelNodeGetParams(nodes[i], &kp) == cudaSuccess && kp.kernelParams &&
          (kp.func == (void*)voxel_adam_kernel ||
           (*(void**)kp.kernelParams[0] == vox && *(void**)kp.kernelParams[1] == grad_vox))) {
        g->adam_node = nodes[i];
        g->adam_params = kp;
      }
    }
  }
  if (!g->h2d_node || !g->adam_node) {
    step_graph_free(g);
    set_error("dvgo_step_graph_create: could not locate the %s node among %zu captured nodes",
              g->h2d_node ? "Adam kernel" : "ray copy", n_nodes);
    return PP_ERR_CUDA;
  }
  g->adam = AdamArgs{(float4*)vox, (float4*)grad_vox, (float4*)exp_avg, (float4*)exp_avg_sq, n_voxels,
                     0.f, 0.f, beta1, beta2, eps, adam_flags};
  void* argv[11] = {&g->adam.vox, &g->adam.grad, &g->adam.m4, &g->adam.v4, &g->adam.n, &g->adam.ss_density,
                    &g->adam.ss_k0, &g->adam.b1, &g->adam.b2, &g->adam.eps, &g->adam.flags};
  for (int i = 0; i < 11; ++i) g->adam_argv[i] = argv[i];
  g->adam_params.kernelParams = g->adam_argv;
  g->adam_params.extra = nullptr;
  e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) {
    step_graph_free(g);
    set_error("dvgo_step_graph_create: instantiate failed: %s", cudaGetErrorString(e));
    return PP_ERR_CUDA;
  }
  *graph_out = g;
  return PP_OK;
}

extern "C" int pp_dvgo_step_graph_launch(void* graph, const float* rays_host, int32_t step, pp_stream_t stream) {
  StepGraph* g = static_cast<StepGraph*>(graph);
  PP_REQUIRE(g && g->exec, PP_ERR_INVALID_ARG, "dvgo_step_graph_launch: null graph");
  PP_REQUIRE(step >= 1, PP_ERR_INVALID_ARG, "dvgo_step_graph_launch: step counts from 1");
  int dev = 0;
  PP_CHECK_CUDA(cudaGetDevice(&dev));
  PP_REQUIRE(dev == g->device, PP_ERR_INVALID_ARG, "dvgo_step_graph_launch: graph belongs to device %d, current is %d",
             g->device, dev);
  if (rays_host && rays_host != g->rays_host) {
    PP_CHECK_CUDA(cudaGraphExecMemcpyNodeSetParams1D(g->exec, g->h2d_node, g->rays_dev, rays_host, g->ray_bytes,
                                                     cudaMemcpyHostToDevice));
    g->rays_host = rays_host;
  }
  // the same host arithmetic as pp_voxel_adam
  g->adam.ss_density = adam_step_size(g->lr_density, g->adam.b1, g->adam.b2, step);
  g->adam.ss_k0 = adam_step_size(g->lr_k0, g->adam.b1, g->adam.b2, step);
  g->adam.eps = (g->adam.flags & 2) ? g->eps * sqrtf(1.f - powf(g->adam.b2, (float)step)) : g->eps;
  PP_CHECK_CUDA(cudaGraphExecKernelNodeSetParams(g->exec, g->adam_node, &g->adam_params));
  PP_CHECK_CUDA(cudaGraphLaunch(g->exec, as_stream(stream)));
  count_launch(g->kernels);
  return PP_OK;
}

extern "C" int pp_dvgo_step_graph_destroy(void* graph) {
  step_graph_free(static_cast<StepGraph*>(graph));
  return PP_OK;
}

// ---- rendering rays that live in host memory ---------------------------------------------------------------
// An image is marched in chunks; issued in line, every chunk pays its PCIe copies (24 B in, 12..20 B out per ray)
// and a host round trip on top of a ~49 us kernel (8192 rays, config B).  Here chunk c+1's copy-in and chunk c-1's
// copy-out run on two copy streams under chunk c's kernel (double-buffered staging blocks, event-ordered), and the
// whole image is issued by this one call - ~8 driver calls per chunk from C instead of ~25 from an interpreter, which
// is what keeps the issue rate above the kernel rate.
namespace pp {

struct RenderHostCtx {
  cudaStream_t in = nullptr, out = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr}, marched[2] = {nullptr, nullptr}, landed[2] = {nullptr, nullptr};
  float* rays[2] = {nullptr, nullptr};   // [2, chunk, 3]
  float* outs[2] = {nullptr, nullptr};   // rgb [chunk,3] | depth [chunk] | last [chunk]
  int64_t chunk = 0;
};

static int render_host_ctx(RenderHostCtx** out, int64_t chunk) {
  static RenderHostCtx ctxs[64];
  int dev = 0;
  PP_CHECK_CUDA(cudaGetDevice(&dev));
  PP_REQUIRE(dev >= 0 && dev < 64, PP_ERR_UNSUPPORTED, "dvgo_render_host: device ordinal %d out of range", dev);
  RenderHostCtx& c = ctxs[dev];
  if (!c.in) {
    PP_CHECK_CUDA(cudaStreamCreateWithFlags(&c.in, cudaStreamNonBlocking));
    PP_CHECK_CUDA(cudaStreamCreateWithFlags(&c.out, cudaStreamNonBlocking));
    for (int j = 0; j < 2; ++j) {
      PP_CHECK_CUDA(cudaEventCreateWithFlags(&c.copied[j], cudaEventDisableTiming));
      PP_CHECK_CUDA(cudaEventCreateWithFlags(&c.marched[j], cudaEventDisableTiming));
      PP_CHECK_CUDA(cudaEventCreateWithFlags(&c.landed[j], cudaEventDisableTiming));
    }
  }
  if (c.chunk < chunk) {
    // every earlier call ended with a synchronisation of the copy-out stream, so the old blocks are idle
    for (int j = 0; j < 2; ++j) {
      if (c.rays[j]) cudaFree(c.rays[j]);
      if (c.outs[j]) cudaFree(c.outs[j]);
      c.rays[j] = c.outs[j] = nullptr;
    }
    c.chunk = 0;
    for (int j = 0; j < 2; ++j) {
      PP_CHECK_CUDA(cudaMalloc((void**)&c.rays[j], (size_t)chunk * 6 * sizeof(float)));
      PP_CHECK_CUDA(cudaMalloc((void**)&c.outs[j], (size_t)chunk * 5 * sizeof(float)));
    }
    c.chunk = chunk;
  }
  *out = &c;
  return PP_OK;
}

}  // namespace pp

extern "C" int pp_dvgo_render_host(const pp_dvgo_scene* scene, const void* vox, const float* rays_host,
                                   int64_t n_rays, int64_t chunk, float* rgb_host, float* depth_host,
                                   float* last_host, pp_stream_t stream) {
  const int rc = check_scene(scene, "dvgo_render_host");
  if (rc != PP_OK) return rc;
  PP_REQUIRE(n_rays >= 0, PP_ERR_INVALID_ARG, "dvgo_render_host: n_rays < 0");
  if (n_rays == 0) return PP_OK;
  PP_REQUIRE(chunk > 0, PP_ERR_INVALID_ARG, "dvgo_render_host: chunk must be positive");
  PP_REQUIRE(vox && rays_host && rgb_host, PP_ERR_INVALID_ARG, "dvgo_render_host: null pointer");
  const void* host_ptrs[4] = {rays_host, rgb_host, depth_host, last_host};
  for (int i = 0; i < 4; ++i) {
    if (!host_ptrs[i]) continue;
    cudaPointerAttributes attr;
    PP_CHECK_CUDA(cudaPointerGetAttributes(&attr, host_ptrs[i]));
    PP_REQUIRE(attr.type == cudaMemoryTypeHost, PP_ERR_INVALID_ARG,
               "dvgo_render_host: host buffer %d is not pinned host memory (no pageable staging here)", i);
  }
  if (chunk > n_rays) chunk = n_rays;
  RenderHostCtx* cx = nullptr;
  const int crc = render_host_ctx(&cx, chunk);
  if (crc != PP_OK) return crc;
  cudaStream_t s = as_stream(stream);
  int64_t c = 0;
  for (int64_t c0 = 0; c0 < n_rays; c0 += chunk, ++c) {
    const int j = (int)(c & 1);
    const int64_t m = min64(chunk, n_rays - c0);
    float* o_dev = cx->rays[j];
    float* d_dev = cx->rays[j] + 3 * cx->chunk;
    float* rgb_dev = cx->outs[j];
    float* depth_dev = cx->outs[j] + 3 * cx->chunk;
    float* last_dev = cx->outs[j] + 4 * cx->chunk;
    PP_CHECK_CUDA(cudaStreamWaitEvent(cx->in, cx->marched[j], 0));   // the kernel that last read this block
    PP_CHECK_CUDA(cudaMemcpyAsync(o_dev, rays_host + 3 * c0, (size_t)m * 12, cudaMemcpyHostToDevice, cx->in));
    PP_CHECK_CUDA(cudaMemcpyAsync(d_dev, rays_host + 3 * (n_rays + c0), (size_t)m * 12, cudaMemcpyHostToDevice, cx->in));
    PP_CHECK_CUDA(cudaEventRecord(cx->copied[j], cx->in));
    PP_CHECK_CUDA(cudaStreamWaitEvent(s, cx->copied[j], 0));
    PP_CHECK_CUDA(cudaStreamWaitEvent(s, cx->landed[j], 0));         // the copy-out that last read these outputs
    const int mrc = launch_march<kModeForward, false>(scene, nullptr, vox, o_dev, d_dev, m, nullptr, nullptr, nullptr,
                                                      nullptr, rgb_dev, last_dev, depth_host ? depth_dev : nullptr,
                                                      nullptr, nullptr, nullptr, nullptr, 0, s, "dvgo_render_host");
    if (mrc != PP_OK) return mrc;
    PP_CHECK_CUDA(cudaEventRecord(cx->marched[j], s));
    PP_CHECK_CUDA(cudaStreamWaitEvent(cx->out, cx->marched[j], 0));
    PP_CHECK_CUDA(cudaMemcpyAsync(rgb_host + 3 * c0, rgb_dev, (size_t)m * 12, cudaMemcpyDeviceToHost, cx->out));
    if (depth_host)
      PP_CHECK_CUDA(cudaMemcpyAsync(depth_host + c0, depth_dev, (size_t)m * 4, cudaMemcpyDeviceToHost, cx->out));
    if (last_host)
      PP_CHECK_CUDA(cudaMemcpyAsync(last_host + c0, last_dev, (size_t)m * 4, cudaMemcpyDeviceToHost, cx->out));
    PP_CHECK_CUDA(cudaEventRecord(cx->landed[j], cx->out));
  }
  PP_CHECK_CUDA(cudaStreamSynchronize(cx->out));
  return PP_OK;
}
