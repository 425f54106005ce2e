// Per-ray alpha compositing (alpha2weight forward / backward) and per-ray segmented reductions.
//
// Replaces render_utils_kernel.cu:577-707.  The reference walks each ray's segment with ONE thread
// (`for(i=i_s; i<i_e_max; ++i)`), so a warp touches 32 different rays per step: every load/store is
// a 32-way uncoalesced access, and 8192 rays occupy 256 warps.  The recurrences themselves are strictly
// sequential *and* their float roundings decide an integer output (the early-termination index written
// back to i_end), so they cannot be re-associated.  The B200 design keeps one lane per ray for the recurrence (every
// instruction of the chain advances 32 rays) and fixes the memory side: the warp moves [32 rays x 8 samples] tiles
// through shared memory, so every global load / store is a 32-byte sector per ray instead of a 4-byte touch of 32
// different lines, and the next tile's loads are in flight while the chain runs.
#include "common.cuh"

namespace pp {


// ---- segment bounds from a sorted ray_id (reference :607-617 and the host line :635) -------------
__global__ void segment_bounds_kernel(const int64_t* __restrict__ ray_id, int64_t n_pts,
                                      int64_t n_rays, int64_t* __restrict__ i_start,
                                      int64_t* __restrict__ i_end) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_pts) return;
  const int64_t cur = ray_id[i];
  if (i > 0) {
    const int64_t prev = ray_id[i - 1];
    if (cur != prev) {
      if ((uint64_t)cur < (uint64_t)n_rays) i_start[cur] = i;
      if ((uint64_t)prev < (uint64_t)n_rays) i_end[prev] = i;
    }
  }
  if (i == n_pts - 1 && (uint64_t)cur < (uint64_t)n_rays) i_end[cur] = n_pts;
}

// ---- forward -------------------------------------------------------------------------------------
// The recurrence T_cum *= (1. - alpha) is double arithmetic narrowed to float at every step (reference :596) and its
// roundings decide the early-termination index, so it runs in the reference's order: ONE lane per ray, as in the
// reference, 32 rays per warp - every instruction of the chain (two conversions and a double multiply, 44 cycles per
// sample) advances 32 rays.  What the reference gets wrong is memory: its lanes walk 32 different rays, so every load /
// store touches 32 lines.  Here a block moves a [32 rays x kT samples] tile per step through shared memory: groups
// of kT lanes copy kT consecutive samples of one ray (64 bytes for fp32), the ray's own lane reads its row back with
// 128-bit loads, runs kT steps of the chain in registers and writes its T / weight rows, which leave as 64-byte runs
// again.  Samples past the termination get the reference's fill values (T = 1, weight = 0) from the same streaming
// stores, so no separate fill pass over [M] is needed.  The kernels below split this between one chain warp and
// several memory warps.
// Tile geometry: kT samples per ray and step (16 for fp32, 8 for fp64), kWarp / kT rays per sector instruction.
// row r, column c of a [32][kT] tile.  A lane reads its whole row with 128-bit loads, 8 lanes per bank cycle; rows that
// would meet in the same banks (32 / kT rows apart) get their 4-element groups XOR-permuted by the row's index within
// that cycle, so the row reads and the column-wise staging stores are both conflict free.
template <int kT>
__device__ __forceinline__ int tile_at(int r, int c) {
  constexpr int kRowsPerCycle = 32 / kT, kQuads = kT / 4;
  return r * kT + (c ^ (((r / kRowsPerCycle) % kQuads) << 2));
}

// a lane's whole row of a tile (its ray's kT samples): 128-bit shared-memory accesses, four elements at a time - the
// XOR permutation of tile_at moves whole 4-element groups, so group q of the row sits at group q ^ swz
template <typename T, int kT>
__device__ __forceinline__ void load_row(const T* tile, int r, T (&v)[kT]) {
  if constexpr (sizeof(T) == 4) {
    constexpr int kRowsPerCycle = 32 / kT, kQuads = kT / 4;
    const int swz = (r / kRowsPerCycle) % kQuads;
    const float4* row = reinterpret_cast<const float4*>(tile + r * kT);
#pragma unroll
    for (int q = 0; q < kQuads; ++q) {
      const float4 x = row[q ^ swz];
      v[4 * q] = x.x;
      v[4 * q + 1] = x.y;
      v[4 * q + 2] = x.z;
      v[4 * q + 3] = x.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kT; ++j) v[j] = tile[tile_at<kT>(r, j)];
  }
}
template <typename T, int kT>
__device__ __forceinline__ void store_row(T* tile, int r, const T (&v)[kT]) {
  if constexpr (sizeof(T) == 4) {
    constexpr int kRowsPerCycle = 32 / kT, kQuads = kT / 4;
    const int swz = (r / kRowsPerCycle) % kQuads;
    float4* row = reinterpret_cast<float4*>(tile + r * kT);
#pragma unroll
    for (int q = 0; q < kQuads; ++q) row[q ^ swz] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < kT; ++j) tile[tile_at<kT>(r, j)] = v[j];
  }
}

// 1 + kFwdMemWarps warps per 32 rays.  Warp 0 ("chain") only runs the recurrence: per step it reads its ray's row of
// the input tile, advances kT samples, writes its T / weight rows - ~100 instructions, and a lone warp issues one
// dependent instruction every ~5 cycles, so instructions on this warp ARE the kernel's duration.  The memory warps do
// all the addressing (each owns 8 / kFwdMemWarps of a step's 8 sector instructions): they know their rays'
// [begin, end), copy tile t+3 from global memory STRAIGHT into shared memory with cp.async (zero-filled outside the
// ray; it is needed two steps later, so DRAM latency is covered - register-staged loads were serialised by the compiler's
// scoreboard waits, profiles/r2_results.md) and stream the outputs of tile t-1 back as 32-byte sectors.  Input tiles
// rotate through five shared-memory slots, output tiles through three; one block barrier per step orders the warps.
constexpr int kFwdMemWarps = 4;
constexpr int kInSlots = 5;

template <int kBytes>
__device__ __forceinline__ void cp_async_elem(void* dst_smem, const void* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(d), "l"(src), "n"(kBytes),
               "r"(valid ? (uint32_t)kBytes : 0u)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

template <typename scalar_t, int kT>
__global__ void __launch_bounds__((1 + kFwdMemWarps) * kWarp)
alpha2weight_fwd_kernel(const scalar_t* __restrict__ alpha, int64_t n_rays,
                        scalar_t* __restrict__ weight, scalar_t* __restrict__ T_out,
                        scalar_t* __restrict__ alphainv_last, const int64_t* __restrict__ i_start,
                        int64_t* __restrict__ i_end) {
  constexpr int kGroups = kWarp / kT;            // rays per sector instruction
  constexpr int kFwdK = kT / kFwdMemWarps;       // sector instructions per memory warp and step
  __shared__ __align__(16) scalar_t s_a[kInSlots][kWarp * kT];
  __shared__ __align__(16) scalar_t s_T[3][kWarp * kT], s_w[3][kWarp * kT];
  __shared__ int s_steps;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane / kT, sl = lane % kT;
  const bool chain = warp == 0;
  const int64_t ray0 = blockIdx.x * (int64_t)kWarp;
  // chain warp: own ray.  (sample indices fit 32 bits: the entry point rejects n_pts > INT32_MAX)
  const int64_t ray = ray0 + lane;
  const bool live = chain && ray < n_rays;
  const int begin = live ? static_cast<int>(i_start[ray]) : 0, end = live ? static_cast<int>(i_end[ray]) : 0;
  // memory warp w serves the sector instructions k = (w - 1) * kFwdK + q of a step: ray 4k + sub
  int bk[kFwdK], ek[kFwdK], rk[kFwdK];
#pragma unroll
  for (int k = 0; k < kFwdK; ++k) {
    rk[k] = ((warp - 1) * kFwdK + k) * kGroups + sub;
    const int64_t r = ray0 + rk[k];
    const bool ok = !chain && r < n_rays;
    bk[k] = ok ? static_cast<int>(i_start[r]) : 0;
    ek[k] = ok ? static_cast<int>(i_end[r]) : 0;
  }
  if (chain) {
    int steps = (end - begin + kT - 1) / kT;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) steps = max(steps, __shfl_xor_sync(kFull, steps, o));
    if (lane == 0) s_steps = steps;
  }
  __syncthreads();
  const int steps = s_steps;
  auto fetch = [&](int t) {      // tile t: global -> shared, asynchronously; one commit group per tile
#pragma unroll
    for (int k = 0; k < kFwdK; ++k) {
      const int idx = bk[k] + t * kT + sl;
      const bool ok = t < steps && idx < ek[k];
      cp_async_elem<sizeof(scalar_t)>(&s_a[t % kInSlots][tile_at<kT>(rk[k], sl)], alpha + (ok ? idx : 0), ok);
    }
    cp_async_commit();
  };
  auto flush = [&](int t) {      // outputs of tile t: shared -> global
#pragma unroll
    for (int k = 0; k < kFwdK; ++k) {
      const int idx = bk[k] + t * kT + sl;
      if (idx < ek[k]) {
        T_out[idx] = s_T[t % 3][tile_at<kT>(rk[k], sl)];
        weight[idx] = s_w[t % 3][tile_at<kT>(rk[k], sl)];
      }
    }
  };
  if (!chain) {
    fetch(0);
    fetch(1);
    fetch(2);
    cp_async_wait<2>();          // tile 0 has landed
  }
  __syncthreads();
  int pos = begin, cut = -1;
  float T_cum = 1.f;
  for (int t = 0; t < steps; ++t) {
    if (chain) {
      scalar_t a[kT];
      load_row<scalar_t, kT>(s_a[t % kInSlots], lane, a);
      // The chain itself carries no predicate: samples beyond the ray's end were staged as alpha = 0, which leaves T
      // unchanged exactly (T * 1.0 in double, narrowed back), and the termination test is taken off the dependent path
      // - all kT transmittances are produced first, the first one below 1e-3 is located afterwards.  What the chain
      // computes past the termination is discarded (the reference stops multiplying there).
      float Tarr[kT + 1];
      Tarr[0] = T_cum;
#pragma unroll
      for (int j = 0; j < kT; ++j)
        Tarr[j + 1] = static_cast<float>(static_cast<double>(Tarr[j]) * (1. - static_cast<double>(a[j])));
      // reference :597 compares the widened float with the double literal 1e-3; 0.001f is the smallest float above
      // it, so for a float operand `< 1e-3` and `< 0.001f` are the same predicate
      int stop = kT;                 // index of the sample after which the ray terminates (kT: not in this tile)
#pragma unroll
      for (int j = kT - 1; j >= 0; --j)
        if (Tarr[j + 1] < 0.001f && pos + j < end) stop = j;
      const bool alive = cut < 0;
      scalar_t Tj[kT], wj[kT];
#pragma unroll
      for (int j = 0; j < kT; ++j) {
        const bool in = alive && j <= stop && pos + j < end;
        Tj[j] = in ? static_cast<scalar_t>(Tarr[j]) : scalar_t(1);
        wj[j] = in ? Tarr[j] * a[j] : scalar_t(0);
      }
      if (alive) {
        if (stop < kT) {
          cut = pos + stop + 1;
          T_cum = Tarr[kT];
#pragma unroll
          for (int j = 0; j < kT; ++j)
            if (j == stop) T_cum = Tarr[j + 1];
        } else {
          T_cum = Tarr[kT];
        }
      }
      store_row<scalar_t, kT>(s_T[t % 3], lane, Tj);
      store_row<scalar_t, kT>(s_w[t % 3], lane, wj);
      pos += kT;
    } else {
      fetch(t + 3);              // slot (t + 3) % 5: last read by the chain warp in step t - 2
      if (t > 0) flush(t - 1);
      cp_async_wait<2>();        // tile t + 1 has landed (tiles t + 2, t + 3 may still be in flight)
    }
    __syncthreads();
  }
  if (!chain && steps > 0) flush(steps - 1);
  if (live) {
    if (cut >= 0) i_end[ray] = cut;  // otherwise i_end already holds i_e_max (or 0 for an empty ray)
    alphainv_last[ray] = T_cum;
  }
}

// ---- backward ------------------------------------------------------------------------------------
// Walks each ray downwards from its (early-terminated) end in tiles [top - 8, top).  The recurrence
// back_cum += grad_weights*weight is an FMA in scalar_t narrowed to float (reference :674) - 8 dependent FMAs per
// step on the chain warp, which hands the value of back_cum in front of every sample to the memory warps through
// shared memory.  Everything else - the cp.async copies of the four inputs, and grad = gw*T - back/(1-alpha+1e-10)
// with its double-precision quotient, which is not part of the recurrence - is done by kBwdMemWarps memory warps in
// the coalesced layout (each owns 8 / kBwdMemWarps of the step's 8 sector instructions).
constexpr int kBwdMemWarps = 8;

template <typename scalar_t, int kT>
__global__ void __launch_bounds__((1 + kBwdMemWarps) * kWarp)
alpha2weight_bwd_kernel(const scalar_t* __restrict__ alpha, const scalar_t* __restrict__ weight,
                        const scalar_t* __restrict__ T, const scalar_t* __restrict__ alphainv_last,
                        const int64_t* __restrict__ i_start, const int64_t* __restrict__ i_end,
                        int64_t n_rays, const scalar_t* __restrict__ grad_weights,
                        const scalar_t* __restrict__ grad_last, scalar_t* __restrict__ grad) {
  constexpr int kGroups = kWarp / kT;
  constexpr int kBwdK = kT / kBwdMemWarps;       // sector instructions per memory warp and step
  __shared__ __align__(16) scalar_t s_in[kInSlots][4][kWarp * kT];   // [slot][gw, w, T, alpha]
  __shared__ __align__(16) float s_back[3][kWarp * kT];              // back_cum in front of every sample
  __shared__ int s_steps;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane / kT, sl = lane % kT;
  const bool chain = warp == 0;
  const int64_t ray0 = blockIdx.x * (int64_t)kWarp;
  const int64_t ray = ray0 + lane;
  const bool live = chain && ray < n_rays;
  const int begin = live ? static_cast<int>(i_start[ray]) : 0;
  const int top0 = live ? static_cast<int>(i_end[ray]) : 0;
  // back_cum = grad_last * alphainv_last (reference :671)
  float back = live ? static_cast<float>(grad_last[ray] * alphainv_last[ray]) : 0.f;
  // memory warp w serves the sector instructions k = (w - 1) * kBwdK + q of a step: ray 4k + sub
  int bk[kBwdK], tk[kBwdK], rk[kBwdK];
#pragma unroll
  for (int q = 0; q < kBwdK; ++q) {
    rk[q] = ((warp - 1) * kBwdK + q) * kGroups + sub;
    const int64_t r = ray0 + rk[q];
    const bool ok = !chain && r < n_rays;
    bk[q] = ok ? static_cast<int>(i_start[r]) : 0;
    tk[q] = ok ? static_cast<int>(i_end[r]) : 0;
  }
  if (chain) {
    int steps = (max(top0 - begin, 0) + kT - 1) / kT;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) steps = max(steps, __shfl_xor_sync(kFull, steps, o));
    if (lane == 0) s_steps = steps;
  }
  __syncthreads();
  const int steps = s_steps;
  const scalar_t* src[4] = {grad_weights, weight, T, alpha};
  // tile t of a ray covers samples [top - 8 (t + 1), top - 8 t); column c of the tile is sample top - 8 (t + 1) + c
  auto fetch = [&](int t) {
#pragma unroll
    for (int q = 0; q < kBwdK; ++q) {
      const int idx = tk[q] - (t + 1) * kT + sl;
      const bool ok = t < steps && idx >= bk[q];
#pragma unroll
      for (int f = 0; f < 4; ++f)
        cp_async_elem<sizeof(scalar_t)>(&s_in[t % kInSlots][f][tile_at<kT>(rk[q], sl)], src[f] + (ok ? idx : 0), ok);
    }
    cp_async_commit();
  };
  auto finish = [&](int t) {     // grad of tile t from the staged inputs and the chain warp's back_cum values
#pragma unroll
    for (int q = 0; q < kBwdK; ++q) {
      const int idx = tk[q] - (t + 1) * kT + sl;
      if (idx >= bk[q]) {
        const int at = tile_at<kT>(rk[q], sl);
        const scalar_t gw = s_in[t % kInSlots][0][at], Tv = s_in[t % kInSlots][2][at], av = s_in[t % kInSlots][3][at];
        // grad = gw*T - back/(1-alpha+1e-10): product in scalar_t, quotient and difference in double
        const scalar_t gT = mul_rn(gw, Tv);
        const scalar_t one_minus = scalar_t(1) - av;
        const double qd = static_cast<double>(s_back[t % 3][at]) / (static_cast<double>(one_minus) + 1e-10);
        grad[idx] = static_cast<scalar_t>(static_cast<double>(gT) - qd);
      }
    }
  };
  if (!chain) {
    fetch(0);
    fetch(1);
    fetch(2);
    cp_async_wait<2>();
  }
  __syncthreads();
  int top = top0;
  for (int t = 0; t < steps; ++t) {
    if (chain) {
      scalar_t gw[kT], wv[kT];
      load_row<scalar_t, kT>(s_in[t % kInSlots][0], lane, gw);
      load_row<scalar_t, kT>(s_in[t % kInSlots][1], lane, wv);
      const int base = top - kT;
      float bj[kT];
#pragma unroll
      for (int j = kT - 1; j >= 0; --j) {
        bj[j] = back;
        if (base + j >= begin && base + j < top)
          back = static_cast<float>(fma_rn(gw[j], wv[j], static_cast<scalar_t>(back)));
      }
      store_row<float, kT>(s_back[t % 3], lane, bj);
      if (top > begin) top -= kT;
    } else {
      fetch(t + 3);              // slot (t + 3) % 5: the chain warp reads slot t % 5, finish(t - 1) slot (t - 1) % 5
      if (t > 0) finish(t - 1);
      cp_async_wait<2>();
    }
    __syncthreads();
  }
  if (!chain && steps > 0) finish(steps - 1);
}

// ---- dense compositing: get_ray_marching_ray / cumprod_exclusive (lib/dvgo_ori.py:478-485) -------------------
//   alphainv_cum = [1, cumprod(clamp_min(1 - alpha, 1e-10))]  [n, S+1],   weights = alpha * alphainv_cum[..., :-1]
// over DENSE [n_rays, S] samples (out-of-box samples carry alpha = 0), with no early termination - the formulation of
// the reference's pure-PyTorch DirectVoxGO.  One warp per ray, lanes = 32 consecutive samples: a shuffle product scan
// per chunk, the running product carried across chunks; the backward is autograd's cumprod rule
// d/dp_j = (sum_{i>j} G_i C_i) / p_j as a reverse pass with a suffix sum.
constexpr int kDenseWarps = 8;

__global__ void __launch_bounds__(kDenseWarps* kWarp)
dense_composite_fwd_kernel(const float* __restrict__ alpha, int64_t n_rays, int S, float* __restrict__ weights,
                           float* __restrict__ cum) {
  const int lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * (int64_t)kDenseWarps + (threadIdx.x >> 5);
  if (r >= n_rays) return;
  const float* a_r = alpha + r * S;
  float* w_r = weights + r * S;
  float* c_r = cum + r * (int64_t)(S + 1);
  float carry = 1.f;
  if (lane == 0) c_r[0] = 1.f;
  for (int base = 0; base < S; base += 32) {
    const int i = base + lane;
    const float a = i < S ? a_r[i] : 0.f;
    float incl = i < S ? fmaxf(1.f - a, 1e-10f) : 1.f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl *= up;
    }
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.f;
    if (i < S) {
      w_r[i] = a * (carry * excl);
      c_r[i + 1] = carry * incl;
    }
    carry *= __shfl_sync(kFull, incl, 31);
  }
}

__global__ void __launch_bounds__(kDenseWarps* kWarp)
dense_composite_bwd_kernel(const float* __restrict__ alpha, const float* __restrict__ cum, int64_t n_rays, int S,
                           const float* __restrict__ g_w, const float* __restrict__ g_cum,
                           float* __restrict__ g_alpha) {
  const int lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * (int64_t)kDenseWarps + (threadIdx.x >> 5);
  if (r >= n_rays) return;
  const float* a_r = alpha + r * S;
  const float* c_r = cum + r * (int64_t)(S + 1);
  const float* gw_r = g_w ? g_w + r * S : nullptr;
  const float* gc_r = g_cum ? g_cum + r * (int64_t)(S + 1) : nullptr;
  float* ga_r = g_alpha + r * S;
  // suffix = sum_{i > j} G_i C_i with G_i = g_cum[i] + g_w[i] alpha_i (i < S); starts with the entry i = S
  float suffix = gc_r ? gc_r[S] * c_r[S] : 0.f;
  const int last_base = ((S - 1) / 32) * 32;
  for (int base = last_base; base >= 0; base -= 32) {
    const int j = base + lane;
    const bool in = j < S;
    const float a = in ? a_r[j] : 0.f;
    const float C = in ? c_r[j] : 0.f;
    const float gw = (in && gw_r) ? gw_r[j] : 0.f;
    const float term = in ? ((gc_r ? gc_r[j] : 0.f) + gw * a) * C : 0.f;   // G_j C_j
    float sfx = term;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float dn = __shfl_down_sync(kFull, sfx, o);
      if (lane + o < 32) sfx += dn;
    }
    float right = __shfl_down_sync(kFull, sfx, 1);
    if (lane == 31) right = 0.f;
    const float after = suffix + right;              // sum over i > j
    suffix += __shfl_sync(kFull, sfx, 0);
    if (in) {
      const float q = 1.f - a;
      const float p = fmaxf(q, 1e-10f);
      const float dp = (q >= 1e-10f) ? -(after / p) : 0.f;   // clamp_min passes the gradient where it does not clamp
      ga_r[j] = gw * C + dp;
    }
  }
}

// ---- per-ray segmented sum (torch_scatter.segment_coo(reduce='sum') call sites) -------------------
// One thread per sample, warp-level segmented reduction keyed by ray_id, one atomic per (warp, ray).
__global__ void segment_sum_kernel(const float* __restrict__ src, const float* __restrict__ weight,
                                   const int64_t* __restrict__ ray_id, int64_t n_pts, int n_ch,
                                   int64_t n_rays, float* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool valid = i < n_pts;
  const int64_t rid = valid ? ray_id[i] : -1;
  const float wv = (valid && weight) ? weight[i] : 1.f;
  // head of a run of equal ray ids inside this warp
  const int64_t prev = __shfl_up_sync(kFull, rid, 1);
  const bool head = (lane == 0) || (prev != rid);
  const unsigned heads = __ballot_sync(kFull, head);
  // index of the lane after the end of my run
  const unsigned after = heads & ~((2u << lane) - 1u);           // heads strictly above my lane
  const int run_end = after ? (__ffs(after) - 1) : 32;            // exclusive
  for (int c = 0; c < n_ch; ++c) {
    float v = valid ? src[i * n_ch + c] * wv : 0.f;
    // segmented suffix-sum by doubling: lane adds lane+o if it is still inside its run
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_down_sync(kFull, v, o);
      if (lane + o < run_end) v += up;
    }
    if (head && valid && (uint64_t)rid < (uint64_t)n_rays) atomicAdd(&out[rid * n_ch + c], v);
  }
}

// ---- per-ray segmented inclusive prefix sum ("segment_cumsum") -------------------------------------
// Two passes over blocks of kScanElems samples.  Pass 1 records, per block, the sum of its trailing
// open segment and whether the block contains a segment head; pass 2 rebuilds each block's carry-in
// by looking back over those summaries and performs the in-block segmented scan.
constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 4;
constexpr int kScanElems = kScanThreads * kScanPerThread;

__device__ __forceinline__ bool is_head(const int64_t* ray_id, int64_t i) {
  return i == 0 || ray_id[i] != ray_id[i - 1];
}

// segmented inclusive scan of (v, f) pairs across a block; returns the inclusive value for each of the
// thread's kScanPerThread consecutive elements. `carry` is added to elements before the first head.
__device__ __forceinline__ void block_segscan(float v[kScanPerThread], bool h[kScanPerThread],
                                              float carry, float& tail_sum, bool& any_head) {
  __shared__ float s_val[kScanThreads / kWarp];
  __shared__ int s_flag[kScanThreads / kWarp];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // thread-local
  float acc = 0.f;
  bool f = false;
#pragma unroll
  for (int e = 0; e < kScanPerThread; ++e) {
    if (h[e]) { acc = 0.f; f = true; }
    acc += v[e];
    v[e] = acc;
  }
  // warp-level segmented scan on (acc, f)
  float x = acc;
  int fl = f;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float ux = __shfl_up_sync(kFull, x, o);
    const int uf = __shfl_up_sync(kFull, fl, o);
    if (lane >= o) {
      if (!fl) x += ux;
      fl |= uf;
    }
  }
  if (lane == 31) { s_val[warp] = x; s_flag[warp] = fl; }
  __syncthreads();
  // prefix from previous warps
  float wpre = 0.f;
  int wflag = 0;
  for (int q = 0; q < warp; ++q) {
    if (s_flag[q]) { wpre = s_val[q]; wflag = 1; } else { wpre += s_val[q]; }
  }
  // exclusive (previous-thread) prefix inside the warp
  float px = __shfl_up_sync(kFull, x, 1);
  int pf = __shfl_up_sync(kFull, fl, 1);
  if (lane == 0) { px = 0.f; pf = 0; }
  float pre = pf ? px : (px + wpre);
  const bool pre_flag = pf || wflag;
  if (!pre_flag) pre += carry;
  // apply to the thread's elements that precede its first local head
#pragma unroll
  for (int e = 0; e < kScanPerThread; ++e) {
    if (h[e]) break;
    v[e] += pre;
  }
  // block summary
  float tot = 0.f;
  int anyf = 0;
  for (int q = 0; q < kScanThreads / kWarp; ++q) {
    if (s_flag[q]) { tot = s_val[q]; anyf = 1; } else { tot += s_val[q]; }
  }
  tail_sum = tot;
  any_head = anyf;
  __syncthreads();
}

__global__ void __launch_bounds__(kScanThreads)
segcumsum_summary_kernel(const float* __restrict__ src, const int64_t* __restrict__ ray_id,
                         int64_t n_pts, float* __restrict__ blk_tail, int* __restrict__ blk_head) {
  const int64_t base = blockIdx.x * (int64_t)kScanElems + threadIdx.x * kScanPerThread;
  float v[kScanPerThread];
  bool h[kScanPerThread];
#pragma unroll
  for (int e = 0; e < kScanPerThread; ++e) {
    const int64_t i = base + e;
    v[e] = i < n_pts ? src[i] : 0.f;
    h[e] = i < n_pts ? is_head(ray_id, i) : false;
  }
  float tail;
  bool any;
  block_segscan(v, h, 0.f, tail, any);
  if (threadIdx.x == 0) {
    blk_tail[blockIdx.x] = tail;
    blk_head[blockIdx.x] = any;
  }
}

__global__ void __launch_bounds__(kScanThreads)
segcumsum_apply_kernel(const float* __restrict__ src, const int64_t* __restrict__ ray_id,
                       int64_t n_pts, const float* __restrict__ blk_tail,
                       const int* __restrict__ blk_head, float* __restrict__ out) {
  __shared__ float s_carry;
  if (threadIdx.x == 0) {
    // carry-in = tails of the previous blocks back to (and including) the nearest one with a head,
    // accumulated in forward order so the result does not depend on the block size
    int first = (int)blockIdx.x;
    while (first > 0) {
      --first;
      if (blk_head[first]) break;
    }
    float c = 0.f;
    for (int b = first; b < (int)blockIdx.x; ++b) c += blk_tail[b];
    s_carry = c;
  }
  __syncthreads();
  const int64_t base = blockIdx.x * (int64_t)kScanElems + threadIdx.x * kScanPerThread;
  float v[kScanPerThread];
  bool h[kScanPerThread];
#pragma unroll
  for (int e = 0; e < kScanPerThread; ++e) {
    const int64_t i = base + e;
    v[e] = i < n_pts ? src[i] : 0.f;
    h[e] = i < n_pts ? is_head(ray_id, i) : false;
  }
  float tail;
  bool any;
  block_segscan(v, h, s_carry, tail, any);
#pragma unroll
  for (int e = 0; e < kScanPerThread; ++e) {
    const int64_t i = base + e;
    if (i < n_pts) out[i] = v[e];
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

template <typename scalar_t>
static __global__ void fill_kernel(scalar_t* p, scalar_t v, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

extern "C" int pp_alpha2weight(pp_dtype dtype, const void* alpha, const int64_t* ray_id,
                               int64_t n_pts, int64_t n_rays, void* weight, void* T,
                               void* alphainv_last, int64_t* i_start, int64_t* i_end,
                               pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0 && n_rays >= 0, PP_ERR_INVALID_ARG, "alpha2weight: negative extent");
  PP_REQUIRE(n_pts <= INT32_MAX, PP_ERR_TOO_LARGE, "alpha2weight: n_pts exceeds 32-bit index space");
  cudaStream_t s = as_stream(stream);
  if (n_rays == 0) return PP_OK;
  PP_REQUIRE(alphainv_last && i_start && i_end, PP_ERR_INVALID_ARG, "alpha2weight: null pointer");
  PP_CHECK_CUDA(cudaMemsetAsync(i_start, 0, sizeof(int64_t) * n_rays, s));
  PP_CHECK_CUDA(cudaMemsetAsync(i_end, 0, sizeof(int64_t) * n_rays, s));
  if (n_pts == 0) {
    // reference early return (:629-631): alphainv_last = 1, i_start = i_end = 0
    PP_DISPATCH(dtype, {
      fill_kernel<scalar_t><<<ceil_div(n_rays, 256), 256, 0, s>>>((scalar_t*)alphainv_last,
                                                                 scalar_t(1), n_rays);
    });
    PP_CHECK_LAUNCH("alpha2weight(fill)");
    return PP_OK;
  }
  PP_REQUIRE(alpha && ray_id && weight && T, PP_ERR_INVALID_ARG, "alpha2weight: null pointer");
  segment_bounds_kernel<<<ceil_div(n_pts, 256), 256, 0, s>>>(ray_id, n_pts, n_rays, i_start, i_end);
  PP_CHECK_LAUNCH("alpha2weight(bounds)");
  PP_DISPATCH(dtype, {
    constexpr int kT = sizeof(scalar_t) == 4 ? 16 : 8;
    alpha2weight_fwd_kernel<scalar_t, kT><<<ceil_div(n_rays, kWarp), (1 + kFwdMemWarps) * kWarp, 0, s>>>(
        (const scalar_t*)alpha, n_rays, (scalar_t*)weight, (scalar_t*)T, (scalar_t*)alphainv_last,
        i_start, i_end);
  });
  PP_CHECK_LAUNCH("alpha2weight");
  return PP_OK;
}

extern "C" int pp_alpha2weight_backward(pp_dtype dtype, const void* alpha, const void* weight,
                                        const void* T, const void* alphainv_last,
                                        const int64_t* i_start, const int64_t* i_end,
                                        int64_t n_rays, int64_t n_pts, const void* grad_weights,
                                        const void* grad_last, void* grad, pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0 && n_rays >= 0, PP_ERR_INVALID_ARG, "alpha2weight_backward: negative extent");
  PP_REQUIRE(n_pts <= INT32_MAX, PP_ERR_TOO_LARGE, "alpha2weight_backward: n_pts exceeds 32-bit index space");
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(grad, PP_ERR_INVALID_ARG, "alpha2weight_backward: grad is null");
  cudaStream_t s = as_stream(stream);
  const size_t esz = dtype == PP_F64 ? 8 : 4;
  // samples past a ray's early-termination index get zero gradient (reference: zeros_like, :684)
  PP_CHECK_CUDA(cudaMemsetAsync(grad, 0, esz * n_pts, s));
  if (n_rays == 0) return PP_OK;
  PP_REQUIRE(alpha && weight && T && alphainv_last && i_start && i_end && grad_weights && grad_last,
             PP_ERR_INVALID_ARG, "alpha2weight_backward: null pointer");
  PP_DISPATCH(dtype, {
    constexpr int kT = sizeof(scalar_t) == 4 ? 16 : 8;
    alpha2weight_bwd_kernel<scalar_t, kT><<<ceil_div(n_rays, kWarp), (1 + kBwdMemWarps) * kWarp, 0, s>>>(
        (const scalar_t*)alpha, (const scalar_t*)weight, (const scalar_t*)T,
        (const scalar_t*)alphainv_last, i_start, i_end, n_rays, (const scalar_t*)grad_weights,
        (const scalar_t*)grad_last, (scalar_t*)grad);
  });
  PP_CHECK_LAUNCH("alpha2weight_backward");
  return PP_OK;
}

extern "C" int pp_dense_composite_forward(const float* alpha, int64_t n_rays, int32_t n_samples, float* weights,
                                          float* alphainv_cum, pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0 && n_samples >= 0, PP_ERR_INVALID_ARG, "dense_composite_forward: negative extent");
  if (n_rays == 0) return PP_OK;
  PP_REQUIRE(alphainv_cum && (n_samples == 0 || (alpha && weights)), PP_ERR_INVALID_ARG,
             "dense_composite_forward: null pointer");
  dense_composite_fwd_kernel<<<ceil_div(n_rays, kDenseWarps), kDenseWarps * kWarp, 0, as_stream(stream)>>>(
      alpha, n_rays, n_samples, weights, alphainv_cum);
  PP_CHECK_LAUNCH("dense_composite_forward");
  return PP_OK;
}

extern "C" int pp_dense_composite_backward(const float* alpha, const float* alphainv_cum, int64_t n_rays,
                                           int32_t n_samples, const float* grad_weights,
                                           const float* grad_alphainv_cum, float* grad_alpha, pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0 && n_samples >= 0, PP_ERR_INVALID_ARG, "dense_composite_backward: negative extent");
  if (n_rays == 0 || n_samples == 0) return PP_OK;
  PP_REQUIRE(alpha && alphainv_cum && grad_alpha, PP_ERR_INVALID_ARG, "dense_composite_backward: null pointer");
  dense_composite_bwd_kernel<<<ceil_div(n_rays, kDenseWarps), kDenseWarps * kWarp, 0, as_stream(stream)>>>(
      alpha, alphainv_cum, n_rays, n_samples, grad_weights, grad_alphainv_cum, grad_alpha);
  PP_CHECK_LAUNCH("dense_composite_backward");
  return PP_OK;
}

extern "C" int pp_segment_sum(const float* src, const float* weight, const int64_t* ray_id,
                              int64_t n_pts, int32_t n_ch, int64_t n_rays, float* out,
                              pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0 && n_rays >= 0 && n_ch >= 1, PP_ERR_INVALID_ARG, "segment_sum: bad extent");
  if (n_rays == 0) return PP_OK;
  PP_REQUIRE(out, PP_ERR_INVALID_ARG, "segment_sum: out is null");
  cudaStream_t s = as_stream(stream);
  PP_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * n_rays * n_ch, s));
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(src && ray_id, PP_ERR_INVALID_ARG, "segment_sum: null pointer");
  segment_sum_kernel<<<ceil_div(n_pts, 256), 256, 0, s>>>(src, weight, ray_id, n_pts, n_ch,
                                                             n_rays, out);
  PP_CHECK_LAUNCH("segment_sum");
  return PP_OK;
}

extern "C" int64_t pp_segment_cumsum_scratch_bytes(int64_t n_pts) {
  const int64_t blocks = (n_pts + kScanElems - 1) / kScanElems;
  return blocks * (int64_t)(sizeof(float) + sizeof(int));
}

extern "C" int pp_segment_cumsum(const float* src, const int64_t* ray_id, int64_t n_pts,
                                 float* out, void* scratch, pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0, PP_ERR_INVALID_ARG, "segment_cumsum: n_pts < 0");
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(src && ray_id && out && scratch, PP_ERR_INVALID_ARG, "segment_cumsum: null pointer");
  cudaStream_t s = as_stream(stream);
  const int64_t blocks = (n_pts + kScanElems - 1) / kScanElems;
  float* blk_tail = reinterpret_cast<float*>(scratch);
  int* blk_head = reinterpret_cast<int*>(blk_tail + blocks);
  segcumsum_summary_kernel<<<(unsigned)blocks, kScanThreads, 0, s>>>(src, ray_id, n_pts, blk_tail,
                                                                     blk_head);
  PP_CHECK_LAUNCH("segment_cumsum(summary)");
  segcumsum_apply_kernel<<<(unsigned)blocks, kScanThreads, 0, s>>>(src, ray_id, n_pts, blk_tail,
                                                                   blk_head, out);
  PP_CHECK_LAUNCH("segment_cumsum");
  return PP_OK;
}
