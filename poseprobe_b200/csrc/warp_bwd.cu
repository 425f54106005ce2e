// Deformation ("warp") network: weight / bias gradients from gradients w.r.t. its outputs AND its input Jacobian,
// on tcgen05 + TMEM, one pass, nothing but the rows and the output gradients read from HBM.
//
// The reference trains DeformedImplicitField (3 -> 128 -> 128 -> 128 -> 128 -> 4, ReLU; lib/deformation/
// deform_net.py:12-29, modules.py:80-161) through torch.autograd with create_graph=True: the loss depends on the
// network's value AND on d net / d x (normal of the warped SDF, grad_deform prior; lib/voxurf_coarse.py:958-985,
// lib/losses.py:6-24), so autograd runs a double backward through five Linear layers and the 60-kernel custom
// sampler.  Here every sample is FOUR rows (value row + three tangent rows, as in the forward kernel, mlp.cu kJac):
// a ReLU network is piecewise linear, so with the gates taken from the value row the four rows are four independent
// LINEAR chains and the exact parameter gradient is the ordinary layer-by-layer backward applied to all four rows
// (oracle/deform_analytic.py::warp_rows_backward; bias gradients from the value rows only).  d loss / d x through
// the network is J^T (d loss / d value) and is formed by the caller from the forward's Jacobian: no dX here.
//
// One persistent CTA per SM (16 epilogue warps + 1 MMA-issue warp) owns tiles of 128 rows = 32 samples:
//   forward recompute  H_0 .. H_3 (fp16 operands like the forward, fp32 accumulation, gates from the value lane)
//   dZ_4               = grad_out * 2^k (power-of-two loss scale that puts the tile into fp16's normal range)
//   per layer l=4..1   dH_{l-1} = dZ_l W_l   (W_l read MN-major from the forward's copy)
//                      dW_l    += dZ_l^T H_{l-1},  db_l += dZ_l^T 1_value   (contraction over the tile's rows)
//                      dZ_{l-1} = dH_{l-1} * gate_{l-1}   in place of H_{l-1}
//   dW_0 (+ db_0)     += dZ_0^T X   (X carries a constant 1 in column 6 of the value rows: its column of dW_0 is db_0)
// Budgets decide the schedule.  TMEM: three 128 x 128 fp32 weight-gradient accumulators are 384 of the 512 columns,
// two 16-column accumulators hold every small gradient (dW_4^T and db_1..3 share one through column-disjoint B tiles,
// dW_0 and db_0 the other), which leaves 96 columns for the working accumulator: every 128-wide product is issued as
// N = 96 (features 0..95) and, once the first 32 columns have been read out, N = 32 (features 96..127) into the same
// columns.  Shared memory: the fp16 weights (104 KB) leave room for three 32 KB activation slots, so H_0 is dropped
// after the forward and recomputed (K = 16, one MMA) just before dW_1 needs it: 9 accumulator drains per tile.
//   slots:  fwd  H_0 -> s0, H_1 -> s1, H_2 -> s2, H_3 -> s0
//           bwd  dZ_3 -> s0, dZ_2 -> s2, dZ_1 -> s1, H_0 -> s0 (recompute), dZ_0 -> s0
#include <stdlib.h>

#include "mlp_common.cuh"

namespace pp {

constexpr int kWbTile = 128;
constexpr int kWbThreads = 544;
constexpr int kWbDepth = 5;
// TMEM columns
constexpr int kWbAcc = 0;           // [0,96)    working accumulator
constexpr int kWbDw1 = 96;          // [96,480)  dW_1, dW_2, dW_3 (lane = output feature, column = input feature)
constexpr int kWbSmallA = 480;      // [480,496) columns 0..3: dW_4^T (lane = input feature); 5, 6, 7: db_1, db_2, db_3
constexpr int kWbSmallB = 496;      // [496,512) columns 0..5: dW_0 (x_hi | x_lo halves); 6: db_0
constexpr long long kWbWaitBudget = 600000000LL;

struct WarpBwdParams {
  const float* w[kWbDepth];
  const float* b[kWbDepth];
  float* gw[kWbDepth];      // gw[0] is [128,6] (hi | lo halves of the doubled first layer), gw[4] is [4,128]
  float* gb[kWbDepth];
};

__device__ __forceinline__ void wb_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ bool wb_wait(uint32_t bar, uint32_t parity, volatile int* dead) {
  if (*dead) return false;
  const long long t0 = clock64();
  for (int spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
    if ((spin & 127) == 127) {
      if (*dead) return false;
      if (clock64() - t0 > kWbWaitBudget) {
        *dead = 1;
        return false;
      }
    }
  }
}

// max |g| over the output gradients (float bit patterns of non-negative values order like unsigned integers)
__global__ void warpnet_absmax_kernel(const float4* __restrict__ g, int64_t n_rows, const int32_t* __restrict__ n_ptr,
                                      unsigned int* __restrict__ out) {
  if (n_ptr) n_rows = min64(n_rows, static_cast<int64_t>(*n_ptr));
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ld_stream(g + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && m < 3.0e38f) atomicMax(out, __float_as_uint(m));
}

__global__ void __launch_bounds__(kWbThreads, 1)
warpnet_backward_kernel(WarpBwdParams P, const __half* __restrict__ x, const float* __restrict__ grad_out, int64_t M,
                        const int32_t* __restrict__ n_ptr, const unsigned int* __restrict__ amax_ptr,
                        int* __restrict__ err_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (n_ptr) M = min64(M, static_cast<int64_t>(*n_ptr));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = (warp & 3) * 32 + lane;   // TMEM lane = tile row (epilogue warps 0..15); 4 rows per sample
  const int cg = (warp >> 2) & 3;           // quarter of the 128 features this warp reads out

  // ---- shared memory map ----
  uint32_t off = 0;
  uint8_t* sX[2] = {smem, smem + 4096};                 // input rows [128 x 16] fp16, double buffered
  off += 8192;
  uint8_t* sG = smem + off;                             // dZ_4 [128 x 16] fp16 (columns 0..3)
  off += 4096;
  uint8_t* sS[3];
  for (int i = 0; i < 3; ++i) {
    sS[i] = smem + off;
    off += kWbTile * kMlpWidth * 2;
  }
  uint8_t* sW[kWbDepth];
  sW[0] = smem + off;                                   // [128 x 16]
  off += 4096;
  for (int l = 1; l < 4; ++l) {
    sW[l] = smem + off;                                 // [128 x 128]
    off += kMlpWidth * kMlpWidth * 2;
  }
  sW[4] = smem + off;                                   // [16 x 128]
  off += 4096;
  uint8_t* sOnes[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int l = 1; l < 4; ++l) {
    sOnes[l] = smem + off;                              // [128 x 16]: 1 at (value rows, column 4 + l)
    off += 4096;
  }
  float* sBias = reinterpret_cast<float*>(smem + off);  // 4 x 128
  off += 4 * kMlpWidth * 4;
  float* sDb4 = reinterpret_cast<float*>(smem + off);
  off += 16;
  uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + off);   // accA, accB, q0, epi, dw, x
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(smem + off + 48);
  volatile int* sDead = reinterpret_cast<volatile int*>(smem + off + 52);

  // ---- prologue: weights -> fp16 canonical tiles, ones tiles, biases, TMEM, barriers ----
  for (int l = 0; l < kWbDepth; ++l) {
    const int rows = (l == 4) ? kMlpOutPad : kMlpWidth;
    const int real_rows = (l == 4) ? 4 : kMlpWidth;
    const int K = (l == 0) ? 16 : kMlpWidth;
    const int real_K = (l == 0) ? 6 : kMlpWidth;
    const int cpr = K >> 3;
    for (int e = tid; e < rows * cpr; e += kWbThreads) {
      const int r = e / cpr, k0 = (e - r * cpr) * 8;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (r < real_rows && k0 + j < real_K) ? __ldg(P.w[l] + r * real_K + k0 + j) : 0.f;
      *reinterpret_cast<uint4*>(sW[l] + canon_off(r, k0, K)) =
          make_uint4(pack_f16(f[0], f[1]), pack_f16(f[2], f[3]), pack_f16(f[4], f[5]), pack_f16(f[6], f[7]));
    }
    if (l < 4)
      for (int e = tid; e < kMlpWidth; e += kWbThreads) sBias[l * kMlpWidth + e] = P.b[l][e];
  }
  for (int l = 1; l < 4; ++l)
    for (int e = tid; e < kWbTile * 2; e += kWbThreads) {
      const int r = e >> 1, c8 = (e & 1) * 8;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if ((r & 3) == 0 && c8 == 0) {     // column 4 + l of the value rows (fp16 1.0 = 0x3C00)
        if (l == 1) v.z = 0x3C000000u;   // column 5
        if (l == 2) v.w = 0x00003C00u;   // column 6
        if (l == 3) v.w = 0x3C000000u;   // column 7
      }
      *reinterpret_cast<uint4*>(sOnes[l] + canon_off(r, c8, 16)) = v;
    }
  if (tid < 4) sDb4[tid] = 0.f;
  if (tid == 0) *sDead = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(smem_u32(sBar + 0), 1);
    mbar_init(smem_u32(sBar + 1), 1);
    mbar_init(smem_u32(sBar + 2), 128);
    mbar_init(smem_u32(sBar + 3), 512);
    mbar_init(smem_u32(sBar + 4), 1);
    mbar_init(smem_u32(sBar + 5), 512);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *sTmem;
  const uint32_t bAccA = smem_u32(sBar + 0), bAccB = smem_u32(sBar + 1), bQ0 = smem_u32(sBar + 2);
  const uint32_t bEpi = smem_u32(sBar + 3), bDw = smem_u32(sBar + 4), bX = smem_u32(sBar + 5);
  const int64_t n_tiles = (M + kWbTile - 1) / kWbTile;
  // power-of-two loss scale: the largest output gradient lands at 2^5..2^6 in fp16
  float scale = 1.f, inv_scale = 1.f;
  {
    const float amax = __uint_as_float(*amax_ptr);
    if (amax > 0.f) {
      int e = static_cast<int>(floorf(log2f(64.f / amax)));
      e = max(-100, min(100, e));
      scale = exp2f(static_cast<float>(e));
      inv_scale = exp2f(static_cast<float>(-e));
    }
  }
  bool first = true;

  if (warp == 16) {
    // =========================== MMA issuer (whole warp, elected lane issues) ===========================
    uint32_t pq = 0, pe = 0, px = 0;
    auto wait_epi = [&]() {
      wb_wait(bEpi, pe, sDead);
      pe ^= 1;
      tc_fence_after();
    };
    auto wait_q0 = [&]() {
      wb_wait(bQ0, pq, sDead);
      pq ^= 1;
      tc_fence_after();
    };
    auto group = [&](uint32_t d_tmem, uint64_t a0, uint32_t a_step, uint64_t b0, uint32_t b_step, uint32_t idesc, int n,
                     uint32_t acc_first) {
#pragma unroll 4
      for (int k = 0; k < n; ++k) {
        umma_bf16(d_tmem, a0, b0, idesc, (k > 0) ? 1u : acc_first);
        a0 += a_step;
        b0 += b_step;
      }
    };
    constexpr uint32_t kStepK = 256 >> 4;               // K-major operand: next 16 columns
    constexpr uint32_t kStepMn128 = (2u * 2048u) >> 4;  // MN-major, 128-column tile: next 16 rows
    constexpr uint32_t kStepMn16 = (2u * 256u) >> 4;    // MN-major, 16-column tile
    constexpr uint32_t kN96K128 = 24576 >> 4;           // K-major [128 x 128] B operand: rows 96..127
    constexpr uint32_t kN96K16 = 3072 >> 4;             // K-major [128 x 16] B operand: rows 96..127
    constexpr uint32_t kN96Mn = 1536 >> 4;              // MN-major B operand: columns 96..127
    const uint32_t id_f96 = umma_idesc_f16(kWbTile, 96), id_f32 = umma_idesc_f16(kWbTile, 32);
    const uint32_t id_b96 = id_f96 | (1u << 16), id_b32 = id_f32 | (1u << 16);          // B MN-major
    const uint32_t id_w128 = umma_idesc_f16(kWbTile, 128) | (1u << 15) | (1u << 16);    // A and B MN-major
    const uint32_t id_w16 = umma_idesc_f16(kWbTile, 16) | (1u << 15) | (1u << 16);
    uint64_t dSk[3], dSmn[3], dWk[4], dWmn[5];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      dSk[i] = desc_kmajor(smem_u32(sS[i]), kMlpWidth, 0);
      dSmn[i] = desc_mnmajor(smem_u32(sS[i]), kMlpWidth, 0);
    }
    dWk[0] = desc_kmajor(smem_u32(sW[0]), 16, 0);
#pragma unroll
    for (int l = 1; l < 4; ++l) {
      dWk[l] = desc_kmajor(smem_u32(sW[l]), kMlpWidth, 0);
      dWmn[l] = desc_mnmajor(smem_u32(sW[l]), kMlpWidth, 0);
    }
    dWmn[4] = desc_mnmajor(smem_u32(sW[4]), kMlpWidth, 0);
    dWmn[0] = 0;
    const uint64_t dGk = desc_kmajor(smem_u32(sG), 16, 0), dGmn = desc_mnmajor(smem_u32(sG), 16, 0);
    const uint64_t dXk[2] = {desc_kmajor(smem_u32(sX[0]), 16, 0), desc_kmajor(smem_u32(sX[1]), 16, 0)};
    const uint64_t dXmn[2] = {desc_mnmajor(smem_u32(sX[0]), 16, 0), desc_mnmajor(smem_u32(sX[1]), 16, 0)};
    const uint64_t dOnes[4] = {0, desc_mnmajor(smem_u32(sOnes[1]), 16, 0), desc_mnmajor(smem_u32(sOnes[2]), 16, 0),
                               desc_mnmajor(smem_u32(sOnes[3]), 16, 0)};
    const uint32_t acc = tmem + kWbAcc;

    // layer 0 (K = 16): H_0 = gate(X W_0^T + b_0)
    auto layer0 = [&](uint64_t xk) {
      umma_bf16(acc, xk, dWk[0], id_f96, 0);
      umma_commit(bAccA);
      wait_q0();
      umma_bf16(acc, xk, dWk[0] + kN96K16, id_f32, 0);
      umma_commit(bAccB);
    };
    // forward hidden layer: acc = S_in W_l^T
    auto fwd_layer = [&](int in, int l) {
      group(acc, dSk[in], kStepK, dWk[l], kStepK, id_f96, 8, 0);
      umma_commit(bAccA);
      wait_q0();
      group(acc, dSk[in], kStepK, dWk[l] + kN96K128, kStepK, id_f32, 8, 0);
      umma_commit(bAccB);
    };
    // backward hidden layer l (dZ_l in slot dz, H_{l-1} in slot hin): acc = dZ_l W_l; dW_l += dZ_l^T H_{l-1}; db_l
    auto bwd_layer = [&](int dz, int hin, int l, uint32_t accf) {
      group(acc, dSk[dz], kStepK, dWmn[l], kStepMn128, id_b96, 8, 0);
      umma_commit(bAccA);
      // the bias-gradient product fills the tensor pipe while quarter 0 reads its columns out; the weight-gradient
      // product runs underneath the last quarter's read-out (its own barrier: the slot it reads is overwritten next)
      group(tmem + kWbSmallA, dSmn[dz], kStepMn128, dOnes[l], kStepMn16, id_w16, 8, 1);
      wait_q0();
      group(acc, dSk[dz], kStepK, dWmn[l] + kN96Mn, kStepMn128, id_b32, 8, 0);
      umma_commit(bAccB);
      group(tmem + kWbDw1 + (l - 1) * kMlpWidth, dSmn[dz], kStepMn128, dSmn[hin], kStepMn128, id_w128, 8, accf);
      umma_commit(bDw);
    };

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int xb = it & 1;
      const uint32_t accf = first ? 0u : 1u;
      wb_wait(bX, px, sDead);
      px ^= 1;
      tc_fence_after();
      layer0(xb ? dXk[1] : dXk[0]);        // S0: H_0 -> s0
      wait_epi();
      fwd_layer(0, 1);                     // S1: H_1 -> s1
      wait_epi();
      fwd_layer(1, 2);                     // S2: H_2 -> s2
      wait_epi();
      fwd_layer(2, 3);                     // S3: H_3 -> s0
      wait_epi();
      // S4: dW_4^T += H_3^T dZ_4 (first, so that the accumulator's commit also covers it); dH_3 = dZ_4 W_4 (K = 16)
      group(tmem + kWbSmallA, dSmn[0], kStepMn128, dGmn, kStepMn16, id_w16, 8, accf);
      umma_bf16(acc, dGk, dWmn[4], id_b96, 0);
      umma_commit(bAccA);
      wait_q0();
      umma_bf16(acc, dGk, dWmn[4] + kN96Mn, id_b32, 0);
      umma_commit(bAccB);
      wait_epi();                          // dZ_3 in s0
      bwd_layer(0, 2, 3, accf);            // S5: dH_2; dW_3 += dZ_3^T H_2
      wait_epi();                          // dZ_2 in s2
      bwd_layer(2, 1, 2, accf);            // S6: dH_1; dW_2 += dZ_2^T H_1
      wait_epi();                          // dZ_1 in s1
      layer0(xb ? dXk[1] : dXk[0]);        // S7: H_0 again -> s0
      wait_epi();
      bwd_layer(1, 0, 1, accf);            // S8: dH_0; dW_1 += dZ_1^T H_0
      wait_epi();                          // dZ_0 in s0
      group(tmem + kWbSmallB, dSmn[0], kStepMn128, xb ? dXmn[1] : dXmn[0], kStepMn16, id_w16, 8, accf);
      umma_commit(bDw);
      first = false;
    }
  } else {
    // =========================== epilogue warps ===========================
    uint32_t pa = 0, pb = 0, pd = 0, pq = 0;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int c0 = cg * 32;                                   // features this thread handles
    const uint32_t t_acc = tmem + kWbAcc + lane_base + (cg < 3 ? c0 : 0);
    const int vlane = lane & ~3;
    const float bsel = (lane & 3) == 0 ? 1.f : 0.f;
    const __half2 zero2 = __floats2half2_rn(0.f, 0.f);

    // accumulator quarter -> registers.  Quarters 0..2 read the N = 96 product; quarter 3 reads the N = 32 product
    // that re-uses columns [0,32) once quarter 0 has them in registers.
    auto load_acc = [&](uint32_t (&r)[32]) {
      if (cg < 3) {
        wb_wait(bAccA, pa, sDead);
        pa ^= 1;
        if (cg != 0) {
          // quarter 0 goes first: the TMEM read port is shared, and the second product (and with it the whole
          // step) starts when quarter 0's columns are in registers - 16 KB alone instead of a third of 48 KB
          wb_wait(bQ0, pq, sDead);
          pq ^= 1;
        }
      } else {
        wb_wait(bAccB, pb, sDead);
        pb ^= 1;
      }
      tc_fence_after();
      tmem_ld32_nowait(t_acc, r);
    };
    auto after_load = [&]() {
      tmem_ld_wait();
      if (cg == 0) {
        tc_fence_before();
        wb_arrive(bQ0);
      }
    };
    // every MMA of the step (second product, weight gradients) has completed once accB has
    auto before_store = [&]() {
      if (cg < 3) {
        wb_wait(bAccB, pb, sDead);
        pb ^= 1;
      }
    };
    auto publish = [&]() {
      tc_fence_before();
      fence_async_smem();
      wb_arrive(bEpi);
    };
    auto stage = [&](int64_t tile, uint8_t* dst) {
      const uint32_t base = smem_u32(dst);
      if (tid < kWbTile * 2) {
        const int r = tid >> 1, c8 = tid & 1;
        const int64_t gr = tile * kWbTile + r;
        const bool ok = gr < M;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(base + canon_off(r, c8 * 8, 16)),
                     "l"(x + (ok ? gr : 0) * 16 + c8 * 8), "r"(ok ? 16u : 0u)
                     : "memory");
      }
    };
    // forward-type epilogue: slot <- gate(acc + bias) in fp16 (value lane: ReLU; tangent lanes: the value lane's gate)
    auto epi_forward = [&](int l, uint8_t* dst, bool wait_dw) {
      uint32_t r[32];
      load_acc(r);
      const float2* bf2 = reinterpret_cast<const float2*>(sBias + l * kMlpWidth + c0);
      float2 bb[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) bb[j] = bf2[j];
      after_load();
      const float2 bs2 = make_float2(bsel, bsel);
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 a = make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
        const __half2 xh = __float22half2_rn(__ffma2_rn(bb[j], bs2, a));
        const uint32_t xu = *reinterpret_cast<const uint32_t*>(&xh);
        const uint32_t vu = __shfl_sync(kFull, xu, vlane);
        const __half2 o = __hmul2(xh, __hgt2(*reinterpret_cast<const __half2*>(&vu), zero2));
        pk[j] = *reinterpret_cast<const uint32_t*>(&o);
      }
      before_store();
      if (wait_dw) {     // the previous tile's dW_0 read s0 and the other input buffer
        wb_wait(bDw, pd, sDead);
        pd ^= 1;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(dst + canon_off(row, c0 + 8 * j, kMlpWidth)) =
            make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      publish();
    };
    // backward-type epilogue: slot <- acc * gate(slot's value lane), in place.  wait_dw: the step's weight-gradient
    // product reads this slot (as H_{l-1}) and is committed on its own barrier
    auto epi_backward = [&](uint8_t* slot, bool wait_dw) {
      uint32_t r[32];
      load_acc(r);
      uint4 act[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) act[j] = *reinterpret_cast<const uint4*>(slot + canon_off(row, c0 + 8 * j, kMlpWidth));
      after_load();
      const uint32_t* au = reinterpret_cast<const uint32_t*>(act);
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const __half2 dh = __floats2half2_rn(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
        const uint32_t vu = __shfl_sync(kFull, au[j], vlane);
        const __half2 o = __hmul2(dh, __hgt2(*reinterpret_cast<const __half2*>(&vu), zero2));
        pk[j] = *reinterpret_cast<const uint32_t*>(&o);
      }
      before_store();
      if (wait_dw) {
        wb_wait(bDw, pd, sDead);
        pd ^= 1;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(slot + canon_off(row, c0 + 8 * j, kMlpWidth)) =
            make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      publish();
    };

    if (blockIdx.x < n_tiles) stage(blockIdx.x, sX[0]);
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int64_t grow = tile * kWbTile + row;
      const bool valid = grow < M;
      asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
      if (cg == 0) {
        // dZ_4 tile: the row's four output gradients, scaled, fp16; db_4 from the value rows (unscaled fp32)
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) g = ld_stream(reinterpret_cast<const float4*>(grad_out) + grow);
        *reinterpret_cast<uint4*>(sG + canon_off(row, 0, 16)) =
            make_uint4(pack_f16(g.x * scale, g.y * scale), pack_f16(g.z * scale, g.w * scale), 0u, 0u);
        *reinterpret_cast<uint4*>(sG + canon_off(row, 8, 16)) = make_uint4(0u, 0u, 0u, 0u);
        const float gv[4] = {g.x * bsel, g.y * bsel, g.z * bsel, g.w * bsel};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float s = warp_sum(gv[j]);
          if (lane == 0) atomicAdd(&sDb4[j], s);
        }
      }
      fence_async_smem();
      wb_arrive(bX);

      epi_forward(0, sS[0], !first);                                               // S0
      // the other input buffer is free now (dW_0 of the previous tile has completed)
      if (tile + gridDim.x < n_tiles) stage(tile + gridDim.x, sX[(it + 1) & 1]);
      epi_forward(1, sS[1], false);                                                // S1
      epi_forward(2, sS[2], false);                                                // S2
      epi_forward(3, sS[0], false);                                                // S3
      epi_backward(sS[0], false);                                                  // S4: dZ_3
      epi_backward(sS[2], true);                                                   // S5: dZ_2 (dW_3 read H_2)
      epi_backward(sS[1], true);                                                   // S6: dZ_1 (dW_2 read H_1)
      epi_forward(0, sS[0], false);                                                // S7: H_0 again
      epi_backward(sS[0], true);                                                   // S8: dZ_0 (dW_1 read H_0)
      first = false;
    }
    if (!first) {    // the last tile's dW_0 before the accumulators are read out
      wb_wait(bDw, pd, sDead);
      pd ^= 1;
    }
  }

  // ---- the CTA's gradients: TMEM -> global (fp32 atomics, once per CTA), loss scale removed ----
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const bool dead = *sDead != 0;
  if (dead && tid == 0 && err_flag) atomicExch(err_flag, 1);
  if (warp < 16 && !first && !dead) {
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int c0 = cg * 32;
    for (int l = 1; l < 4; ++l) {
      float v[32];
      tmem_ld32(tmem + kWbDw1 + (l - 1) * kMlpWidth + lane_base + c0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) atomicAdd(P.gw[l] + row * kMlpWidth + c0 + j, v[j] * inv_scale);
    }
    if (cg == 0) {
      float v[16];
      tmem_ld16(tmem + kWbSmallA + lane_base, v);
#pragma unroll
      for (int o = 0; o < 4; ++o) atomicAdd(P.gw[4] + o * kMlpWidth + row, v[o] * inv_scale);
#pragma unroll
      for (int l = 1; l < 4; ++l) atomicAdd(P.gb[l] + row, v[4 + l] * inv_scale);
    } else if (cg == 1) {
      float v[16];
      tmem_ld16(tmem + kWbSmallB + lane_base, v);
#pragma unroll
      for (int c = 0; c < 6; ++c) atomicAdd(P.gw[0] + row * 6 + c, v[c] * inv_scale);
      atomicAdd(P.gb[0] + row, v[6] * inv_scale);
    }
    if (tid < 4) atomicAdd(P.gb[4] + tid, sDb4[tid]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace pp

using namespace pp;

static int* g_wb_err_flag = nullptr;

extern "C" int pp_warpnet_backward_stalled(void) {
  if (!g_wb_err_flag) return 0;
  int v = 0;
  if (cudaMemcpy(&v, g_wb_err_flag, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v;
}

extern "C" int64_t pp_warpnet_backward_workspace_bytes(void) { return 16; }

extern "C" int pp_warpnet_backward_packed(const void* x_rows, const float* grad_out, int64_t n_rows,
                                          const int32_t* n_rows_ptr, int32_t depth, int32_t width,
                                          const float* const* weights, const float* const* biases,
                                          float* const* grad_weights, float* const* grad_biases, void* workspace,
                                          pp_stream_t stream) {
  PP_REQUIRE(n_rows >= 0 && (n_rows & 3) == 0, PP_ERR_INVALID_ARG, "warpnet_backward: n_rows must be a multiple of 4");
  PP_REQUIRE(width == kMlpWidth, PP_ERR_UNSUPPORTED, "warpnet_backward: hidden width must be %d", kMlpWidth);
  PP_REQUIRE(depth == kWbDepth, PP_ERR_UNSUPPORTED,
             "warpnet_backward: the deformation network has %d Linear layers (3 -> 128 x 4 -> 4), got %d", kWbDepth,
             depth);
  if (n_rows == 0) return PP_OK;
  PP_REQUIRE(x_rows && grad_out && weights && biases && grad_weights && grad_biases && workspace, PP_ERR_INVALID_ARG,
             "warpnet_backward: null pointer");
  WarpBwdParams P;
  for (int l = 0; l < kWbDepth; ++l) {
    P.w[l] = weights[l];
    P.b[l] = biases[l];
    P.gw[l] = grad_weights[l];
    P.gb[l] = grad_biases[l];
    PP_REQUIRE(P.w[l] && P.b[l] && P.gw[l] && P.gb[l], PP_ERR_INVALID_ARG, "warpnet_backward: null layer %d", l);
  }
  const size_t smem = 8192 + 4096 + 3 * 32768 + 4096 + 3 * 32768 + 4096 + 3 * 4096 + 2048 + 16 + 64;
  static_assert(8192 + 4096 + 3 * 32768 + 4096 + 3 * 32768 + 4096 + 3 * 4096 + 2048 + 16 + 64 <= 227 * 1024,
                "warpnet_backward: shared memory");
  static int* err_flag = nullptr;
  if (!err_flag) {
    PP_CHECK_CUDA(cudaMalloc(&err_flag, sizeof(int)));
    PP_CHECK_CUDA(cudaMemset(err_flag, 0, sizeof(int)));
    g_wb_err_flag = err_flag;
  }
  unsigned int* amax = static_cast<unsigned int*>(workspace);
  PP_CHECK_CUDA(cudaMemsetAsync(amax, 0, sizeof(unsigned int), as_stream(stream)));
  {
    const int blocks = (int)min64((n_rows + 255) / 256, (int64_t)sm_count() * 8);
    warpnet_absmax_kernel<<<blocks, 256, 0, as_stream(stream)>>>((const float4*)grad_out, n_rows, n_rows_ptr, amax);
    PP_CHECK_LAUNCH("warpnet_absmax");
  }
  const int64_t n_tiles = (n_rows + kWbTile - 1) / kWbTile;
  const int blocks = (int)min64(n_tiles, sm_count());
  PP_CHECK_CUDA(cudaFuncSetAttribute(warpnet_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  warpnet_backward_kernel<<<blocks, kWbThreads, smem, as_stream(stream)>>>(P, (const __half*)x_rows, grad_out, n_rows,
                                                                          n_rows_ptr, amax, err_flag);
  PP_CHECK_LAUNCH("warpnet_backward");
  return PP_OK;
}
