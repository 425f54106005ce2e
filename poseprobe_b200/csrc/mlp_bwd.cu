// rgbnet (colour MLP) training step on the 5th-generation tensor cores: forward recompute + full backward
// (input gradient, weight gradients, bias gradients) of one 128-sample tile without leaving the SM.
//
// The reference trains the colour network through autograd over nn.Linear / ReLU (cuBLAS sgemm) with the
// [M,d0] feature matrix, three [M,128] activations and their gradients round-tripping HBM
// (lib/voxurf_coarse.py:208-216, :1032-1033).  Here one persistent CTA per SM (16 epilogue warps + 1 MMA-issue
// warp) owns tiles of 128 samples:
//   forward   : as mlp.cu, but every hidden activation tile H_l stays in shared memory (bf16, UMMA canonical
//               layout) - they are the operands of the backward;
//   dZ_L      : (d loss / d rgb) * rgb (1 - rgb), rounded to bf16 into a [128 x 16] tile;
//   per layer : dH_{l-1} = dZ_l W_l            A = dZ_l (K-major), B = W_l read MN-major from the SAME
//                                              shared-memory copy the forward used (no transposed copy)
//               dW_l    += dZ_l^T H_{l-1}      both operands read MN-major from the activation tiles: the
//                                              contraction runs over the tile's 128 samples
//               db_l    += dZ_l^T 1            an N=16 MMA against a tile of ones
//               dZ_{l-1} = dH_{l-1} * (H_{l-1} > 0) is written in place of H_{l-1};
//   the weight / bias gradients of ALL the CTA's tiles accumulate in TMEM (fp32, up to 496 of the 512 columns)
//   and are added to global memory once per CTA at the end.
// MMAs that nothing waits for (dW, db) are issued behind the dH MMA of the same layer and run while the
// epilogue warps turn dH into the next dZ.  Input rows arrive as bf16 [M, k0p] (k0p = 64 or 128; the feature
// kernel of the fused path writes that format directly, the fp32 [n,d0] entry point converts once) and are
// moved with 16-byte cp.async straight into the UMMA layout of the NEXT tile's buffer while the current tile
// computes; the input gradient leaves as fp32 [M, k0p] rows.  HBM traffic per sample: k0p*2 B in, 12 B
// gradient in, k0p*4 B input gradient out; activations and their gradients never leave the SM.
#include <stdlib.h>

#include "mlp_common.cuh"

namespace pp {

constexpr int kBwdMaxDepth = 4;
constexpr int kTile = 128;
// TMEM columns (fp32, 128 lanes)
constexpr int kColAcc = 0;        // [0,128)   working accumulator: layer outputs, dH, dX
constexpr int kColOut = 0;        // [0,16)    logits: aliases the working accumulator (free at that point)
constexpr int kColDb = 128;       // [128,176) hidden-layer bias gradients, 16 columns per layer (all equal)
constexpr int kColDwLast = 176;   // [176,192) dW_last^T  (lane = input feature, column = output)
constexpr int kColDw0 = 192;      // [192,192+k0p) dW_0; then 128 columns per further hidden layer

// optional fused inputs / outputs of the fine-stage step (all null / 0 for the plain operator)
struct MlpFusedIO {
  const int32_t* n_ptr;     // device row count (compacted samples), clamps M
  const float* w;           // per-sample compositing weight
  const int32_t* ray_id;    // per-sample ray
  const float* g_ray;       // [N,3] d loss / d rgb_marched: grad_out_i = w_i * g_ray[ray_i]
  float* gw_out;            // [M] sum_c g_ray[ray_i,c] * rgb_i,c  (d loss / d w_i through the colour sum)
  int dx_cols;              // columns of the input gradient to store (row pitch): k0p or 16
};

struct MlpBwdParams {
  const float* w[kBwdMaxDepth];
  const float* b[kBwdMaxDepth];
  float* gw[kBwdMaxDepth];
  float* gb[kBwdMaxDepth];
  int depth;      // number of Linear layers (2..4)
  int d0;         // input features
  int k0p;        // d0 padded to 64 or 128
};

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// =====================================================================================================
// Schedule: warp-specialised and half-pipelined.
//
// The epilogue of a 128 x 128 fp32 accumulator is bounded by the TMEM read port (64 B/clk = 1024 cycles), the
// layer's MMAs take 516 cycles; a schedule that runs the two strictly one after the other with a block barrier in
// between (the first version of this kernel, 1.79 ms for 2.73 M samples against 1.77 ms here) leaves both idle
// half of the time.  Here warp 16 only issues MMAs and the 16 epilogue warps are split in two halves by accumulator
// column (0..63 / 64..127), each with its own "drained / written" barrier.  The next layer's N=128 MMAs overwrite the
// whole working accumulator, so they are issued once both halves have published; the layer's dW / db MMAs are
// queued behind them and run underneath the next drain.  (Splitting every layer into N=64 groups so that half 0's
// successor can start early was measured: same time, 50 % more MMA issue slots - N=64 costs 48 cycles, N=128 64.)
// Synchronisation is mbarrier-only (no block barrier inside a tile):
//     acc_ready[h]  MMA -> epilogue half h   (tcgen05.commit)
//     epi_done[h]   epilogue half h -> MMA   (256 arrivals: operand tile written / accumulator half drained)
//     dw_done       MMA -> epilogue          (the layer's dW / db MMAs finished reading the tiles they overwrite next)
//     x_done        epilogue -> MMA          (512 arrivals: the tile's input rows landed)
// Every wait carries a cycle-count bail-out that raises a flag instead of hanging the device.
constexpr int kBwd2Threads = 544;
constexpr long long kWaitBudgetCycles = 600000000LL;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ bool mbar_wait_guard(uint32_t bar, uint32_t parity, volatile int* dead) {
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
      if (clock64() - t0 > kWaitBudgetCycles) {
        *dead = 1;
        return false;
      }
    }
  }
}

__global__ void __launch_bounds__(kBwd2Threads, 1)
rgbnet_backward2_kernel(MlpBwdParams P, const __nv_bfloat16* __restrict__ x, const float* __restrict__ grad_out,
                        int64_t M, int apply_sigmoid, float* __restrict__ grad_x, float* __restrict__ rgb_out,
                        MlpFusedIO F, int* __restrict__ err_flag, long long* __restrict__ prof) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (F.n_ptr) M = min64(M, static_cast<int64_t>(*F.n_ptr));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = P.depth;
  long long pw[6] = {0, 0, 0, 0, 0, 0};     // cycle counters (PP_MLP_PROF): waits by category, total
  const long long t_begin = clock64();
  const int row = (warp & 3) * 32 + lane;   // TMEM lane = sample row (epilogue warps 0..15)
  const int cg = (warp >> 2) & 3;           // quarter of the accumulator's columns
  const int h = cg >> 1;                    // half (commit-barrier group)
  const int k0p = P.k0p, kq = k0p >> 2;

  // ---- shared memory map (as v1, plus the mbarriers) ----
  uint32_t off = 0;
  uint8_t* sXbuf[2];
  sXbuf[0] = smem;
  sXbuf[1] = smem + kTile * k0p * 2;
  off += 2 * kTile * k0p * 2;
  uint8_t* sH[kBwdMaxDepth];
  for (int l = 0; l < L - 1; ++l) {
    sH[l] = smem + off;
    off += kTile * kMlpWidth * 2;
  }
  uint8_t* sG = smem + off;
  off += kTile * kMlpOutPad * 2;
  uint8_t* sW[kBwdMaxDepth];
  for (int l = 0; l < L; ++l) {
    sW[l] = smem + off;
    const int rows = (l == L - 1) ? kMlpOutPad : kMlpWidth;
    const int K = (l == 0) ? k0p : kMlpWidth;
    off += rows * K * 2;
  }
  uint8_t* sOnes = smem + off;
  off += kTile * kMlpOutPad * 2;
  float* sBias = reinterpret_cast<float*>(smem + off);
  off += L * kMlpWidth * 4;
  float* sDbLast = reinterpret_cast<float*>(smem + off);
  off += 16;
  uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + off);   // [0,1] acc_ready, [2,3] epi_done, [4] dw_done, [5] x_done
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(smem + off + 48);
  volatile int* sDead = reinterpret_cast<volatile int*>(smem + off + 52);

  // ---- prologue ----
  for (int l = 0; l < L; ++l) {
    const bool last = (l == L - 1);
    const int rows = last ? kMlpOutPad : kMlpWidth;
    const int real_rows = last ? 3 : kMlpWidth;
    const int K = (l == 0) ? k0p : kMlpWidth;
    const int real_K = (l == 0) ? P.d0 : kMlpWidth;
    const int cpr = K >> 3;
    for (int e = tid; e < rows * cpr; e += kBwd2Threads) {
      const int r = e / cpr, k0 = (e - r * cpr) * 8;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (r < real_rows && k0 + j < real_K) ? __ldg(P.w[l] + r * real_K + k0 + j) : 0.f;
      uint4 pk;
      pk.x = pack_bf16(f[0], f[1]);
      pk.y = pack_bf16(f[2], f[3]);
      pk.z = pack_bf16(f[4], f[5]);
      pk.w = pack_bf16(f[6], f[7]);
      *reinterpret_cast<uint4*>(sW[l] + canon_off(r, k0, K)) = pk;
    }
    for (int e = tid; e < kMlpWidth; e += kBwd2Threads)
      sBias[l * kMlpWidth + e] = (e < real_rows) ? P.b[l][e] : 0.f;
  }
  for (int e = tid; e < kTile * kMlpOutPad; e += kBwd2Threads)
    reinterpret_cast<__nv_bfloat16*>(sOnes)[e] = __float2bfloat16_rn(1.f);
  if (tid < 4) sDbLast[tid] = 0.f;
  if (tid == 0) *sDead = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(smem_u32(sBar + 0), 1);
    mbar_init(smem_u32(sBar + 1), 1);
    mbar_init(smem_u32(sBar + 2), 256);
    mbar_init(smem_u32(sBar + 3), 256);
    mbar_init(smem_u32(sBar + 4), 1);
    mbar_init(smem_u32(sBar + 5), 512);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *sTmem;
  const uint32_t bAcc[2] = {smem_u32(sBar + 0), smem_u32(sBar + 1)};
  const uint32_t bEpi[2] = {smem_u32(sBar + 2), smem_u32(sBar + 3)};
  const uint32_t bDw = smem_u32(sBar + 4), bX = smem_u32(sBar + 5);
  const int64_t n_tiles = (M + kTile - 1) / kTile;
  bool first = true;

  if (warp == 16) {
    // =========================== MMA issuer ===========================
    // The whole warp runs the script on warp-uniform values and an elected lane issues (see umma_bf16): a loop run
    // by a single thread would cap the tensor pipe at one MMA per ~120 cycles.  The two shared-memory descriptors
    // of a group are built once and advanced by adding the slice stride to their address field (bits [0,14) of
    // the low word; shared-memory addresses never carry out of it).
    {
      uint32_t pe[2] = {0, 0}, px = 0;
      auto wait_epi = [&](int hh) {
        const long long t0 = clock64();
        mbar_wait_guard(bEpi[hh], pe[hh], sDead);
        pw[hh] += clock64() - t0;
        pe[hh] ^= 1;
        tc_fence_after();
      };
      // n MMAs over consecutive slices: (a0, b0) descriptors of the first slice, strides in 16-byte units
      auto group = [&](uint32_t d_tmem, uint64_t a0, uint32_t a_step, uint64_t b0, uint32_t b_step, uint32_t idesc,
                       int n, uint32_t acc_first) {
#pragma unroll 4
        for (int k = 0; k < n; ++k) {
          umma_bf16(d_tmem, a0, b0, idesc, (k > 0) ? 1u : acc_first);
          a0 += a_step;
          b0 += b_step;
        }
      };
      constexpr uint32_t kStepK = 256 >> 4;                        // K-major: next 16 columns
      const uint32_t kStepMn128 = (2u * 2048u) >> 4;               // MN-major, 128-column tile: next 16 rows
      const uint32_t kStepMn16 = (2u * 256u) >> 4;                 // MN-major, 16-column tile
      const uint32_t stepMnX = (2u * (static_cast<uint32_t>(k0p >> 3) << 7)) >> 4;   // MN-major, input tile / W_0
      const uint32_t id_f128 = umma_idesc_bf16(kTile, kMlpWidth);
      const uint32_t id_f16 = umma_idesc_bf16(kTile, kMlpOutPad);
      const uint32_t id_b128 = umma_idesc_bf16_major(kTile, kMlpWidth, 0, 1);
      const uint32_t id_w16 = umma_idesc_bf16_major(kTile, kMlpOutPad, 1, 1);
      const uint32_t id_w128 = umma_idesc_bf16_major(kTile, kMlpWidth, 1, 1);
      const uint32_t id_wx = umma_idesc_bf16_major(kTile, k0p, 1, 1);
      const uint32_t id_dx = umma_idesc_bf16_major(kTile, k0p, 0, 1);
      // every descriptor of the per-tile script, built once (registers; the loops below are fully unrolled)
      uint64_t dHk[kBwdMaxDepth], dHmn[kBwdMaxDepth], dWk[kBwdMaxDepth], dWmn[kBwdMaxDepth];
#pragma unroll
      for (int l = 0; l < kBwdMaxDepth; ++l) {
        const bool hid = l < L - 1;
        dHk[l] = hid ? desc_kmajor(smem_u32(sH[l]), kMlpWidth, 0) : 0;      // H_l / dZ_l, contraction over its columns
        dHmn[l] = hid ? desc_mnmajor(smem_u32(sH[l]), kMlpWidth, 0) : 0;    // same tile, contraction over the samples
        const int K = (l == 0) ? k0p : kMlpWidth;
        dWk[l] = (l < L) ? desc_kmajor(smem_u32(sW[l]), K, 0) : 0;          // forward: W_l K-major
        dWmn[l] = (l < L) ? desc_mnmajor(smem_u32(sW[l]), K, 0) : 0;        // backward: same bytes, MN-major
      }
      const uint64_t dGk = desc_kmajor(smem_u32(sG), kMlpOutPad, 0), dGmn = desc_mnmajor(smem_u32(sG), kMlpOutPad, 0);
      const uint64_t dOnes = desc_mnmajor(smem_u32(sOnes), kMlpOutPad, 0);
      const uint64_t dXk[2] = {desc_kmajor(smem_u32(sXbuf[0]), k0p, 0), desc_kmajor(smem_u32(sXbuf[1]), k0p, 0)};
      const uint64_t dXmn[2] = {desc_mnmajor(smem_u32(sXbuf[0]), k0p, 0), desc_mnmajor(smem_u32(sXbuf[1]), k0p, 0)};
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int xb = it & 1;
        const uint32_t accf = first ? 0u : 1u;      // persistent accumulators start from zero on the first tile
        {
          const long long t0 = clock64();
          mbar_wait_guard(bX, px, sDead);
          pw[2] += clock64() - t0;
        }
        px ^= 1;
        if (!first) {   // the previous tile's dX drain released the working accumulator
          wait_epi(0);
          wait_epi(1);
        }
        tc_fence_after();
        // ---- forward layer 0 ----
        group(tmem + kColAcc, xb ? dXk[1] : dXk[0], kStepK, dWk[0], kStepK, id_f128, k0p >> 4, 0);
        umma_commit(bAcc[0]);
        umma_commit(bAcc[1]);
        // ---- forward layers 1 .. L-2.  An N=128 MMA overwrites all 128 accumulator columns, so BOTH halves of
        //      the previous layer must have been drained (and written) before the first one is issued ----
#pragma unroll
        for (int l = 1; l < kBwdMaxDepth - 1; ++l) {
          if (l < L - 1) {
            wait_epi(0);
            wait_epi(1);
            group(tmem + kColAcc, dHk[l - 1], kStepK, dWk[l], kStepK, id_f128, 8, 0);
            umma_commit(bAcc[0]);
            umma_commit(bAcc[1]);
          }
        }
        // ---- logits ----
        wait_epi(0);
        wait_epi(1);
#pragma unroll
        for (int l = 1; l < kBwdMaxDepth; ++l)
          if (l == L - 1) group(tmem + kColOut, dHk[l - 1], kStepK, dWk[l], kStepK, id_f16, 8, 0);
        umma_commit(bAcc[0]);
        // ---- backward: last layer ----
        wait_epi(0);   // dZ_L written, logits drained
#pragma unroll
        for (int l = 1; l < kBwdMaxDepth; ++l) {
          if (l == L - 1) {
            // dW_last^T += H_{L-2}^T dZ_L ; dH_{L-2} = dZ_L W_L (one K slice)
            group(tmem + kColDwLast, dHmn[l - 1], kStepMn128, dGmn, kStepMn16, id_w16, 8, accf);
            umma_bf16(tmem + kColAcc, dGk, dWmn[l], id_b128, 0);
          }
        }
        umma_commit(bAcc[0]);
        umma_commit(bAcc[1]);
        // ---- backward: hidden layers ----
#pragma unroll
        for (int l = kBwdMaxDepth - 2; l >= 0; --l) {
          if (l < L - 1) {
            const bool want = (l > 0) || (grad_x != nullptr);
            const uint32_t stepB = (l == 0) ? stepMnX : kStepMn128;
            const uint32_t id0 = (l == 0) ? id_dx : id_b128;
            wait_epi(0);
            wait_epi(1);
            if (want) group(tmem + kColAcc, dHk[l], kStepK, dWmn[l], stepB, id0, 8, 0);
            umma_commit(bAcc[0]);
            umma_commit(bAcc[1]);
            // dW_l += dZ_l^T H_{l-1}, db_l += dZ_l^T 1 : nothing waits for them except the tiles' next writers
            const uint32_t col = (l == 0) ? kColDw0 : kColDw0 + k0p + (l - 1) * kMlpWidth;
            const uint64_t dB = (l == 0) ? (xb ? dXmn[1] : dXmn[0]) : dHmn[l > 0 ? l - 1 : 0];
            group(tmem + col, dHmn[l], kStepMn128, dB, stepB, (l == 0) ? id_wx : id_w128, 8, accf);
            group(tmem + kColDb + l * kMlpOutPad, dHmn[l], kStepMn128, dOnes, kStepMn16, id_w16, 8, accf);
            umma_commit(bDw);
          }
        }
        first = false;
      }
    }
  } else {
    // =========================== epilogue warps ===========================
    uint32_t pa = 0, pd = 0;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int c0 = cg * 32;
    auto wait_acc = [&]() {
      const long long t0 = clock64();
      mbar_wait_guard(bAcc[h], pa, sDead);
      pw[0] += clock64() - t0;
      pa ^= 1;
      tc_fence_after();
    };
    auto publish = [&]() {   // tile / accumulator half handed back to the MMA issuer
      tc_fence_before();
      fence_async_smem();
      mbar_arrive(bEpi[h]);
    };
    auto stage = [&](int64_t tile, uint8_t* dst) {
      const int chunks_per_row = k0p >> 3;
      const uint32_t base = smem_u32(dst);
      for (int id = tid; id < kTile * chunks_per_row; id += 512) {
        const int r = id / chunks_per_row, c8 = id - r * chunks_per_row;
        const int64_t gr = tile * kTile + r;
        const bool ok = gr < M;
        cp_async16(base + canon_off(r, c8 * 8, k0p), x + (ok ? gr : 0) * k0p + c8 * 8, ok ? 16u : 0u);
      }
    };
    if (blockIdx.x < n_tiles) stage(blockIdx.x, sXbuf[0]);
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int64_t grow = tile * kTile + row;
      const bool valid = grow < M;
      {
        const long long t0 = clock64();
        cp_async_wait_all();
        pw[3] += clock64() - t0;
      }
      float gz[3] = {0.f, 0.f, 0.f}, gr3[3] = {0.f, 0.f, 0.f};
      if (cg == 0 && valid) {
        if (F.w) {
          const float wv = __ldg(F.w + grow);
          const int64_t rid = __ldg(F.ray_id + grow);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            gr3[j] = __ldg(F.g_ray + rid * 3 + j);
            gz[j] = wv * gr3[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 3; ++j) gz[j] = __ldg(grad_out + grow * 3 + j);
        }
      }
      fence_async_smem();
      mbar_arrive(bX);
      // the other input buffer was last read by the previous tile's dW_0 (waited for at the end of that tile)
      if (tile + gridDim.x < n_tiles) stage(tile + gridDim.x, sXbuf[(it + 1) & 1]);

      // ---- forward epilogues ----
      for (int l = 0; l < L - 1; ++l) {
        wait_acc();
        uint32_t r[32];
        tmem_ld32_nowait(tmem + kColAcc + lane_base + c0, r);
        const float4* b4 = reinterpret_cast<const float4*>(sBias + l * kMlpWidth + c0);
        float4 bb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bb[j] = b4[j];
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(sH[l] + canon_off(row, c0 + 8 * j, kMlpWidth)) =
              bias_relu_pack8(r + 8 * j, bb[2 * j], bb[2 * j + 1]);
        publish();
      }
      // ---- logits -> dZ_L (first half only) ----
      if (h == 0) {
        wait_acc();
        if (cg == 0) {
          float v[16];
          tmem_ld16(tmem + kColOut + lane_base, v);
          const float* bias = sBias + (L - 1) * kMlpWidth;
          float dz[3];
          float gw_acc = 0.f;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float z = v[j] + bias[j];
            const float y = apply_sigmoid ? 1.f / (1.f + __expf(-z)) : z;
            dz[j] = apply_sigmoid ? gz[j] * y * (1.f - y) : gz[j];
            gw_acc += gr3[j] * y;
            if (rgb_out && valid) rgb_out[grow * 3 + j] = y;
          }
          if (F.gw_out && valid) F.gw_out[grow] = gw_acc;
          uint4 pk;
          pk.x = pack_bf16(dz[0], dz[1]);
          pk.y = pack_bf16(dz[2], 0.f);
          pk.z = 0u;
          pk.w = 0u;
          *reinterpret_cast<uint4*>(sG + canon_off(row, 0, kMlpOutPad)) = pk;
          *reinterpret_cast<uint4*>(sG + canon_off(row, 8, kMlpOutPad)) = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float s = warp_sum(dz[j]);
            if (lane == 0) atomicAdd(&sDbLast[j], s);
          }
        }
        publish();
      }
      // ---- backward epilogues: dZ_l = dH_l * (H_l > 0), in place ----
      for (int l = L - 2; l >= 0; --l) {
        wait_acc();
        uint4 pk[4];
        {
          uint32_t r[32];
          tmem_ld32_nowait(tmem + kColAcc + lane_base + c0, r);
          uint4 act[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            act[j] = *reinterpret_cast<const uint4*>(sH[l] + canon_off(row, c0 + 8 * j, kMlpWidth));
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) pk[j] = relu_grad_pack8(r + 8 * j, act[j]);
        }
        if (l < L - 2) {   // dW_{l+1} (reads H_l) must have finished before H_l is overwritten
          const long long t0 = clock64();
          mbar_wait_guard(bDw, pd, sDead);
          pw[1] += clock64() - t0;
          pd ^= 1;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(sH[l] + canon_off(row, c0 + 8 * j, kMlpWidth)) = pk[j];
        publish();
      }
      // ---- input gradient ----
      wait_acc();
      if (grad_x) {
        const int cx = cg * kq;
        if (kq == 32) {
          float v[32];
          tmem_ld32(tmem + kColAcc + lane_base + cx, v);
          if (valid && cx < F.dx_cols) {
            float4* dst = reinterpret_cast<float4*>(grad_x + grow * F.dx_cols + cx);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        } else {
          float v[16];
          tmem_ld16(tmem + kColAcc + lane_base + cx, v);
          if (valid && cx < F.dx_cols) {
            float4* dst = reinterpret_cast<float4*>(grad_x + grow * F.dx_cols + cx);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
      }
      publish();
      {
        const long long t0 = clock64();
        mbar_wait_guard(bDw, pd, sDead);   // dW_0 read this tile's input rows and dZ_0
        pw[2] += clock64() - t0;
      }
      pd ^= 1;
      first = false;
    }
  }

  if (prof && blockIdx.x == 0 && lane == 0 && (warp == 16 || warp == 0 || warp == 8)) {
    // [mma warp: wait epi0, wait epi1, wait x | epilogue warp 0 and 8: wait acc, wait dw (in loop), wait dw (tile end), cp.async] + total
    long long* dst = prof + (warp == 16 ? 0 : warp == 0 ? 8 : 16);
    for (int q = 0; q < 6; ++q) dst[q] = pw[q];
    dst[6] = clock64() - t_begin;
  }
  // ---- the CTA's weight / bias gradients: TMEM -> global (fp32 atomics, once per CTA) ----
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const bool dead = *sDead != 0;
  if (dead && tid == 0 && err_flag) atomicExch(err_flag, 1);
  if (warp < 16 && !first && !dead) {
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int c0 = cg * 32;
    for (int l = 1; l < L - 1; ++l) {
      const uint32_t col = kColDw0 + k0p + (l - 1) * kMlpWidth;
      float v[32];
      tmem_ld32(tmem + col + lane_base + c0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) atomicAdd(P.gw[l] + row * kMlpWidth + c0 + j, v[j]);
    }
    {
      const int cx = cg * kq;
      if (kq == 32) {
        float v[32];
        tmem_ld32(tmem + kColDw0 + lane_base + cx, v);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (cx + j < P.d0) atomicAdd(P.gw[0] + row * P.d0 + cx + j, v[j]);
      } else {
        float v[16];
        tmem_ld16(tmem + kColDw0 + lane_base + cx, v);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (cx + j < P.d0) atomicAdd(P.gw[0] + row * P.d0 + cx + j, v[j]);
      }
    }
    if (cg == 0) {
      float v[16];
      tmem_ld16(tmem + kColDwLast + lane_base, v);
#pragma unroll
      for (int o = 0; o < 3; ++o) atomicAdd(P.gw[L - 1] + o * kMlpWidth + row, v[o]);
    } else if (cg == 1) {
      for (int l = 0; l < L - 1; ++l) {
        float v[16];
        tmem_ld16(tmem + kColDb + l * kMlpOutPad + lane_base, v);
        atomicAdd(P.gb[l] + row, v[0]);
      }
    }
    if (tid < 3) atomicAdd(P.gb[L - 1] + tid, sDbLast[tid]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// fp32 [n,d0] rows -> bf16 [n,k0p] rows (zero padded): one thread per 8 output columns
__global__ void rgbnet_pack_kernel(const float* __restrict__ x, int64_t n, int d0, int k0p,
                                   __nv_bfloat16* __restrict__ out) {
  const int cpr = k0p >> 3;
  for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < n * cpr;
       id += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = id / cpr;
    const int c0 = static_cast<int>(id - r * cpr) * 8;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (c0 + j < d0) ? __ldg(x + r * d0 + c0 + j) : 0.f;
    uint4 pk;
    pk.x = pack_bf16(f[0], f[1]);
    pk.y = pack_bf16(f[2], f[3]);
    pk.z = pack_bf16(f[4], f[5]);
    pk.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + r * k0p + c0) = pk;
  }
}

// fp32 [n,k0p] rows -> fp32 [n,d0] rows
__global__ void rgbnet_unpack_kernel(const float* __restrict__ gx, int64_t n, int d0, int k0p,
                                     float* __restrict__ out) {
  for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < n * d0;
       id += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = id / d0;
    const int c = static_cast<int>(id - r * d0);
    out[id] = gx[r * k0p + c];
  }
}

}  // namespace pp

using namespace pp;

extern "C" int32_t pp_rgbnet_padded_width(int32_t d0) { return d0 <= 64 ? 64 : 128; }

extern "C" int64_t pp_rgbnet_workspace_bytes(int64_t n_pts, int32_t d0) {
  if (n_pts <= 0 || d0 <= 0) return 0;
  const int64_t k0p = pp_rgbnet_padded_width(d0);
  return n_pts * k0p * 2 + n_pts * k0p * 4 + 256;
}

extern "C" int pp_rgbnet_pack_input(const float* x, int64_t n_pts, int32_t d0, void* x_packed,
                                    pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0 && d0 >= 1 && d0 <= kMlpWidth, PP_ERR_INVALID_ARG, "rgbnet_pack_input: bad extent");
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(x && x_packed, PP_ERR_INVALID_ARG, "rgbnet_pack_input: null pointer");
  const int k0p = pp_rgbnet_padded_width(d0);
  const int64_t work = n_pts * (k0p >> 3);
  const int blocks = (int)min64((work + 255) / 256, (int64_t)sm_count() * 16);
  rgbnet_pack_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, n_pts, d0, k0p, (__nv_bfloat16*)x_packed);
  PP_CHECK_LAUNCH("rgbnet_pack_input");
  return PP_OK;
}

extern "C" int pp_rgbnet_unpack_grad(const float* grad_x_padded, int64_t n_pts, int32_t d0, float* grad_x,
                                     pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0 && d0 >= 1 && d0 <= kMlpWidth, PP_ERR_INVALID_ARG, "rgbnet_unpack_grad: bad extent");
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(grad_x_padded && grad_x, PP_ERR_INVALID_ARG, "rgbnet_unpack_grad: null pointer");
  const int k0p = pp_rgbnet_padded_width(d0);
  const int blocks = (int)min64((n_pts * d0 + 255) / 256, (int64_t)sm_count() * 16);
  rgbnet_unpack_kernel<<<blocks, 256, 0, as_stream(stream)>>>(grad_x_padded, n_pts, d0, k0p, grad_x);
  PP_CHECK_LAUNCH("rgbnet_unpack_grad");
  return PP_OK;
}

static int* g_bwd_err_flag = nullptr;
static long long* g_bwd_prof = nullptr;

// PP_MLP_PROF diagnostic: cycle counters of CTA 0 of the last backward launch (24 values, see the kernel)
extern "C" int pp_rgbnet_backward_profile(long long* out24) {
  if (!g_bwd_prof || !out24) return -1;
  return cudaMemcpy(out24, g_bwd_prof, 24 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}

// 1 if a barrier wait of the backward kernel ever ran out of its cycle budget (a scheduling bug, never expected);
// synchronises the device.
extern "C" int pp_rgbnet_backward_stalled(void) {
  if (!g_bwd_err_flag) return 0;
  int v = 0;
  if (cudaMemcpy(&v, g_bwd_err_flag, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v;
}

static int launch_rgbnet_backward(const void* x_packed, const float* grad_out, int64_t n_pts, int32_t d0,
                                  int32_t depth, int32_t width, const float* const* weights,
                                  const float* const* biases, int apply_sigmoid, float* grad_x_padded,
                                  float* const* grad_weights, float* const* grad_biases, float* rgb_out,
                                  MlpFusedIO F, pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0, PP_ERR_INVALID_ARG, "rgbnet_backward: n_pts < 0");
  PP_REQUIRE(width == kMlpWidth, PP_ERR_UNSUPPORTED, "rgbnet_backward: hidden width must be %d", kMlpWidth);
  PP_REQUIRE(depth >= 2 && depth <= kBwdMaxDepth, PP_ERR_UNSUPPORTED, "rgbnet_backward: depth must be 2..%d",
             kBwdMaxDepth);
  PP_REQUIRE(d0 >= 1 && d0 <= kMlpWidth, PP_ERR_UNSUPPORTED, "rgbnet_backward: 1 <= d0 <= %d", kMlpWidth);
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(x_packed && (grad_out || F.w) && weights && biases && grad_weights && grad_biases, PP_ERR_INVALID_ARG,
             "rgbnet_backward: null pointer");
  MlpBwdParams P;
  P.depth = depth;
  P.d0 = d0;
  P.k0p = pp_rgbnet_padded_width(d0);
  if (F.dx_cols == 0) F.dx_cols = P.k0p;
  PP_REQUIRE(F.dx_cols == P.k0p || (F.dx_cols == 16 && P.k0p == 64), PP_ERR_INVALID_ARG,
             "rgbnet_backward: dx_cols must be the padded width or 16");
  // TMEM budget: 192 fixed columns + dW_0 (k0p) + 128 per further hidden layer <= 512
  PP_REQUIRE(kColDw0 + P.k0p + (depth - 2) * kMlpWidth <= 512, PP_ERR_UNSUPPORTED,
             "rgbnet_backward: depth %d with %d input features exceeds the 512 TMEM columns "
             "(supported: d0 <= 64 up to depth 4, d0 <= 128 up to depth 3)", depth, d0);
  size_t smem = (size_t)2 * kTile * P.k0p * 2 + (size_t)(depth - 1) * kTile * kMlpWidth * 2 +
                (size_t)2 * kTile * kMlpOutPad * 2 + (size_t)depth * kMlpWidth * 4 + 128;
  for (int l = 0; l < kBwdMaxDepth; ++l) {
    P.w[l] = l < depth ? weights[l] : nullptr;
    P.b[l] = l < depth ? biases[l] : nullptr;
    P.gw[l] = l < depth ? grad_weights[l] : nullptr;
    P.gb[l] = l < depth ? grad_biases[l] : nullptr;
    PP_REQUIRE(l >= depth || (P.w[l] && P.b[l] && P.gw[l] && P.gb[l]), PP_ERR_INVALID_ARG,
               "rgbnet_backward: null layer %d", l);
    if (l < depth) {
      const int rows = (l == depth - 1) ? kMlpOutPad : kMlpWidth;
      const int K = (l == 0) ? P.k0p : kMlpWidth;
      smem += (size_t)rows * K * 2;
    }
  }
  PP_REQUIRE(smem <= 227 * 1024, PP_ERR_UNSUPPORTED, "rgbnet_backward: %zu B of shared memory needed", smem);
  const int64_t n_tiles = (n_pts + kTile - 1) / kTile;
  const int blocks = (int)min64(n_tiles, sm_count());
  {
    static int* err_flag = nullptr;   // device flag raised by a wait that ran out of its cycle budget
    static long long* prof = nullptr;
    if (!err_flag) {
      PP_CHECK_CUDA(cudaMalloc(&err_flag, sizeof(int)));
      PP_CHECK_CUDA(cudaMemset(err_flag, 0, sizeof(int)));
      if (getenv("PP_MLP_PROF")) {
        PP_CHECK_CUDA(cudaMalloc(&prof, 24 * sizeof(long long)));
        PP_CHECK_CUDA(cudaMemset(prof, 0, 24 * sizeof(long long)));
        g_bwd_prof = prof;
      }
    }
    g_bwd_err_flag = err_flag;
    PP_CHECK_CUDA(cudaFuncSetAttribute(rgbnet_backward2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rgbnet_backward2_kernel<<<blocks, kBwd2Threads, smem, as_stream(stream)>>>(
        P, (const __nv_bfloat16*)x_packed, grad_out, n_pts, apply_sigmoid, grad_x_padded, rgb_out, F, err_flag, prof);
  }
  PP_CHECK_LAUNCH("rgbnet_backward");
  return PP_OK;
}

extern "C" int pp_rgbnet_backward_packed(const void* x_packed, const float* grad_out, int64_t n_pts, int32_t d0,
                                         int32_t depth, int32_t width, const float* const* weights,
                                         const float* const* biases, int apply_sigmoid, float* grad_x_padded,
                                         float* const* grad_weights, float* const* grad_biases, float* rgb_out,
                                         pp_stream_t stream) {
  MlpFusedIO F = {nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  return launch_rgbnet_backward(x_packed, grad_out, n_pts, d0, depth, width, weights, biases, apply_sigmoid,
                                grad_x_padded, grad_weights, grad_biases, rgb_out, F, stream);
}

extern "C" int pp_rgbnet_backward_fused(const void* x_packed, int64_t capacity, const int32_t* n_ptr, int32_t d0,
                                        int32_t depth, int32_t width, const float* const* weights,
                                        const float* const* biases, const float* sample_weight,
                                        const int32_t* sample_ray, const float* g_ray, float* grad_weight_out,
                                        float* grad_feat, int32_t grad_feat_cols, float* const* grad_weights,
                                        float* const* grad_biases, pp_stream_t stream) {
  PP_REQUIRE(sample_weight && sample_ray && g_ray && grad_weight_out && grad_feat, PP_ERR_INVALID_ARG,
             "rgbnet_backward_fused: null pointer");
  MlpFusedIO F = {n_ptr, sample_weight, sample_ray, g_ray, grad_weight_out, grad_feat_cols};
  return launch_rgbnet_backward(x_packed, nullptr, capacity, d0, depth, width, weights, biases, 1, grad_feat,
                                grad_weights, grad_biases, nullptr, F, stream);
}

extern "C" int pp_rgbnet_backward(const float* x, const float* grad_out, int64_t n_pts, int32_t d0,
                                  int32_t depth, int32_t width, const float* const* weights,
                                  const float* const* biases, int apply_sigmoid, float* grad_x,
                                  float* const* grad_weights, float* const* grad_biases, float* rgb_out,
                                  void* workspace, int64_t workspace_bytes, pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0, PP_ERR_INVALID_ARG, "rgbnet_backward: n_pts < 0");
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(d0 >= 1 && d0 <= kMlpWidth, PP_ERR_UNSUPPORTED, "rgbnet_backward: 1 <= d0 <= %d", kMlpWidth);
  PP_REQUIRE(workspace && workspace_bytes >= pp_rgbnet_workspace_bytes(n_pts, d0), PP_ERR_INVALID_ARG,
             "rgbnet_backward: workspace of %lld bytes required (pp_rgbnet_workspace_bytes)",
             (long long)pp_rgbnet_workspace_bytes(n_pts, d0));
  const int64_t k0p = pp_rgbnet_padded_width(d0);
  uint8_t* ws = (uint8_t*)workspace;
  void* xp = ws;
  float* gxp = (float*)(ws + (n_pts * k0p * 2 + 255) / 256 * 256);
  int rc = pp_rgbnet_pack_input(x, n_pts, d0, xp, stream);
  if (rc != PP_OK) return rc;
  rc = pp_rgbnet_backward_packed(xp, grad_out, n_pts, d0, depth, width, weights, biases, apply_sigmoid,
                                 grad_x ? gxp : nullptr, grad_weights, grad_biases, rgb_out, stream);
  if (rc != PP_OK) return rc;
  if (grad_x) rc = pp_rgbnet_unpack_grad(gxp, n_pts, d0, grad_x, stream);
  return rc;
}
