// rgbnet (colour MLP) forward on the 5th-generation tensor cores: tcgen05.mma + TMEM.
//
// The reference evaluates the colour network with nn.Linear / ReLU (cuBLAS sgemm) on a materialised
// [M, d0] fp32 feature matrix and three [M,128] activations that round-trip HBM
// (lib/voxurf_coarse.py:208-216 definition, :1032-1033 use; lib/dvgo_ori.py:82-89, :355-360).  This is the
// only dense contraction on the path, so it is the only place tensor cores are used.
//
// One CTA (4 warps = 128 threads = 128 TMEM lanes) owns a tile of 128 samples and runs the WHOLE network
// on it without leaving the SM:
//   * weights are converted once per CTA to bf16 and kept in shared memory in the UMMA canonical K-major
//     (no-swizzle) layout: 8x8 core matrices of 128 B, LBO = 128 B along K, SBO = (K/8)*128 B along M/N;
//   * the 128 x K activation tile lives in shared memory in the same layout (A operand);
//   * each layer is K/16 tcgen05.mma (M=128, N=128 or 16, K=16, bf16 x bf16 -> fp32) issued by one thread,
//     accumulating in TMEM; completion is signalled through tcgen05.commit -> mbarrier;
//   * the epilogue reads the accumulator with tcgen05.ld (thread t owns row t), adds the bias, applies ReLU,
//     converts to bf16 and writes the row straight back into the A tile for the next layer; the last layer
//     applies the sigmoid and writes [M,3].
// Activations never touch HBM: traffic is d0*4 B in and 12 B out per sample.
#include <stdlib.h>
#include <string.h>

#include <cuda.h>   // CUtensorMap (type and enums only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "mlp_common.cuh"

namespace pp {

constexpr int kFwdThreads = 512;   // kPipes independent pipelines of 16 / kPipes warps per CTA
constexpr int kFwdTile = 128;

struct MlpParams {
  const float* w[kMlpMaxDepth];
  const float* b[kMlpMaxDepth];
  int depth;      // number of Linear layers
  int d0;         // input features
  int k0p;        // input row width of the packed bf16 input: 64 or 128 (16 for the warp network)
  int n_out;      // real outputs of the last layer (3 colour channels; 4 for the warp network)
};

template <int kThreadsPerPipe>
__device__ __forceinline__ void group_sync(int g) {
  asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(kThreadsPerPipe) : "memory");
}

// kPipes pipelines (4 with 4 warps each when the weights leave room for four activation tiles, else 2 with 8
// warps) share the weights in shared memory; each owns an activation tile, a 128-column TMEM accumulator (the
// logits alias its first 16 columns) and an mbarrier and walks its own tiles.  The epilogue of a 128 x 128 fp32
// accumulator is bounded by the TMEM read port (64 B/clk: 1024 cycles against 544 cycles of MMA for the layer), so
// the goal is to keep tcgen05.ld busy: while one pipeline waits (MMA latency, input tile, barrier) the others
// drain their accumulators.
// shared memory map: kPipes A tiles (128 x 128 bf16 each), weights, biases, kPipes mbarriers, tmem base.
//
// kJac (the deformation network with its input Jacobian, lib/deformation/deform_net.py:12-29 and the autograd.grad
// calls of lib/voxurf_coarse.py:966-985): every sample occupies FOUR consecutive rows - the value row (x, with
// biases) and three tangent rows (the unit vectors e_k, no biases).  A ReLU network is piecewise linear, so the
// tangent of a hidden layer is the layer's linear map gated by the VALUE row's activation pattern:
// t_l = (z_l > 0) * (W_l t_{l-1}).  The four rows are four lanes of one warp in the epilogue, the gate is one
// shuffle per accumulator element.  Row 4s carries net(x), rows 4s+1..3 carry d net / d x_k.
// Operands are fp16 here (11 significant bits, same MMA rate as bf16): a gate that flips because z_l was rounded
// changes the Jacobian of that sample by a whole rank-1 term, and the flip rate scales with the operand rounding;
// the values of this small network (|h| < ~1e2) are far inside the fp16 range.
// kTma: the input tiles arrive by TMA (cp.async.bulk.tensor, one elected thread, completion on an mbarrier) in the
// chunked layout of desc_kmajor_chunked instead of 16-byte cp.async from every thread.
template <int kPipes, bool kJac, bool kTma>
__global__ void __launch_bounds__(kFwdThreads, 1)
rgbnet_forward_kernel(MlpParams P, const __nv_bfloat16* __restrict__ x, int64_t M, int apply_sigmoid,
                      float* __restrict__ out, const float* __restrict__ wgt, const int32_t* __restrict__ ray_id,
                      float* __restrict__ rgb_marched, const int32_t* __restrict__ n_ptr,
                      const __grid_constant__ CUtensorMap x_map) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (n_ptr) M = min64(M, static_cast<int64_t>(*n_ptr));   // row count produced on the device (compacted samples)
  constexpr int kPipeThreads = kFwdThreads / kPipes;       // 256 or 128
  constexpr int kColsPerThread = kMlpWidth / (kPipeThreads / kFwdTile);   // 64 or 128 accumulator columns
  const int tid = threadIdx.x, g = tid / kPipeThreads, gt = tid % kPipeThreads, warp = gt >> 5, lane = tid & 31;
  const int L = P.depth, k0p = P.k0p;
  const int row = (warp & 3) * 32 + lane, half = warp >> 2;
  uint8_t* sA = smem + g * (kFwdTile * kMlpWidth * 2);
  uint32_t off = kPipes * kFwdTile * kMlpWidth * 2;
  uint8_t* sW[kMlpMaxDepth];
  for (int l = 0; l < L; ++l) {
    sW[l] = smem + off;
    const int rows = (l == L - 1) ? kMlpOutPad : kMlpWidth;
    const int K = (l == 0) ? k0p : kMlpWidth;
    off += rows * K * 2;
  }
  float* sBias = reinterpret_cast<float*>(smem + off);
  off += L * kMlpWidth * 4;
  uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + off);
  // (behind ALL kPipes mbarriers: at + 16 it shared its bytes with the third pipeline's barrier, and the deallocation
  // at the end read whatever state that barrier was in - cudaErrorTensorMemoryLeak once per few thousand launches)
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(smem + off + 8 * kPipes);
  uint64_t* sBarX = reinterpret_cast<uint64_t*>(smem + off + 64);      // kPipes "input tile has landed" barriers (TMA)
  static_assert(8 * kPipes + 4 <= 64 && 64 + 8 * kPipes <= 128, "the host reserves 128 bytes for barriers + TMEM base");

  // ---- prologue (whole CTA): weights -> bf16 canonical layout, biases, TMEM, mbarriers ----
  for (int l = 0; l < L; ++l) {
    const bool last = (l == L - 1);
    const int rows = last ? kMlpOutPad : kMlpWidth;
    const int real_rows = last ? P.n_out : kMlpWidth;
    const int K = (l == 0) ? k0p : kMlpWidth;
    const int real_K = (l == 0) ? P.d0 : kMlpWidth;
    const int cpr = K >> 3;
    for (int e = tid; e < rows * cpr; e += kFwdThreads) {
      const int r = e / cpr, k0 = (e - r * cpr) * 8;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (r < real_rows && k0 + j < real_K) ? __ldg(P.w[l] + r * real_K + k0 + j) : 0.f;
      uint4 pk;
      if constexpr (kJac) {   // the warp network runs on fp16 operands, see the kernel comment
        pk.x = pack_f16(f[0], f[1]);
        pk.y = pack_f16(f[2], f[3]);
        pk.z = pack_f16(f[4], f[5]);
        pk.w = pack_f16(f[6], f[7]);
      } else {
        pk.x = pack_bf16(f[0], f[1]);
        pk.y = pack_bf16(f[2], f[3]);
        pk.z = pack_bf16(f[4], f[5]);
        pk.w = pack_bf16(f[6], f[7]);
      }
      *reinterpret_cast<uint4*>(sW[l] + canon_off(r, k0, K)) = pk;
    }
    for (int e = tid; e < kMlpWidth; e += kFwdThreads)
      sBias[l * kMlpWidth + e] = (e < real_rows) ? P.b[l][e] : 0.f;
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int q = 0; q < kPipes; ++q) {
      mbar_init(smem_u32(sBar + q), 1);
      mbar_init(smem_u32(sBarX + q), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sTmem;
  const uint32_t tmem = tmem_base + g * kMlpWidth;   // this pipeline's accumulator: 128 columns, logits in the first 16
  const uint32_t bar = smem_u32(sBar + g), bar_x = smem_u32(sBarX + g);
  uint32_t phase = 0, phase_x = 0;
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t idesc_h = kJac ? umma_idesc_f16(kFwdTile, kMlpWidth) : umma_idesc_bf16(kFwdTile, kMlpWidth);
  const uint32_t idesc_o = kJac ? umma_idesc_f16(kFwdTile, kMlpOutPad) : umma_idesc_bf16(kFwdTile, kMlpOutPad);
  const int chunks_per_row = k0p >> 3;

  const int64_t n_tiles = (M + kFwdTile - 1) / kFwdTile;
  const int64_t tile_stride = (int64_t)gridDim.x * kPipes;
  auto stage = [&](int64_t tile) {   // request a tile of packed rows straight into the A buffer (UMMA layout)
    const uint32_t base = smem_u32(sA);
    if constexpr (kTma) {
      if (gt == 0) {
        mbar_expect_tx(bar_x, static_cast<uint32_t>(kFwdTile * k0p * 2));
        for (int c8 = 0; c8 < chunks_per_row; ++c8)
          tma_load_2d(base + c8 * 2048, &x_map, c8 * 8, static_cast<int>(tile * kFwdTile), bar_x);
      }
    } else {
      for (int id = gt; id < kFwdTile * chunks_per_row; id += kPipeThreads) {
        const int r = id / chunks_per_row, c8 = id - r * chunks_per_row;
        const int64_t grow = tile * kFwdTile + r;
        const bool ok = grow < M;
        const uint32_t dst = base + canon_off(r, c8 * 8, k0p);
        const void* src = x + (ok ? grow : 0) * k0p + c8 * 8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
      }
    }
  };
  int64_t tile = (int64_t)blockIdx.x * kPipes + g;
  if (tile < n_tiles) stage(tile);
  for (; tile < n_tiles; tile += tile_stride) {
    const int64_t grow = tile * kFwdTile + row;
    const bool valid = grow < M;
    if constexpr (kTma) {
      mbar_wait_or_trap(bar_x, phase_x);     // the tile's bytes have landed (rows past the buffer's end: zeros;
      phase_x ^= 1;                          // rows between M and the buffer's end: stale, their outputs are dropped)
    } else {
      asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
      fence_async_smem();
    }
    group_sync<kPipeThreads>(g);

    for (int l = 0; l < L; ++l) {
      const bool last = (l == L - 1);
      const int K = (l == 0) ? k0p : kMlpWidth;
      const uint32_t d_tmem = tmem;
      if (gt < 32) {   // the pipeline's first warp issues (warp-uniform loop, one elected lane)
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sW[l]);
        for (int k = 0; k < (K >> 4); ++k)
          umma_bf16(d_tmem, (kTma && l == 0) ? desc_kmajor_chunked(a0, k) : desc_kmajor(a0, K, k),
                    desc_kmajor(b0, K, k), last ? idesc_o : idesc_h, k > 0);
        umma_commit(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      if (!last) {
        const float* bias = sBias + l * kMlpWidth;
#pragma unroll 1
        for (int i = 0; i < kColsPerThread / 32; ++i) {
          const int c0 = half * kColsPerThread + i * 32;
          uint32_t r[32];
          tmem_ld32_nowait(d_tmem + lane_base + c0, r);
          const float4* b4 = reinterpret_cast<const float4*>(bias + c0);
          float4 bb[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bb[j] = b4[j];
          tmem_ld_wait();
          if constexpr (kJac) {
            // X = acc + bias on the value lane, acc on the tangent lanes; both become X * (z_value > 0): ReLU for the
            // value row (z * [z > 0]), the gate for the tangents.  Packed throughout: per two accumulator columns one
            // FFMA2, one f32x2 -> f16x2 conversion, one shuffle of the packed pair, one HSET2, one HMUL2.
            const float2* bf2 = reinterpret_cast<const float2*>(bb);
            const float bsel = (lane & 3) == 0 ? 1.f : 0.f;
            const float2 bs2 = make_float2(bsel, bsel);
            const __half2 zero2 = __floats2half2_rn(0.f, 0.f);
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 a = make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
              const __half2 xh = __float22half2_rn(__ffma2_rn(bf2[j], bs2, a));
              const uint32_t xu = *reinterpret_cast<const uint32_t*>(&xh);
              const uint32_t vu = __shfl_sync(kFull, xu, lane & ~3);
              const __half2 gate = __hgt2(*reinterpret_cast<const __half2*>(&vu), zero2);
              const __half2 o = __hmul2(xh, gate);
              pk[j] = *reinterpret_cast<const uint32_t*>(&o);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(sA + canon_off(row, c0 + 8 * j, kMlpWidth)) =
                  make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(sA + canon_off(row, c0 + 8 * j, kMlpWidth)) =
                  bias_relu_pack8(r + 8 * j, bb[2 * j], bb[2 * j + 1]);
          }
        }
        tc_fence_before();
        fence_async_smem();
        group_sync<kPipeThreads>(g);
      } else {
        // the A tile is free: the next tile's rows start moving while the outputs are written
        if (tile + tile_stride < n_tiles) stage(tile + tile_stride);
        if (half == 0) {
          float v[16];
          tmem_ld16(d_tmem + lane_base, v);
          const float* bias = sBias + l * kMlpWidth;
          if constexpr (kJac) {
            // [rows, 4] fp32: net(x) on the value rows (with bias), d net / d x_k on the tangent rows
            if (valid) {
              const float bsel = (lane & 3) == 0 ? 1.f : 0.f;
              *reinterpret_cast<float4*>(out + grow * 4) =
                  make_float4(fma_rn(bias[0], bsel, v[0]), fma_rn(bias[1], bsel, v[1]), fma_rn(bias[2], bsel, v[2]),
                              fma_rn(bias[3], bsel, v[3]));
            }
            tc_fence_before();
            continue;
          }
          float y[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float z = v[j] + bias[j];
            y[j] = apply_sigmoid ? 1.f / (1.f + __expf(-z)) : z;
          }
          if (out && valid) {
#pragma unroll
            for (int j = 0; j < 3; ++j) out[grow * 3 + j] = y[j];
          }
          if (rgb_marched) {
            // fused compositing (segment_coo(w * rgb), voxurf_coarse.py:1034): the 32 rows of a warp are
            // consecutive samples, almost always of one ray -> one atomic per warp and channel
            const float wv = valid ? wgt[grow] : 0.f;
            const int rid = valid ? ray_id[grow] : -1;
            const int rid0 = __shfl_sync(kFull, rid, 0);
            const bool uniform = __all_sync(kFull, rid == rid0 || rid < 0);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const float c = wv * y[j];
              if (uniform) {
                const float sum = warp_sum(c);
                if (lane == 0 && rid0 >= 0) atomicAdd(rgb_marched + (int64_t)rid0 * 3 + j, sum);
              } else if (rid >= 0) {
                atomicAdd(rgb_marched + (int64_t)rid * 3 + j, c);
              }
            }
          }
        }
        tc_fence_before();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tid < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// rows of pp_warpnet_forward_packed for free-standing points: x fp32 [n,3] -> fp16 [4n,16]
__global__ void warpnet_pack_points_kernel(const float* __restrict__ x, int64_t n, uint4* __restrict__ rows) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float hi[3], lo[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float v = x[3 * i + a];
    hi[a] = __half2float(__float2half_rn(v));
    lo[a] = v - hi[a];
  }
  uint4* dst = rows + i * 8;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  dst[0] = make_uint4(pack_f16(hi[0], hi[1]), pack_f16(hi[2], lo[0]), pack_f16(lo[1], lo[2]), 0x00003C00u);   // column 6 = 1: db_0 column of the backward's dW_0 product
  dst[1] = zero;
  dst[2] = make_uint4(0x00003C00u, 0u, 0u, 0u);   // e_0 (fp16 1.0 = 0x3C00)
  dst[3] = zero;
  dst[4] = make_uint4(0x3C000000u, 0u, 0u, 0u);   // e_1
  dst[5] = zero;
  dst[6] = make_uint4(0u, 0x00003C00u, 0u, 0u);   // e_2
  dst[7] = zero;
}

}  // namespace pp

using namespace pp;

// cuTensorMapEncodeTiled, fetched through the runtime (the library links cudart only).  nullptr: not available.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    (void)cudaGetLastError();
  }
  return fn;
}

// [n_rows, k0p] bf16 rows as a 2-D tensor, box = {8 columns, 128 rows} (see desc_kmajor_chunked)
static bool make_rows_map(CUtensorMap* map, const void* rows, int64_t n_rows, int k0p) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc || (reinterpret_cast<uintptr_t>(rows) & 15)) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)k0p, (cuuint64_t)n_rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)k0p * 2};
  const cuuint32_t box[2] = {8, (cuuint32_t)kFwdTile};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(rows), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern "C" int pp_warpnet_pack_points(const float* x, int64_t n_pts, void* rows, pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0, PP_ERR_INVALID_ARG, "warpnet_pack_points: n_pts < 0");
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(x && rows, PP_ERR_INVALID_ARG, "warpnet_pack_points: null pointer");
  warpnet_pack_points_kernel<<<ceil_div(n_pts, 256), 256, 0, as_stream(stream)>>>(x, n_pts, (uint4*)rows);
  PP_CHECK_LAUNCH("warpnet_pack_points");
  return PP_OK;
}

extern "C" int pp_rgbnet_forward_packed(const void* x_packed, int64_t n_pts, int32_t d0, int32_t depth,
                                        int32_t width, const float* const* weights, const float* const* biases,
                                        int apply_sigmoid, float* out, const float* sample_weight,
                                        const int32_t* ray_id, float* rgb_marched, const int32_t* n_ptr,
                                        pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0, PP_ERR_INVALID_ARG, "rgbnet_forward: n_pts < 0");
  PP_REQUIRE(width == kMlpWidth, PP_ERR_UNSUPPORTED, "rgbnet_forward: hidden width must be %d", kMlpWidth);
  PP_REQUIRE(depth >= 2 && depth <= kMlpMaxDepth, PP_ERR_UNSUPPORTED, "rgbnet_forward: depth must be 2..%d",
             kMlpMaxDepth);
  PP_REQUIRE(d0 >= 1 && d0 <= kMlpWidth, PP_ERR_UNSUPPORTED, "rgbnet_forward: 1 <= d0 <= %d", kMlpWidth);
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(x_packed && weights && biases && (out || rgb_marched), PP_ERR_INVALID_ARG,
             "rgbnet_forward: null pointer");
  PP_REQUIRE(!rgb_marched || (sample_weight && ray_id), PP_ERR_INVALID_ARG,
             "rgbnet_forward: compositing needs sample_weight and ray_id");
  MlpParams P;
  P.depth = depth;
  P.d0 = d0;
  P.k0p = d0 <= 64 ? 64 : 128;
  P.n_out = 3;
  size_t smem = 0;
  for (int l = 0; l < kMlpMaxDepth; ++l) {
    P.w[l] = l < depth ? weights[l] : nullptr;
    P.b[l] = l < depth ? biases[l] : nullptr;
    PP_REQUIRE(l >= depth || (P.w[l] && P.b[l]), PP_ERR_INVALID_ARG, "rgbnet_forward: null layer %d", l);
    if (l < depth) {
      const int rows = (l == depth - 1) ? kMlpOutPad : kMlpWidth;
      const int K = (l == 0) ? P.k0p : kMlpWidth;
      smem += (size_t)rows * K * 2;
    }
  }
  smem += (size_t)depth * kMlpWidth * 4 + 128;
  const size_t tile_bytes = (size_t)kFwdTile * kMlpWidth * 2;
  const char* pe = getenv("PP_MLP_PIPES");
  int pipes = (smem + 4 * tile_bytes <= 227 * 1024) ? 4 : 2;
  if (pe && atoi(pe) == 2) pipes = 2;
  smem += pipes * tile_bytes;
  PP_REQUIRE(smem <= 227 * 1024, PP_ERR_UNSUPPORTED, "rgbnet_forward: %zu B of shared memory needed", smem);
  const int64_t n_tiles = (n_pts + kFwdTile - 1) / kFwdTile;
  const int blocks = (int)min64((n_tiles + pipes - 1) / pipes, sm_count());
  // input tiles by TMA when a tensor map can be built (PP_MLP_TMA=0 keeps the cp.async staging)
  CUtensorMap x_map;
  memset(&x_map, 0, sizeof(x_map));
  const char* te = getenv("PP_MLP_TMA");
  const bool tma = !(te && atoi(te) == 0) && make_rows_map(&x_map, x_packed, n_pts, P.k0p);
#define PP_LAUNCH_FWD(PIPES, TMA)                                                                                   \
  do {                                                                                                              \
    PP_CHECK_CUDA(cudaFuncSetAttribute(rgbnet_forward_kernel<PIPES, false, TMA>,                                    \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
    rgbnet_forward_kernel<PIPES, false, TMA><<<blocks, kFwdThreads, smem, as_stream(stream)>>>(                     \
        P, (const __nv_bfloat16*)x_packed, n_pts, apply_sigmoid, out, sample_weight, ray_id, rgb_marched, n_ptr,    \
        x_map);                                                                                                     \
  } while (0)
  if (pipes == 4) {
    if (tma) PP_LAUNCH_FWD(4, true); else PP_LAUNCH_FWD(4, false);
  } else {
    if (tma) PP_LAUNCH_FWD(2, true); else PP_LAUNCH_FWD(2, false);
  }
#undef PP_LAUNCH_FWD
  PP_CHECK_LAUNCH("rgbnet_forward");
  return PP_OK;
}

// Deformation network with its input Jacobian (kJac).  x_rows: fp16 [n_rows, 16], four rows per sample: the value
// row {x_hi[3], x_lo[3], 0...} (the point split into two fp16 terms so that the input keeps 22 mantissa bits) and the
// three tangent rows e_k; weights[0] is [128, 6] = [W_0 | W_0] (the same columns multiply x_hi and x_lo), the other
// layers as the reference stores them, the last one [4, 128].  out: fp32 [n_rows, 4].
extern "C" int pp_warpnet_forward_packed(const void* x_rows, int64_t n_rows, int32_t depth, int32_t width,
                                         const float* const* weights, const float* const* biases, float* out,
                                         const int32_t* n_rows_ptr, pp_stream_t stream) {
  PP_REQUIRE(n_rows >= 0 && (n_rows & 3) == 0, PP_ERR_INVALID_ARG, "warpnet_forward: n_rows must be a multiple of 4");
  PP_REQUIRE(width == kMlpWidth, PP_ERR_UNSUPPORTED, "warpnet_forward: hidden width must be %d", kMlpWidth);
  PP_REQUIRE(depth >= 2 && depth <= kMlpMaxDepth, PP_ERR_UNSUPPORTED, "warpnet_forward: depth must be 2..%d",
             kMlpMaxDepth);
  if (n_rows == 0) return PP_OK;
  PP_REQUIRE(x_rows && weights && biases && out, PP_ERR_INVALID_ARG, "warpnet_forward: null pointer");
  MlpParams P;
  P.depth = depth;
  P.d0 = 6;
  P.k0p = 16;
  P.n_out = 4;
  size_t smem = 0;
  for (int l = 0; l < kMlpMaxDepth; ++l) {
    P.w[l] = l < depth ? weights[l] : nullptr;
    P.b[l] = l < depth ? biases[l] : nullptr;
    PP_REQUIRE(l >= depth || (P.w[l] && P.b[l]), PP_ERR_INVALID_ARG, "warpnet_forward: null layer %d", l);
    if (l < depth) {
      const int rows = (l == depth - 1) ? kMlpOutPad : kMlpWidth;
      const int K = (l == 0) ? P.k0p : kMlpWidth;
      smem += (size_t)rows * K * 2;
    }
  }
  smem += (size_t)depth * kMlpWidth * 4 + 128;
  const size_t tile_bytes = (size_t)kFwdTile * kMlpWidth * 2;
  smem += 2 * tile_bytes;
  PP_REQUIRE(smem <= 227 * 1024, PP_ERR_UNSUPPORTED, "warpnet_forward: %zu B of shared memory needed", smem);
  const int64_t n_tiles = (n_rows + kFwdTile - 1) / kFwdTile;
  const int blocks = (int)min64((n_tiles + 1) / 2, sm_count());
  CUtensorMap no_map;
  memset(&no_map, 0, sizeof(no_map));
  PP_CHECK_CUDA(cudaFuncSetAttribute(rgbnet_forward_kernel<2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
  rgbnet_forward_kernel<2, true, false><<<blocks, kFwdThreads, smem, as_stream(stream)>>>(
      P, (const __nv_bfloat16*)x_rows, n_rows, 0, out, nullptr, nullptr, nullptr, n_rows_ptr, no_map);
  PP_CHECK_LAUNCH("warpnet_forward");
  return PP_OK;
}

extern "C" int pp_rgbnet_forward(const float* x, int64_t n_pts, int32_t d0, int32_t depth, int32_t width,
                                 const float* const* weights, const float* const* biases,
                                 int apply_sigmoid, float* out, void* workspace, int64_t workspace_bytes,
                                 pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0, PP_ERR_INVALID_ARG, "rgbnet_forward: n_pts < 0");
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(d0 >= 1 && d0 <= kMlpWidth, PP_ERR_UNSUPPORTED, "rgbnet_forward: 1 <= d0 <= %d", kMlpWidth);
  const int64_t need = n_pts * pp_rgbnet_padded_width(d0) * 2;
  PP_REQUIRE(workspace && workspace_bytes >= need, PP_ERR_INVALID_ARG,
             "rgbnet_forward: workspace of %lld bytes required (n * pp_rgbnet_padded_width(d0) * 2)",
             (long long)need);
  int rc = pp_rgbnet_pack_input(x, n_pts, d0, workspace, stream);
  if (rc != PP_OK) return rc;
  return pp_rgbnet_forward_packed(workspace, n_pts, d0, depth, width, weights, biases, apply_sigmoid, out,
                                  nullptr, nullptr, nullptr, nullptr, stream);
}
