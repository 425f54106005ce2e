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
#include <cuda_bf16.h>

#include "common.cuh"

namespace pp {

constexpr int kMlpThreads = 128;
constexpr int kMlpWidth = 128;     // hidden width (rgbnet_width, default_fine_s.py / scan1.py:116)
constexpr int kMlpOutPad = 16;     // 3 outputs padded to the smallest legal UMMA N for M = 128
constexpr int kMlpMaxDepth = 5;
constexpr int kTmemCols = 256;     // [0,128): hidden accumulator, [128,144): output accumulator

struct MlpParams {
  const float* w[kMlpMaxDepth];
  const float* b[kMlpMaxDepth];
  int depth;      // number of Linear layers
  int d0;         // input features
  int k0p;        // d0 padded to a multiple of 16
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// byte offset of element (r, k) of an [R x K] K-major operand in the canonical no-swizzle layout
__device__ __forceinline__ uint32_t canon_off(int r, int k, int K) {
  return static_cast<uint32_t>((((r >> 3) * (K >> 3) + (k >> 3)) << 7) + ((r & 7) << 4) + ((k & 7) << 1));
}

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=0 (SWIZZLE_NONE) [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF) |
         (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) |
         (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// UMMA instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: D=f32, A=B=bf16, both K-major
__device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar_smem) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_smem)
               : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar_smem, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_smem), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar_smem, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_%=;\n\t"
      "}\n" ::"r"(bar_smem),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);   // .x (low half) = lo
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared memory map (bytes, all 128-byte aligned):
//   [0, 32K)            A tile, 128 x up-to-128 bf16
//   then                layer 0 weights 128 x k0p, hidden weights 128 x 128 each, last weights 16 x 128
//   then                biases fp32 (depth x 128)
//   then                mbarrier (8 B), tmem base (4 B)
__global__ void __launch_bounds__(kMlpThreads, 1)
rgbnet_forward_kernel(MlpParams P, const float* __restrict__ x, int64_t M, int apply_sigmoid,
                      float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sA = smem;
  uint8_t* sW[kMlpMaxDepth];
  uint32_t off = kMlpThreads * kMlpWidth * 2;
  for (int l = 0; l < P.depth; ++l) {
    sW[l] = smem + off;
    const int rows = (l == P.depth - 1) ? kMlpOutPad : kMlpWidth;
    const int K = (l == 0) ? P.k0p : kMlpWidth;
    off += rows * K * 2;
  }
  float* sBias = reinterpret_cast<float*>(smem + off);
  off += P.depth * kMlpWidth * 4;
  uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + off);
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(smem + off + 8);

  // ---- prologue: weights -> bf16 canonical layout, biases, TMEM, mbarrier ----
  for (int l = 0; l < P.depth; ++l) {
    const bool last = (l == P.depth - 1);
    const int rows = last ? kMlpOutPad : kMlpWidth;
    const int real_rows = last ? 3 : kMlpWidth;
    const int K = (l == 0) ? P.k0p : kMlpWidth;
    const int real_K = (l == 0) ? P.d0 : kMlpWidth;
    for (int e = tid; e < rows * K; e += kMlpThreads) {
      const int r = e / K, k = e - r * K;
      const float wv = (r < real_rows && k < real_K) ? P.w[l][r * real_K + k] : 0.f;
      *reinterpret_cast<__nv_bfloat16*>(sW[l] + canon_off(r, k, K)) = __float2bfloat16_rn(wv);
    }
    for (int e = tid; e < kMlpWidth; e += kMlpThreads)
      sBias[l * kMlpWidth + e] = (e < real_rows) ? P.b[l][e] : 0.f;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(smem_u32(sBar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *sTmem;
  const uint32_t bar = smem_u32(sBar);
  uint32_t phase = 0;

  const int64_t n_tiles = (M + kMlpThreads - 1) / kMlpThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row = tile * kMlpThreads + tid;
    const bool valid = row < M;
    // ---- stage the input rows (fp32 -> bf16) into the A tile ----
    {
      const float* xr = x + row * P.d0;
      for (int k8 = 0; k8 < (P.k0p >> 3); ++k8) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = k8 * 8 + j;
          f[j] = (valid && k < P.d0) ? __ldg(xr + k) : 0.f;
        }
        uint4 pk;
        pk.x = pack_bf16(f[0], f[1]);
        pk.y = pack_bf16(f[2], f[3]);
        pk.z = pack_bf16(f[4], f[5]);
        pk.w = pack_bf16(f[6], f[7]);
        *reinterpret_cast<uint4*>(sA + canon_off(tid, k8 * 8, P.k0p)) = pk;
      }
    }
    fence_async_smem();
    __syncthreads();

    for (int l = 0; l < P.depth; ++l) {
      const bool last = (l == P.depth - 1);
      const int K = (l == 0) ? P.k0p : kMlpWidth;
      const int N = last ? kMlpOutPad : kMlpWidth;
      const uint32_t d_tmem = tmem + (last ? kMlpWidth : 0);
      if (tid == 0) {
        tc_fence_after();
        const uint32_t sbo = static_cast<uint32_t>(K >> 3) << 7;   // bytes between 8-row groups
        const uint32_t idesc = umma_idesc_bf16(kMlpThreads, N);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sW[l]);
        for (int k = 0; k < (K >> 4); ++k) {
          // one K=16 slice = two core matrices along K = 256 B further into both operands
          umma_bf16(d_tmem, umma_desc(a0 + k * 256, 128, sbo), umma_desc(b0 + k * 256, 128, sbo), idesc, k > 0);
        }
        umma_commit(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
      if (!last) {
        const float* bias = sBias + l * kMlpWidth;
#pragma unroll 1
        for (int c0 = 0; c0 < kMlpWidth; c0 += 32) {
          float v[32];
          tmem_ld32(d_tmem + lane_base + c0, v);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float a[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = fmaxf(v[j + q] + bias[c0 + j + q], 0.f);
            uint4 pk;
            pk.x = pack_bf16(a[0], a[1]);
            pk.y = pack_bf16(a[2], a[3]);
            pk.z = pack_bf16(a[4], a[5]);
            pk.w = pack_bf16(a[6], a[7]);
            *reinterpret_cast<uint4*>(sA + canon_off(tid, c0 + j, kMlpWidth)) = pk;
          }
        }
        tc_fence_before();
        fence_async_smem();
        __syncthreads();
      } else {
        float v[16];
        tmem_ld16(d_tmem + lane_base, v);
        if (valid) {
          const float* bias = sBias + l * kMlpWidth;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float z = v[j] + bias[j];
            out[row * 3 + j] = apply_sigmoid ? 1.f / (1.f + expf(-z)) : z;
          }
        }
        tc_fence_before();
        __syncthreads();
      }
    }
  }

  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace pp

using namespace pp;

extern "C" int pp_rgbnet_forward(const float* x, int64_t n_pts, int32_t d0, int32_t depth, int32_t width,
                                 const float* const* weights, const float* const* biases,
                                 int apply_sigmoid, float* out, pp_stream_t stream) {
  PP_REQUIRE(n_pts >= 0, PP_ERR_INVALID_ARG, "rgbnet_forward: n_pts < 0");
  PP_REQUIRE(width == kMlpWidth, PP_ERR_UNSUPPORTED, "rgbnet_forward: hidden width must be %d", kMlpWidth);
  PP_REQUIRE(depth >= 2 && depth <= kMlpMaxDepth, PP_ERR_UNSUPPORTED, "rgbnet_forward: depth must be 2..%d",
             kMlpMaxDepth);
  PP_REQUIRE(d0 >= 1 && d0 <= kMlpWidth, PP_ERR_UNSUPPORTED, "rgbnet_forward: 1 <= d0 <= %d", kMlpWidth);
  if (n_pts == 0) return PP_OK;
  PP_REQUIRE(x && weights && biases && out, PP_ERR_INVALID_ARG, "rgbnet_forward: null pointer");
  MlpParams P;
  P.depth = depth;
  P.d0 = d0;
  P.k0p = (d0 + 15) / 16 * 16;
  size_t smem = (size_t)kMlpThreads * kMlpWidth * 2;
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
  smem += (size_t)depth * kMlpWidth * 4 + 16;
  PP_REQUIRE(smem <= 227 * 1024, PP_ERR_UNSUPPORTED, "rgbnet_forward: %zu B of shared memory needed", smem);
  PP_CHECK_CUDA(cudaFuncSetAttribute(rgbnet_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_tiles = (n_pts + kMlpThreads - 1) / kMlpThreads;
  const int blocks = (int)min64(n_tiles, sm_count());
  rgbnet_forward_kernel<<<blocks, kMlpThreads, smem, as_stream(stream)>>>(P, x, n_pts, apply_sigmoid, out);
  PP_CHECK_LAUNCH("rgbnet_forward");
  return PP_OK;
}
