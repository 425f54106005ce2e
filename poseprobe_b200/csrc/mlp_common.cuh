// tcgen05 / TMEM / mbarrier helpers shared by the rgbnet kernels (mlp.cu, mlp_bwd.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "common.cuh"

namespace pp {

constexpr int kMlpWidth = 128;     // hidden width (rgbnet_width, default_fine_s.py / scan1.py:116)
constexpr int kMlpOutPad = 16;     // 3 outputs padded to the smallest legal UMMA N for M = 128
constexpr int kMlpMaxDepth = 5;

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

// same with fp16 operands (a_format = b_format = 0): 11 significant bits instead of 8 at the same rate
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// MMA issue.  MUST be called by all 32 lanes of a warp with identical arguments: the loop around it then runs on
// warp-uniform values (descriptors live in uniform registers) and `elect.sync` picks the one lane that issues.
// Measured on B200 (tools/probe/umma_probe.cu): a loop executed by a single thread (`if (lane == 0)`) issues one
// tcgen05.mma per ~120 cycles whatever its size; the warp-uniform form reaches the execution rate (64 cycles for
// M=128, N=128, K=16; 48 for N=64; ~40 for N=16), K-major and MN-major operands alike.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// commit of everything the warp's elected lane issued so far (same calling convention as umma_bf16)
__device__ __forceinline__ void umma_commit(uint32_t bar_smem) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(bar_smem)
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

__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- operand descriptors for a [rows x cols] bf16 tile stored in the canonical no-swizzle layout ---------
// (8x8 core matrices of 128 B; element (r, c) at canon_off(r, c, cols)).
// K-major view: the contraction runs over the tile's columns (forward: A = activations, B = weights).
//   slice = which 16 columns; LBO = 128 B between the two core matrices of a slice, SBO = row-group stride.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int cols, int slice) {
  return umma_desc(base + slice * 256, 128, static_cast<uint32_t>(cols >> 3) << 7);
}
// MN-major view: the contraction runs over the tile's ROWS (samples for the weight gradients, output
// features for dH = dZ W), the tile's columns are the M / N index.  Canonical layout (cute, Major::MN,
// INTERLEAVE): ((8,1,m),(8,k)) : ((1,8,SBO),(8,LBO)) elements - a core matrix holds 8 k-rows of 16 B, each
// 8 consecutive MN elements, which is exactly how a row-group of the tile is stored.  SBO = 128 B (next
// 8 columns), LBO = row-group stride (next 8 rows); slice = which 16 rows.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int cols, int slice) {
  const uint32_t row_group = static_cast<uint32_t>(cols >> 3) << 7;
  return umma_desc(base + slice * 2 * row_group, row_group, 128);
}

// instruction descriptor with operand majors (bit 15: A is MN-major, bit 16: B is MN-major)
__device__ __forceinline__ uint32_t umma_idesc_bf16_major(int M, int N, int a_mn, int b_mn) {
  return umma_idesc_bf16(M, N) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16);
}

// ---- TMA (cp.async.bulk.tensor) pieces for the input tiles of the forward kernel ---------------------------
// A [rows x k0p] bf16 tile arrives as k0p / 8 two-dimensional boxes of {8 columns (16 B), 128 rows}: TMA writes a box
// row after row, 16 B each, which stacks the tile's 8 x 8 core matrices of one column chunk contiguously (128 B per
// core matrix, 2 KB per chunk).  Element (r, c) therefore sits at (c / 8) * 2048 + (r / 8) * 128 + (r % 8) * 16 +
// (c % 8) * 2: still a no-swizzle K-major operand, with LBO = 2048 B between the two column chunks of a K = 16 slice and
// SBO = 128 B between row groups.
__device__ __forceinline__ uint64_t desc_kmajor_chunked(uint32_t base, int slice) {
  return umma_desc(base + slice * 2 * 2048, 2048, 128);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar_smem, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const void* tensor_map, int c0, int c1,
                                            uint32_t bar_smem) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst_smem), "l"(tensor_map), "r"(c0), "r"(c1), "r"(bar_smem)
      : "memory");
}
// mbarrier wait with a cycle budget: a transfer that never completes (a bad tensor map) traps instead of hanging
__device__ __forceinline__ void mbar_wait_or_trap(uint32_t bar_smem, uint32_t parity) {
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
        : "r"(bar_smem), "r"(parity)
        : "memory");
    if (ok) return;
    if ((spin & 255) == 255 && clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// bias + ReLU + bf16 rounding of 8 accumulator columns: packed fp32 adds, the ReLU on the packed bf16 pair
// (max commutes with the rounding, so this equals round(relu(acc + b)))
__device__ __forceinline__ uint4 bias_relu_pack8(const uint32_t* r, const float4 b0, const float4 b1) {
  const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
  const float2 s0 = __fadd2_rn(make_float2(__uint_as_float(r[0]), __uint_as_float(r[1])), make_float2(b0.x, b0.y));
  const float2 s1 = __fadd2_rn(make_float2(__uint_as_float(r[2]), __uint_as_float(r[3])), make_float2(b0.z, b0.w));
  const float2 s2 = __fadd2_rn(make_float2(__uint_as_float(r[4]), __uint_as_float(r[5])), make_float2(b1.x, b1.y));
  const float2 s3 = __fadd2_rn(make_float2(__uint_as_float(r[6]), __uint_as_float(r[7])), make_float2(b1.z, b1.w));
  __nv_bfloat162 p[4] = {__hmax2(__floats2bfloat162_rn(s0.x, s0.y), zero2), __hmax2(__floats2bfloat162_rn(s1.x, s1.y), zero2),
                         __hmax2(__floats2bfloat162_rn(s2.x, s2.y), zero2), __hmax2(__floats2bfloat162_rn(s3.x, s3.y), zero2)};
  return *reinterpret_cast<uint4*>(p);
}

// dZ = dH * (H > 0) for 8 columns, rounded to bf16
__device__ __forceinline__ uint4 relu_grad_pack8(const uint32_t* r, const uint4 act) {
  const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
  const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(&act);
  __nv_bfloat162 p[4];
#pragma unroll
  for (int q = 0; q < 4; ++q)
    p[q] = __hmul2(__floats2bfloat162_rn(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1])), __hgt2(a[q], zero2));
  return *reinterpret_cast<uint4*>(p);
}

}  // namespace pp
