// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, bf16 -> fp32, M = 128, K = 16) for the operand
// layouts / majors / N the rgbnet kernels use (UMMA canonical no-swizzle layout).  One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/umma_probe tools/probe/umma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../poseprobe_b200/csrc/mlp_common.cuh"

using namespace pp;

// single-thread issue (the slow form, kept here for the comparison)
__device__ __forceinline__ void umma_bf16_single(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
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
__device__ __forceinline__ void umma_commit_single(uint32_t bar_smem) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_smem) : "memory");
}

// warp-uniform issue: every lane runs the loop with identical values (descriptors stay in uniform registers), one
// elected lane issues the instruction
__device__ __forceinline__ void umma_bf16_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
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

__global__ void __launch_bounds__(128, 1) probe(int a_mn, int b_mn, int N, int reps, int swizzle_none, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // 128 x 128 bf16 tile
  uint8_t* sB = smem + 32768;         // 128 x 128 bf16 tile (or [128 x 16])
  uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(smem + 65536 + 16);
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(sBar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *sTmem;
  if (swizzle_none >= 2 && threadIdx.x < 32) {
    // mode 2: warp-uniform loop + elect; mode 3: same with constant descriptors (issue floor)
    const int bcols = (N == 16) ? 16 : 128;
    const uint32_t idesc = umma_idesc_bf16_major(128, N, a_mn, b_mn);
    uint32_t phase = 0;
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
    for (int round = 0; round < 2; ++round) {
      const long long t0 = clock64();
      uint64_t da = a_mn ? desc_mnmajor(a_base, 128, 0) : desc_kmajor(a_base, 128, 0);
      uint64_t db = b_mn ? desc_mnmajor(b_base, bcols, 0) : desc_kmajor(b_base, 128, 0);
      const uint32_t sa = a_mn ? 256u : 16u, sb = b_mn ? (bcols == 16 ? 32u : 256u) : 16u;
      for (int r = 0; r < reps; ++r) {
        umma_bf16_elect(tmem, da, db, idesc, r > 0);
        if (swizzle_none == 2) {
          if ((r & 7) == 7) { da -= 7 * sa; db -= 7 * sb; } else { da += sa; db += sb; }
        }
      }
      umma_commit(smem_u32(sBar));
      mbar_wait(smem_u32(sBar), phase);
      phase ^= 1;
      const long long t1 = clock64();
      if (round == 1 && blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    }
  } else if (swizzle_none < 2 && threadIdx.x == 0) {
    const int bcols = (N == 16) ? 16 : 128;
    const uint32_t idesc = umma_idesc_bf16_major(128, N, a_mn, b_mn);
    uint32_t phase = 0;
    for (int round = 0; round < 2; ++round) {   // round 0 = warm-up
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const int k = r & 7;
        const uint64_t da = a_mn ? desc_mnmajor(smem_u32(sA), 128, k) : desc_kmajor(smem_u32(sA), 128, k);
        const uint64_t db = b_mn ? desc_mnmajor(smem_u32(sB), bcols, k) : desc_kmajor(smem_u32(sB), b_mn ? bcols : 128, k);
        umma_bf16_single(tmem, da, db, idesc, r > 0);
      }
      umma_commit_single(smem_u32(sBar));
      mbar_wait(smem_u32(sBar), phase);
      phase ^= 1;
      const long long t1 = clock64();
      if (round == 1 && blockIdx.x == 0) out[0] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  const int reps = 512;
  struct V { int a, b, n; const char* name; } vs[] = {
      {0, 0, 128, "A K-major, B K-major, N=128 (forward layers)"}, {0, 0, 64, "A K, B K, N=64"}, {0, 0, 16, "A K, B K, N=16 (logits)"},
      {0, 1, 128, "A K, B MN, N=128 (dH = dZ W)"}, {0, 1, 64, "A K, B MN, N=64"},
      {1, 1, 128, "A MN, B MN, N=128 (dW = dZ^T H)"}, {1, 1, 64, "A MN, B MN, N=64 (dW_0)"}, {1, 1, 16, "A MN, B MN, N=16 (db, dW_last)"}};
  for (int mode : {1, 2, 3}) {
    const int grid = 148;
    printf("mode %d (%s), grid = %d CTAs, %d MMAs back to back, M=128 K=16 bf16 (ideal: N/128 * 64 cycles)\n", mode,
           mode == 1 ? "single thread issues, descriptors rebuilt per MMA"
                     : mode == 2 ? "warp-uniform loop + elect.sync, descriptors advanced by adds"
                                 : "warp-uniform loop + elect.sync, constant descriptors", grid, reps);
    for (auto& v : vs) {
      probe<<<grid, 128, 70000>>>(v.a, v.b, v.n, reps, mode, d);
      long long c = 0;
      cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaGetLastError();
      printf("  %-48s %8.1f cycles / MMA   %s\n", v.name, (double)c / reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
