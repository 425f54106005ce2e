// Multi-GPU voxel exchange fused with the optimiser (see include/poseprobe_b200_fused.h).
//
// One kernel per rank = reduce-scatter (peer loads of every rank's gradient slice over NVLink 5) +
// masked Adam on the slice + all-gather (peer stores of the updated records into every rank's
// parameters).  Per rank and step the NVLink traffic is (world-1)/world of the gradient records in and
// the same amount of parameter records out - the two directions of the link are used concurrently -
// and the dense optimiser tail shrinks by 1/world because every voxel is updated exactly once.
#include <stdlib.h>
#include <string.h>

#include "../../include/poseprobe_b200_fused.h"
#include "common.cuh"

namespace pp {

constexpr int kMaxWorld = 16;
struct PeerPtrs {
  float4* vox[kMaxWorld];
  const float4* grad[kMaxWorld];
};

struct SyncBlock {
  unsigned int flags[kMaxWorld];
  float payload[kMaxWorld][2];
  unsigned int timed_out;
};
struct SyncPtrs {
  SyncBlock* blk[kMaxWorld];
};

// Cycle budget of a barrier wait.  A rank that stays away longer (dead process, debugger) makes its peers raise the
// STICKY timed_out flag of their sync block: from then on the exchange kernels do nothing and report NaN as the global
// loss, and the host side (dist.py, pp_peer_sync_timed_out) raises - replicas are never updated from partial sums.
constexpr long long kPeerWaitCycles = 120000000000LL;

__device__ __forceinline__ bool peer_timed_out(const SyncPtrs& sp, int me) {
  return *reinterpret_cast<volatile unsigned int*>(&sp.blk[me]->timed_out) != 0u;
}

// announce epoch E to every peer (optionally with a 2-float payload) and wait until every peer announced it
// (release / acquire at system scope on flags in peer memory).  Executed by threads r < world of a warp.
__device__ __forceinline__ void peer_announce_and_wait(const SyncPtrs& sp, int world, int me, unsigned int epoch,
                                                       const float* payload_in, bool announce) {
  const int r = threadIdx.x;
  SyncBlock* mine = sp.blk[me];
  if (r < world) {
    if (announce) {
      SyncBlock* peer = sp.blk[r];
      if (payload_in) {
        peer->payload[me][0] = payload_in[0];
        peer->payload[me][1] = payload_in[1];
      }
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&peer->flags[me]), "r"(epoch) : "memory");
    }
    const long long t0 = clock64();
    for (;;) {
      unsigned int v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(&mine->flags[r]) : "memory");
      if ((int)(v - epoch) >= 0) break;
      if (clock64() - t0 > kPeerWaitCycles) {   // ~60 s: a peer died; do not hang the device ...
        mine->timed_out = 1;                     // ... and never run the exchange on half-written gradients:
        break;                                   // every caller checks the flag (sticky) and skips its work
      }
    }
  }
}

__device__ __forceinline__ void adam_elem_x(float& p, float g, float& m, float& v, float step_size,
                                            float b1, float b2, float eps, bool dense = false) {
  if (!dense && g == 0.f) return;   // masked_adam_upd (adam_upd_kernel.cu:35); dense: lib/utils.py:176-198
  m = fma_rn(b1, m, mul_rn(1.f - b1, g));
  v = fma_rn(b2, v, mul_rn(mul_rn(1.f - b2, g), g));
  p = p - mul_rn(step_size, m) / (sqrtf(v) + eps);
}

// Software pipeline.  A slice is only a few MB, so "every thread loads, then every thread stores" (one wave)
// uses the two directions of the link one after the other: all gradient records come in, then all parameter
// records go out, ~2x the one-direction time.  Instead a thread walks kRecordsPerThread records with its loads
// issued kDepth records ahead: while record i is updated and stored to the peers, the loads of records
// i+1..i+kDepth are in flight, so both directions carry traffic for the whole kernel.
constexpr int kXchgThreads = 256;
constexpr int kRecordsPerThread = 8;

template <int kWorld>
__global__ void __launch_bounds__(kXchgThreads)
voxel_adam_exchange_kernel(PeerPtrs pp, float4* __restrict__ m4, float4* __restrict__ v4,
                           int64_t begin, int64_t end, int rank, float ss_density, float ss_k0,
                           float b1, float b2, float eps, int dbg, const unsigned int* __restrict__ timed_out) {
  constexpr int kDepth = 2;
  if (timed_out && *reinterpret_cast<const volatile unsigned int*>(timed_out) != 0u) return;   // see kPeerWaitCycles
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t first = begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  // one 16-byte load per rank and record (fixed rank order keeps the sum deterministic run to run)
  float4 gr[kDepth][kWorld], pq[kDepth], mq[kDepth], vq[kDepth];
#pragma unroll
  for (int d = 0; d < kDepth; ++d) {
    const int64_t i = first + d * stride;
    const bool in = i < end;
#pragma unroll
    for (int r = 0; r < kWorld; ++r) gr[d][r] = in ? pp.grad[(dbg & 2) ? rank : r][i] : zero;
    pq[d] = in ? pp.vox[rank][i] : zero;
    mq[d] = in ? m4[i - begin] : zero;
    vq[d] = in ? v4[i - begin] : zero;
  }
  for (int64_t i = first; i < end; i += stride) {
    float4 g = gr[0][0];
#pragma unroll
    for (int r = 1; r < kWorld; ++r) {
      g.x += gr[0][r].x;
      g.y += gr[0][r].y;
      g.z += gr[0][r].z;
      g.w += gr[0][r].w;
    }
    float4 p = pq[0], m = mq[0], v = vq[0];
    // rotate the pipeline and issue the loads of record i + kDepth*stride
#pragma unroll
    for (int d = 0; d + 1 < kDepth; ++d) {
#pragma unroll
      for (int r = 0; r < kWorld; ++r) gr[d][r] = gr[d + 1][r];
      pq[d] = pq[d + 1];
      mq[d] = mq[d + 1];
      vq[d] = vq[d + 1];
    }
    {
      const int64_t n = i + kDepth * stride;
      const bool in = n < end;
#pragma unroll
      for (int r = 0; r < kWorld; ++r) gr[kDepth - 1][r] = in ? pp.grad[(dbg & 2) ? rank : r][n] : zero;
      pq[kDepth - 1] = in ? pp.vox[rank][n] : zero;
      mq[kDepth - 1] = in ? m4[n - begin] : zero;
      vq[kDepth - 1] = in ? v4[n - begin] : zero;
    }
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
    adam_elem_x(p.x, g.x, m.x, v.x, ss_density, b1, b2, eps);
    adam_elem_x(p.y, g.y, m.y, v.y, ss_k0, b1, b2, eps);
    adam_elem_x(p.z, g.z, m.z, v.z, ss_k0, b1, b2, eps);
    adam_elem_x(p.w, g.w, m.w, v.w, ss_k0, b1, b2, eps);
    m4[i - begin] = m;
    v4[i - begin] = v;
    // gather: the updated record goes to every replica
#pragma unroll
    for (int r = 0; r < kWorld; ++r)
      if (!(dbg & 1) || r == rank) pp.vox[r][i] = p;
  }
}

// ---- NVSwitch variant: in-switch reduction and multicast ----------------------------------------------
// With plain peer loads/stores every link direction carries (w-1)/w of the gradient records AND (w-1)/w of
// the parameter records.  On an NVSwitch system the same two steps can be expressed on a multicast address:
// multimem.ld_reduce makes the switch fetch the record from every replica, add them (fp32) and return one
// 16-byte result; multimem.st makes the switch replicate one 16-byte store into every replica.  Per GPU and
// direction the traffic drops from 2(w-1)/w to (w+1)/w of the record array.
__device__ __forceinline__ float4 mc_ld_reduce_add(const float4* mc_addr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc_addr)
               : "memory");
  return r;
}
__device__ __forceinline__ void mc_st(float4* mc_addr, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_addr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

template <int kDepth>
__global__ void __launch_bounds__(kXchgThreads)
voxel_adam_exchange_mc_kernel(const float4* __restrict__ vox_local, float4* vox_mc, const float4* grad_mc,
                              float4* __restrict__ m4, float4* __restrict__ v4, int64_t begin, int64_t end,
                              float ss_density, float ss_k0, float b1, float b2, float eps, SyncPtrs sp, int world,
                              int me, unsigned int epoch, const float* __restrict__ payload_in,
                              float* __restrict__ payload_sum, int dense) {
  // optional fused "all gradients complete" barrier: block 0 announces, every block waits for all peers before
  // its first in-switch reduction (saves the separate one-warp barrier launch in front of this kernel)
  if (epoch != 0u) {
    if (threadIdx.x < 32) peer_announce_and_wait(sp, world, me, epoch, payload_in, blockIdx.x == 0);
    __threadfence();
    __syncthreads();
    const bool dead = peer_timed_out(sp, me);
    if (payload_sum && blockIdx.x == 0 && threadIdx.x == 0) {
      const SyncBlock* mine = sp.blk[me];
      float a = 0.f, b = 0.f;
      for (int q = 0; q < world; ++q) {
        a += mine->payload[q][0];
        b += mine->payload[q][1];
      }
      payload_sum[0] = dead ? __int_as_float(0x7fc00000) : a;
      payload_sum[1] = dead ? __int_as_float(0x7fc00000) : b;
    }
    if (dead) return;   // a peer never announced: its gradient records may be half written
  } else if (sp.blk[me] && peer_timed_out(sp, me)) {
    return;             // the separate barrier in front of this launch timed out
  }
  // same software pipeline as the peer-load variant: the in-switch reductions of the next records are in
  // flight (up-link: every GPU's contribution) while this record's multicast store goes out (down-link)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t first = begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 gq[kDepth], pq[kDepth], mq[kDepth], vq[kDepth];
#pragma unroll
  for (int d = 0; d < kDepth; ++d) {
    const int64_t i = first + d * stride;
    const bool in = i < end;
    gq[d] = in ? mc_ld_reduce_add(grad_mc + i) : zero;
    pq[d] = in ? vox_local[i] : zero;
    mq[d] = in ? m4[i - begin] : zero;
    vq[d] = in ? v4[i - begin] : zero;
  }
  for (int64_t i = first; i < end; i += stride) {
    const float4 g = gq[0];
    float4 p = pq[0], m = mq[0], v = vq[0];
#pragma unroll
    for (int d = 0; d + 1 < kDepth; ++d) {
      gq[d] = gq[d + 1];
      pq[d] = pq[d + 1];
      mq[d] = mq[d + 1];
      vq[d] = vq[d + 1];
    }
    {
      const int64_t n = i + kDepth * stride;
      const bool in = n < end;
      gq[kDepth - 1] = in ? mc_ld_reduce_add(grad_mc + n) : zero;
      pq[kDepth - 1] = in ? vox_local[n] : zero;
      mq[kDepth - 1] = in ? m4[n - begin] : zero;
      vq[kDepth - 1] = in ? v4[n - begin] : zero;
    }
    if (!dense && g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
    adam_elem_x(p.x, g.x, m.x, v.x, ss_density, b1, b2, eps, dense);
    adam_elem_x(p.y, g.y, m.y, v.y, ss_k0, b1, b2, eps, dense);
    adam_elem_x(p.z, g.z, m.z, v.z, ss_k0, b1, b2, eps, dense);
    adam_elem_x(p.w, g.w, m.w, v.w, ss_k0, b1, b2, eps, dense);
    m4[i - begin] = m;
    v4[i - begin] = v;
    mc_st(vox_mc + i, p);
  }
}

// In-place sum all-reduce of float4 records through the switch: every rank reduces its slice of the multicast
// address (multimem.ld_reduce) and writes the sum back to every replica (multimem.st) - the same (w+1)/w traffic per
// direction as the exchange kernel, no optimiser.  Used for the small or stencil-coupled gradients of the fine stage
// (SDF gradient records, rgbnet gradients), which every rank then applies identically.
__global__ void __launch_bounds__(kXchgThreads)
allreduce_mc_kernel(float4* buf_mc, int64_t begin, int64_t end, SyncPtrs sp, int me) {
  if (sp.blk[me] && peer_timed_out(sp, me)) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int kDepth = 4;
  for (int64_t i0 = begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < end; i0 += kDepth * stride) {
    float4 v[kDepth];
#pragma unroll
    for (int d = 0; d < kDepth; ++d)
      if (i0 + d * stride < end) v[d] = mc_ld_reduce_add(buf_mc + i0 + d * stride);
#pragma unroll
    for (int d = 0; d < kDepth; ++d)
      if (i0 + d * stride < end) mc_st(buf_mc + i0 + d * stride, v[d]);
  }
}

// ---- hybrid: in-switch reduction AND plain peer loads / stores, side by side ------------------------------
// Measured at 8 ranks (profiles/r1_results.md): the multimem path moves (w+1)/w of the record array per link
// direction but is held at ~460 GB/s by the switch's reduction / multicast rate, while plain peer loads / stores
// move 2(w-1)/w of it at ~600 GB/s - neither fills the links.  The two limits are different resources, so one
// launch runs both: the first mc_records of the slice go through the switch (blocks [0, mc_blocks)), the rest of
// the slice through peer loads (fixed rank order) and peer stores (the remaining blocks), concurrently.
template <int kWorld>
__global__ void __launch_bounds__(kXchgThreads, 2)
voxel_adam_exchange_hybrid_kernel(PeerPtrs pp, float4* vox_mc, const float4* grad_mc, float4* __restrict__ m4,
                                  float4* __restrict__ v4, int64_t begin, int64_t split, int64_t end, int mc_blocks,
                                  int rank, float ss_density, float ss_k0, float b1, float b2, float eps, SyncPtrs sp,
                                  unsigned int epoch, const float* __restrict__ payload_in,
                                  float* __restrict__ payload_sum) {
  if (epoch != 0u) {
    if (threadIdx.x < 32) peer_announce_and_wait(sp, kWorld, rank, epoch, payload_in, blockIdx.x == 0);
    __threadfence();
    __syncthreads();
    const bool dead = peer_timed_out(sp, rank);
    if (payload_sum && blockIdx.x == 0 && threadIdx.x == 0) {
      const SyncBlock* mine = sp.blk[rank];
      float a = 0.f, b = 0.f;
      for (int q = 0; q < kWorld; ++q) {
        a += mine->payload[q][0];
        b += mine->payload[q][1];
      }
      payload_sum[0] = dead ? __int_as_float(0x7fc00000) : a;
      payload_sum[1] = dead ? __int_as_float(0x7fc00000) : b;
    }
    if (dead) return;
  } else if (sp.blk[rank] && peer_timed_out(sp, rank)) {
    return;
  }
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* vox_local = pp.vox[rank];
  if ((int)blockIdx.x < mc_blocks) {
    // ---- switch path over [begin, split) ----
    constexpr int kDepth = 3;
    const int64_t stride = (int64_t)mc_blocks * blockDim.x;
    const int64_t first = begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    float4 gq[kDepth], pq[kDepth], mq[kDepth], vq[kDepth];
#pragma unroll
    for (int d = 0; d < kDepth; ++d) {
      const int64_t i = first + d * stride;
      const bool in = i < split;
      gq[d] = in ? mc_ld_reduce_add(grad_mc + i) : zero;
      pq[d] = in ? vox_local[i] : zero;
      mq[d] = in ? m4[i - begin] : zero;
      vq[d] = in ? v4[i - begin] : zero;
    }
    for (int64_t i = first; i < split; i += stride) {
      const float4 g = gq[0];
      float4 p = pq[0], m = mq[0], v = vq[0];
#pragma unroll
      for (int d = 0; d + 1 < kDepth; ++d) {
        gq[d] = gq[d + 1];
        pq[d] = pq[d + 1];
        mq[d] = mq[d + 1];
        vq[d] = vq[d + 1];
      }
      {
        const int64_t n = i + kDepth * stride;
        const bool in = n < split;
        gq[kDepth - 1] = in ? mc_ld_reduce_add(grad_mc + n) : zero;
        pq[kDepth - 1] = in ? vox_local[n] : zero;
        mq[kDepth - 1] = in ? m4[n - begin] : zero;
        vq[kDepth - 1] = in ? v4[n - begin] : zero;
      }
      if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
      adam_elem_x(p.x, g.x, m.x, v.x, ss_density, b1, b2, eps);
      adam_elem_x(p.y, g.y, m.y, v.y, ss_k0, b1, b2, eps);
      adam_elem_x(p.z, g.z, m.z, v.z, ss_k0, b1, b2, eps);
      adam_elem_x(p.w, g.w, m.w, v.w, ss_k0, b1, b2, eps);
      m4[i - begin] = m;
      v4[i - begin] = v;
      mc_st(vox_mc + i, p);
    }
  } else {
    // ---- peer path over [split, end) ----
    constexpr int kDepth = 2;
    const int nb = (int)gridDim.x - mc_blocks;
    const int64_t stride = (int64_t)nb * blockDim.x;
    const int64_t first = split + (blockIdx.x - mc_blocks) * (int64_t)blockDim.x + threadIdx.x;
    float4 gr[kDepth][kWorld], pq[kDepth], mq[kDepth], vq[kDepth];
#pragma unroll
    for (int d = 0; d < kDepth; ++d) {
      const int64_t i = first + d * stride;
      const bool in = i < end;
#pragma unroll
      for (int r = 0; r < kWorld; ++r) gr[d][r] = in ? pp.grad[r][i] : zero;
      pq[d] = in ? vox_local[i] : zero;
      mq[d] = in ? m4[i - begin] : zero;
      vq[d] = in ? v4[i - begin] : zero;
    }
    for (int64_t i = first; i < end; i += stride) {
      float4 g = gr[0][0];
#pragma unroll
      for (int r = 1; r < kWorld; ++r) {
        g.x += gr[0][r].x;
        g.y += gr[0][r].y;
        g.z += gr[0][r].z;
        g.w += gr[0][r].w;
      }
      float4 p = pq[0], m = mq[0], v = vq[0];
#pragma unroll
      for (int d = 0; d + 1 < kDepth; ++d) {
#pragma unroll
        for (int r = 0; r < kWorld; ++r) gr[d][r] = gr[d + 1][r];
        pq[d] = pq[d + 1];
        mq[d] = mq[d + 1];
        vq[d] = vq[d + 1];
      }
      {
        const int64_t n = i + kDepth * stride;
        const bool in = n < end;
#pragma unroll
        for (int r = 0; r < kWorld; ++r) gr[kDepth - 1][r] = in ? pp.grad[r][n] : zero;
        pq[kDepth - 1] = in ? vox_local[n] : zero;
        mq[kDepth - 1] = in ? m4[n - begin] : zero;
        vq[kDepth - 1] = in ? v4[n - begin] : zero;
      }
      if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
      adam_elem_x(p.x, g.x, m.x, v.x, ss_density, b1, b2, eps);
      adam_elem_x(p.y, g.y, m.y, v.y, ss_k0, b1, b2, eps);
      adam_elem_x(p.z, g.z, m.z, v.z, ss_k0, b1, b2, eps);
      adam_elem_x(p.w, g.w, m.w, v.w, ss_k0, b1, b2, eps);
      m4[i - begin] = m;
      v4[i - begin] = v;
#pragma unroll
      for (int r = 0; r < kWorld; ++r) pp.vox[r][i] = p;
    }
  }
}

// ---- cross-rank stream barrier over peer memory -----------------------------------------------------
// Each rank owns a small IPC-shared sync block.  Rank `me` announces epoch E by storing E into slot `me`
// of every peer's block (release, system scope) and then waits until all slots of its own block reached E
// (acquire).  An optional 2-float payload travels with the announcement and is summed in rank order, which
// is how the global loss is formed without a collective.  One warp; launched in stream order, so it
// separates "my gradient scatter is complete" from "I start reading everybody's gradients".
__global__ void peer_sync_kernel(SyncPtrs sp, int world, int me, unsigned int epoch,
                                 const float* __restrict__ payload_in, float* __restrict__ payload_sum) {
  peer_announce_and_wait(sp, world, me, epoch, payload_in, true);
  __threadfence();
  __syncwarp();
  if (payload_sum && threadIdx.x == 0) {
    const bool dead = peer_timed_out(sp, me);
    const SyncBlock* mine = sp.blk[me];
    float a = 0.f, b = 0.f;
    for (int q = 0; q < world; ++q) {
      a += mine->payload[q][0];
      b += mine->payload[q][1];
    }
    payload_sum[0] = dead ? __int_as_float(0x7fc00000) : a;
    payload_sum[1] = dead ? __int_as_float(0x7fc00000) : b;
  }
}

}  // namespace pp

using namespace pp;

// threads walk kRecordsPerThread records each (PP_XCHG_RPT overrides, for timing experiments)
static int exchange_blocks(int64_t records) {
  const char* e = getenv("PP_XCHG_RPT");
  const int rpt = (e && atoi(e) > 0) ? atoi(e) : kRecordsPerThread;
  const int64_t per_block = (int64_t)kXchgThreads * rpt;
  return (int)min64((records + per_block - 1) / per_block, (int64_t)sm_count() * 8);
}

static float adam_step_size_x(float lr, float beta1, float beta2, int step) {
  return lr * sqrtf(1.f - powf(beta2, (float)step)) / (1.f - powf(beta1, (float)step));
}


extern "C" int pp_shared_alloc(int64_t bytes, void** ptr_out, void* handle_out) {
  PP_REQUIRE(bytes > 0 && ptr_out && handle_out, PP_ERR_INVALID_ARG, "shared_alloc: bad argument");
  void* p = nullptr;
  PP_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return PP_ERR_CUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == PP_IPC_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  *ptr_out = p;
  return PP_OK;
}

extern "C" int pp_shared_open(const void* handle, void** ptr_out) {
  PP_REQUIRE(handle && ptr_out, PP_ERR_INVALID_ARG, "shared_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  PP_CHECK_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return PP_OK;
}

extern "C" int pp_shared_close(void* ptr) {
  if (ptr) PP_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return PP_OK;
}

extern "C" int pp_shared_free(void* ptr) {
  if (ptr) PP_CHECK_CUDA(cudaFree(ptr));
  return PP_OK;
}

static int exchange_mc_impl(const void* vox_local, void* vox_mc, const void* grad_mc, void* exp_avg,
                            void* exp_avg_sq, int64_t n_voxels, int32_t world, int32_t rank, int32_t step,
                            float beta1, float beta2, float lr_density, float lr_k0, float eps,
                            void* const* sync_ptrs, uint32_t epoch, const float* payload_in,
                            float* payload_sum_out, int dense, pp_stream_t stream) {
  PP_REQUIRE(vox_local && vox_mc && grad_mc && exp_avg && exp_avg_sq, PP_ERR_INVALID_ARG,
             "voxel_adam_exchange_mc: null pointer");
  PP_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, PP_ERR_INVALID_ARG,
             "voxel_adam_exchange_mc: bad world/rank %d/%d", rank, world);
  PP_REQUIRE(epoch == 0 || sync_ptrs, PP_ERR_INVALID_ARG, "voxel_adam_exchange_mc: fused barrier needs sync_ptrs");
  const int64_t shard = pp_voxel_shard_size(n_voxels, world);
  const int64_t begin = shard * rank < n_voxels ? shard * rank : n_voxels;
  const int64_t end = begin + shard < n_voxels ? begin + shard : n_voxels;
  SyncPtrs sp;
  for (int r = 0; r < kMaxWorld; ++r) sp.blk[r] = (sync_ptrs && r < world) ? (SyncBlock*)sync_ptrs[r] : nullptr;
  // an empty slice still takes part in the fused barrier (one block, no records)
  if (end <= begin && epoch == 0) return PP_OK;
  const int64_t n_rec = end > begin ? end - begin : 0;
  const int blocks = n_rec > 0 ? exchange_blocks(n_rec) : 1;
  const char* de = getenv("PP_XCHG_DEPTH");
  const int depth = de ? atoi(de) : 3;
  const float ssd = adam_step_size_x(lr_density, beta1, beta2, step), ssk = adam_step_size_x(lr_k0, beta1, beta2, step);
#define PP_LAUNCH_MC(D)                                                                                          \
  voxel_adam_exchange_mc_kernel<D><<<blocks, kXchgThreads, 0, as_stream(stream)>>>(                              \
      (const float4*)vox_local, (float4*)vox_mc, (const float4*)grad_mc, (float4*)exp_avg, (float4*)exp_avg_sq, \
      begin, end, ssd, ssk, beta1, beta2, dense ? eps * sqrtf(1.f - powf(beta2, (float)step)) : eps, sp, world,     \
      rank, epoch, payload_in, payload_sum_out, dense)
  if (depth <= 2) PP_LAUNCH_MC(2);
  else if (depth == 3) PP_LAUNCH_MC(3);
  else if (depth == 4) PP_LAUNCH_MC(4);
  else PP_LAUNCH_MC(6);
#undef PP_LAUNCH_MC
  PP_CHECK_LAUNCH("voxel_adam_exchange_mc");
  return PP_OK;
}

extern "C" int pp_voxel_adam_exchange_mc(const void* vox_local, void* vox_mc, const void* grad_mc,
                                         void* exp_avg, void* exp_avg_sq, int64_t n_voxels,
                                         int32_t world, int32_t rank, int32_t step, float beta1,
                                         float beta2, float lr_density, float lr_k0, float eps,
                                         void* const* sync_ptrs, uint32_t epoch, const float* payload_in,
                                         float* payload_sum_out, pp_stream_t stream) {
  return exchange_mc_impl(vox_local, vox_mc, grad_mc, exp_avg, exp_avg_sq, n_voxels, world, rank, step, beta1, beta2,
                          lr_density, lr_k0, eps, sync_ptrs, epoch, payload_in, payload_sum_out, 0, stream);
}

extern "C" int pp_records_adam_exchange_mc(const void* rec_local, void* rec_mc, const void* grad_mc, void* exp_avg,
                                           void* exp_avg_sq, int64_t n_records, int32_t world, int32_t rank,
                                           int32_t step, float beta1, float beta2, float lr, float eps, int dense,
                                           void* const* sync_ptrs, uint32_t epoch, const float* payload_in,
                                           float* payload_sum_out, pp_stream_t stream) {
  return exchange_mc_impl(rec_local, rec_mc, grad_mc, exp_avg, exp_avg_sq, n_records, world, rank, step, beta1, beta2,
                          lr, lr, eps, sync_ptrs, epoch, payload_in, payload_sum_out, dense ? 1 : 0, stream);
}

extern "C" int pp_allreduce_mc(void* buf_mc, int64_t n_records, int32_t world, int32_t rank,
                               void* const* sync_ptrs, pp_stream_t stream) {
  PP_REQUIRE(buf_mc && n_records >= 0, PP_ERR_INVALID_ARG, "allreduce_mc: bad argument");
  PP_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, PP_ERR_INVALID_ARG,
             "allreduce_mc: bad world/rank %d/%d", rank, world);
  const int64_t shard = pp_voxel_shard_size(n_records, world);
  const int64_t begin = shard * rank < n_records ? shard * rank : n_records;
  const int64_t end = begin + shard < n_records ? begin + shard : n_records;
  if (end <= begin) return PP_OK;
  SyncPtrs sp;
  for (int r = 0; r < kMaxWorld; ++r) sp.blk[r] = (sync_ptrs && r < world) ? (SyncBlock*)sync_ptrs[r] : nullptr;
  allreduce_mc_kernel<<<exchange_blocks(end - begin), kXchgThreads, 0, as_stream(stream)>>>((float4*)buf_mc, begin,
                                                                                             end, sp, rank);
  PP_CHECK_LAUNCH("allreduce_mc");
  return PP_OK;
}

extern "C" int pp_voxel_adam_exchange_hybrid(void* const* vox_ptrs, void* const* grad_ptrs, void* vox_mc,
                                             const void* grad_mc, void* exp_avg, void* exp_avg_sq, int64_t n_voxels,
                                             int32_t world, int32_t rank, int32_t step, float beta1, float beta2,
                                             float lr_density, float lr_k0, float eps, void* const* sync_ptrs,
                                             uint32_t epoch, const float* payload_in, float* payload_sum_out,
                                             float mc_fraction, pp_stream_t stream) {
  PP_REQUIRE(vox_ptrs && grad_ptrs && vox_mc && grad_mc && exp_avg && exp_avg_sq, PP_ERR_INVALID_ARG,
             "voxel_adam_exchange_hybrid: null pointer");
  PP_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, PP_ERR_INVALID_ARG,
             "voxel_adam_exchange_hybrid: bad world/rank %d/%d", rank, world);
  PP_REQUIRE(epoch == 0 || sync_ptrs, PP_ERR_INVALID_ARG, "voxel_adam_exchange_hybrid: fused barrier needs sync_ptrs");
  PP_REQUIRE(mc_fraction >= 0.f && mc_fraction <= 1.f, PP_ERR_INVALID_ARG,
             "voxel_adam_exchange_hybrid: mc_fraction %f outside [0, 1]", (double)mc_fraction);
  const int64_t shard = pp_voxel_shard_size(n_voxels, world);
  const int64_t begin = shard * rank < n_voxels ? shard * rank : n_voxels;
  const int64_t end = begin + shard < n_voxels ? begin + shard : n_voxels;
  if (end <= begin && epoch == 0) return PP_OK;
  const int64_t n_rec = end > begin ? end - begin : 0;
  // the split is the same on every rank relative to its slice and falls on a 256-record (4 KB) boundary
  int64_t n_mc = (int64_t)((double)n_rec * (double)mc_fraction) / 256 * 256;
  if (mc_fraction >= 1.f) n_mc = n_rec;
  const int64_t split = begin + n_mc;
  PeerPtrs pp;
  SyncPtrs sp;
  for (int r = 0; r < kMaxWorld; ++r) {
    pp.vox[r] = r < world ? (float4*)vox_ptrs[r] : nullptr;
    pp.grad[r] = r < world ? (const float4*)grad_ptrs[r] : nullptr;
    PP_REQUIRE(r >= world || (pp.vox[r] && pp.grad[r]), PP_ERR_INVALID_ARG,
               "voxel_adam_exchange_hybrid: null peer pointer for rank %d", r);
    sp.blk[r] = (sync_ptrs && r < world) ? (SyncBlock*)sync_ptrs[r] : nullptr;
  }
  const int total = n_rec > 0 ? exchange_blocks(n_rec) : 1;
  int mc_blocks = n_rec > 0 ? (int)((double)total * (double)n_mc / (double)n_rec + 0.5) : 1;
  if (n_mc > 0 && mc_blocks < 1) mc_blocks = 1;
  if (n_mc == 0) mc_blocks = 0;
  int p2p_blocks = total - mc_blocks;
  if (n_rec - n_mc > 0 && p2p_blocks < 1) p2p_blocks = 1;
  if (n_rec - n_mc == 0) p2p_blocks = 0;
  const int blocks = mc_blocks + p2p_blocks > 0 ? mc_blocks + p2p_blocks : 1;
  const float ssd = adam_step_size_x(lr_density, beta1, beta2, step), ssk = adam_step_size_x(lr_k0, beta1, beta2, step);
#define PP_LAUNCH_H(W)                                                                                              \
  voxel_adam_exchange_hybrid_kernel<W><<<blocks, kXchgThreads, 0, as_stream(stream)>>>(                             \
      pp, (float4*)vox_mc, (const float4*)grad_mc, (float4*)exp_avg, (float4*)exp_avg_sq, begin, split, end,        \
      n_mc > 0 ? (mc_blocks > 0 ? mc_blocks : 1) : 0, rank, ssd, ssk, beta1, beta2, eps, sp, epoch, payload_in,     \
      payload_sum_out)
  switch (world) {
    case 2: PP_LAUNCH_H(2); break;
    case 4: PP_LAUNCH_H(4); break;
    case 8: PP_LAUNCH_H(8); break;
    default:
      set_error("voxel_adam_exchange_hybrid: world size %d not supported (2, 4, 8)", world);
      return PP_ERR_UNSUPPORTED;
  }
#undef PP_LAUNCH_H
  PP_CHECK_LAUNCH("voxel_adam_exchange_hybrid");
  return PP_OK;
}

extern "C" int64_t pp_peer_sync_bytes(void) { return 4096; }

extern "C" int pp_peer_sync(void* const* sync_ptrs, int32_t world, int32_t rank, uint32_t epoch,
                            const float* payload_in, float* payload_sum_out, pp_stream_t stream) {
  PP_REQUIRE(sync_ptrs, PP_ERR_INVALID_ARG, "peer_sync: null pointer table");
  PP_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, PP_ERR_INVALID_ARG,
             "peer_sync: bad world/rank %d/%d", rank, world);
  static_assert(sizeof(SyncBlock) <= 4096, "sync block size");
  SyncPtrs sp;
  for (int r = 0; r < kMaxWorld; ++r) {
    sp.blk[r] = r < world ? (SyncBlock*)sync_ptrs[r] : nullptr;
    PP_REQUIRE(r >= world || sp.blk[r], PP_ERR_INVALID_ARG, "peer_sync: null sync block for rank %d", r);
  }
  peer_sync_kernel<<<1, 32, 0, as_stream(stream)>>>(sp, world, rank, epoch, payload_in, payload_sum_out);
  PP_CHECK_LAUNCH("peer_sync");
  return PP_OK;
}

// 1 if a barrier wait of this rank ever ran out of its budget (sticky; synchronises the device), else 0
extern "C" int pp_peer_sync_timed_out(const void* sync_block) {
  if (!sync_block) return 0;
  unsigned int v = 0;
  if (cudaMemcpy(&v, &static_cast<const SyncBlock*>(sync_block)->timed_out, sizeof(v), cudaMemcpyDeviceToHost) !=
      cudaSuccess)
    return -1;
  return v != 0u;
}

extern "C" int64_t pp_voxel_shard_size(int64_t n_voxels, int32_t world) {
  if (world <= 0 || n_voxels <= 0) return 0;
  const int64_t s = (n_voxels + world - 1) / world;
  return (s + 255) / 256 * 256;   // whole 4 KB pages of records per slice
}

extern "C" int pp_voxel_adam_exchange(void* const* vox_ptrs, void* const* grad_ptrs, void* exp_avg,
                                      void* exp_avg_sq, int64_t n_voxels, int32_t world, int32_t rank,
                                      int32_t step, float beta1, float beta2, float lr_density,
                                      float lr_k0, float eps, const void* sync_block, pp_stream_t stream) {
  PP_REQUIRE(vox_ptrs && grad_ptrs && exp_avg && exp_avg_sq, PP_ERR_INVALID_ARG,
             "voxel_adam_exchange: null pointer");
  PP_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, PP_ERR_INVALID_ARG,
             "voxel_adam_exchange: bad world/rank %d/%d", rank, world);
  PP_REQUIRE(n_voxels >= 0, PP_ERR_INVALID_ARG, "voxel_adam_exchange: n_voxels < 0");
  const int64_t shard = pp_voxel_shard_size(n_voxels, world);
  const int64_t begin = shard * rank < n_voxels ? shard * rank : n_voxels;
  const int64_t end = begin + shard < n_voxels ? begin + shard : n_voxels;
  if (end <= begin) return PP_OK;
  PeerPtrs pp;
  for (int r = 0; r < kMaxWorld; ++r) {
    pp.vox[r] = r < world ? (float4*)vox_ptrs[r] : nullptr;
    pp.grad[r] = r < world ? (const float4*)grad_ptrs[r] : nullptr;
    PP_REQUIRE(r >= world || (pp.vox[r] && pp.grad[r]), PP_ERR_INVALID_ARG,
               "voxel_adam_exchange: null peer pointer for rank %d", r);
  }
  const float ssd = adam_step_size_x(lr_density, beta1, beta2, step);
  const float ssk = adam_step_size_x(lr_k0, beta1, beta2, step);
  const int blocks = exchange_blocks(end - begin);
  cudaStream_t s = as_stream(stream);
  const unsigned int* timed_out = sync_block ? &static_cast<const SyncBlock*>(sync_block)->timed_out : nullptr;
  const char* dbg_env = getenv("PP_XCHG_FLAGS");   // timing experiments only: 1 = no peer stores, 2 = no peer loads
  const int dbg = dbg_env ? atoi(dbg_env) : 0;
#define PP_LAUNCH_X(W)                                                                         \
  voxel_adam_exchange_kernel<W><<<blocks, kXchgThreads, 0, s>>>(pp, (float4*)exp_avg, (float4*)exp_avg_sq, \
                                                       begin, end, rank, ssd, ssk, beta1, beta2, eps, dbg, timed_out)
  switch (world) {
    case 1: PP_LAUNCH_X(1); break;
    case 2: PP_LAUNCH_X(2); break;
    case 4: PP_LAUNCH_X(4); break;
    case 8: PP_LAUNCH_X(8); break;
    default:
      set_error("voxel_adam_exchange: world size %d not supported (1, 2, 4, 8)", world);
      return PP_ERR_UNSUPPORTED;
  }
#undef PP_LAUNCH_X
  PP_CHECK_LAUNCH("voxel_adam_exchange");
  return PP_OK;
}
