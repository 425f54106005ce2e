// Dependent-chain latency of the instructions on alpha2weight's recurrence (one warp, one block):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_latency tools/probe/fp64_latency.cu && /tmp/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe(float t0, const float* a, long long* out, float* sink) {
  const int n = 4096;
  float T = t0 + threadIdx.x * 1e-9f;
  double q = 1.0 - (double)a[threadIdx.x & 7];
  long long c0, c1;
  // (1) the reference chain: F2F.F64.F32 -> DMUL -> F2F.F32.F64
  c0 = clock64();
#pragma unroll 8
  for (int i = 0; i < n; ++i) T = (float)((double)T * q);
  c1 = clock64();
  if (threadIdx.x == 0) out[0] = (c1 - c0) / n;
  sink[threadIdx.x] = T;
  // (2) DMUL only
  double d = (double)T + 1.0;
  c0 = clock64();
#pragma unroll 8
  for (int i = 0; i < n; ++i) d = d * q;
  c1 = clock64();
  if (threadIdx.x == 0) out[1] = (c1 - c0) / n;
  sink[32 + threadIdx.x] = (float)d;
  // (3) conversions only: float -> double -> float (+ a cheap dependency so it is not folded)
  float f = T + 0.5f;
  c0 = clock64();
#pragma unroll 8
  for (int i = 0; i < n; ++i) f = (float)((double)f) * 0.999f;
  c1 = clock64();
  if (threadIdx.x == 0) out[2] = (c1 - c0) / n;
  sink[64 + threadIdx.x] = f;
  // (4) FMUL only
  c0 = clock64();
#pragma unroll 8
  for (int i = 0; i < n; ++i) f = f * 0.999f;
  c1 = clock64();
  if (threadIdx.x == 0) out[3] = (c1 - c0) / n;
  sink[96 + threadIdx.x] = f;
  // (5) chain with the float -> double widening done with integer ops (exact for normal numbers)
  T = t0 + threadIdx.x * 1e-9f;
  c0 = clock64();
#pragma unroll 8
  for (int i = 0; i < n; ++i) {
    const unsigned int b = __float_as_uint(T);
    const unsigned long long w = ((unsigned long long)b << 29) + (896ull << 52);
    T = (float)(__longlong_as_double((long long)w) * q);
  }
  c1 = clock64();
  if (threadIdx.x == 0) out[4] = (c1 - c0) / n;
  sink[128 + threadIdx.x] = T;
  // (6) DFMA-free: DADD chain
  c0 = clock64();
#pragma unroll 8
  for (int i = 0; i < n; ++i) d = d + q;
  c1 = clock64();
  if (threadIdx.x == 0) out[5] = (c1 - c0) / n;
  sink[160 + threadIdx.x] = (float)d;
}

int main() {
  float h[8] = {0.001f, 0.002f, 0.0015f, 0.003f, 0.001f, 0.002f, 0.0015f, 0.003f};
  float *a, *sink;
  long long* out;
  cudaMalloc(&a, sizeof(h));
  cudaMalloc(&sink, 1024 * sizeof(float));
  cudaMalloc(&out, 8 * sizeof(long long));
  cudaMemcpy(a, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int threads : {32, 128, 512}) {
    probe<<<1, threads>>>(0.9f, a, out, sink);
    long long r[8];
    cudaMemcpy(r, out, sizeof(r), cudaMemcpyDeviceToHost);
    printf("threads %4d: chain(F2F,DMUL,F2F) %lld  DMUL %lld  F2F+F2F+FMUL %lld  FMUL %lld  chain(int widen) %lld  DADD %lld cycles/step\n",
           threads, r[0], r[1], r[2], r[3], r[4], r[5]);
  }
  return 0;
}
