// Camera-ray generation for selected pixels and the fused per-image pose-gradient reduction.
//
// The reference regenerates the rays of EVERY pixel of every training view each step
// (lib/recon_scene.py:596 -> lib/voxurf_coarse.py:1518-1549 -> get_rays :1339-1368, get_rays_of_a_view
// :1402-1407), picks N_rand of them, and back-propagates the ray gradients to the 3x4 camera matrices
// through dozens of tiny autograd kernels.  Here one kernel generates only the selected rays and one
// kernel reduces d loss / d (rays_o, rays_d) [N,6] to d loss / d c2w [n_img,3,4] (block-level shared
// accumulation, one global atomic per block, image and entry).
#include "common.cuh"

namespace pp {

// dirs = ((i - cx)/fx, -(j - cy)/fy, -1)  [inverse_y: (.., (j - cy)/fy, 1)]   (voxurf_coarse.py:1359-1362)
__device__ __forceinline__ void pixel_dir(const float* K, float px, float py, int inverse_y, float dir[3]) {
  dir[0] = (px - K[2]) / K[0];
  const float v = (py - K[5]) / K[4];
  dir[1] = inverse_y ? v : -v;
  dir[2] = inverse_y ? 1.f : -1.f;
}

__global__ void rays_from_pixels_kernel(const float* __restrict__ Ks, const float* __restrict__ c2w,
                                        const int32_t* __restrict__ img_id, const float* __restrict__ px,
                                        const float* __restrict__ py, int inverse_y, int64_t n,
                                        float* __restrict__ rays_o, float* __restrict__ rays_d) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int im = img_id[i];
  const float* K = Ks + 9 * im;
  const float* P = c2w + 12 * im;
  float dir[3];
  pixel_dir(K, px[i], py[i], inverse_y, dir);
  float u[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)   // sum(dirs * c2w[r,:3]) : three rounded products, summed left to right (:1364)
    u[r] = add_rn(add_rn(mul_rn(dir[0], P[4 * r]), mul_rn(dir[1], P[4 * r + 1])), mul_rn(dir[2], P[4 * r + 2]));
  // rays_d / rays_d.norm() (:1404); torch's norm = sqrt(sum of squares)
  const float nrm = sqrtf(add_rn(add_rn(mul_rn(u[0], u[0]), mul_rn(u[1], u[1])), mul_rn(u[2], u[2])));
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    rays_d[3 * i + r] = u[r] / nrm;
    rays_o[3 * i + r] = P[4 * r + 3];     // c2w[:3,3] (:1366)
  }
}

constexpr int kMaxImages = 64;

__global__ void __launch_bounds__(256)
pose_grad_reduce_kernel(const float* __restrict__ Ks, const float* __restrict__ c2w,
                        const int32_t* __restrict__ img_id, const float* __restrict__ px,
                        const float* __restrict__ py, int inverse_y, int64_t n, int n_img,
                        const float* __restrict__ grad_rays, float* __restrict__ grad_c2w) {
  __shared__ float s_acc[kMaxImages * 12];
  for (int t = threadIdx.x; t < n_img * 12; t += blockDim.x) s_acc[t] = 0.f;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int im = img_id[i];
    const float* K = Ks + 9 * im;
    const float* P = c2w + 12 * im;
    float dir[3];
    pixel_dir(K, px[i], py[i], inverse_y, dir);
    float u[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) u[r] = dir[0] * P[4 * r] + dir[1] * P[4 * r + 1] + dir[2] * P[4 * r + 2];
    const float inv = rsqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    const float d[3] = {u[0] * inv, u[1] * inv, u[2] * inv};
    const float* go = grad_rays + 6 * i;
    const float* gd = go + 3;
    // d = u/|u|  =>  g_u = (g_d - d (d.g_d)) / |u|
    const float dot = d[0] * gd[0] + d[1] * gd[1] + d[2] * gd[2];
    float* acc = s_acc + 12 * im;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float gu = (gd[r] - d[r] * dot) * inv;
      atomicAdd(acc + 4 * r, gu * dir[0]);
      atomicAdd(acc + 4 * r + 1, gu * dir[1]);
      atomicAdd(acc + 4 * r + 2, gu * dir[2]);
      atomicAdd(acc + 4 * r + 3, go[r]);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n_img * 12; t += blockDim.x)
    if (s_acc[t] != 0.f) atomicAdd(grad_c2w + t, s_acc[t]);
}

}  // namespace pp

using namespace pp;

extern "C" int pp_rays_from_pixels(const float* Ks, const float* c2w, const int32_t* img_id,
                                   const float* px, const float* py, int inverse_y, int64_t n_rays,
                                   float* rays_o, float* rays_d, pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0, PP_ERR_INVALID_ARG, "rays_from_pixels: n_rays < 0");
  if (n_rays == 0) return PP_OK;
  PP_REQUIRE(Ks && c2w && img_id && px && py && rays_o && rays_d, PP_ERR_INVALID_ARG,
             "rays_from_pixels: null pointer");
  rays_from_pixels_kernel<<<ceil_div(n_rays, 256), 256, 0, as_stream(stream)>>>(
      Ks, c2w, img_id, px, py, inverse_y, n_rays, rays_o, rays_d);
  PP_CHECK_LAUNCH("rays_from_pixels");
  return PP_OK;
}

extern "C" int pp_pose_grad_reduce(const float* Ks, const float* c2w, const int32_t* img_id,
                                   const float* px, const float* py, int inverse_y, int64_t n_rays,
                                   int32_t n_img, const float* grad_rays, float* grad_c2w,
                                   pp_stream_t stream) {
  PP_REQUIRE(n_rays >= 0 && n_img >= 0, PP_ERR_INVALID_ARG, "pose_grad_reduce: negative extent");
  PP_REQUIRE(n_img <= kMaxImages, PP_ERR_UNSUPPORTED, "pose_grad_reduce: at most %d images", kMaxImages);
  if (n_img == 0) return PP_OK;
  PP_REQUIRE(grad_c2w, PP_ERR_INVALID_ARG, "pose_grad_reduce: grad_c2w is null");
  cudaStream_t s = as_stream(stream);
  PP_CHECK_CUDA(cudaMemsetAsync(grad_c2w, 0, sizeof(float) * 12 * n_img, s));
  if (n_rays == 0) return PP_OK;
  PP_REQUIRE(Ks && c2w && img_id && px && py && grad_rays, PP_ERR_INVALID_ARG,
             "pose_grad_reduce: null pointer");
  const int blocks = (int)min64((n_rays + 255) / 256, (int64_t)sm_count());
  pose_grad_reduce_kernel<<<blocks, 256, 0, s>>>(Ks, c2w, img_id, px, py, inverse_y, n_rays, n_img,
                                                 grad_rays, grad_c2w);
  PP_CHECK_LAUNCH("pose_grad_reduce");
  return PP_OK;
}
