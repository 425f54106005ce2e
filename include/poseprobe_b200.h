/*
 * poseprobe_b200 -- C ABI of the B200-native (sm_100a) voxel-grid volume-rendering operators.
 *
 * This is the drop-in boundary for PoseProbe's `lib/cuda` extensions.  The reference binds its
 * kernels through pybind11 + torch::Tensor (render_utils.cpp:170-184, total_variation.cpp:29-32,
 * adam_upd.cpp:79-86, ub360_utils.cpp:21-23).  Here every entry point takes plain device pointers,
 * extents and a CUDA stream, so any host (ctypes, pybind, cgo, ...) can bind it; the Python shim in
 * poseprobe_b200/ re-creates the reference's four modules (same names, same 19 functions, same
 * argument order and return arity) on top of it.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - tensors are contiguous, row-major, in the reference's layouts ([N,3] rays, [M] samples,
 *     [1,C,X,Y,Z] grids, int64 indices, 1-byte bools);
 *   - `dtype` selects the storage type of the floating tensors (PP_F32 / PP_F64), mirroring the
 *     reference's AT_DISPATCH_FLOATING_TYPES; scalars stay `float` exactly as in the reference
 *     signatures (their fp32 rounding is part of bit-exactness);
 *   - `stream` is a cudaStream_t (the shim passes torch's current stream; the reference used the
 *     legacy default stream);
 *   - every function returns PP_OK (0) or a negative pp_status; nothing is allocated inside, no
 *     host synchronisation happens inside unless documented;
 *   - outputs are fully written by the callee (callers need not pre-fill), including the
 *     semantically relevant fill values of the reference (weight=0, T=1 past early termination;
 *     alphainv_last=1, i_start=i_end=0 for rays without samples; mask=false out of range);
 *   - there is NO CPU fallback: a host pointer is an error (PP_ERR_NOT_DEVICE where detectable).
 */
#ifndef POSEPROBE_B200_H_
#define POSEPROBE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_ABI_VERSION 1

#if defined(__GNUC__)
#define PP_API __attribute__((visibility("default")))
#else
#define PP_API
#endif

typedef void* pp_stream_t; /* cudaStream_t */

typedef enum pp_dtype { PP_F32 = 0, PP_F64 = 1 } pp_dtype;

typedef enum pp_status {
  PP_OK = 0,
  PP_ERR_INVALID_ARG = -1,  /* null pointer, negative extent, unknown enum */
  PP_ERR_CUDA = -2,         /* a CUDA runtime call / launch failed; see pp_last_error() */
  PP_ERR_TOO_LARGE = -3,    /* extent exceeds the 32-bit sample index space of the reference */
  PP_ERR_NOT_DEVICE = -4,   /* pointer is not device memory (no CPU fallback) */
  PP_ERR_UNSUPPORTED = -5   /* combination not implemented (e.g. channel count) */
} pp_status;

/* --- library -------------------------------------------------------------------------------- */
PP_API int pp_abi_version(void);
PP_API const char* pp_last_error(void);          /* thread-local, human-readable text of the last failure */
PP_API const char* pp_build_info(void);          /* "sm_100a nvcc x.y ..." */
PP_API int pp_check_device_ptr(const void* p);   /* PP_OK if p is device/managed memory */
PP_API uint64_t pp_launch_count(void);           /* kernels launched by this library so far (process-wide) */

/* =============================================================================================
 * render_utils_cuda  (reference: lib/cuda/render_utils.cpp + render_utils_kernel.cu)
 * ============================================================================================= */

/* replaces infer_t_minmax (render_utils.cpp:50; kernel :12-35).  out: t_min[n], t_max[n] */
PP_API int pp_infer_t_minmax(pp_dtype dtype, const void* rays_o, const void* rays_d, const void* xyz_min,
                      const void* xyz_max, float near, float far, int64_t n_rays, void* t_min,
                      void* t_max, pp_stream_t stream);

/* replaces infer_n_samples (render_utils.cpp:60; kernel :38-55).  out: n_samples[n] int64 */
PP_API int pp_infer_n_samples(pp_dtype dtype, const void* rays_d, const void* t_min, const void* t_max,
                       float stepdist, int64_t n_rays, int64_t* n_samples, pp_stream_t stream);

/* replaces infer_ray_start_dir (render_utils.cpp:67; kernel :58-79).  out: 2 x [n,3] */
PP_API int pp_infer_ray_start_dir(pp_dtype dtype, const void* rays_o, const void* rays_d,
                           const void* t_min, int64_t n_rays, void* rays_start, void* rays_dir,
                           pp_stream_t stream);

/* sample_pts_on_rays (render_utils.cpp:74; host :196-242) is split in two because the number of
 * samples M is data dependent and the caller owns all memory.
 *
 * Phase 1 (one kernel): per-ray t_min, t_max, N_steps, rays_start, rays_dir, the inclusive prefix
 * sum of N_steps (`n_steps_cumsum`, int64[n], scratch the caller keeps until phase 2) and the total
 * M written to `total_dev` (int64[1], device) AND, if `total_host_pinned` is non-null, to that
 * pinned host word.  `scan_scratch` is int64[pp_sample_scan_scratch_elems(n_rays)] device
 * scratch.  No host sync inside; the caller synchronises the stream and reads the total.
 */
PP_API int64_t pp_sample_scan_scratch_elems(int64_t n_rays);
PP_API int pp_sample_pts_prepare(pp_dtype dtype, const void* rays_o, const void* rays_d,
                          const void* xyz_min, const void* xyz_max, float near, float far,
                          float stepdist, int64_t n_rays, void* t_min, void* t_max,
                          int64_t* n_steps, int64_t* n_steps_cumsum, void* rays_start,
                          void* rays_dir, int64_t* scan_scratch, int64_t* total_dev,
                          int64_t* total_host_pinned, pp_stream_t stream);
/* Phase 2 (one kernel): rays_pts[M,3], mask_outbbox[M] (bool), ray_id[M], step_id[M] (int64).
 * Any of the four outputs may be null to skip it. */
PP_API int pp_sample_pts_fill(pp_dtype dtype, const void* rays_start, const void* rays_dir,
                       const void* xyz_min, const void* xyz_max, const int64_t* n_steps,
                       const int64_t* n_steps_cumsum, float stepdist, int64_t n_rays,
                       int64_t total, void* rays_pts, uint8_t* mask_outbbox, int64_t* ray_id,
                       int64_t* step_id, pp_stream_t stream);

/* replaces sample_ndc_pts_on_rays (render_utils.cpp:87; kernel :245-270) */
PP_API int pp_sample_ndc_pts_on_rays(pp_dtype dtype, const void* rays_o, const void* rays_d,
                              const void* xyz_min, const void* xyz_max, int32_t n_samples,
                              int64_t n_rays, void* rays_pts, uint8_t* mask_outbbox,
                              pp_stream_t stream);

/* replaces sample_bg_pts_on_rays (render_utils.cpp:100; kernel :301-340) */
PP_API int pp_sample_bg_pts_on_rays(pp_dtype dtype, const void* rays_o, const void* rays_d,
                             const void* t_max, float bg_preserve, int32_t n_samples,
                             int64_t n_rays, void* rays_pts, pp_stream_t stream);

/* replaces maskcache_lookup (render_utils.cpp:109; kernel :374-392).  world: bool[sz_i,sz_j,sz_k] */
PP_API int pp_maskcache_lookup(pp_dtype dtype, const uint8_t* world, const void* xyz,
                        const void* xyz2ijk_scale, const void* xyz2ijk_shift, int32_t sz_i,
                        int32_t sz_j, int32_t sz_k, int64_t n_pts, uint8_t* out,
                        pp_stream_t stream);

/* replaces raw2alpha / raw2alpha_nonuni (render_utils.cpp:120,:125; kernels :431-458).
 * `interval_pts` non-null selects the per-point (nonuni) variant and `interval` is ignored. */
PP_API int pp_raw2alpha(pp_dtype dtype, const void* density, float shift, float interval,
                 const void* interval_pts, int64_t n_pts, void* exp_d, void* alpha,
                 pp_stream_t stream);

/* replaces raw2alpha_backward / raw2alpha_nonuni_backward (render_utils.cpp:131,:136; :507-530) */
PP_API int pp_raw2alpha_backward(pp_dtype dtype, const void* exp_d, const void* grad_back,
                          float interval, const void* interval_pts, int64_t n_pts, void* grad,
                          pp_stream_t stream);

/* replaces alpha2weight (render_utils.cpp:142; kernels :577-617, host :619-651).
 * ray_id must be sorted (all samples of a ray contiguous), as the reference requires. */
PP_API int pp_alpha2weight(pp_dtype dtype, const void* alpha, const int64_t* ray_id, int64_t n_pts,
                    int64_t n_rays, void* weight, void* T, void* alphainv_last, int64_t* i_start,
                    int64_t* i_end, pp_stream_t stream);

/* replaces alpha2weight_backward (render_utils.cpp:151; kernel :654-677) */
PP_API int pp_alpha2weight_backward(pp_dtype dtype, const void* alpha, const void* weight, const void* T,
                             const void* alphainv_last, const int64_t* i_start,
                             const int64_t* i_end, int64_t n_rays, int64_t n_pts,
                             const void* grad_weights, const void* grad_last, void* grad,
                             pp_stream_t stream);

/* =============================================================================================
 * total_variation_cuda  (reference: lib/cuda/total_variation.cpp + total_variation_kernel.cu)
 * param/grad/mask: [1,C,X,Y,Z] viewed as [n_planes = C, sz_i, sz_j, sz_k]; in place on grad.
 * `mask` null  -> total_variation_add_grad      (:16; kernel :13-35; k,j,i weights = wz,wy,wz)
 * `mask` given -> total_variation_add_grad_new  (:22; kernel :38-66; k,j,i weights = wx,wy,wz)
 * wx,wy,wz are the caller's values; the /6 of the reference host code (:76-78) happens inside.
 * ============================================================================================= */
PP_API int pp_total_variation_add_grad(pp_dtype dtype, const void* param, void* grad, const void* mask,
                                float wx, float wy, float wz, int dense_mode, int64_t n_planes,
                                int64_t sz_i, int64_t sz_j, int64_t sz_k, pp_stream_t stream);

/* =============================================================================================
 * adam_upd_cuda  (reference: lib/cuda/adam_upd.cpp + adam_upd_kernel.cu), all in place
 * mode 0: adam_upd (:36)   mode 1: masked_adam_upd (:50, skip grad==0)
 * mode 2: adam_upd_with_perlr (:64, needs perlr)
 * mode 3: the optimiser the reference actually RUNS - the Python Adam of lib/utils.py:176-198 (the kernels above
 *         are never called by it): dense like mode 0, but denom = sqrt(v) / sqrt(1 - beta2^t) + eps
 * ============================================================================================= */
PP_API int pp_adam_upd(pp_dtype dtype, int mode, void* param, const void* grad, void* exp_avg,
                void* exp_avg_sq, const void* perlr, int64_t numel, int32_t step, float beta1,
                float beta2, float lr, float eps, pp_stream_t stream);

/* =============================================================================================
 * ub360_utils_cuda  (reference: lib/cuda/ub360_utils.cpp:15 + ub360_utils_kernel.cu:12-47)
 * ============================================================================================= */
PP_API int pp_cumdist_thres(pp_dtype dtype, const void* dist, float thres, int64_t n_rays, int64_t n_pts,
                     uint8_t* mask, pp_stream_t stream);

/* =============================================================================================
 * Operators the reference runs as Python/ATen/torch_scatter composites on the same path.
 * ============================================================================================= */

/* Per-ray segmented sums: out[r,:] (+)= sum_{i: ray_id[i]==r} src[i,:]  -- the reference's
 * torch_scatter.segment_coo(reduce='sum') call sites (voxurf_coarse.py:1034-1056, :1179-1200).
 * `weight` (optional, [M]) multiplies src rows first (fuses the `weights.unsqueeze(-1)*rgb`
 * products the reference materialises).  `out` is [n_rays, n_ch], overwritten. fp32 only. */
PP_API int pp_segment_sum(const float* src, const float* weight, const int64_t* ray_id, int64_t n_pts,
                   int32_t n_ch, int64_t n_rays, float* out, pp_stream_t stream);

/* Dense compositing of the reference's pure-PyTorch DirectVoxGO: get_ray_marching_ray / cumprod_exclusive
 * (lib/dvgo_ori.py:478-485).  alpha [n_rays, S] (dense samples, 0 outside the box) ->
 * alphainv_cum [n_rays, S+1] = [1, cumprod(clamp_min(1 - alpha, 1e-10))], weights [n_rays, S] = alpha *
 * alphainv_cum[:, :-1].  No early termination.  Backward: d loss / d alpha from d loss / d weights and
 * d loss / d alphainv_cum (each nullable = zero), torch.cumprod's rule (sum_{i>j} G_i C_i) / p_j. */
PP_API int pp_dense_composite_forward(const float* alpha, int64_t n_rays, int32_t n_samples, float* weights,
                                      float* alphainv_cum, pp_stream_t stream);
PP_API int pp_dense_composite_backward(const float* alpha, const float* alphainv_cum, int64_t n_rays,
                                       int32_t n_samples, const float* grad_weights,
                                       const float* grad_alphainv_cum, float* grad_alpha, pp_stream_t stream);

/* Inclusive prefix sum of src within each ray segment (the north star's `segment_cumsum`, which
 * the reference names but does not ship; defined in SURVEY.md section 8 a12). fp32 only.
 * `scratch` is pp_segment_cumsum_scratch_bytes(n_pts) bytes of device memory. */
PP_API int64_t pp_segment_cumsum_scratch_bytes(int64_t n_pts);
PP_API int pp_segment_cumsum(const float* src, const int64_t* ray_id, int64_t n_pts, float* out,
                      void* scratch, pp_stream_t stream);

/* Dense-grid trilinear interpolation == F.grid_sample(grid[1,C,X,Y,Z], mode='bilinear',
 * align_corners=True) on ((xyz-xyz_min)/(xyz_max-xyz_min)).flip(-1)*2-1 as DenseGrid.forward does
 * (lib/grid.py:47-58).  padding: 0 = zeros (DenseGrid), 1 = border (Voxurf.grid_sampler,
 * voxurf_coarse.py:540), 2 = grid_sample_3d (see pp_trilinear_backward2).  out: [n, C] (row per point).  grad_xyz_out (optional, [n,C,3]) receives
 * d out / d xyz analytically (what the reference obtains with autograd.grad, :972-984). fp32. */
PP_API int pp_trilinear_forward(const float* grid, int32_t n_ch, int32_t sz_x, int32_t sz_y, int32_t sz_z,
                         const float* xyz, int64_t n_pts, const float* xyz_min_host,
                         const float* xyz_max_host, int padding, float* out, float* dout_dxyz,
                         pp_stream_t stream);

/* Backward of the above: grad_grid[1,C,X,Y,Z] += scatter(grad_out[n,C]) (atomic accumulate, the
 * caller zeroes grad_grid), and optionally grad_xyz[n,3] = sum_c grad_out[n,c] * d out_c / d xyz.
 * Either output may be null. */
PP_API int pp_trilinear_backward(const float* grid, int32_t n_ch, int32_t sz_x, int32_t sz_y,
                          int32_t sz_z, const float* xyz, int64_t n_pts,
                          const float* xyz_min_host, const float* xyz_max_host, int padding,
                          const float* grad_out, float* grad_grid, float* grad_xyz,
                          pp_stream_t stream);

/* padding 2 (all three entry points) = the reference's own twice-differentiable sampler grid_sample_3d
 * (lib/voxurf_coarse.py:545-659): trilinear weights from the UNCLAMPED voxel coordinate, the 8 corner indices clamped
 * into the grid, no derivative clipping.
 * pp_trilinear_backward2 = backward of pp_trilinear_backward's grad_xyz (what torch.autograd.grad(...,
 * create_graph=True) + backward reach in the reference, :966-985; Eikonal / deformation losses): given
 * grad_grad_xyz [n,3] = d loss / d grad_xyz and the first backward's grad_out [n,C]:
 *   grad_grad_out [n,C] (overwritten) = d loss / d grad_out,   grad_grid += d loss / d grid (cell-derivative weights),
 *   grad_xyz [n,3] (overwritten) = d loss / d xyz through the cell's off-diagonal Hessian.  Outputs are nullable. */
PP_API int pp_trilinear_backward2(const float* grid, int32_t n_ch, int32_t sz_x, int32_t sz_y, int32_t sz_z,
                                  const float* xyz, int64_t n_pts, const float* xyz_min_host,
                                  const float* xyz_max_host, int padding, const float* grad_out,
                                  const float* grad_grad_xyz, float* grad_grad_out, float* grad_grid,
                                  float* grad_xyz, pp_stream_t stream);

/* NeuS SDF -> alpha (voxurf_coarse.py:483-519, mid-point variant) and its backward.
 * alpha = clip((sig(prev*inv_s) - sig(next*inv_s) + 1e-5) / (sig(prev*inv_s) + 1e-5), 0, 1),
 * prev/next = sdf -/+ iter_cos*dist*0.5, iter_cos = -relu(-(viewdirs[ray_id] . gradient)). */
PP_API int pp_neus_alpha_forward(const float* sdf, const float* gradient, const float* viewdirs,
                          const int64_t* ray_id, float dist, float inv_s, int64_t n_pts,
                          float* alpha, pp_stream_t stream);
PP_API int pp_neus_alpha_backward(const float* sdf, const float* gradient, const float* viewdirs,
                           const int64_t* ray_id, float dist, float inv_s, int64_t n_pts,
                           const float* grad_alpha, float* grad_sdf, float* grad_gradient,
                           pp_stream_t stream);

/* Python-side total variation loss the reference actually uses (`total_variation`, lib/voxurf_coarse.py:1298-1313,
 * via k0_total_variation :443-456 / density_total_variation :407-411): sum over the three axes of |forward
 * difference|, optionally restricted to pairs whose two voxels are both set in `mask` ([X,Y,Z] bool, shared by all
 * planes = nonempty_mask.repeat).  Forward writes the un-normalised sum to *sum_out (double, device); the caller
 * divides by 3 * count (count = numel, or planes * mask.sum()).  Backward: grad += scale * d(sum)/dv. */
PP_API int pp_total_variation_loss(const float* v, const uint8_t* mask, int64_t n_planes, int64_t sz_i,
                                   int64_t sz_j, int64_t sz_k, double* sum_out, pp_stream_t stream);
PP_API int pp_total_variation_loss_backward(const float* v, const uint8_t* mask, int64_t n_planes, int64_t sz_i,
                                            int64_t sz_j, int64_t sz_k, float scale, float* grad,
                                            pp_stream_t stream);

/* rgbnet / colour MLP forward on tcgen05 tensor cores (lib/voxurf_coarse.py:208-216, :1032-1033;
 * lib/dvgo_ori.py:82-89): out[n,3] = [sigmoid](W_L relu(... relu(W_1 x + b_1) ...) + b_L).
 * x [n, d0] fp32 row-major (the concatenated feature rows the reference builds);
 * weights / biases: HOST arrays of `depth` DEVICE pointers, nn.Linear layout ([out,in] row-major fp32);
 * hidden width must be 128, the last layer has 3 outputs, 2 <= depth <= 5, d0 <= 128.
 * workspace: n * pp_rgbnet_padded_width(d0) * 2 bytes (the bf16 copy of x).
 * Operands are rounded to bf16, accumulation is fp32 in TMEM; activations never leave the SM. */
PP_API int pp_rgbnet_forward(const float* x, int64_t n_pts, int32_t d0, int32_t depth, int32_t width,
                             const float* const* weights, const float* const* biases, int apply_sigmoid,
                             float* out, void* workspace, int64_t workspace_bytes, pp_stream_t stream);
/* Same on packed bf16 [n, k0p] rows; `out` may be null when the compositing epilogue is used:
 * rgb_marched[ray_id[i], :] += sample_weight[i] * out[i, :]  (segment_coo of w * rgb, voxurf_coarse.py:1034).
 * n_ptr (nullable): DEVICE int32 row count produced by an earlier kernel; n_pts is then the capacity. */
PP_API int pp_rgbnet_forward_packed(const void* x_packed, int64_t n_pts, int32_t d0, int32_t depth,
                                    int32_t width, const float* const* weights, const float* const* biases,
                                    int apply_sigmoid, float* out, const float* sample_weight,
                                    const int32_t* ray_id, float* rgb_marched, const int32_t* n_ptr,
                                    pp_stream_t stream);

/* Deformation ("warp") network WITH its input Jacobian, forward only: DeformedImplicitField (lib/deformation/
 * deform_net.py:12-29; SingleBVPNet 3 -> 128 x4 -> 4, ReLU, lib/deformation/modules.py:43-124) and the
 * torch.autograd.grad(..., [ray_pts]) calls around it (lib/voxurf_coarse.py:966-985, :1116-1135).
 * A ReLU network is piecewise linear: the tangent of each hidden layer is the layer's linear map gated by the value
 * row's activation pattern, so one pass over FOUR rows per sample yields net(x) and d net / d x.
 * x_rows: fp16 [n_rows,16], n_rows = 4 * samples; per sample the value row {x_hi[3], x_lo[3], 0 ...} (x = x_hi + x_lo:
 * the point keeps 22 mantissa bits) and the tangent rows e_0, e_1, e_2 (pp_voxurf_deform_rows writes this format).
 * weights[0]: [128,6] = [W_0 | W_0]; weights[1..depth-2]: [128,128]; weights[depth-1]: [4,128]; biases as nn.Linear.
 * out: fp32 [n_rows,4]: row 4s = net(x_s) (NOT yet multiplied by the output range), row 4s+1+k = d net / d x_k.
 * n_rows_ptr (nullable): DEVICE int32 row count (n_rows is then the capacity). */
PP_API int pp_warpnet_forward_packed(const void* x_rows, int64_t n_rows, int32_t depth, int32_t width,
                                     const float* const* weights, const float* const* biases, float* out,
                                     const int32_t* n_rows_ptr, pp_stream_t stream);
/* x fp32 [n,3] -> the four fp16 [16] rows per point that pp_warpnet_forward_packed consumes ([4n,16]). */
PP_API int pp_warpnet_pack_points(const float* x, int64_t n_pts, void* rows, pp_stream_t stream);

/* Training the deformation network: parameter gradients from gradients w.r.t. net(x) AND its input Jacobian - what
 * the reference's double backward (torch.autograd.grad(..., create_graph=True), lib/voxurf_coarse.py:966-985, then
 * loss.backward(), lib/recon_scene.py:649) delivers to warp_network.*.grad.  One tcgen05 pass per 128-row tile:
 * forward recompute, layer-by-layer backward of all four rows with the value row's gates, weight / bias gradients
 * accumulated in TMEM (bias gradients from the value rows only).  The loss's gradient w.r.t. x THROUGH the network is
 * J^T (d loss / d value row) and is formed by the caller from the forward's Jacobian; nothing is written per row.
 * x_rows: as pp_warpnet_forward_packed, with fp16 1.0 in column 6 of every value row (pp_voxurf_deform_rows and
 * pp_warpnet_pack_points write it; the forward ignores the column).  grad_out: fp32 [n_rows,4], the gradient w.r.t.
 * pp_warpnet_forward_packed's output in the same four-row layout.  depth must be 5, width 128.
 * grad_weights / grad_biases: HOST arrays of 5 DEVICE pointers, ACCUMULATED into (+=): grad_weights[0] is [128,6]
 * like weights[0] (the gradient of the nn.Linear weight [128,3] is the sum of its two halves), the rest nn.Linear
 * layouts.  workspace: pp_warpnet_backward_workspace_bytes() bytes on the device.
 * pp_warpnet_backward_stalled: 1 if a barrier wait ever ran out of its cycle budget (never expected; synchronises). */
PP_API int64_t pp_warpnet_backward_workspace_bytes(void);
PP_API int pp_warpnet_backward_packed(const void* x_rows, const float* grad_out, int64_t n_rows,
                                      const int32_t* n_rows_ptr, int32_t depth, int32_t width,
                                      const float* const* weights, const float* const* biases,
                                      float* const* grad_weights, float* const* grad_biases, void* workspace,
                                      pp_stream_t stream);
PP_API int pp_warpnet_backward_stalled(void);

/* rgbnet training step on tcgen05: forward recompute + backward of the whole network per 128-sample tile
 * (autograd of lib/voxurf_coarse.py:1032-1033 `torch.sigmoid(self.rgbnet(rgb_feat))`).
 * grad_out [n,3]: d loss / d out (post-sigmoid when apply_sigmoid, else d loss / d logits).
 * grad_x [n,d0] (nullable: skip the input gradient) is overwritten; grad_weights / grad_biases (HOST arrays of
 * `depth` DEVICE pointers, nn.Linear layouts) are ACCUMULATED into (+=), like autograd's .grad;
 * rgb_out [n,3] (nullable) receives the recomputed forward output.
 * workspace: pp_rgbnet_workspace_bytes(n, d0) bytes of device memory (the bf16 copy of x and the padded
 * input gradient); nothing is allocated inside.
 * Supported: hidden width 128, 3 outputs, d0 <= 64 with depth <= 4 or d0 <= 128 with depth <= 3 (the 512
 * TMEM columns hold every layer's weight-gradient accumulator).  bf16 operands, fp32 accumulation. */
PP_API int pp_rgbnet_backward(const float* x, const float* grad_out, int64_t n_pts, int32_t d0,
                              int32_t depth, int32_t width, const float* const* weights,
                              const float* const* biases, int apply_sigmoid, float* grad_x,
                              float* const* grad_weights, float* const* grad_biases, float* rgb_out,
                              void* workspace, int64_t workspace_bytes, pp_stream_t stream);
PP_API int64_t pp_rgbnet_workspace_bytes(int64_t n_pts, int32_t d0);
/* Diagnostic: 1 if a barrier wait inside the backward kernel ever exceeded its cycle budget (never expected; the
 * kernel then finishes without hanging the device and its results are invalid).  Synchronises the device. */
PP_API int pp_rgbnet_backward_stalled(void);
/* Diagnostic (env PP_MLP_PROF=1): cycle counters of CTA 0 of the last backward launch: [0..6] MMA warp (wait for
 * epilogue half 0, half 1, input tile, -, -, -, total), [8..14] epilogue warp 0, [16..22] epilogue warp 8 (wait for
 * accumulator, for dW inside the tile, for dW at the tile end, for cp.async, -, -, total). */
PP_API int pp_rgbnet_backward_profile(long long* out24);

/* The pieces of pp_rgbnet_backward, for callers that keep features in the kernels' own format (the fused
 * fine-stage path): x_packed = bf16 [n, k0p] rows, grad_x_padded = fp32 [n, k0p] rows,
 * k0p = pp_rgbnet_padded_width(d0) (64 or 128), padding columns zero. */
PP_API int32_t pp_rgbnet_padded_width(int32_t d0);
PP_API int pp_rgbnet_pack_input(const float* x, int64_t n_pts, int32_t d0, void* x_packed, pp_stream_t stream);
PP_API int pp_rgbnet_unpack_grad(const float* grad_x_padded, int64_t n_pts, int32_t d0, float* grad_x,
                                 pp_stream_t stream);
PP_API int pp_rgbnet_backward_packed(const void* x_packed, const float* grad_out, int64_t n_pts, int32_t d0,
                                     int32_t depth, int32_t width, const float* const* weights,
                                     const float* const* biases, int apply_sigmoid, float* grad_x_padded,
                                     float* const* grad_weights, float* const* grad_biases, float* rgb_out,
                                     pp_stream_t stream);

/* Fine-stage form: grad_out_i = sample_weight[i] * g_ray[sample_ray[i], :] is formed in the kernel (sigmoid
 * output), grad_weight_out[i] = sum_c g_ray[sample_ray[i],c] * rgb_i,c, and grad_feat_cols = 16 (only the leading
 * input-gradient columns, row pitch 16) or the padded width (all columns) are stored in grad_feat; the row count
 * is read from the device (n_ptr). */
PP_API int pp_rgbnet_backward_fused(const void* x_packed, int64_t capacity, const int32_t* n_ptr, int32_t d0,
                                    int32_t depth, int32_t width, const float* const* weights,
                                    const float* const* biases, const float* sample_weight,
                                    const int32_t* sample_ray, const float* g_ray, float* grad_weight_out,
                                    float* grad_feat, int32_t grad_feat_cols, float* const* grad_weights,
                                    float* const* grad_biases, pp_stream_t stream);

/* Camera rays for selected pixels (get_rays / get_rays_of_a_view, lib/voxurf_coarse.py:1339-1368, :1402-1407,
 * mode='center' is px = i + 0.5 etc. supplied by the caller; flips are applied to px/py by the caller):
 *   dirs = ((px-cx)/fx, -(py-cy)/fy, -1)   [inverse_y: (.., +(py-cy)/fy, +1)]
 *   rays_d = normalize(c2w[:3,:3] . dirs),  rays_o = c2w[:3,3]
 * Ks [n_img,3,3], c2w [n_img,3,4], img_id [N] int32, px/py [N].  Outputs [N,3]. */
PP_API int pp_rays_from_pixels(const float* Ks, const float* c2w, const int32_t* img_id, const float* px,
                               const float* py, int inverse_y, int64_t n_rays, float* rays_o,
                               float* rays_d, pp_stream_t stream);

/* Fused per-image pose-gradient reduction: grad_rays [N,6] (d loss/d rays_o | d loss/d rays_d, as emitted
 * by pp_dvgo_train_step / pp_dvgo_render_backward) -> grad_c2w [n_img,3,4] (overwritten), i.e. the backward
 * of pp_rays_from_pixels summed over the rays of each image (the autograd reduction of SURVEY.md a21). */
PP_API int pp_pose_grad_reduce(const float* Ks, const float* c2w, const int32_t* img_id, const float* px,
                               const float* py, int inverse_y, int64_t n_rays, int32_t n_img,
                               const float* grad_rays, float* grad_c2w, pp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* POSEPROBE_B200_H_ */
