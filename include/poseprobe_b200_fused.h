/*
 * poseprobe_b200 -- fused B200-native entry points for the DVGO render / train step.
 *
 * The reference runs this path as a chain of operators with every per-sample tensor materialised in
 * HBM: sample_pts_on_rays -> DenseGrid(density) -> raw2alpha -> alpha2weight -> DenseGrid(k0) ->
 * sigmoid -> segment_coo sums -> loss -> autograd of all of it -> (masked) Adam
 * (lib/dvgo_ori.py:289-379 for the dense pure-PyTorch form, lib/voxurf_coarse.py:661-695 and
 * :1094-1222 for the CUDA-sampler form, the .cu files under lib/cuda for the kernels).  The entry points below
 * compute the same quantities with ONE kernel per direction: a warp marches a ray, gathers 16-byte
 * voxel records, composites with a warp scan, and (backward) scatters with 128-bit reductions.  No
 * per-sample tensor ever reaches HBM.
 *
 * Voxel record layout ("vox"): float4 per voxel, [X,Y,Z] row-major (Z contiguous) =
 *   { density, k0_r, k0_g, k0_b }.  pp_voxels_pack / pp_voxels_unpack convert from / to the
 * reference's parameter layout (density [1,1,X,Y,Z], k0 [1,3,X,Y,Z]; lib/dvgo_ori.py:50,65) so
 * checkpoints stay interchangeable.  Gradients and Adam moments use the same record layout.
 */
#ifndef POSEPROBE_B200_FUSED_H_
#define POSEPROBE_B200_FUSED_H_

#include "poseprobe_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pp_dvgo_scene {
  int32_t sx, sy, sz;       /* grid extents X, Y, Z (world_size) */
  float xyz_min[3];
  float xyz_max[3];
  float near, far;          /* ray clamp passed to the sampler (render_kwargs) */
  float stepdist;           /* stepsize * voxel_size (voxurf_coarse.py:676) */
  float act_shift;          /* density bias, log(1/(1-alpha_init) - 1) (dvgo_ori.py:43) */
  float interval;           /* stepsize * voxel_size_ratio (dvgo_ori.py:297) */
  float bg;                 /* background colour (render_kwargs['bg']) */
  float fast_color_thres;   /* colour is evaluated for samples with weight > thres (dvgo_ori.py:328) */
} pp_dvgo_scene;

typedef struct pp_dvgo_loss {
  float w_rgb;              /* weight of the photometric MSE (mean over rays x 3) */
  float w_entropy_last;     /* weight of the background-entropy term on alphainv_last (0 = off) */
  float w_rgbper;           /* weight of the per-point colour term, weights detached (0 = off) */
  float mse_scale;          /* factor on the reported photometric MSE (loss_out[1]); 1/world when ranks sum */
} pp_dvgo_loss;

/* density [V], k0 [3,V] (channels first) -> vox [V] float4 records, and back.  Either plane pointer of
 * unpack may be null. */
PP_API int pp_voxels_pack(const float* density, const float* k0, int64_t n_voxels, void* vox,
                          pp_stream_t stream);
PP_API int pp_voxels_unpack(const void* vox, int64_t n_voxels, float* density, float* k0,
                            pp_stream_t stream);

/* Forward render (no gradient state kept).  Outputs, each optional:
 *   rgb_marched [N,3] = sum_i w_i * sigmoid(k0_i) + alphainv_last * bg      (dvgo_ori.py:363)
 *   alphainv_last [N]  final transmittance                                   (alpha2weight's 3rd output)
 *   depth [N]          sum_i w_i * (i * stepdist)                            (voxurf_coarse.py:1198)
 *   n_marched [N] i32  samples composited before termination (in-box ones only) */
PP_API int pp_dvgo_render_forward(const pp_dvgo_scene* scene, const void* vox, const float* rays_o,
                                  const float* rays_d, int64_t n_rays, float* rgb_marched,
                                  float* alphainv_last, float* depth, int32_t* n_marched,
                                  pp_stream_t stream);

/* Scratch the backward / train entry points need: one float4 per step for every ray in flight (the
 * per-sample colour and d sigma/d alpha computed by the forward half, re-read by the reverse half; it
 * is written and read by the same warp within microseconds, so it lives in L2).  Depends only on the
 * scene (steps per ray) and the device (resident warps). */
PP_API int64_t pp_dvgo_workspace_bytes(const pp_dvgo_scene* scene);

/* Backward of the render for arbitrary losses on (rgb_marched, alphainv_last): recomputes the march,
 * then walks the ray in reverse and accumulates d loss / d vox into grad_vox (float4 records, +=).
 * grad_alphainv_last may be null.  grad_rays (optional, [N,6] = d loss/d rays_o | d loss/d rays_d)
 * receives the per-ray pose gradient through the sample positions. */
PP_API int pp_dvgo_render_backward(const pp_dvgo_scene* scene, const void* vox, const float* rays_o,
                                   const float* rays_d, int64_t n_rays, const float* grad_rgb_marched,
                                   const float* grad_alphainv_last, void* grad_vox, float* grad_rays,
                                   void* workspace, int64_t workspace_bytes, pp_stream_t stream);

/* Fused forward + loss + backward for the DVGO training losses against target [N,3]:
 *   loss = w_rgb * mean((rgb_marched - target)^2) [+ entropy_last + rgbper terms]
 * Accumulates d loss / d vox into grad_vox (+=), writes rgb_marched [N,3] (optional) and ADDS the
 * batch loss into loss_out[0] and the photometric MSE into loss_out[1] (device floats the caller
 * zeroes).  grad_rays as above. */
PP_API int pp_dvgo_train_step(const pp_dvgo_scene* scene, const pp_dvgo_loss* loss, const void* vox,
                              const float* rays_o, const float* rays_d, const float* target,
                              int64_t n_rays, void* grad_vox, float* rgb_marched, float* loss_out,
                              float* grad_rays, void* workspace, int64_t workspace_bytes,
                              pp_stream_t stream);

/* Masked Adam on voxel records, per element exactly masked_adam_upd (adam_upd_kernel.cu:25-40, step size
 * :72) with separate learning rates for the density channel and the three k0 channels; gradients are
 * reset to zero in the same pass (the zero_grad the reference does separately).  In place.
 * flags: bit 0 (1) = zero the gradients; bit 1 (2) = the reference's LIVE optimiser instead of the dormant masked
 * kernel: dense Python Adam of lib/utils.py:176-198 - every element updated each step (zero gradients too) and
 * eps added after the bias correction (denom = sqrt(v)/sqrt(1-beta2^t) + eps). */
PP_API int pp_voxel_adam(void* vox, void* grad_vox, void* exp_avg, void* exp_avg_sq, int64_t n_voxels,
                         int32_t step, float beta1, float beta2, float lr_density, float lr_k0,
                         float eps, int flags, pp_stream_t stream);

/* Same, preceded by the total-variation gradient (total_variation_add_grad, old kernel weights
 * k,j,i = wz,wy,wz; total_variation_kernel.cu:13-35) evaluated on the OLD parameters: reads vox_in,
 * writes vox_out (ping-pong; vox_in is left untouched so neighbours always see pre-step values).
 * tv_density / tv_k0 are the caller's wx=wy=wz for the density and k0 channels (0 = off);
 * dense_mode as in the reference (0: only where the ray gradient is non-zero). */
PP_API int pp_voxel_tv_adam(const pp_dvgo_scene* scene, const void* vox_in, void* vox_out,
                            void* grad_vox, void* exp_avg, void* exp_avg_sq, int32_t step, float beta1,
                            float beta2, float lr_density, float lr_k0, float eps, float tv_density,
                            float tv_k0, int dense_mode, pp_stream_t stream);

/* ---- the whole step as ONE CUDA-graph launch ---------------------------------------------------------
 * A trainer that reads the loss of every step (lib/recon_scene.py:607 does: `mse2psnr(...).item()`) issues six stream
 * operations per 0.27 ms step; a host that is slow to issue them (an interpreter busy with the rest of the trainer)
 * leaves the GPU waiting.  Captured once, the step is ONE launch.  Measured on an otherwise idle host (B200, config B,
 * profiles/r2_results.md section 8): 0.301 ms per step against 0.300 ms for the eager sequence - there the ~30 us
 * between steps is the PCIe copy plus the host round trip, which the look-ahead form (HostPipeline in dvgo.py:
 * next batch copied on a copy stream under the step in flight) hides: 0.275 ms.  pp_dvgo_step_graph_create captures
 *     rays_host (pinned, [3,N,3] = origins | directions | target colours)  ->  rays_dev      (H2D copy)
 *     loss_dev[0..1] = 0, pp_dvgo_train_step, pp_voxel_adam(adam_flags)                      (the step)
 *     loss_dev -> loss_host (pinned, 2 floats: loss, photometric MSE)                        (D2H copy)
 * once; pp_dvgo_step_graph_launch replays it as one launch on `stream` after patching the two things that change
 * per step: the pinned buffer the rays come from (null = same as last time) and Adam's step sizes for optimiser
 * step `step` (1-based; same host arithmetic as pp_voxel_adam, so replay and eager calls give the same update).
 * The caller synchronises the stream (or an event) before reading loss_host.  All device pointers must stay valid
 * and fixed for the life of the graph; TV and the pose gradient are not part of it (use the eager calls). */
PP_API int pp_dvgo_step_graph_create(const pp_dvgo_scene* scene, const pp_dvgo_loss* loss, void* vox,
                                     void* grad_vox, void* exp_avg, void* exp_avg_sq, const float* rays_host,
                                     float* rays_dev, int64_t n_rays, float* loss_dev, float* loss_host,
                                     void* workspace, int64_t workspace_bytes, float beta1, float beta2,
                                     float lr_density, float lr_k0, float eps, int adam_flags, void** graph_out);
PP_API int pp_dvgo_step_graph_launch(void* graph, const float* rays_host, int32_t step, pp_stream_t stream);
PP_API int pp_dvgo_step_graph_destroy(void* graph);

/* Forward render of rays that live in HOST memory, results into host memory, in ONE call: rays_host is pinned
 * [2,N,3] (rays_o | rays_d), rgb_host pinned [N,3], depth_host / last_host (alphainv_last) pinned [N] or null.  The
 * rays are marched `chunk` at a time on `stream`; chunk c+1's host->device copy and chunk c-1's device->host copy run
 * on the library's two copy streams under chunk c's kernel (double-buffered device staging, event-ordered).  Returns
 * once everything has landed in the host buffers.  This is lib/nvs_fun.py:83-91's chunked full-image loop
 * (rays.split(4096) -> model.inference(...) -> torch.cat -> .cpu()) with the copies and the host round trips taken off
 * the critical path.
 * One call at a time per device (the staging blocks belong to the device, not to the call). */
PP_API int pp_dvgo_render_host(const pp_dvgo_scene* scene, const void* vox, const float* rays_host, int64_t n_rays,
                               int64_t chunk, float* rgb_host, float* depth_host, float* last_host,
                               pp_stream_t stream);

/* ---- multi-GPU: ray-sharded replicas, voxel-sharded optimiser --------------------------------------
 * The reference is single-GPU; rays are independent, so each rank marches its own batch
 * into its own gradient records and the only exchange per step is the voxel gradient / parameter
 * traffic.  Instead of all-reduce + a full Adam pass on every rank, ONE kernel per rank owns a 1/world
 * slice of the voxels: it reads that slice of every rank's gradient over NVLink (peer loads = the
 * reduce-scatter), applies masked Adam once (moments exist only for the slice), and writes the updated
 * records into every rank's parameter buffer (peer stores = the all-gather).
 *
 * Buffers that peers touch must come from pp_shared_alloc (cudaMalloc + IPC handle); handles are
 * exchanged by the host framework (torch.distributed) and opened with pp_shared_open. */
#define PP_IPC_HANDLE_BYTES 64
PP_API int pp_shared_alloc(int64_t bytes, void** ptr_out, void* handle_out);
PP_API int pp_shared_open(const void* handle, void** ptr_out);
PP_API int pp_shared_close(void* ptr);
PP_API int pp_shared_free(void* ptr);

/* Cross-rank stream barrier over peer memory (one warp; no collective library on the critical path).
 * sync_ptrs: HOST array of `world` device pointers to each rank's sync block (pp_peer_sync_bytes() bytes from
 * pp_shared_alloc, zero-initialised once).  `epoch` must increase by one per call on every rank.  If
 * payload_in (2 device floats) is given, it is delivered to every rank and the rank-ordered sum of all
 * ranks' payloads is written to payload_sum_out - used to form the global loss for free. */
PP_API int64_t pp_peer_sync_bytes(void);
PP_API int pp_peer_sync(void* const* sync_ptrs, int32_t world, int32_t rank, uint32_t epoch,
                        const float* payload_in, float* payload_sum_out, pp_stream_t stream);

/* vox_ptrs / grad_ptrs: HOST arrays of `world` device pointers (entry `rank` is the local buffer).
 * exp_avg / exp_avg_sq: local moments of the slice only, pp_voxel_shard_size(n_voxels, world) records.
 * Slice of rank r: [r*S, min((r+1)*S, n_voxels)), S = pp_voxel_shard_size.  Gradients are summed over
 * ranks (scale the per-rank loss by 1/world for a global mean).  The caller brackets the launch with
 * cross-rank stream barriers (all gradients complete before; all slices written after) and zeroes its own
 * gradient records afterwards. */
PP_API int64_t pp_voxel_shard_size(int64_t n_voxels, int32_t world);
PP_API int pp_voxel_adam_exchange(void* const* vox_ptrs, void* const* grad_ptrs, void* exp_avg,
                                  void* exp_avg_sq, int64_t n_voxels, int32_t world, int32_t rank,
                                  int32_t step, float beta1, float beta2, float lr_density,
                                  float lr_k0, float eps, const void* sync_block, pp_stream_t stream);
/* A barrier wait that runs out of its budget (~60 s: a peer died or stalled) raises a STICKY flag in this rank's sync
 * block instead of hanging the device; from then on the exchange kernels leave parameters and moments untouched
 * (sync_block = this rank's block, nullable = no check; the NVLS variant reads it from sync_ptrs) and the summed
 * payload (the global loss) is NaN.  pp_peer_sync_timed_out reads the flag (synchronises): 1 = timed out. */
PP_API int pp_peer_sync_timed_out(const void* sync_block);

/* NVSwitch (NVLS) variant of pp_voxel_adam_exchange: vox_mc / grad_mc are MULTICAST addresses bound to every
 * rank's parameter / gradient records (e.g. torch symmetric memory's multicast_ptr).  The gradient slice is
 * reduced in the switch (multimem.ld_reduce.add.v4.f32) and the updated records are replicated by the switch
 * (multimem.st.v4.f32): per GPU and link direction (w+1)/w of the record array instead of 2(w-1)/w.
 * epoch != 0 fuses the "all gradients complete" barrier of pp_peer_sync into the kernel's prologue (sync_ptrs,
 * payload_in, payload_sum_out as for pp_peer_sync); epoch == 0: the caller synchronised the ranks itself. */
PP_API int pp_voxel_adam_exchange_mc(const void* vox_local, void* vox_mc, const void* grad_mc,
                                     void* exp_avg, void* exp_avg_sq, int64_t n_voxels, int32_t world,
                                     int32_t rank, int32_t step, float beta1, float beta2,
                                     float lr_density, float lr_k0, float eps, void* const* sync_ptrs,
                                     uint32_t epoch, const float* payload_in, float* payload_sum_out,
                                     pp_stream_t stream);

/* The same exchange for any array of float4 records with ONE learning rate (the fine stage's 12-channel feature
 * records = 3 records per voxel): reduce-scatter in the switch + Adam on this rank's slice (moments for the slice
 * only) + multicast store.  dense != 0: the reference's live dense Python Adam (lib/utils.py:176-198, see
 * pp_voxel_adam flags bit 1) instead of masked_adam_upd. */
PP_API int pp_records_adam_exchange_mc(const void* rec_local, void* rec_mc, const void* grad_mc, void* exp_avg,
                                       void* exp_avg_sq, int64_t n_records, int32_t world, int32_t rank,
                                       int32_t step, float beta1, float beta2, float lr, float eps, int dense,
                                       void* const* sync_ptrs, uint32_t epoch, const float* payload_in,
                                       float* payload_sum_out, pp_stream_t stream);
/* In-place sum all-reduce of n_records float4 records at a multicast address (multimem.ld_reduce + multimem.st on
 * this rank's slice).  The caller brackets it with the cross-rank barriers (all contributions written before; all
 * slices written after).  sync_ptrs (nullable): skip when this rank's barrier timed out. */
PP_API int pp_allreduce_mc(void* buf_mc, int64_t n_records, int32_t world, int32_t rank, void* const* sync_ptrs,
                           pp_stream_t stream);

/* Hybrid of the two above in ONE launch: the first mc_fraction of the slice (rounded down to 256 records) goes
 * through the switch (vox_mc / grad_mc), the rest through peer loads and stores (vox_ptrs / grad_ptrs, which must
 * address the same buffers the multicast addresses are bound to).  The multimem path is held by the switch's
 * reduction rate and the peer path by link bytes, so together they finish sooner than either alone (DESIGN.md
 * section 7).  Rank order of the fp32 sum differs between the two halves (switch order vs 0..w-1) but is the same on
 * every rank for a given voxel, so replicas stay bit-identical.  world in {2, 4, 8}. */
PP_API int pp_voxel_adam_exchange_hybrid(void* const* vox_ptrs, void* const* grad_ptrs, void* vox_mc,
                                         const void* grad_mc, void* exp_avg, void* exp_avg_sq, int64_t n_voxels,
                                         int32_t world, int32_t rank, int32_t step, float beta1, float beta2,
                                         float lr_density, float lr_k0, float eps, void* const* sync_ptrs,
                                         uint32_t epoch, const float* payload_in, float* payload_sum_out,
                                         float mc_fraction, pp_stream_t stream);

/* ---- Voxurf fine stage (SDF grid + 12-channel feature grid + rgbnet), fused -------------------------------
 * Reference: Voxurf.forward, non-deform branch (lib/voxurf_coarse.py:922-1091).  A train step is
 *   pp_voxurf_march_features -> pp_rgbnet_forward_packed (compositing epilogue) -> pp_voxurf_ray_loss ->
 *   pp_rgbnet_backward_fused -> pp_voxurf_backward -> pp_voxurf_sdf_adam + pp_voxel_adam (feature records) ->
 *   pp_voxurf_build_records.
 * SDF records: float4 per voxel {s', d s'/dx, d s'/dy, d s'/dz} with s' = a (sigmoid(b s) - 0.5) (the per-step
 * grid remap :946-949; a = 0 means s' = s) and the central-difference gradient grid of neus_sdf_gradient
 * ('interpolate', :463-467).  Feature records: 3 float4 per voxel = the 12 channels of k0.
 * Feature row of the MLP (bf16, 64 columns): [0,12) k0 | [12,15) normal | 15: 0 | [16,19) rays_xyz |
 * [19,34) sin(xyz 2^k) (coordinate-major, stride 5) | [34,49) cos | [49,52) viewdirs | [52,55) sin | [55,58) cos |
 * [58,64) 0, BARF weights applied; the caller permutes the first layer's columns accordingly (reference order,
 * :1005-1030: k0 | xyz | sin | cos | viewdirs | sin | cos | normal). */
typedef struct pp_voxurf_cfg {
  float inv_s;              /* 1 / s_val (voxurf_coarse.py:495) */
  float sdf_scale_a;        /* softplus10(sdf_alpha); 0 disables the mapping function */
  float sdf_scale_b;        /* softplus10(sdf_beta) */
  float voxel_size;         /* divisor of the finite-difference gradient grid */
  int32_t pos_pe;           /* posbase_pe, <= 5 */
  int32_t view_pe;          /* viewbase_pe, <= 1 in the fused path */
  float pe_w_pos[8];        /* BARF coarse-to-fine weight per frequency (:721-732), 1 = off */
  float pe_w_view[8];
} pp_voxurf_cfg;

PP_API int pp_voxurf_build_records(const pp_dvgo_scene* scene, const pp_voxurf_cfg* cfg, const float* sdf,
                                   void* sdf_records, pp_stream_t stream);
PP_API int pp_voxurf_k0_pack(const float* k0, int64_t n_voxels, void* k0_records, pp_stream_t stream);
PP_API int pp_voxurf_k0_unpack(const void* k0_records, int64_t n_voxels, float* k0, pp_stream_t stream);

/* Forward march of n_rays rays.  Composited samples (in the box, before early termination) are written to
 * compacted slots [0, counter[0]) in ray-arrival order: feat [capacity,64] bf16, sample_weight, sample_ray
 * (int32), sample_rec (float4: interpolated SDF record).  Per ray: ray_off / ray_cnt (its slot range),
 * ray_end (exclusive step index where the march stopped), cum_w (sum of weights), alphainv_last.
 * counter: 2 x int32 on the device; counter[0] = samples written, counter[1] = 1 if `capacity` was too small
 * (rays that did not fit contribute nothing).  jitter (nullable): per-ray step offset in [0,1) (:712-713).
 * depth [N] (nullable) = sum_i w_i * step_i with step_i = stepdist * (i + jitter) (n_step of :1052-1056; the render
 * depth of :1198-1202), normal_marched [N,3] (nullable) = sum_i w_i * gradient_i / (|gradient_i| + 1e-6) (:1044-1048). */
PP_API int pp_voxurf_march_features(const pp_dvgo_scene* scene, const pp_voxurf_cfg* cfg,
                                    const void* sdf_records, const void* k0_records, const float* rays_o,
                                    const float* rays_d, const float* viewdirs, const float* jitter,
                                    int64_t n_rays, int32_t capacity, void* feat, float* sample_weight,
                                    int32_t* sample_ray, void* sample_rec, int32_t* ray_off, int32_t* ray_cnt,
                                    int32_t* ray_end, float* cum_w, float* alphainv_last, int32_t* counter,
                                    float* depth, float* normal_marched, pp_stream_t stream);

/* ---- implicit deformation, forward (Voxurf.inference :1093-1215 and forward(use_deform=True) :958-985) ---------
 *   (deform, correction) = warp_network(x) * out_range;  x' = x + deform
 *   sdf_final = grid_sample_3d(sdf', x') + correction;  gradient = d sdf_final / d x
 * pp_voxurf_deform_rows: one run of slots per ray for ALL its in-box samples (the reference's compacted ray_pts
 * order) and the warp network's input rows: rows fp16 [4 * sample_capacity, 16]; warp_ray_off [N] (first slot of
 * the ray, -1 if it did not fit); counter: 4 x int32 on the device: [0] samples, [1] overflow flag, [2] rows = 4 *
 * samples (pass &counter[2] to pp_warpnet_forward_packed as its device row count); sample_pts (nullable): float4
 * [sample_capacity] = the sample positions, kept for pp_voxurf_deform_sample_backward.  rows may be null when only
 * the enumeration (warp_ray_off, counter, sample_pts) is wanted.  warp_ray_cnt (nullable) [N]: in-box samples per ray
 * (runs are handed out in completion order, so the reference's ray-major sample order is offset + count per ray).
 * pp_voxurf_march_features_deform: pp_voxurf_march_features with the SDF value / normal of every sample taken from
 * the warped lookup: warp_out = the network's fp32 [rows,4] output.  The template SDF is read with the arithmetic of
 * the reference's custom sampler (weights from the unclamped floor, clamped corner indices, :545-659) and its
 * analytic gradient is chained through the warp Jacobian - the closed form of the reference's autograd.grad. */
PP_API int pp_voxurf_deform_rows(const pp_dvgo_scene* scene, const float* rays_o, const float* rays_d,
                                 const float* jitter, int64_t n_rays, int32_t sample_capacity, void* rows,
                                 int32_t* warp_ray_off, int32_t* counter, void* sample_pts,
                                 int32_t* warp_ray_cnt, pp_stream_t stream);
PP_API int pp_voxurf_march_features_deform(const pp_dvgo_scene* scene, const pp_voxurf_cfg* cfg,
                                           const void* sdf_records, const void* k0_records, const float* rays_o,
                                           const float* rays_d, const float* viewdirs, const float* jitter,
                                           int64_t n_rays, int32_t capacity, void* feat, float* sample_weight,
                                           int32_t* sample_ray, void* sample_rec, int32_t* ray_off,
                                           int32_t* ray_cnt, int32_t* ray_end, float* cum_w, float* alphainv_last,
                                           int32_t* counter, float* depth, float* normal_marched,
                                           const float* warp_out, const int32_t* warp_ray_off, float out_range,
                                           pp_stream_t stream);

/* ---- implicit deformation, backward pieces (training through forward(use_deform=True), :958-985) ----------------
 * The SDF grid is frozen in this mode (voxurf_coarse.py:136-138); gradients flow to the deformation network, the
 * feature grid, the rgbnet (and the rays).  After pp_voxurf_march_features_deform + rgbnet forward / loss / backward:
 * pp_voxurf_backward_deform = pp_voxurf_backward with the per-sample {d loss / d sdf_final, d loss / d gradient}
 * written to grad_val (float4 [sample_capacity], indexed like warp_out / 16; zero-fill it first: samples past early
 * termination keep zero) instead of a scatter into the SDF records; the feature-grid scatter is unchanged.  grad_rays /
 * grad_viewdirs (nullable, need 64 gradient columns and the records): the colour-side pose gradient of the composited
 * samples (features, positional encoding, NeuS cosine); pp_voxurf_deform_ray_grad adds the SDF side.
 * pp_voxurf_deform_sample_backward: per in-box sample, chains grad_val through the warped lookup (cell gradient and
 * off-diagonal Hessian of the custom sampler) and the warp Jacobian, adds the regularisers of deform_implicit_loss
 * (lib/losses.py:6-24; w_* = their weights, each term a mean over all in-box samples; reg_terms (nullable, 4 floats,
 * accumulated): the unweighted term values) and writes grad_warp_out fp32 [4 * samples, 4] = the gradient w.r.t.
 * pp_warpnet_forward_packed's output in the same four-row layout (the grad_out of the network's backward);
 * grad_pts (nullable) [samples,3]: d loss / d x through x' and sdf(x).  Formulas: oracle/deform_analytic.py.
 * The network's own backward is pp_warpnet_backward_packed (poseprobe_b200.h). */
PP_API int pp_voxurf_backward_deform(const pp_dvgo_scene* scene, const pp_voxurf_cfg* cfg, const float* rays_o,
                                     const float* rays_d, const float* viewdirs, const float* jitter,
                                     int64_t n_rays, const void* sample_rec, const float* grad_weight,
                                     const float* grad_feat, int32_t grad_feat_cols, const int32_t* ray_off,
                                     const int32_t* ray_cnt, const int32_t* ray_end, const float* g_ray,
                                     void* grad_k0_records, const float* grad_cum_weights,
                                     const float* grad_alphainv_last, void* grad_val,
                                     const int32_t* warp_ray_off, const void* sdf_records, const void* k0_records,
                                     float* grad_rays, float* grad_viewdirs, pp_stream_t stream);
PP_API int pp_voxurf_deform_sample_backward(const pp_dvgo_scene* scene, const void* sdf_records,
                                            const void* sample_pts, const float* warp_out, const void* grad_val,
                                            const int32_t* n_samples_ptr, int64_t sample_capacity, float out_range,
                                            float w_eikonal, float w_jacobian, float w_correction, float w_deform,
                                            float* grad_warp_out, float* grad_pts, float* reg_terms,
                                            const float* ext_grad, pp_stream_t stream);
/* Forward half of the same per-sample computation, for callers that keep the loss in Python (lib/losses.py:6-24
 * reads gradient, grad_deform, sdf_correct, sdf_deform of ALL in-box samples, voxurf_coarse.py:1085-1090):
 * sample_out fp32 [samples,16] = sdf_final | gradient[3] | grad_deform[3][3] ([i][k] = d x'_k / d x_i) | sdf_correct |
 * sdf_deform | 0.  ext_grad of pp_voxurf_deform_sample_backward (nullable) is the way back: fp32 [samples,14] =
 * d loss / d (gradient[3] | grad_deform[9] | sdf_correct | sdf_deform), added to the built-in terms.
 * grad_pts of pp_voxurf_deform_sample_backward is the complete d loss / d x of the sample through the deformation
 * branch: x' and the warped lookup, the unwarped lookup of sdf_deform, and J^T (d loss / d net(x)). */
PP_API int pp_voxurf_deform_sample_forward(const pp_dvgo_scene* scene, const void* sdf_records, const void* sample_pts,
                                           const float* warp_out, const int32_t* n_samples_ptr,
                                           int64_t sample_capacity, float out_range, float* sample_out,
                                           float* reg_terms, pp_stream_t stream);
/* Non-deform branch, per-sample entries of ALL in-box samples (the reference's `gradient` [M,3], voxurf_coarse.py:
 * 986-989, :1082, on which lib/losses.py:7-10 puts the live Eikonal term): sample_pts = the list
 * pp_voxurf_deform_rows writes (rows may be null there: only the enumeration is wanted).
 * pp_voxurf_sample_records: sample_rec_out float4 [samples] = interpolated {s', d s'/dx, d s'/dy, d s'/dz}.
 * pp_voxurf_sample_records_backward: grad_sdf_records += scatter of grad_sample_rec (float4 [samples], nullable) plus
 * the built-in Eikonal gradient w_eikonal * d mean|‖gradient‖ - 1|; eikonal_term (nullable, accumulated) += the
 * unweighted term. */
PP_API int pp_voxurf_sample_records(const pp_dvgo_scene* scene, const void* sdf_records, const void* sample_pts,
                                    const int32_t* n_samples_ptr, int64_t sample_capacity, void* sample_rec_out,
                                    pp_stream_t stream);
PP_API int pp_voxurf_sample_records_backward(const pp_dvgo_scene* scene, const void* sdf_records,
                                             const void* sample_pts, const int32_t* n_samples_ptr,
                                             int64_t sample_capacity, const void* grad_sample_rec, float w_eikonal,
                                             void* grad_sdf_records, float* eikonal_term, pp_stream_t stream);
/* d loss / d (rays_o | rays_d) [N,6] through the positions of all in-box samples: grad_pts [samples,3] chained through
 * sample_pts_on_rays' geometry (t_min slab, normalised direction), ADDED to grad_rays (which pp_voxurf_backward_deform
 * filled with the colour-side terms of the composited samples). */
PP_API int pp_voxurf_deform_ray_grad(const pp_dvgo_scene* scene, const float* rays_o, const float* rays_d,
                                     const float* jitter, int64_t n_rays, const int32_t* warp_ray_off,
                                     const float* grad_pts, float* grad_rays, pp_stream_t stream);

/* rgb_marched = clamp(rgb_acc + (1 - cum_w) bg, 0, 1) (:1040-1042); photometric MSE (lib/losses.py:26-29):
 * g_ray [N,3] = d (w_rgb * mse) / d rgb_acc, loss_out[0] += w_rgb * mse, loss_out[1] += mse. */
PP_API int pp_voxurf_ray_loss(int64_t n_rays, const float* rgb_acc, const float* cum_w, const float* target,
                              float bg, float w_rgb, float* g_ray, float* rgb_marched, float* loss_out,
                              pp_stream_t stream);

/* The reference's per-ray object losses (lib/losses.py:34-74) on the per-ray outputs, value and gradient:
 *   img_render = sum((rgb_marched m - target m)^2) / (3 sum m)        (weight w_main;   mask null: m = 1)
 *   weight_entropy_last on clamp(alphainv_last, 1e-6, 1-1e-6)          (weight w_entropy, 0.01 in the reference)
 *   mask_render = BCE(clip(cum_w, 1e-3, 1-1e-3), m)                    (weight w_mask)
 * g_acc [N,3], g_cum_w [N], g_last [N] = d (weighted sum) / d (rgb_acc, cum_weights, alphainv_last): the inputs of
 * pp_rgbnet_backward_fused (g_ray) and pp_voxurf_backward(_deform) (grad_cum_weights, grad_alphainv_last).
 * loss_out (nullable, 4 floats, accumulated): weighted sum, img_render, entropy, mask_render.  scratch: one device
 * float (needed with a mask: holds sum m). */
PP_API int pp_voxurf_object_loss(int64_t n_rays, const float* rgb_acc, const float* cum_w, const float* alphainv_last,
                                 const float* target, const float* mask, float bg, float w_main, float w_entropy,
                                 float w_mask, float* g_acc, float* g_cum_w, float* g_last, float* rgb_marched,
                                 float* loss_out, float* scratch, pp_stream_t stream);
/* Voxurf.k0_total_variation (voxurf_coarse.py:443-455 -> total_variation :1298-1313, unmasked) on the feature
 * records: tv_out (nullable, accumulated) += sum of |forward differences| (the loss term is tv_out / (3 * 12 V));
 * grad_k0_records (nullable) += weight * d tv / d k0. */
PP_API int pp_voxurf_k0_tv(const pp_dvgo_scene* scene, const void* k0_records, void* grad_k0_records, float weight,
                           float* tv_out, pp_stream_t stream);

/* Reverse march: alpha2weight / NeuS / normal backward per ray and the scatter into grad_sdf_records
 * ([V] float4, same layout as the SDF records) and grad_k0_records ([V][3] float4); both are accumulated into.
 * grad_feat: the MLP's input gradient rows, grad_feat_cols = 16 (features + normal only) or 64 (all columns).
 * Pose gradient (optional, needs 64 columns): grad_rays [N,6] = d loss / d (rays_o | rays_d) through the sample
 * positions (SDF value and gradient interpolation, feature interpolation, positional encoding), grad_viewdirs [N,3]
 * = d loss / d viewdirs (NeuS cosine, view encoding); the caller chains viewdirs = rays_d / |rays_d| and
 * pp_pose_grad_reduce (the autograd reduction of SURVEY.md a21).  Pass null to skip.
 * grad_cum_weights [N] / grad_alphainv_last [N] (nullable): d loss / d cum_weights and d loss / d alphainv_last of
 * an external (autograd) loss; when grad_cum_weights is null the fused loss' background term -(sum g_ray) * bg is
 * used, when grad_alphainv_last is null it is zero. */
PP_API int pp_voxurf_backward(const pp_dvgo_scene* scene, const pp_voxurf_cfg* cfg, const float* rays_o,
                              const float* rays_d, const float* viewdirs, const float* jitter, int64_t n_rays,
                              const void* sample_rec, const float* grad_weight, const float* grad_feat,
                              int32_t grad_feat_cols, const int32_t* ray_off, const int32_t* ray_cnt,
                              const int32_t* ray_end, const float* g_ray, void* grad_sdf_records,
                              void* grad_k0_records, const void* sdf_records, const void* k0_records,
                              float* grad_rays, float* grad_viewdirs, const float* grad_cum_weights,
                              const float* grad_alphainv_last, pp_stream_t stream);

/* Dense tail of the SDF grid: folds the gradient-grid components back onto the SDF (transposed central
 * difference), applies the mapping function's chain rule and masked Adam (adam_upd_kernel.cu:35-39) in place.
 * exp_avg / exp_avg_sq null: no update (gradient only).  dense != 0: the reference's live dense Python Adam instead
 * (lib/utils.py:176-198, see pp_voxel_adam).  grad_sdf_out (nullable) receives d loss / d sdf;
 * grad_ab (nullable, 2 floats, accumulated) receives d loss / d (a, b).  The caller zeroes grad_sdf_records
 * afterwards and rebuilds the records. */
PP_API int pp_voxurf_sdf_adam(const pp_dvgo_scene* scene, const pp_voxurf_cfg* cfg, float* sdf,
                              const void* grad_sdf_records, float* exp_avg, float* exp_avg_sq, int32_t step,
                              float beta1, float beta2, float lr, float eps, float* grad_sdf_out, float* grad_ab,
                              int dense, pp_stream_t stream);

/* ---- Scene branch: the dense-sample NeRF compositing the reference runs next to the object
 * branch in every training step (lib/recon_scene.py:645-649), csrc/scene.cu.  All tensors fp32, contiguous;
 * n_rays = B * R of the reference's [B, R, ...] tensors, n_samples = N. ---- */

/* NeRF.composite (lib/bg_nerf/source/models/frequency_nerf.py:283-343): intervals (last = 1e10), sigma * delta,
 * exclusive exp-cumsum transmittance, weights and the seven per-ray outputs in one pass.
 * ray [n,3], rgb_samples [n,N,3] (after sigmoid), density_samples [n,N] (after the density activation),
 * depth_samples [n,N] -> rgb [n,3] (+ (1 - opacity) when setbg_opaque), depth [n], opacity [n], weights [n,N];
 * rgb_var, depth_var, all_cumulated [n]: nullable. */
PP_API int pp_nerf_composite_forward(const float* ray, const float* rgb_samples, const float* density_samples,
                                     const float* depth_samples, int64_t n_rays, int32_t n_samples,
                                     int setbg_opaque, float* rgb, float* rgb_var, float* depth, float* depth_var,
                                     float* opacity, float* weights, float* all_cumulated, pp_stream_t stream);
/* autograd's mirror image of the above: d loss / d (rgb_samples [n,N,3], density_samples [n,N], ray [n,3]) from
 * d loss / d (rgb [n,3], depth [n], opacity [n], weights [n,N]) (each nullable = zero).  Outputs nullable.
 * n_samples <= 1024. */
PP_API int pp_nerf_composite_backward(const float* ray, const float* rgb_samples, const float* density_samples,
                                      const float* depth_samples, int64_t n_rays, int32_t n_samples,
                                      int setbg_opaque, const float* grad_rgb, const float* grad_depth,
                                      const float* grad_opacity, const float* grad_weights, float* grad_rgb_samples,
                                      float* grad_density_samples, float* grad_ray, pp_stream_t stream);
/* Graph.sample_depth_from_pdf (lib/bg_nerf/source/models/renderer.py:703-738): inverse-CDF resampling of the coarse
 * weights [n,N]; unif [Nf] = the interval mid points 0.5 (grid[:-1] + grid[1:]) the reference shares between all rays
 * (grid = linspace(0,1,Nf+1) or ONE torch.rand(Nf+1) draw) -> depth_samples_fine [n,Nf].  N <= 4095. */
PP_API int pp_nerf_sample_depth_from_pdf(const float* weights, int64_t n_rays, int32_t n_samples_coarse,
                                         const float* unif, int32_t n_samples_fine, float depth_min,
                                         float depth_max, float* depth_samples_fine, pp_stream_t stream);
/* torch.cat([depth_samples, depth_samples_fine], dim=2).sort(dim=2).values (renderer.py:612-614): per ray the n_a
 * coarse and n_b resampled depths merged in ascending order -> out [n, n_a + n_b].  n_a + n_b <= 1024. */
PP_API int pp_nerf_sort_depth(const float* depth_a, int32_t n_a, const float* depth_b, int32_t n_b, int64_t n_rays,
                              float* out, pp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* POSEPROBE_B200_FUSED_H_ */
