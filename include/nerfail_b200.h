/*
 * nerfail_b200.h — C ABI of libnerfail_b200.so (sm_100a only).
 *
 * The reference (jiang-wenxiang/NeRFail) has no FFI layer: its hot path is eager
 * PyTorch.  Each entry point below replaces the group of ATen ops that one
 * reference function dispatches; the citation after "replaces:" is the
 * reference file:line (relative to the reference root) whose arithmetic the
 * kernel reproduces.  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add to call these from the unmodified Python signatures.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all tensors are contiguous, row-major, fp32 unless stated;
 *   - nothing is allocated behind the caller's back: outputs and workspaces are
 *     caller-owned (the only owned state is the opaque nfb_mlp_t handle);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - every function returns 0 on success or a negative NFB_E_* code; the text
 *     of the last failure on the calling thread is nfb_last_error();
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point returns NFB_E_CUDA.
 */
#ifndef NERFAIL_B200_H
#define NERFAIL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NFB_OK             0
#define NFB_E_ARG         -1   /* bad argument (null pointer, non-positive size, misalignment) */
#define NFB_E_UNSUPPORTED -2   /* shape outside what the sm_100a kernels are built for          */
#define NFB_E_CUDA        -3   /* CUDA runtime error (text in nfb_last_error)                    */

#define NFB_ABI_VERSION    2

#if defined(__GNUC__)
#define NFB_API __attribute__((visibility("default")))
#else
#define NFB_API
#endif

/* Library plumbing, no reference counterpart (the reference is eager PyTorch: errors are Python exceptions). */
NFB_API int         nfb_abi_version(void);
NFB_API const char* nfb_last_error(void);
/* number of kernels launched by this library in this process (bench.py's gpu_launches); no reference counterpart */
NFB_API uint64_t    nfb_launch_count(void);
/* compute capability of the current device as major*10+minor, or NFB_E_CUDA; no reference counterpart */
NFB_API int         nfb_device_cc(void);

/* ------------------------------------------------------------------------- */
/* A. NeRF render path                                                        */
/* ------------------------------------------------------------------------- */

/* Pinhole rays for one camera.
 * replaces: run_nerf_helpers.py:157-166 (get_rays) + run_nerf.py:102-123 (viewdirs, [N,11] ray batch)
 * K_host: 3x3 row-major doubles on the HOST (fx, 0, cx, 0, fy, cy, ...); c2w_host: 3x4 row-major floats on the HOST.
 * rays: [H*W, 11] = o(3) d(3) near far viewdir(3).                                                    */
NFB_API int nfb_get_rays(int H, int W, const double* K_host, const float* c2w_host,
                 float near_, float far_, float* rays, void* stream);

/* The ray batch of render(rays=(rays_o, rays_d)) for use_viewdirs = True, ndc = False (the training step's call).
 * replaces: run_nerf.py:95-123 (viewdirs = rays_d / norm, near / far columns, two concatenations) in one launch.
 * rays_o, rays_d [N,3] -> rays [N,11].                                                                   */
NFB_API int nfb_rays_from_batch(const float* rays_o, const float* rays_d, int64_t N, float near_, float far_, float* rays,
                                void* stream);

/* The training loss with its gradient in one pass.
 * replaces: img2mse(rgb, target) + img2mse(rgb0, target) (run_nerf.py:781-789, run_nerf_helpers.py:9) and their autograd.
 * rgb, rgb0 (or NULL), target: n floats; out5 = (loss, mse(rgb), mse(rgb0), psnr(rgb), psnr(rgb0)) with psnr = -10 log10(mse)
 * (mse2psnr, run_nerf_helpers.py:10); g_rgb / g_rgb0 = 2 (x - target) / n.                                              */
NFB_API int nfb_mse_loss2(const float* rgb, const float* rgb0, const float* target, int64_t n, float* out5, float* g_rgb,
                          float* g_rgb0, void* stream);

/* Coarse depths. replaces: run_nerf.py:357-379 (linspace / lindisp / stratified jitter)
 * rays [R,11]; t_rand [R,S] uniform numbers or NULL (perturb == 0); z_vals out [R,S].               */
NFB_API int nfb_coarse_z(const float* rays, int R, int S, int lindisp, const float* t_rand,
                 float* z_vals, void* stream);
/* The same with the stratified jitter drawn IN the kernel (replaces the torch.rand of run_nerf.py:371 and its [R,S] HBM
 * tensor): element p = r*S + i is word p & 3 of Philox4x32-10 at counter (p >> 2, stream 0, offset) under key `seed`,
 * mapped to [0,1) with 24 bits like torch.rand.  Same seed + offset -> same depths; advance offset per call.          */
/* The generator by itself: out[p] for p < n of stream `stream_id` (0 = coarse jitter, 1 = inverse-CDF u); lets a caller or a
 * test reproduce exactly the numbers nfb_coarse_z_rng / nfb_hierarchical_rng consume in place of the torch.rand calls of
 * run_nerf.py:371 and run_nerf_helpers.py:213.                                                                          */
NFB_API int nfb_philox_uniform(uint64_t seed, uint64_t offset, uint32_t stream_id, int64_t n, float* out, void* stream);
NFB_API int nfb_coarse_z_rng(const float* rays, int R, int S, int lindisp, uint64_t seed, uint64_t offset,
                             float* z_vals, void* stream);

/* Opaque network handle: bf16 weight images pre-swizzled for tcgen05 + fp32 biases and heads. */
typedef struct nfb_mlp nfb_mlp_t;

/* replaces: run_nerf_helpers.py:72-98 (NeRF.__init__).  The fused kernel is built for the
 * reference's one shipped architecture: D=8, W=256, input_ch=63, input_ch_views=27, skips=[4],
 * use_viewdirs=True (configs/lego.txt via run_nerf.py:181-198); anything else -> NFB_E_UNSUPPORTED
 * (the host side then uses the layer-wise nfb_linear_* path).                                       */
NFB_API int nfb_mlp_create(nfb_mlp_t** out, int D, int W, int input_ch, int input_ch_views, int skip);
/* params: flat fp32 device buffer holding the state_dict tensors in this order, each row-major
 * [out,in] then bias: pts_linears.0..7, views_linears.0, feature_linear, alpha_linear, rgb_linear
 * (the nn.Module registration order of run_nerf_helpers.py:82-96).  Call again after optimizer.step. */
NFB_API int nfb_mlp_update(nfb_mlp_t* h, const float* params, int64_t n_params, void* stream);
NFB_API int nfb_mlp_destroy(nfb_mlp_t* h);
NFB_API int64_t nfb_mlp_param_count(const nfb_mlp_t* h);

/* Fused positional-encoding + 8x256 skip MLP, bf16 tcgen05 MMA with fp32 TMEM accumulation.
 * replaces: run_nerf.py:37-51 (run_network) + run_nerf_helpers.py:36-50 (embed) + :100-123 (NeRF.forward)
 *   mode 0: pts  [R*S,3] explicit points,  dirs [R,3] unit view directions
 *   mode 1: pts  = NULL; rays [R,11], z_vals [R,S]  -> pts = o + d*z formed in-kernel (run_nerf.py:381)
 * raw out [R*S,4] = (rgb logits, sigma), fp32.                                                        */
NFB_API int nfb_mlp_fwd(const nfb_mlp_t* h, int mode, const float* pts, const float* dirs,
                const float* rays, const float* z_vals, int R, int S, float* raw, void* stream);

/* 0 = healthy.  NFB_E_CUDA if a pipeline barrier inside a fused kernel of this network timed out since the last call
 * (the kernel bails out instead of hanging the GPU; its outputs are then invalid).  Synchronises the device and CLEARS
 * the flag.  The flag lives in mapped pinned host memory, so it is also visible without a synchronisation:
 *   - nfb_mlp_poll reads it (no sync, no clear): NFB_E_CUDA once a kernel that has already run timed out;
 *   - every launch entry point of the network (nfb_mlp_fwd, nfb_mlp_fwd_train, nfb_mlp_bwd_*, nfb_render_rays_fwd) polls it
 *     first and refuses to run (NFB_E_CUDA) while it is raised, so one time-out cannot silently poison later results.
 * Any number of networks may be alive per device; the 4 constant-memory entries that hold their head weights are shared
 * LRU (a network that lost its entry re-acquires one at its next launch; that costs one host-side wait).
 * Health / validation / profiling entries: no reference counterpart. */
NFB_API int nfb_mlp_status(nfb_mlp_t* h);
NFB_API int nfb_mlp_poll(const nfb_mlp_t* h);
/* Test hook (no reference counterpart): raises the flag from the host exactly as a timed-out barrier wait does. */
NFB_API int nfb_mlp_debug_raise_abort(nfb_mlp_t* h);
/* Validation entry (no reference counterpart; the values are NeRF.forward's hidden activations, run_nerf_helpers.py:108-117):
 * run only the first nsteps (1..10) MMA steps of the fused kernel and dump the fp32
 * post-activation values of the last executed step to dbg [R*S,256] (128 columns for the view layer).     */
NFB_API int nfb_mlp_fwd_debug(const nfb_mlp_t* h, int mode, const float* pts, const float* dirs,
                      const float* rays, const float* z_vals, int R, int S, float* raw, int nsteps, float* dbg,
                      void* stream);

/* Training on tensor cores (bf16 operands, fp32 accumulate).
 * replaces: NeRF.forward under autograd + the data-gradient half of loss.backward() (run_nerf.py:776-791).
 * nfb_mlp_fwd_train = nfb_mlp_fwd (mode 1) that also leaves, per 128-row tile, every layer's bf16 activation as a tile
 * image act_img [tiles][40][16 KB] and the relu masks as bit words mask [tiles][9][8][128] (layout: csrc/mlp_train.inl).
 * nfb_mlp_bwd_data runs the data-gradient chain from g_raw [M,4] (d loss / d raw) and leaves every layer's dY as a tile
 * image dy_img [tiles][39][16 KB] (chunk 38 is reserved and left unwritten).
 * nfb_mlp_bwd_weights computes every weight / bias gradient of the network from the two images in ONE grouped launch
 * (14 tensor-core products dW = dY^T X, see nfb_wgrad_bf16; alpha_linear / rgb_linear as fp32 side sums of g_raw against
 * operands those products load anyway) and ACCUMULATES them into grad [nfb_mlp_param_count] in state_dict
 * order: zero grad once per step.  tiles = nfb_mlp_train_tiles(M).  A barrier time-out is reported by nfb_mlp_status.
 * nfb_mlp_bwd = both of them as two CONCURRENT kernels on disjoint SMs (the loss.backward() of run_nerf.py:791 for one
 * network): the weight-gradient CTAs consume each tile's dY out of L2 as soon as the data-gradient CTAs have published it
 * through ready [tiles] int32 (workspace, zeroed by the call), so dY is never read back from HBM.  Same outputs as the two
 * separate calls up to fp32 summation order; the work after the call on `stream` is ordered behind both kernels. */
NFB_API int64_t nfb_mlp_train_tiles(int64_t M);
NFB_API int nfb_mlp_fwd_train(const nfb_mlp_t* h, const float* rays, const float* z_vals, int R, int S, float* raw,
                              void* act_img, uint32_t* mask, void* stream);
NFB_API int nfb_mlp_bwd_data(const nfb_mlp_t* h, const float* g_raw, int64_t M, const uint32_t* mask, void* dy_img,
                             void* stream);
NFB_API int nfb_mlp_bwd_weights(const nfb_mlp_t* h, const void* act_img, const void* dy_img, const float* g_raw, int64_t M,
                                float* grad, void* stream);
NFB_API int nfb_mlp_bwd(const nfb_mlp_t* h, const float* g_raw, int64_t M, const uint32_t* mask, const void* act_img,
                        void* dy_img, float* grad, int* ready, void* stream);

/* Profiling aid (no reference counterpart): the full forward (mode 1) while CTA 0 records a timeline of its barrier waits into
 * trace [3 roles][2048 events][4] uint64 = (tag, clock begin, clock end, aux); roles: 0 weight producer, 1 MMA warp,
 * 2 epilogue warps.  scripts/trace_mlp.py decodes it.                                                     */
NFB_API int nfb_mlp_fwd_trace(const nfb_mlp_t* h, const float* rays, const float* z_vals, int R, int S, float* raw,
                              unsigned long long* trace, void* stream);

/* Positional encoding to HBM (layer-wise fp32 path only).
 * replaces: run_nerf_helpers.py:36-50.  x [M,3] -> out [M, 3+6*L] at column offset col0 of a row pitch ld. */
NFB_API int nfb_embed(const float* x, int64_t M, int L, float* out, int ld, int col0, int64_t row_repeat, void* stream);

/* fp32 layer primitives (exact-parity / training path; CUDA-core FFMA, fp32 accumulate).
 * replaces: the addmm / relu / autograd pairs of run_nerf_helpers.py:100-123.
 * Every matrix carries its row pitch (ld*) so that torch.cat along features and column slices of a
 * weight are views: e.g. the skip layer's dX only for h uses (Wt + 63, ldw = 319, K = 256).
 *   fwd       : Y[M,N] = act(X[M,K] * W[N,K]^T + b)
 *   bwd_data  : dX[M,K] (+)= (dY o mask) * W         mask = (Y > 0) when relu (Y = the layer's output)
 *   bwd_weight: dW[N,K] = (dY o mask)^T * X ; db[N] = column sums of (dY o mask)   (db may be NULL)
 *               split over M into a caller workspace and reduced in a fixed order (deterministic).    */
NFB_API int nfb_linear_fwd(const float* X, int ldx, const float* Wt, int ldw, const float* b, int64_t M, int N, int K,
                   int relu, float* Y, int ldy, void* stream);
NFB_API int nfb_linear_bwd_data(const float* dY, int lddy, const float* Y, int ldy, int relu, const float* Wt, int ldw,
                        int64_t M, int N, int K, float* dX, int lddx, int accumulate, void* stream);
NFB_API int nfb_linear_bwd_weight(const float* dY, int lddy, const float* Y, int ldy, int relu, const float* X, int ldx,
                          int64_t M, int N, int K, float* dW, int lddw, float* db,
                          float* workspace, int64_t workspace_bytes, void* stream);
NFB_API int64_t nfb_linear_bwd_weight_workspace(int64_t M, int N, int K);

/* bf16 tensor-core weight gradient over "tile images" (per 128-row tile: 64-column chunks of 128 rows x 128 bytes with
 * the 128-byte swizzle, i.e. the layout the fused kernels use for their MMA operands and leave in HBM when training).
 * replaces: the wgrad GEMMs of loss.backward() (run_nerf.py:791) for one layer:
 *   dW[n, k] = sum_rows dY[row, n] * X[row, k],   db[n] = sum_rows dY[row, n]
 * dy: ndy in {2,4} chunks per tile (n = 64*ndy), x: nx in {1,2,4} chunks (k = 64*nx); *_tile_pitch = bytes between tiles.
 * The per-CTA partial sums are ACCUMULATED with L2 float reductions into out_w[(n - row_begin)*ld + col0 + k] for
 * n in [row_begin, row_end), k < cols_valid, and into out_b[n - row_begin] (or NULL): zero the gradients once per step.
 * status: device int that the kernel raises if a pipeline barrier timed out.                                        */
NFB_API int nfb_wgrad_bf16(const void* dy, int64_t dy_tile_pitch, int ndy, const void* x, int64_t x_tile_pitch, int nx,
                           int64_t ntiles, float* out_w, int ld, int col0, int cols_valid, int row_begin, int row_end,
                           float* out_b, int* status, void* stream);

/* Fused multi-tensor Adam (no amsgrad / weight decay), one launch for every parameter tensor of both networks.
 * replaces: optimizer.step() of torch.optim.Adam at run_nerf.py:792 (created at :213, betas (0.9, 0.999), eps 1e-8);
 * the learning-rate decay of :796-800 is a host scalar passed as lr.  step = 1 for the first update (bias corrections
 * 1 - beta^step).  grad_scale multiplies every gradient first (1/world_size of a data-parallel mean, else 1).
 * State layout = torch.optim.Adam's state_dict (exp_avg, exp_avg_sq per parameter), so reference checkpoints resume.  */
typedef struct nfb_adam_tensor {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq; int64_t numel;
} nfb_adam_tensor;
NFB_API int nfb_adam_step(const nfb_adam_tensor* tensors, int count, int64_t step, double lr, double beta1, double beta2,
                          double eps, double grad_scale, void* stream);
/* CUDA-graph form of the same update (optimizer.step() of run_nerf.py:792): the two step-dependent constants (lr / (1 - beta1^step), 1 / sqrt(1 - beta2^step);
 * nfb_adam_step_scalars computes them on the host exactly as nfb_adam_step does) are read from step_scalars [2] in device
 * memory when the kernel RUNS, so a captured training step is replayed with the current step and learning rate. */
NFB_API int nfb_adam_step_scalars(int64_t step, double lr, double beta1, double beta2, float* out2);
NFB_API int nfb_adam_step_dev(const nfb_adam_tensor* tensors, int count, const float* step_scalars, double beta1,
                              double beta2, double eps, double grad_scale, void* stream);

/* Alpha compositing. replaces: run_nerf.py:262-305 (raw2outputs) and, when pts_max != NULL,
 * nerf_to_coord.py:418-421 (arg-max-weight point o + d*z[argmax], first maximum wins).
 * raw [R,S,4], z_vals [R,S], rays_d [R,3] (taken from rays+3 with pitch ray_pitch floats), noise [R,S] or NULL.
 * Outputs (any may be NULL except weights): rgb_map [R,3], disp [R], acc [R], weights [R,S], depth [R],
 * pts_max [R,3] (needs rays_o = rays_d - 3 floats, i.e. pass a [R,11] ray batch).                    */
NFB_API int nfb_composite_fwd(const float* raw, const float* z_vals, const float* rays_d, int ray_pitch,
                      const float* noise, int R, int S, int white_bkgd,
                      float* rgb_map, float* disp, float* acc, float* weights, float* depth,
                      float* pts_max, void* stream);
/* Analytic backward of the above (suffix-sum form): what autograd derives for run_nerf.py:262-305 when loss.backward()
 * (run_nerf.py:790) runs.  Any g_* may be NULL (= zero). g_raw out [R,S,4]. */
NFB_API int nfb_composite_bwd(const float* raw, const float* z_vals, const float* rays_d, int ray_pitch,
                      const float* noise, int R, int S, int white_bkgd,
                      const float* g_rgb, const float* g_disp, const float* g_acc,
                      const float* g_weights, const float* g_depth, float* g_raw, void* stream);

/* Inverse-CDF sampling. replaces: run_nerf_helpers.py:200-243 (sample_pdf).
 * bins [R,nb], weights [R,nb-1] (row pitch w_pitch floats), u [R,N] or NULL (det: linspace(0,1,N)).
 * samples out [R,N]; inds out [R,N] int32 (the searchsorted result, for the index-parity tests) or NULL. */
NFB_API int nfb_sample_pdf(const float* bins, const float* weights, int w_pitch, const float* u,
                   int R, int nb, int N, float* samples, int32_t* inds, void* stream);

/* Fused hierarchical step. replaces: run_nerf.py:392-396 + :412
 * (z_mid, sample_pdf on weights[...,1:-1], sort(cat(z_vals, z_samples)), z_std).
 * z_coarse [R,Sc], weights [R,Sc], u [R,N] or NULL -> z_fine [R,Sc+N] ascending, z_samples [R,N] (or NULL),
 * z_std [R] (or NULL).  Requires 3 <= Sc <= 128, Sc+N <= 512.                                         */
NFB_API int nfb_hierarchical(const float* z_coarse, const float* weights, const float* u, int R, int Sc, int N,
                     float* z_fine, float* z_samples, float* z_std, void* stream);
/* The same with u drawn in the kernel (replaces the torch.rand of run_nerf_helpers.py:213): Philox stream 1 of the
 * generator described at nfb_coarse_z_rng, element r*N + k.                                                           */
NFB_API int nfb_hierarchical_rng(const float* z_coarse, const float* weights, uint64_t seed, uint64_t offset, int R, int Sc,
                                 int N, float* z_fine, float* z_samples, float* z_std, void* stream);

/* One ray batch, coarse + fine, without autograd: the whole kernel sequence of render_rays in one call.
 * replaces: run_nerf.py:308-418 (render_rays; nerf_to_coord.py:320-433 when pts_max != NULL) under torch.no_grad() with
 * raw_noise_std = 0: coarse depths -> fused MLP -> compositing -> inverse-CDF resampling + merge -> fused MLP -> compositing.
 * rays [R,11]; t_rand [R,N_samples] / u [R,N_importance]: the caller's uniform numbers, or NULL: deterministic sampling when
 * perturb == 0, else numbers drawn in the kernels (nfb_coarse_z_rng / nfb_hierarchical_rng with rng_seed, rng_offset);
 * outputs rgb [R,3], disp [R], acc [R] (+ rgb0 / disp0 / acc0 / z_std [R] when N_importance > 0, else ignored),
 * pts_max [R,3] or NULL.  fine may be NULL (the coarse network is queried twice, run_nerf.py:399).
 * workspace: nfb_render_rays_workspace_bytes(R, N_samples, N_importance) bytes, 256-byte aligned, caller-owned.     */
NFB_API size_t nfb_render_rays_workspace_bytes(int R, int N_samples, int N_importance);
NFB_API int nfb_render_rays_fwd(const nfb_mlp_t* coarse, const nfb_mlp_t* fine, const float* rays, int R, int N_samples,
                                int N_importance, int lindisp, int white_bkgd, const float* t_rand, const float* u,
                                int perturb, uint64_t rng_seed, uint64_t rng_offset, float* rgb, float* disp, float* acc, float* rgb0, float* disp0, float* acc0, float* z_std,
                                float* pts_max, void* workspace, size_t workspace_bytes, void* stream);

/* The same ray batch UNDER AUTOGRAD (the optimisation step of run_nerf.py:776-791): forward that keeps what the backward
 * needs, and the backward, one call each.
 * replaces: render_rays (run_nerf.py:308-418) with requires_grad parameters + loss.backward() through it, for
 * raw_noise_std = 0, N_importance > 0 and a separate fine network (the NeRFail configuration).
 * nfb_render_rays_train_fwd: coarse depths -> training forward of the coarse network (saves bf16 activation images and
 * relu masks) -> compositing -> resampling + merge (detached, :394) -> training forward of the fine network ->
 * compositing.  Outputs as nfb_render_rays_fwd (no pts_max); raw of the fine pass (retraw) stays in the workspace at byte
 * offset nfb_render_rays_train_raw_offset(...) as [R, N_samples + N_importance, 4].
 * nfb_render_rays_bwd: from the gradients of the six images (any may be NULL = zero; a pass whose three are NULL is
 * skipped) through compositing, the data-gradient chain and the grouped weight-gradient kernel of each network;
 * ACCUMULATES into grad_coarse / grad_fine [nfb_mlp_param_count] in state_dict order (zero them once per step).
 * The workspace (nfb_render_rays_train_workspace_bytes, 256-byte aligned, caller-owned) carries the saved state from the
 * forward to the backward and must not be touched in between; ~10.6 KB per sample.                                   */
NFB_API size_t nfb_render_rays_train_workspace_bytes(int R, int N_samples, int N_importance);
NFB_API size_t nfb_render_rays_train_raw_offset(int R, int N_samples, int N_importance);
NFB_API int nfb_render_rays_train_fwd(const nfb_mlp_t* coarse, const nfb_mlp_t* fine, const float* rays, int R, int N_samples,
                                      int N_importance, int lindisp, int white_bkgd, const float* t_rand, const float* u,
                                      float* rgb, float* disp, float* acc, float* rgb0, float* disp0, float* acc0, float* z_std,
                                      void* workspace, size_t workspace_bytes, void* stream);
NFB_API int nfb_render_rays_bwd(const nfb_mlp_t* coarse, const nfb_mlp_t* fine, const float* rays, int R, int N_samples,
                                int N_importance, int white_bkgd, const float* g_rgb, const float* g_disp, const float* g_acc,
                                const float* g_rgb0, const float* g_disp0, const float* g_acc0, float* grad_coarse,
                                float* grad_fine, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- */
/* B. GaussNet path                                                            */
/* ------------------------------------------------------------------------- */

/* Exact brute-force 8-NN (fp32 direct-difference distances, ascending, ties -> lowest index).
 * replaces: create_index_and_dist.py:126-145 (cdist + sort + running top-8 merge).
 * query [Q,3], cand [C,3] -> out_dist [Q,8] fp32, out_idx [Q,8] fp32 (the reference stores indices as
 * float32, :148-151; C < 2^24 required) and/or out_idx_i32 [Q,8].                                      */
NFB_API int nfb_knn8(const float* query, int64_t Q, const float* cand, int64_t C,
             float* out_dist, float* out_idx, int32_t* out_idx_i32, void* stream);

/* The same exact 8-NN through a uniform grid (identical outputs; ~1000x fewer distance evaluations on surface points).
 * nfb_knn_grid_build buckets the candidates once (they are the fixed P base views of a data set,
 * create_index_and_dist.py:96-108) into a 256^3 Morton-ordered grid: origin bbox_min (host, 3 floats), cell edge h with
 * bbox_min + 256 h >= bbox_max; sorted [C,4] fp32, cell_start [nfb_knn_grid_cells() + 1] int32, workspace
 * [nfb_knn_grid_cells() + 4096] int32.  nfb_knn8_grid answers queries against it; margin_abs = absolute slack of the
 * pruning bound (4e-6 x the largest |coordinate| is ample); stats: optional device uint64 += distance evaluations. */
NFB_API int64_t nfb_knn_grid_cells(void);
NFB_API int nfb_knn_grid_build(const float* cand, int64_t C, const float* bbox_min_host, float h, float* sorted,
                               int32_t* cell_start, int32_t* workspace, void* stream);
NFB_API int nfb_knn8_grid(const float* query, int64_t Q, const float* sorted, const int32_t* cell_start,
                          const float* bbox_min_host, float h, float margin_abs, float* out_dist, float* out_idx,
                          int32_t* out_idx_i32, unsigned long long* stats, void* stream);

/* replaces: model/GaussNet.py:169-186 (create_gauss_w.forward).
 * dist_idx [B,2,HW,8] -> i_w [B,2,HW,8] (ch0 = weights, ch1 = indices copied).                        */
NFB_API int nfb_gauss_weights(const float* dist_idx, int64_t B, int64_t HW, float c, float* i_w, void* stream);

/* replaces: model/GaussNet.py:53-119 (gather, weight, sum, alpha, eps-clip, composite on the original).
 * table [T,4]; w_idx [B,2,HW,8]; ori [B,HW,4] uint8; eps < 0 means "no clip" (epsilon=None).
 * x out [B,HW,4]; x_rgba out [B,HW,4]; minmax out [2] (running min / max of alpha*x_rgb where alpha>0,
 * updated with atomics, replaces the host-synchronising :89-103) or NULL.                             */
NFB_API int nfb_gauss_gather_fwd(const float* table, int64_t T, const float* w_idx, const uint8_t* ori,
                         int64_t B, int64_t HW, float eps, float* x, float* x_rgba, float* minmax,
                         void* stream);
/* Backward of the above w.r.t. table (autograd of model/GaussNet.py:53-119 under total_loss.backward(),
 * attack_NeRFail_S.py:346): one red.global.add.v4.f32 per (pixel, neighbour) straight into the L2-resident table.  (north_star asks for warp-aggregated atomics; the match.any merge of equal rows was built and measured SLOWER
 * on B200 — 64-73 us against 42-47 us per 800x800 view, the L2 atomic units absorb duplicates faster than eight
 * match.any rounds per pixel remove them — so it is opt-in, NERFAIL_B200_SCATTER_AGG=1; csrc/gauss.cu.)
 * g_x, g_xrgba [B,HW,4] (either may be NULL); x is the saved forward output; g_table [T,4] is
 * ACCUMULATED into (caller zeroes it).                                                               */
NFB_API int nfb_gauss_scatter_bwd(const float* g_x, const float* g_xrgba, const float* x, const float* w_idx,
                          const uint8_t* ori, int64_t B, int64_t HW, float eps, int64_t T,
                          float* g_table, void* stream);

/* The same backward for NC cotangents of x_rgba in ONE launch — the per-class gradients of DeepFool (deepfool.py:72-86: 14
 * torch.autograd.grad calls per iteration through the same forward).  g_xrgba [NC,B,HW,4]; weights, indices, the original
 * pixel and the saved x are read once per pixel; g_table [NC,T,4] is ACCUMULATED into (caller zeroes it).                */
NFB_API int nfb_gauss_scatter_bwd_batched(const float* g_xrgba, int NC, const float* x, const float* w_idx, const uint8_t* ori,
                                          int64_t B, int64_t HW, float eps, int64_t T, float* g_table, void* stream);

/* replaces: model/GaussNet.py:121-145 — NHWC RGBA -> NCHW RGB, white where alpha is 0 (the classifier's input).
 * img_f32 [B,HW,4] float or img_u8 [B,HW,4] uint8 (exactly one non-NULL); alpha_src [B,HW,4] float whose channel 3
 * decides (NULL: the image's own channel 3); out [B,3,HW] = alpha > 0 ? rgb : fill (255 in the reference).
 * nfb_chw_to_rgba is its adjoint (and, with fill = 0, its own second derivative): g_img [B,HW,4] =
 * alpha > 0 ? (g_out[b,0..2,q], 0) : 0 — so the pair stays differentiable twice for deepfool.py:76-77.          */
NFB_API int nfb_rgba_to_chw(const float* img_f32, const uint8_t* img_u8, const float* alpha_src, int64_t B, int64_t HW,
                            float fill, float* out, void* stream);
NFB_API int nfb_chw_to_rgba(const float* g_out, const float* alpha_src, int64_t B, int64_t HW, float* g_img, void* stream);

/* nfb_rgba_to_chw with the classifier's bilinear Resize fused behind it.
 * replaces: model/GaussNet.py:121-145 + :147-154 (torchvision.transforms.Resize([299,299]) / ([224,224]) on the float NCHW
 * tensor = ATen upsample_bilinear2d, align_corners = False; antialias selects _upsample_bilinear2d_aa, the default of
 * current torchvision for tensors, off in the torchvision 0.15 the reference pins).
 * A resize is separable: per axis and OUTPUT index a first tap, a tap count and normalised weights.  nfb_resize_weights
 * fills them on the HOST (fp32, ATen's formulas) for one axis: start / count [n], weights [n][maxk] with n = out_size, or,
 * transposed != 0, per INPUT index (n = in_size) the contiguous range of outputs it feeds — the table of the adjoint.
 * maxk >= nfb_resize_max_taps(...).  The caller uploads the tables once per (in, out, antialias).
 * nfb_rgba_to_chw_resized: out [B,3,OH,OW] = resize(alpha > 0 ? rgb : fill); alpha from channel 3 of the image itself or,
 * alpha_src != NULL, of alpha_src [alpha_batch,H*W,4] (image b uses b % alpha_batch).
 * nfb_chw_resized_to_rgba: its adjoint, g_img [B,H*W,4] = alpha > 0 ? (resize^T g_out, 0) : 0 — a gather through the
 * transposed tables, deterministic.  The pair is linear, each the other's derivative: differentiable twice (deepfool.py:76-77).
 * With B = NC * alpha_batch the adjoint serves NC cotangents of the same alpha_batch images in one launch.                 */
NFB_API int nfb_resize_max_taps(int in_size, int out_size, int antialias, int transposed);
NFB_API int nfb_resize_weights(int in_size, int out_size, int antialias, int transposed, int maxk,
                               int* start_host, int* count_host, float* weights_host);
NFB_API int nfb_rgba_to_chw_resized(const float* img_f32, const uint8_t* img_u8, const float* alpha_src, int64_t alpha_batch,
                                    int64_t B, int H, int W, int OH, int OW, float fill,
                                    const int* y_start, const int* y_count, const float* y_w, int y_maxk,
                                    const int* x_start, const int* x_count, const float* x_w, int x_maxk,
                                    float* out, void* stream);
NFB_API int nfb_chw_resized_to_rgba(const float* g_out, const float* alpha_src, int64_t alpha_batch, int64_t B, int H, int W,
                                    int OH, int OW,
                                    const int* yt_start, const int* yt_count, const float* yt_w, int yt_maxk,
                                    const int* xt_start, const int* xt_count, const float* xt_w, int xt_maxk,
                                    float* g_img, void* stream);

/* The I-FGSM update of attack_NeRFail_S.py:357-392 restricted to the rows it can change (alpha > 0; active_idx [n] int64 row
 * numbers of the [T,4] perturbation table, fixed for a whole attack).  nfb_attack_pack_rgb packs the RGB gradient of those
 * rows, [n,3], which is all a data-parallel attack has to all-reduce; nfb_attack_sign_step applies
 * rgb <- clamp(rgb - signed_step * sign(g), init - eps, init + eps) to them in place (signed_step > 0 descends).     */
NFB_API int nfb_attack_pack_rgb(const float* grad, const int64_t* active_idx, int64_t n, float* packed, void* stream);
NFB_API int nfb_attack_sign_step(float* table, const float* init, const int64_t* active_idx, const float* packed_grad,
                                 int64_t n, float signed_step, float eps, void* stream);

/* ------------------------------------------------------------------------- */
/* C. Multi-GPU exchange steps over NVLink peer memory (one process per GPU)   */
/* ------------------------------------------------------------------------- */

/* Peer memory (no reference counterpart: the reference is single-process; the steps built on it replace
 * attack_NeRFail_S.py:348-392 and run_nerf.py:791-792 of a data-parallel run): a cudaMalloc allocation (zero-filled) that
 * the other ranks of the node open through a 64-byte CUDA IPC handle
 * (exchanged by the host side, e.g. torch.distributed.all_gather_object); peer access is enabled on first open.
 * nfb_peer_create binds, for a group of G <= 8 ranks, every rank's gradient buffer, every rank's copy of the quantity
 * being updated and every rank's flag words (nfb_peer_flag_bytes() bytes, zero-filled) — arrays of G pointers indexed by
 * rank, own buffers included.  All ranks must call the exchange steps in the same order (collective semantics).
 * nfb_peer_status: NFB_E_CUDA once a flag wait timed out (a rank never arrived); sticky, no synchronisation.          */
typedef struct nfb_peer nfb_peer_t;
NFB_API int nfb_peer_alloc(size_t bytes, void** out);
NFB_API int nfb_peer_free(void* p);
NFB_API int nfb_peer_export(void* p, void* handle64);
NFB_API int nfb_peer_import(const void* handle64, void** out);
NFB_API int nfb_peer_close(void* p);
NFB_API int nfb_peer_flag_bytes(void);
NFB_API int nfb_peer_create(nfb_peer_t** out, int rank, int G, void* const* grad_ptrs, void* const* value_ptrs,
                            void* const* flag_ptrs);
NFB_API int nfb_peer_destroy(nfb_peer_t* h);
NFB_API int nfb_peer_status(const nfb_peer_t* h);

/* One attack iteration's exchange as ONE kernel per GPU.
 * replaces: the gradient reduction a data-parallel run of attack_NeRFail_S.py:348 needs + the I-FGSM update of :357-392.
 * grad buffers [T,4] hold each rank's partial grad_spatial_rgb (its views' scatter), value buffers [T,4] each rank's copy
 * of the perturbation table.  Rank r owns rows [r*ceil(T/G), ...): for its rows with A > 0 it sums the G partial
 * gradients through peer loads, applies rgb <- clamp(rgb - signed_step * sign(g), init - eps, init + eps) and stores the
 * row into every rank's table.  On return (stream order) this rank's table is final and its gradient buffer may be zeroed. */
NFB_API int nfb_attack_exchange_step(nfb_peer_t* h, const float* init, int64_t T, float signed_step, float eps, void* stream);

/* One retraining step's exchange + optimiser as ONE kernel per GPU.
 * replaces: the gradient average a data-parallel run of run_nerf.py:791 needs + optimizer.step() of :792 (torch.optim.Adam,
 * arithmetic of nfb_adam_step).  grad / value buffers are the flat [n] gradients / parameters (padded to a multiple of 4
 * floats) of ALL parameters in optimiser order; exp_avg / exp_avg_sq [n] are local (only this rank's slice is used: the
 * optimiser state is sharded).  step_scalars: device [2] as nfb_adam_step_dev.  grad_scale = 1 / G for a mean.           */
NFB_API int nfb_adam_exchange_step(nfb_peer_t* h, float* exp_avg, float* exp_avg_sq, int64_t n, const float* step_scalars,
                                   double beta1, double beta2, double eps, double grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERFAIL_B200_H */
