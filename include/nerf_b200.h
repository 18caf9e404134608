/*
 * nerf_b200.h -- C ABI of the B200-native NeRF ray-render hot path.
 *
 * This is the drop-in boundary: the reference (johnfay11/CV-Nerf) is pure Python/PyTorch and has
 * no FFI of its own, so each entry point below names the reference *function* it replaces
 * (file:line in /root/reference).  The Python host in cv-nerf_b200/ binds these with ctypes and
 * re-exposes the reference's call surface (compute_rays, get_ndc, render, render_rays,
 * process_volume_info, inv_transform_sampling, Model, net_forward, create_model).
 *
 * Conventions
 *  - every pointer is a CUDA *device* pointer owned by the caller unless the name says `host`;
 *    the library never allocates or frees caller-visible memory;
 *  - all tensors are contiguous fp32, row-major, shapes in the comments;
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream;
 *  - return value 0 = ok; negative = argument error; positive = cudaError_t of the launch.
 *    nerf_b200_last_error() returns a static description of the last failure on this thread.
 *  - there is no CPU path: without a CUDA device every call fails with a cudaError_t.
 */
#ifndef NERF_B200_H
#define NERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NERF_B200_ABI_VERSION 1

#define NERF_RAY_STRIDE 11      /* [o(3) d(3) near far viewdir(3)], main.py:71-76 */
#define NERF_N_PARAM_TENSORS 24 /* 12 x (weight, bias) in Model registration order, model.py:57-71 */

/* argument-error codes */
#define NERF_ERR_ARG (-1)
#define NERF_ERR_UNSUPPORTED (-2)

int nerf_b200_abi_version(void);
const char* nerf_b200_last_error(void);
/* number of SMs of the current device (grid sizing for callers that want to report it) */
int nerf_b200_sm_count(void);

/* ---------------------------------------------------------------- K1: rays (bit-exact fp32) */

/* compute_rays, main.py:19-46.  pose: [3,4] row-major (device).  Rows [row0,row1) of the H x W
 * image are generated; dirs_out / origins_out are [(row1-row0)*W, 3] (origins_out may be NULL:
 * the reference returns a stride-0 view of pose[:,3]). */
int nerf_compute_rays(int H, int W, float focal, const float* pose, int row0, int row1,
                      float* origins_out, float* dirs_out, void* stream);

/* get_ndc, data_helpers.py:327-344 (quirks reproduced).  cw = fp32(-1/(W/(2 focal))) and ch are
 * folded by the host exactly as Python/numpy folds them.  o,d,o_out,d_out: [n,3]. */
int nerf_get_ndc(float cw, float ch, float near_plane, const float* o, const float* d, long n,
                 float* o_out, float* d_out, void* stream);

/* render() front end, main.py:55-76, fused: ray generation for rows [row0,row1) (pose != NULL) or
 * a caller-provided batch (rays_o/rays_d [n,3], pose == NULL), view-dir normalisation from the
 * pre-NDC direction, optional NDC warp (plane near = 1), packing to rays_out [n,11]. */
int nerf_pack_rays(int H, int W, float focal, float cw, float ch, const float* pose, int row0,
                   int row1, const float* rays_o, const float* rays_d, long n, int ndc,
                   float near, float far, float* rays_out, void* stream);

/* coarse depths, main.py:221-234.  rays [n,11] (near/far read from columns 6,7), t_rand [n,S]
 * or NULL (perturb == 0), z_out [n,S]. */
int nerf_sample_coarse(const float* rays, long n, int S, const float* t_rand, float* z_out,
                       void* stream);

/* ---------------------------------------------------------------- K3/K4: compositing, resampling */

/* process_volume_info, main.py:174-204.  raw [n,S,4], z [n,S], dirs: row i at dirs + i*dir_stride
 * (3 floats; pass rays+3 with stride 11, or a [n,3] tensor with stride 3), noise [n,S] already
 * multiplied by the noise scale or NULL, rgb_out [n,3], weights_out [n,S] or NULL. */
int nerf_composite_fwd(const float* raw, const float* z, const float* dirs, int dir_stride,
                       const float* noise, long n, int S, int white_bkg, float* rgb_out,
                       float* weights_out, void* stream);

/* backward of the above w.r.t. raw: grad_rgb [n,3], grad_weights [n,S] or
 * NULL -> grad_raw [n,S,4].  grad_scale multiplies the result (loss normalisation). */
int nerf_composite_bwd(const float* raw, const float* z, const float* dirs, int dir_stride,
                       const float* noise, long n, int S, int white_bkg, const float* grad_rgb,
                       const float* grad_weights, float* grad_raw, void* stream);

/* Extra maps from the compositing weights: maps_out [n,3] = (depth = sum w z, acc = sum w,
 * disp = 1 / max(1e-10, depth / acc)).  The reference forms only sum(w) (main.py:199) and returns
 * none of them; definitions follow nerf-pytorch's raw2outputs. */
int nerf_composite_maps(const float* weights, const float* z, long n, int S, float* maps_out,
                        void* stream);

/* to_byte / cont_to_byte8_im, model.py:134, utils.py:57: out[i] = uint8(255 * clip(x[i], 0, 1)). */
int nerf_to_byte(const float* x, long n, unsigned char* out, void* stream);

/* inv_transform_sampling, utils.py:4-53, with the uniform draws supplied by the caller.
 * bins [n,B], weights [n,B-1], u [n,m] -> samples_out [n,m] (unsorted). B <= 256. */
int nerf_sample_pdf(const float* bins, const float* weights, const float* u, long n, int B, int m,
                    float* samples_out, void* stream);

/* main.py:248-251 fused: midpoints of z_c, pdf over w_c[:,1:-1], inverse-CDF samples for u [n,m],
 * sort(cat(z_c, samples)) -> z_f [n,S+m].  S+m <= 256. */
int nerf_resample_merge(const float* z_c, const float* w_c, const float* u, long n, int S, int m,
                        float* z_f, void* stream);

/* ---------------------------------------------------------------- in-kernel random draws
 * The reference draws torch.rand / torch.randn tensors for the stratified jitter (main.py:233), the
 * density noise (main.py:188) and the inverse-CDF uniforms (utils.py:23, drawn even at test time).
 * The entry points above take those tensors from the caller (parity runs inject the numbers the
 * oracle consumed).  The *_rng variants draw them inside the consuming kernel from Philox4x32-10
 * keyed by (seed; ray0 + ray, value index / 4, stream), so no [n,S] tensor of random numbers ever
 * touches HBM and a row-sharded render (ray0 = first ray of the shard) draws the same numbers as the
 * unsharded one.  csrc/rng.cuh states the mapping; nerf_rng_fill writes the same numbers out. */
#define NERF_RNG_STREAM_T_RAND 0  /* uniforms [n,S_c]: stratified jitter              */
#define NERF_RNG_STREAM_U 1       /* uniforms [n,m]:   inverse-CDF resampling          */
#define NERF_RNG_STREAM_NOISE_C 2 /* normals  [n,S_c]: density noise, coarse pass      */
#define NERF_RNG_STREAM_NOISE_F 3 /* normals  [n,S_c+m]: density noise, fine pass      */

/* nerf_sample_coarse with t_rand drawn in place (perturb > 0, main.py:227-234). */
int nerf_sample_coarse_rng(const float* rays, long n, int S, unsigned long long seed, long ray0,
                           float* z_out, void* stream);

/* nerf_resample_merge with u drawn in place (utils.py:23). */
int nerf_resample_merge_rng(const float* z_c, const float* w_c, unsigned long long seed, long ray0,
                            long n, int S, int m, float* z_f, void* stream);

/* nerf_composite_fwd / _bwd with noise = noise_scale * N(0,1) drawn in place (main.py:186-189);
 * rng_stream is NERF_RNG_STREAM_NOISE_C or _F.  Forward and backward regenerate the same values. */
int nerf_composite_fwd_rng(const float* raw, const float* z, const float* dirs, int dir_stride,
                           float noise_scale, unsigned long long seed, int rng_stream, long ray0,
                           long n, int S, int white_bkg, float* rgb_out, float* weights_out,
                           void* stream);
int nerf_composite_bwd_rng(const float* raw, const float* z, const float* dirs, int dir_stride,
                           float noise_scale, unsigned long long seed, int rng_stream, long ray0,
                           long n, int S, int white_bkg, const float* grad_rgb,
                           const float* grad_weights, float* grad_raw, void* stream);

/* out [n,cols] = the draws of rows ray0..ray0+n-1 of a stream; kind 0: uniforms, 1: normals. */
int nerf_rng_fill(int kind, unsigned long long seed, int rng_stream, long ray0, long n, int cols,
                  float* out, void* stream);

/* ---------------------------------------------------------------- whole chain
 * render_rays, main.py:207-261 (and, with pose != NULL, the ray generation / NDC / packing front end
 * of render, main.py:49-87, for image rows [row0,row1)): every launch of the chain -- coarse depths,
 * view terms, coarse field, compositing, resampling + merge, fine field, compositing -- sequenced on
 * `stream` by one call.  Inference only (no activation records); random draws in-kernel from
 * (seed, ray0).  packed_* / host_tail_*: nerf_pack_model blobs and their nerf_model_host_tail copies.
 * rays_in [n,11] is read when pose == NULL.  scratch: nerf_render_scratch_bytes(n, S_c, n_fine) bytes
 * of device memory (n = (row1-row0)*W with a pose).  rgb_out, rgb_c_out: [n,3].
 * field_events: NULL, or a HOST array of four cudaEvent_t recorded on `stream` before / after the
 * coarse and the fine field-network launch (how bench.py times the dominant kernel inside a step). */
size_t nerf_render_scratch_bytes(long n, int S_c, int n_fine);
int nerf_render_fused(const void* packed_coarse, const void* host_tail_coarse, const void* packed_fine,
                      const void* host_tail_fine, int H, int W, float focal, float cw, float ch,
                      const float* pose, int row0, int row1, const float* rays_in, long n, int ndc,
                      float near, float far, int S_c, int n_fine, float perturb, float noise,
                      int white_bkg, unsigned long long seed, long ray0, void* scratch, float* rgb_out,
                      float* rgb_c_out, void* const* field_events, void* stream);

/* ---------------------------------------------------------------- K2: the field network */

/* Bytes of the packed parameter blob of one Model (BF16 UMMA-layout weight stages + fp32 tail). */
size_t nerf_packed_model_bytes(void);

/* Model parameters (model.py:57-71) -> packed blob.  host_params: HOST array of 24 DEVICE
 * pointers in registration order l1.weight, l1.bias, ..., l9, l_alpha, l10, l11. */
int nerf_pack_model(const float* const* host_params, void* packed_out, void* stream);

/* Per-ray (or per-row) view-direction term of layer l10: W10[:,256:283] . PE4(dir) + b10,
 * model.py:103-104 hoisted out of the per-sample loop.  dirs: row i at dirs + i*dir_stride;
 * embedded = 0: rows are 3-vectors (PE computed here); 1: rows are 27-wide encodings.
 * out [count,128]. */
int nerf_viewdir_term(const void* packed, const float* dirs, int dir_stride, int embedded,
                      long count, float* out, void* stream);

/* FreqEmbedding.embed, model.py:9-31: x [n,dim] -> out [n, dim*(1+2*n_freq)] =
 * [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]. */
int nerf_freq_encode(const float* x, long n, int dim, int n_freq, float* out, void* stream);

#define NERF_IN_RAYS 0     /* in0 = rays [n,11], in1 = z [n,S]; point = o + d*z        */
#define NERF_IN_POINTS 1   /* in0 = points [M,3]                                       */
#define NERF_IN_EMBEDDED 2 /* in0 = rows [M,in_stride], first 63 columns = PE10(point) */

/* net_forward + Model.forward, model.py:77-131, fused: point -> PE -> 8x256 trunk (skip at l6)
 * -> sigma head, l9, l10 (+vterm), l11 -> raw_out [M,4] = (rgb_raw, sigma_raw).
 * M rows; S samples per ray (row r belongs to ray r / S, only used by NERF_IN_RAYS);
 * vterm [ceil(M / vterm_div), 128] from nerf_viewdir_term.
 * act_save: NULL, or a buffer of nerf_mlp_act_bytes(M) bytes that receives the BF16 activations
 * the backward pass needs.
 * row0: global index of the first row when this call is one shard of a larger batch (0 otherwise).
 * The two 128-row sub-tiles a CTA works on consume a layer's K chunks in opposite orders (the second
 * re-uses the weight slots of the first, csrc/mlp_fwd.cu chunk_at), chosen by the parity of the row's
 * GLOBAL tile index; shards that start at a multiple of 128 rows and pass row0 reproduce the unsharded
 * result bit for bit (other shard points agree to fp32 summation order, ~1e-7). */
int nerf_mlp_fwd(const void* packed, int in_mode, const float* in0, const float* in1, int in_stride,
                 long M, int S, const float* vterm, int vterm_div, float* raw_out, void* act_save,
                 long row0, void* stream);

size_t nerf_mlp_act_bytes(long M);

/* Inference fast path: the biases, l_alpha and l11 (11.8 KB of the blob's fp32 tail) travel in the
 * kernel parameters, so the epilogues read them through the constant bank instead of the L1 data
 * pipe the tensor core reads its operands through.  nerf_model_host_tail copies them to a HOST
 * buffer of nerf_model_host_tail_bytes() bytes (asynchronously on `stream`; synchronise before
 * use); nerf_mlp_fwd_host_tail is nerf_mlp_fwd (without act_save) taking that host copy.  Results
 * are identical to nerf_mlp_fwd. */
size_t nerf_model_host_tail_bytes(void);
int nerf_model_host_tail(const void* packed, void* host_tail_out, void* stream);
int nerf_mlp_fwd_host_tail(const void* packed, const void* host_tail, int in_mode, const float* in0,
                           const float* in1, int in_stride, long M, int S, const float* vterm,
                           int vterm_div, float* raw_out, long row0, void* stream);
/* Test support: nerf_mlp_fwd that additionally writes the FP32 post-activation output of tensor-core
 * layer `probe_layer` (0 = l1 ... 8 = l9, 9 = l10) to probe_out [M,256] (l10 fills the first 128
 * columns).  The layer-by-layer parity tests compare it with a CPU emulation of the kernel's rounding
 * points (tests/test_gpu_kernels.py::test_field_embedded_mode_layers). */
int nerf_mlp_fwd_probe(const void* packed, int in_mode, const float* in0, const float* in1,
                       int in_stride, long M, int S, const float* vterm, int vterm_div,
                       float* raw_out, int probe_layer, float* probe_out, void* stream);

/* ---------------------------------------------------------------- training: backward of the field
 * Replaces autograd through Model.forward for loss.backward(), main.py:385.  Three stages:
 *   nerf_mlp_bwd_dz     grad_raw [M,4] + saved activations -> dZ of every layer (BF16 tile images,
 *                       nerf_mlp_dz_bytes(M) bytes), tensor cores against the transposed weights
 *   nerf_mlp_bwd_dw     dW/db of l1..l10 = dZ^T . X over the sample axis, tensor cores
 *   nerf_mlp_bwd_heads  dW/db of l_alpha and l11; nerf_viewdir_term_bwd: l10's view columns + bias
 * All of them ACCUMULATE (+=) into a padded fp32 gradient blob of nerf_grad_blob_bytes() bytes
 * that the caller zeroes; nerf_grad_unpack scatters it into the 24 .grad tensors (registration
 * order, like nerf_pack_model), adding to them when accumulate != 0. */
size_t nerf_packed_model_bwd_bytes(void);
int nerf_pack_model_bwd(const float* const* host_params, void* packed_bwd_out, void* stream);
/* nerf_pack_model + nerf_pack_model_bwd for n_models (1 or 2) Models in one launch (what a train
 * step re-packs after the optimizer).  host_params: n_models x 24 device pointers. */
int nerf_pack_models_train(int n_models, const float* const* host_params, void* const* packed_out,
                           void* const* packed_bwd_out, void* stream);
size_t nerf_mlp_dz_bytes(long M);
int nerf_mlp_bwd_dz(const void* packed_bwd, const float* grad_raw, const void* act_save, long M,
                    void* dz_out, void* stream);
size_t nerf_grad_blob_bytes(void);
int nerf_mlp_bwd_dw(const void* act_save, const void* dz, long M, float* grad_blob, void* stream);
int nerf_mlp_bwd_heads(const void* act_save, const float* grad_raw, long M, float* grad_blob,
                       void* stream);
/* Deterministic accumulation (parity / debugging runs).  nerf_mlp_bwd_dw, nerf_mlp_bwd_heads,
 * nerf_viewdir_term_bwd and nerf_mse_loss_grad add per-CTA sums with floating-point atomics, so two runs
 * differ in the last bits of a gradient (the reference's CPU autograd is bitwise reproducible).  The _det
 * variants write the per-CTA partial sums to `scratch` and add them in CTA order with a second launch:
 * bitwise reproducible on a given GPU model.  Scratch sizes: nerf_mlp_bwd_dw_det_scratch_bytes();
 * nerf_bwd_det_scratch_bytes(kind, size) with kind 1 = heads (size = M), 2 = view columns (size = number
 * of rays), 3 = loss (size = n). */
size_t nerf_mlp_bwd_dw_det_scratch_bytes(void);
size_t nerf_bwd_det_scratch_bytes(int kind, long size);
int nerf_mlp_bwd_dw_det(const void* act_save, const void* dz, long M, float* grad_blob, void* scratch,
                        void* stream);
int nerf_mlp_bwd_heads_det(const void* act_save, const float* grad_raw, long M, float* grad_blob,
                           void* scratch, void* stream);
int nerf_viewdir_term_bwd_det(const void* dz, const float* dirs, int dir_stride, int embedded, long M,
                              int vterm_div, float* grad_blob, void* scratch, void* stream);
int nerf_mse_loss_grad_det(const float* x, const float* target, long n, float* grad_out,
                           float* loss_accum, void* scratch, void* stream);

/* l9 is folded into l10 in the forward pass and in the dZ chain (feat = l9(h8) feeds l10 without an
 * activation, model.py:100-104; csrc/mlp_layout.h), so nerf_mlp_bwd_dw leaves G = dZ10^T . h8 in the
 * blob's scratch region instead of the gradients of l9 and of l10's first 256 columns.  This call forms
 * them by the chain rule -- dW10[:, :256] += G W9^T + db10 (x) b9, dW9 += W10[:, :256]^T G,
 * db9 += W10[:, :256]^T db10 -- from the FP32 parameters l9.weight [256,256], l9.bias [256],
 * l10.weight [128,283].  Call it ONCE per backward pass, after nerf_mlp_bwd_dw and
 * nerf_viewdir_term_bwd (which produces db10) have completed on `stream`. */
int nerf_mlp_bwd_unfold(float* grad_blob, const float* l9_weight, const float* l9_bias,
                        const float* l10_weight, void* stream);

/* dirs / dir_stride / embedded / vterm_div exactly as passed to nerf_viewdir_term + nerf_mlp_fwd */
int nerf_viewdir_term_bwd(const void* dz, const float* dirs, int dir_stride, int embedded, long M,
                          int vterm_div, float* grad_blob, void* stream);
int nerf_grad_unpack(const float* grad_blob, float* const* host_grads, int accumulate, void* stream);

/* mean((x - target)^2) over n elements, main.py:380-383: *loss_accum += the mean (may be NULL),
 * grad_out[i] = 2 (x[i] - target[i]) / n (may be NULL). */
int nerf_mse_loss_grad(const float* x, const float* target, long n, float* grad_out,
                       float* loss_accum, void* stream);

/* ---------------------------------------------------------------- training: optimizer, ray batch */

/* torch.optim.Adam(lr, betas=(beta1,beta2), eps) of main.py:144,386 for n_tensors tensors in one
 * launch.  All pointer arrays are HOST arrays of DEVICE pointers; sizes in elements; step is the
 * 1-based step count after this update (bias correction); grads are multiplied by grad_scale. */
int nerf_adam_step(int n_tensors, float* const* params, const float* const* grads,
                   float* const* exp_avg, float* const* exp_avg_sq, const long* sizes, float lr,
                   float beta1, float beta2, float eps, long step, float grad_scale, void* stream);
/* Same for the 24 tensors of one Model, gradient read from the padded gradient blob. */
int nerf_adam_step_blob(const float* grad_blob, float* const* params, float* const* exp_avg,
                        float* const* exp_avg_sq, float lr, float beta1, float beta2, float eps,
                        long step, float grad_scale, void* stream);

/* Data-parallel Adam fused with the gradient exchange: peer_blobs is a HOST array of `world`
 * DEVICE pointers, entry k being rank k's gradient blob of this Model mapped into this process
 * (symmetric / peer memory; entry `rank` is the local blob).  The kernel sums the blobs in rank
 * order over NVLink and applies nerf_adam_step_blob's update with grad_scale (1/world for a mean).
 * The caller provides the cross-rank barriers before (all blobs complete) and after (blobs may be
 * reused) the launch. */
int nerf_adam_step_blob_peers(const float* const* peer_blobs, int world, float* const* params,
                              float* const* exp_avg, float* const* exp_avg_sq, float lr, float beta1,
                              float beta2, float eps, long step, float grad_scale, void* stream);

/* One train iteration's batch, main.py:351-374: n rays (packed [n,11] like nerf_pack_rays) and
 * their target pixels [n,3] from image [H,W,3].  pix != NULL: the caller's linear pixel indices
 * (i*W + j).  pix == NULL: n DISTINCT pixels of the crop window drawn by a keyed permutation
 * (np.random.choice(replace=False) of main.py:368; pre-crop window of main.py:354-361).
 * target_out / image / pix_out may be NULL. */
int nerf_train_rays(int H, int W, float focal, float cw, float ch, const float* pose, const int* pix,
                    unsigned long long seed, int crop_r0, int crop_c0, int crop_h, int crop_w, long n,
                    int ndc, float near, float far, const float* image, float* rays_out,
                    float* target_out, int* pix_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_H */
