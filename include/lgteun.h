/* lgteun.h — C ABI of the B200-native LGTEUN forward hot path.
 *
 * Drop-in boundary: the reference calls
 *     self.module_dict['core_module'](input_lr, input_pan)          models/unlg_former.py:80-85
 * i.e. Pansharpening.forward(ms, pan)                               models/unlg_former.py:50-67
 * The reference has no native FFI (it is pure PyTorch); these entry points are what a ctypes/cffi
 * binding of that call binds (see INTEGRATION.md for the reference-side stub).
 *
 * Conventions
 *  - plain C types only: device pointers are `const float*` / `float*` into CUDA global memory of the
 *    handle's device, `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *  - all tensors are fp32, contiguous; ms [N,B,h,w], pan [N,1,4h,4w], out [N,B,4h,4w] (NCHW) exactly as
 *    the reference's forward(LrMS, PAN) receives/returns them;
 *  - every function returns 0 on success or a negative LGTEUN_E* code; the message is available through
 *    lgteun_last_error() (thread-local).  Nothing throws across the ABI.  Unsupported shapes are errors,
 *    never a fallback: 4h and 4w must be powers of two in [16, 1024] (window 8 at two U-Net levels,
 *    power-of-two FFT passes), B in {4, 8}.
 *  - a handle is bound to one device and must be used by one host thread at a time (the reference's
 *    nn.DataParallel path uses one replica, hence one handle, per GPU: models/base/base_model.py:95-97).
 *  - the caller owns ms/pan/out; inputs are never written (base_model.py:304-305 reuses them).
 */
#ifndef LGTEUN_H_
#define LGTEUN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lgteun_ctx lgteun_t;

enum {
  LGTEUN_OK = 0,
  LGTEUN_EINVAL = -1,   /* bad argument / unsupported shape            */
  LGTEUN_ECUDA = -2,    /* a CUDA runtime call failed                  */
  LGTEUN_ESTATE = -3,   /* weights not loaded / missing state_dict key */
  LGTEUN_ENOMEM = -4
};

/* forward() flags */
enum {
  LGTEUN_RUN_DEAD_PRIORS = 1, /* also execute prior_module[0..K-2], whose output the reference discards
                                 (models/unlg_former.py:63-67); the returned tensor is identical        */
  LGTEUN_NO_GRAPH = 2         /* launch kernels directly instead of replaying the cached CUDA graph      */
};

/* ABI version of this header (bumped on any signature change). */
int lgteun_abi_version(void);

/* Thread-local message of the last failing call ("" if none). */
const char* lgteun_last_error(void);

/* Pansharpening.__init__(cfg, logger, stage)  — models/unlg_former.py:22-48.
 * bands = cfg.ms_chans (4: GF-2/WV-2, 8: WV-3), stages = K (config: 2, configs/unlg_former.py:92-94). */
int lgteun_create(int device, int bands, int stages, lgteun_t** out);
void lgteun_destroy(lgteun_t* ctx);

/* Number of state_dict tensors the handle expects and the i-th key / element count
 * (the weight ABI: SURVEY.md Appendix B; nn.Module.state_dict() of the reference class). */
int lgteun_num_weights(const lgteun_t* ctx);
const char* lgteun_weight_name(const lgteun_t* ctx, int i);
int64_t lgteun_weight_numel(const lgteun_t* ctx, int i);

/* load_state_dict: snapshot n fp32 device tensors (names = reference state_dict keys) into the handle's
 * packed weight arena.  Every expected key must be present with the expected element count.
 * Mirrors base_model.py:102-114 (module.load_state_dict(ckpt[name].state_dict())). */
int lgteun_load_weights(lgteun_t* ctx, const char* const* names, const float* const* dev_ptrs,
                        const int64_t* numels, int n, void* stream);

/* Bytes of device workspace the handle holds for a given problem size (allocated lazily, grown on demand). */
int64_t lgteun_workspace_bytes(const lgteun_t* ctx, int N, int h, int w);

/* Pansharpening.forward(ms, pan) -> HrMS   — models/unlg_former.py:50-67.
 * ms [N,B,h,w], pan [N,1,4h,4w], out [N,B,4h,4w]; device pointers; enqueued on `stream`. */
int lgteun_forward(lgteun_t* ctx, const float* ms, const float* pan, float* out,
                   int N, int h, int w, int flags, void* stream);

/* Same call with HOST buffers (pinned or pageable): H2D of ms/pan, forward, D2H of out, all on `stream`,
 * returns after the result is in out_host.  This is the end-to-end path of base_model.py:293-305
 * (set_batch_cuda -> get_model_output -> torch2np). */
int lgteun_forward_host(lgteun_t* ctx, const float* ms_host, const float* pan_host, float* out_host,
                        int N, int h, int w, int flags, void* stream);

/* How many kernels one lgteun_forward of this size launches (graph nodes), for the bench's launch count. */
int lgteun_forward_launches(lgteun_t* ctx, int N, int h, int w, int flags);

/* ---- per-operator entry points (same kernels the forward chains; used by tests and ncu) ----------
 * `prior` selects prior_module[prior]; `lgb` selects 0 = encoder_layers.0.0, 1 = bottleneck,
 * 2 = decoder_layers.0.2; `block` the block inside that LGB.  NHWC tensors are [N,H,W,c]. */

/* bmu.sampling_(x, s_factor) for s in {4, 2, 0.5} on NCHW planes  — basic_module_unformer_v2.py:21-23.
 * scale_num/scale_den = 4/1, 2/1 or 1/2. */
int lgteun_op_bicubic(lgteun_t* ctx, const float* x, float* y, int planes, int h, int w,
                      int scale_num, int scale_den, void* stream);

/* One data-module step  Z - eta[stage]*(DT(D(Z)-ms) + RT(R(Z)-pan))  — unlg_former.py:58-61.
 * z_in/z_out [N,B,4h,4w] (must not alias). */
int lgteun_op_data_step(lgteun_t* ctx, int stage, const float* z_in, const float* ms, const float* pan,
                        float* z_out, int N, int h, int w, void* stream);

/* patch_embedding.forward — LGT.py:85-88: NCHW [N,B,H,W] -> NHWC [N,H,W,4B]. */
int lgteun_op_patch_embed(lgteun_t* ctx, int prior, const float* x_nchw, float* y_nhwc,
                          int N, int H, int W, void* stream);

/* residual(pre_norm(LGMixer))  — LGT.py:45-61,200-219: y = LGMixer(LN(x)) + x, NHWC [N,H,W,c]. */
int lgteun_op_mixer(lgteun_t* ctx, int prior, int lgb, int block, const float* x, float* y,
                    int N, int H, int W, void* stream);

/* the two halves of the mixer on their own, input = LN'd half [N,H,W,c/2] as the reference modules see it:
 * local_mixer.forward + window merge (LGT.py:130-146,207-208) and global_mixer.forward (LGT.py:162-180). */
int lgteun_op_local_mixer(lgteun_t* ctx, int prior, int lgb, int block, const float* x_half, float* y_half,
                          int N, int H, int W, void* stream);
int lgteun_op_global_mixer(lgteun_t* ctx, int prior, int lgb, int block, const float* x_half, float* y_half,
                           int N, int H, int W, void* stream);

/* residual(pre_norm(feed_forward))  — LGT.py:95-109: y = FFN(LN(x)) + x, NHWC. */
int lgteun_op_ffn(lgteun_t* ctx, int prior, int lgb, int block, const float* x, float* y,
                  int N, int H, int W, void* stream);

/* LGT.forward — LGT.py:314-344: NCHW [N,B,H,W] -> NCHW [N,B,H,W]. */
int lgteun_op_prior(lgteun_t* ctx, int prior, const float* x_nchw, float* y_nchw,
                    int N, int H, int W, void* stream);

/* ---- eval glue next to the hot path -------------------------------------------------
 * Full-reference metrics of Base_model.test (models/base/base_model.py:304-327 -> models/base/metrics.py: psnr :39-48,
 * sam :22-35, ergas :166-182) on the device, in float64 like the reference: pred / gt are normalised NCHW tensors
 * [N,B,H,W]; they are de-normalised with max_value = 2**bit_depth - .5 (dataset/utils.py:252-263) first.
 * out_dev receives N x {PSNR, SAM, ERGAS} doubles (device memory). */
int lgteun_op_metrics(lgteun_t* ctx, const float* pred, const float* gt, double* out_dev, int N, int H, int W,
                      float max_value, void* stream);

/* data_normalize (dataset/utils.py:232-248): out[i] = raw[i] / max_value, max_value = 2**bit_depth - .5; n elements,
 * device pointers (in place allowed). */
int lgteun_op_normalize(lgteun_t* ctx, const float* raw, float* out, int64_t n, float max_value, void* stream);

/* torch2np + data_denormalize (models/base/utils.py:28-39, dataset/utils.py:252-263): NCHW [N,C,H,W] -> NHWC
 * [N,H,W,C] times `scale` (max_value, or 1 to only change the layout); C in {1, 4, 8}; device pointers, no aliasing. */
int lgteun_op_to_nhwc(lgteun_t* ctx, const float* nchw, float* nhwc, int N, int C, int H, int W, float scale,
                      void* stream);

/* ---- training step (BASELINE.json configs[4]) ----------------------------------------------------
 * UnlgFormer.train_iter (models/unlg_former.py:87-113): output = G(lr, pan) in train() mode, rec_loss = nn.L1Loss
 * (models/base/losses.py:19-40) * loss_cfg.rec_loss.w, zero_grad / backward / Adam.step (base_model.py:116-131).
 *
 * Parameters and gradients live in ONE flat fp32 device buffer each, laid out as the handle's weight table:
 * weight i occupies [lgteun_weight_offset(i), +lgteun_weight_numel(i)) of lgteun_flat_numel() floats (offsets are
 * 16-byte aligned; padding floats are never read and get zero gradient).  Data-parallel training all-reduces the flat
 * gradient once per step.  As in the reference only prior_module[K-1] reaches the output
 * (unlg_former.py:63-67): the other priors are not executed and their gradient is zero (torch: .grad is None). */
int64_t lgteun_flat_numel(const lgteun_t* ctx);
int64_t lgteun_weight_offset(const lgteun_t* ctx, int i);

/* Bytes of the activation tape + backward scratch for one step of this size (allocated lazily by train_forward). */
int64_t lgteun_train_workspace_bytes(lgteun_t* ctx, int N, int h, int w);

/* Pansharpening.forward in train() mode (models/unlg_former.py:50-67 with nn.Dropout(dropout_p) after the mixer
 * projection, LGT.py:198,216; dropout_p = 0 gives the eval-mode function).  Reads the weights from flat_param (NOT from
 * lgteun_load_weights' snapshot) and records the activations the backward needs inside the handle.  The keep mask of
 * LGB block `layer` (0,1 encoder; 2 bottleneck; 3,4 decoder) is a counter-based hash of (seed, layer, NHWC element
 * index); lgteun_dropout_mask returns the same mask (values 0 or 1/(1-p)) for parity tests. */
int lgteun_train_forward(lgteun_t* ctx, const float* flat_param, const float* ms, const float* pan, float* out,
                         int N, int h, int w, float dropout_p, uint64_t seed, void* stream);

/* loss.backward() for the tape of the last lgteun_train_forward: dout [N,B,4h,4w] is dLoss/dOutput; flat_grad
 * (lgteun_flat_numel floats) is OVERWRITTEN with dLoss/dParameters (zero_grad + backward).  flat_param, ms and pan of the
 * forward must still be valid and unchanged.  One backward per forward. */
int lgteun_train_backward(lgteun_t* ctx, const float* dout, float* flat_grad, void* stream);

/* A handle keeps ONE activation tape.  lgteun_train_generation returns the serial number of the last lgteun_train_forward
 * (0 before the first); lgteun_train_backward_of runs the backward only if the tape still is that forward's and fails with
 * LGTEUN_ESTATE otherwise (forward A, forward B, backward A would silently mix A's dout with B's activations). */
uint64_t lgteun_train_generation(const lgteun_t* ctx);
int lgteun_train_backward_of(lgteun_t* ctx, uint64_t generation, const float* dout, float* flat_grad, void* stream);

/* Kernels launched by the last train_forward + train_backward pair. */
int lgteun_train_launches(const lgteun_t* ctx);

/* ReconstructionLoss('l1') (models/base/losses.py:29,39) times `weight`: loss_dev[0] = weight * mean|out - gt|;
 * dout (may be NULL) = its gradient wrt out.  n = element count. */
int lgteun_l1_loss(lgteun_t* ctx, const float* out, const float* gt, int64_t n, float weight, float* loss_dev,
                   float* dout, void* stream);

/* torch.optim.Adam.step (base_model.py:121-122; no weight decay / amsgrad) fused over a flat buffer: grad is multiplied
 * by grad_scale first (1/world_size after a summing all-reduce); step counts from 1. */
int lgteun_adam_step(lgteun_t* ctx, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                     float lr, float beta1, float beta2, float eps, int step, float grad_scale, void* stream);

int lgteun_dropout_mask(lgteun_t* ctx, uint64_t seed, int layer, float p, float* mask_out, int64_t n, void* stream);

/* Replay a recorded step: masks = five device pointers (NHWC [N,H,W,c] of each LGB block in execution order, values 0 or
 * 1/(1-p)) used by the following train_forward / train_backward pairs instead of the generated masks; NULL switches back. */
int lgteun_train_set_masks(lgteun_t* ctx, const float* const* masks);

/* CUDA-graph support of the training step: with seed_dev != NULL the dropout kernels of the following train_forward /
 * train_backward calls read their seed from *seed_dev (one uint64 in device memory) when they execute, instead of the by-value
 * seed of lgteun_train_forward, so a captured fwd + loss + bwd can be replayed with the seed of each step (the reference
 * draws a fresh nn.Dropout mask per iteration, LGT.py:198,216).  NULL switches back. */
int lgteun_train_set_seed_ptr(lgteun_t* ctx, const uint64_t* seed_dev);

/* ---- companion operators -----------------------------------------------------------------------
 * SFIIN.Freprocess.forward(msf, panf) (models/SFIIN.py:210-236), the FFT amplitude / phase fusion of a network the
 * reference repository ships beside LGTEUN: rfft2 of the two pre-convolved maps, amp_fuse / pha_fuse perceptrons per
 * spectrum bin, |irfft2|, post conv.  msf, panf, out: NCHW [N,C,H,W] device tensors; C in {4, 8, 16} (SFIIN: 8), H and W
 * powers of two in [8, 1024].  weights = 14 device pointers in state_dict order: pre1.weight, pre1.bias, pre2.weight,
 * pre2.bias, amp_fuse.0.weight [C,2C,1,1], amp_fuse.0.bias, amp_fuse.2.weight, amp_fuse.2.bias, pha_fuse.0.weight,
 * pha_fuse.0.bias, pha_fuse.2.weight, pha_fuse.2.bias, post.weight, post.bias.  No handle: the operator keeps no state;
 * the caller provides `workspace` (device memory of at least lgteun_op_freprocess_workspace_bytes bytes).  The first call on a
 * device fills a twiddle table and synchronises `stream` once: make one call before capturing the operator into a CUDA graph. */
int64_t lgteun_op_freprocess_workspace_bytes(int N, int C, int H, int W);
int lgteun_op_freprocess(int device, const float* msf, const float* panf, float* out, int N, int C, int H, int W,
                         const float* const* weights, float* workspace, int64_t workspace_bytes, void* stream);

/* PanFormer's WindowAttention.forward(x, y=None) (models/common/modules.py:341-422; built by SwinBlock :425-455 with
 * dim 64, 4 heads of 16, window 4: models/panformer.py:22): optional cyclic shift, bias-free q/k/v projections (self
 * attention: y = NULL, w_q = to_qkv.weight, w_kv = to_qkv.weight + heads*head_dim*dim; cross attention: q from y with
 * to_q.weight, k/v from x with to_kv.weight), window attention with the relative position table [2ws-1, 2ws-1]
 * (relative_pos_embedding != 0) or a dense [ws^2, ws^2] bias, the two -inf masks [ws^2, ws^2] on the last window row /
 * column of a shifted block, to_out (weight [dim, heads*head_dim], bias), shift back.  x, y, out: [b, n_h, n_w, dim]
 * device tensors (channels last, as the reference passes them); scale = head_dim ** -0.5.
 * window_size must be 4, head_dim in {8, 16, 32}, dim % 4 == 0, dim and heads*head_dim <= 128 with the
 * projection weights (4 * dim * heads*head_dim floats) resident in shared memory (PanFormer's 64 x 64: 93 KB of 227 KB). */
int lgteun_op_window_attention(int device, const float* x, const float* y, float* out, int b, int n_h, int n_w, int dim, int heads,
                               int head_dim, int window_size, int shifted, int relative_pos_embedding, float scale,
                               const float* w_q, const float* w_kv, const float* w_out, const float* b_out,
                               const float* pos_embedding, const float* upper_lower_mask, const float* left_right_mask,
                               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LGTEUN_H_ */
