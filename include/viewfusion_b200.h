/*
 * viewfusion_b200 — C ABI of the B200-native (sm_100a) ViewFusion hot path.
 *
 * This is the drop-in boundary: everything the reference computes on the path
 *     ViewFusion.forward / generate / p_sample / p_mean_variance   (model/view_fusion.py:86-300)
 *     UNet.forward and the blocks it calls                          (model/unet.py:114-303)
 * is reachable through the entry points below with plain pointers and sizes.  The reference has no FFI of
 * its own (it is pure PyTorch); the host side that binds these symbols is `view_fusion_b200/_lib.py` (ctypes),
 * and INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns 0 on success or a negative vf_status; the message is in vf_last_error()
 *    (thread-local).  Nothing throws, prints or exits.
 *  - every pointer is a DEVICE pointer unless the name ends in `_host`.  The library never allocates or
 *    frees device memory: buffers, workspaces and packed weights are owned by the caller (PyTorch).
 *  - every launch is asynchronous on the `stream` argument (a cudaStream_t), performs no host sync and is
 *    CUDA-graph capturable.
 *  - activations inside the library are NHWC ("pixel-major") in fp32 (VF_F32, CUDA-core reference-precision
 *    mode) or bf16 (VF_BF16, tcgen05 tensor-core mode); the reference's NCHW fp32 tensors are converted at
 *    the edges by vf_pack_views / vf_nhwc_to_nchw.
 *  - two row orders exist for an activation of `images` x H x W pixels with C channels:
 *      FLAT   : row m = (img*H + y)*W + x,                          images*H*W rows
 *      PADDED : row p = img*P + (y+1)*(W+1) + (x+1), P = (H+1)*(W+1), images*P rows; rows with y+1 == 0 or
 *               x+1 == 0 are padding.  One zero row above each image and one zero column left of each row serve
 *               as the 3x3 halo of BOTH neighbours (the row after the last pixel of a line is the next line's
 *               padding), so the nine taps of a 3x3 convolution are nine constant row offsets
 *               (kh-1)*(W+1) + (kw-1) of ONE tensor and a tile of it can be loaded once for all taps.
 *    Producers of convolution inputs (vf_gn_apply, vf_upsample2x) write exact zeros into the padding rows;
 *    convolution outputs leave their padding rows unwritten (they are never read as data).
 */
#ifndef VIEWFUSION_B200_H_
#define VIEWFUSION_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VF_ABI_VERSION 2

typedef enum vf_status {
  VF_OK = 0,
  VF_ERR_ARG = -1,    /* bad shape / alignment / null pointer */
  VF_ERR_CUDA = -2,   /* a CUDA runtime or driver call failed */
  VF_ERR_ARCH = -3,   /* device is not sm_100 */
  VF_ERR_STATE = -4   /* call order (e.g. forward before weights were packed) */
} vf_status;

typedef enum vf_dtype { VF_F32 = 0, VF_BF16 = 1 } vf_dtype;

typedef void* vf_stream; /* cudaStream_t */

const char* vf_last_error(void);
int vf_abi_version(void);
/* 0 when the current device is sm_100 (B200); VF_ERR_ARCH / VF_ERR_CUDA otherwise. */
int vf_device_check(void);

/* ------------------------------------------------------------------------------------------------------
 * UNet description — mirrors the constructor of model/unet.py:9-21 (dropout is always 0 in the reference
 * configs and is not supported; with_noise_level_emb is always true).
 * ---------------------------------------------------------------------------------------------------- */
#define VF_MAX_LEVELS 8
typedef struct vf_unet_config {
  int in_channel;
  int out_channel;
  int inner_channel;
  int norm_groups;
  int n_mults;
  int channel_mults[VF_MAX_LEVELS];
  int n_attn_res;
  int attn_res[VF_MAX_LEVELS];
  int res_blocks;
  int image_size;
} vf_unet_config;

typedef struct vf_unet vf_unet; /* opaque host-side plan; owns no device memory */

/* Builds the layer table of unet.py:38-112 for `cfg`. `act_dtype` selects the arithmetic:
 * VF_F32 = fp32 activations + CUDA-core FFMA GEMMs (reference-precision mode, 1e-4 bar),
 * VF_BF16 = bf16 activations + tcgen05/TMEM implicit-GEMM (performance mode, 1e-2 bar). */
int vf_unet_create(const vf_unet_config* cfg, int act_dtype, vf_unet** out);
void vf_unet_destroy(vf_unet* u);

/* Parameter table in the reference's state_dict order (no prefix). */
int vf_unet_num_params(const vf_unet* u);
/* name is written NUL-terminated into name_buf; shape gets up to 4 dims, *ndim their count. */
int vf_unet_param_info(const vf_unet* u, int index, char* name_buf, int name_cap, int64_t shape[4], int* ndim);

/* Total channel count of the per-block additive embeddings (sum of Cout over all ResnetBlocks). */
int vf_unet_emb_channels(const vf_unet* u);

/* Packed (GEMM-ready) weights: K-major [Cout][tap][Cin] in the activation dtype, fused fp32 biases, the
 * concatenated embedding matrix.  `params` is an array of vf_unet_num_params() device pointers to the fp32
 * master tensors in table order.  Re-run after every optimizer step / state_dict load. */
size_t vf_unet_packed_bytes(const vf_unet* u);
int vf_unet_pack_weights(vf_unet* u, const float* const* params_host, void* packed, vf_stream stream);

/* Scratch for one forward over up to `max_images` view-images (activations are not recycled, so the same
 * workspace doubles as the activation stash of the training backward). */
size_t vf_unet_workspace_bytes(const vf_unet* u, int max_images);

/* UNet.forward (unet.py:114-138) on pre-packed input.
 *   x0       : [images*H*W, K0] activation dtype, the im2col'd first-layer input written by vf_pack_views
 *              (K0 = vf_unet_k0()).
 *   level/angle : [rows] fp32 — the noise level ("time") and target angle of each embedding row
 *   img_row  : [images] int32 — embedding row used by each view-image (views of one sample share a row)
 *   out      : [images*H*W, 8] fp32, channels 0..out_channel-1 valid (NHWC, padded to 8)
 */
int vf_unet_k0(const vf_unet* u);
int vf_unet_act_dtype(const vf_unet* u);   /* VF_F32 / VF_BF16 given to vf_unet_create */
int vf_unet_forward(vf_unet* u, const void* packed, void* workspace, size_t workspace_bytes, int images,
                    const void* x0, const float* level, const float* angle, int rows, const int* img_row,
                    float* out, vf_stream stream);

/* Arena layout capacity.  vf_unet_forward / vf_unet_backward size every buffer of their bump-allocated workspaces for
 * max(capacity, images) view-images, so with a workspace of vf_unet_workspace_bytes(u, capacity) bytes the position of every
 * buffer — and with it the zero padding rows the caller established by zero-filling the workspace once — stays put when the
 * batch size or the view counts change from call to call (the reference draws view_count per step, experiment.py:277).
 * The backward workspace (vf_unet_backward_workspace_bytes) follows the layout of the last forward. */
int vf_unet_set_capacity(vf_unet* u, int max_images);
/* Monotonic counter bumped by every vf_unet_forward on this plan: vf_unet_backward differentiates the LAST forward, so a
 * caller that stashed a forward checks this before running its backward. */
unsigned long long vf_unet_forward_generation(const vf_unet* u);

/* Training (default, on = 1): the forward also writes what only vf_unet_backward reads (the transposed V of every
 * attention block).  Inference (on = 0): those extras are skipped — the attention kernel takes V row-major from the qkv
 * tensor — and vf_unet_backward after such a forward returns VF_ERR_STATE.  Activations and results are identical. */
int vf_unet_set_stash(vf_unet* u, int on);

/* Number of kernels enqueued by the last vf_unet_forward on this plan (bench.py's gpu_launches claim). */
int vf_unet_last_launches(const vf_unet* u);

/* Per-kernel-class device timing of one forward (CUDA events around every launch).  Off by default; used only for
 * the roofline breakdown in bench.py.  ms[6] / counts[6] = conv, gn_stats, gn_apply, attention, upsample, embed. */
int vf_unet_set_profiling(vf_unet* u, int on);
int vf_unet_profile_read(vf_unet* u, float* ms_host, int* counts_host);
/* Per-launch table of the same profiled forward, in launch order: ms[i], kind[i] (class index as above) and
 * desc[6*i..6*i+5] = images, H, C_in (conv: total K), C_out, ksize, stride (zeros where not recorded).
 * Returns the number of launches written (<= cap), negative on error. */
int vf_unet_profile_launches(vf_unet* u, float* ms_host, int* kind_host, int* desc_host, int cap);

/* Debug/parity tap: copies the output activation of module `name` ("downs.3", "mid.0", "ups.17", ...) of the
 * LAST vf_unet_forward on this plan into `dst` as NCHW fp32.  dst must hold images*C*H*W floats. */
int vf_unet_read_tap(vf_unet* u, const void* workspace, const char* name, float* dst, int64_t* chw, vf_stream stream);

/* ------------------------------------------------------------------------------------------------------
 * View stacking (view_fusion.py:95-115 / :244-263) fused with the first layer's im2col.
 *   y_cond      : (B, Nmax, Cc, H, W) fp32 NCHW;   y_t : (B, 3, H, W) fp32
 *   view_offset : [B+1] int32 prefix sum of view_count;  only the first V_b views of sample b are used
 *   x0          : [images*H*W, K0] in `x0_dtype`, k = (kh*3+kw)*(Cc+3) + c, zero padded (pad 1 halo and K tail)
 *   img_sample  : [images] int32 out — sample index of each stacked view (the `img_row` of vf_unet_forward)
 * ---------------------------------------------------------------------------------------------------- */
int vf_pack_views(const float* y_cond, const float* y_t, const int* view_offset, int B, int n_max, int cond_channels,
                  int H, int W, int images, int k0, int x0_dtype, void* x0, int* img_sample, vf_stream stream);

/* Generic NCHW fp32 (R,C,H,W) -> im2col'd first-layer input, for the stand-alone UNet.forward API. */
int vf_pack_nchw(const float* x, int R, int C, int H, int W, int k0, int x0_dtype, void* x0, vf_stream stream);

/* [R*H*W, ld] fp32 NHWC -> (R, C, H, W) fp32 NCHW (first C channels). */
int vf_nhwc_to_nchw(const float* src, int ld, int R, int C, int H, int W, float* dst, vf_stream stream);

/* q_sample (view_fusion.py:162-164): y = sqrt(g_b) * y0 + sqrt(1 - g_b) * noise, g per sample. */
int vf_q_sample(const float* y0, const float* noise, const float* gammas, int B, int chw, float* y, vf_stream stream);

/* ------------------------------------------------------------------------------------------------------
 * Composition + DDPM update (view_fusion.py:116-160 + :70-84 + :166-177), one pass over HBM.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct vf_schedule {       /* the six (T,) fp32 buffers of set_new_noise_schedule (view_fusion.py:50-68) */
  const float* gammas;
  const float* sqrt_recip_gammas;
  const float* sqrt_recipm1_gammas;
  const float* posterior_log_variance_clipped;
  const float* posterior_mean_coef1;
  const float* posterior_mean_coef2;
  int num_timesteps;
} vf_schedule;

typedef struct vf_compose_args {
  const float* unet_out;     /* [images*H*W, 8] fp32: eps in 0..2, logits in 3..5 */
  const int* view_offset;    /* [B+1] */
  const int* t;              /* [B] int32 time-step of each sample */
  const float* y_t;          /* (B,3,H,W) fp32 NCHW */
  float* y_prev;             /* (B,3,H,W) out; may alias y_t */
  const float* z;            /* (B,3,H,W) injected N(0,1) draw, or NULL -> Philox(seed, offset) in-kernel */
  uint64_t seed;
  uint64_t offset;
  int add_noise;             /* the reference's `any(t > 0)` (view_fusion.py:176): 0 / 1 decided by the caller, 2 = read from `step` */
  int clip_denoised;
  int weighting;             /* 1: softmax over views (6-channel UNet); 0: plain mean (ablation, :141-150) */
  int B, H, W;
  float* eps_out;            /* optional (B,3,H,W): the composed noise prediction */
  float* weights_out;        /* optional (B,max_v,3,H,W): softmax weights, zero in padded slots */
  int max_v;
  float* logits_out;         /* optional (images,3,H,W): un-padded logits (p_sample's 2nd return value) */
  const struct vf_step_record* step; /* optional DEVICE record written by vf_step_prepare: when set, add_noise == 2 takes the
                                * `any(t > 0)` decision from step->any_t_positive, the Philox offset from step->noise_offset and
                                * (seed == 0) the seed from step->seed
                                * (nothing about the step is baked into the launch: CUDA-graph replays advance on the device) */
} vf_compose_args;

/* Device-resident state of a reverse loop (view_fusion.py:196-206), so that one captured CUDA graph can be replayed for every
 * step: vf_step_prepare reads t_state[b], publishes the step's constants and (advance != 0) decrements t_state / bumps the
 * noise counter for the next replay. */
typedef struct vf_step_record {
  int any_t_positive;              /* the reference's `any(t > 0)` over the batch (view_fusion.py:176) */
  int reserved;
  unsigned long long noise_offset; /* Philox offset of this step's draw */
  unsigned long long seed;         /* Philox seed of the loop (copied from noise_ctr[1]) */
} vf_step_record;

/*   t_state [B] int32 (in/out), noise_ctr [2] u64 = {Philox offset (in/out), seed} (may be NULL), gammas [T] ->
 *   level [B] = gammas[t] (the UNet's noise-level input, view_fusion.py:98), t_cur [B] = t, rec = the step record */
int vf_step_prepare(int* t_state, int B, const float* gammas, int num_timesteps, int advance, unsigned long long* noise_ctr,
                    float* level, int* t_cur, vf_step_record* rec, vf_stream stream);

/* One whole reverse step, enqueued on `stream` without any host decision (graph-capturable): vf_step_prepare -> vf_pack_views ->
 * vf_unet_forward -> vf_compose_ddpm_step.  This is ViewFusion.p_sample (view_fusion.py:166-177) over a batch whose time-steps
 * live on the device. */
typedef struct vf_sample_step_args {
  const void* packed;              /* vf_unet_pack_weights output */
  void* workspace;                 /* vf_unet_workspace_bytes(u, capacity) bytes */
  size_t workspace_bytes;
  const float* y_cond;             /* (B, n_max, cond_channels, H, W) fp32 */
  int B, n_max, cond_channels, H, W, images;
  const int* view_offset;          /* [B+1] int32 */
  const float* angle;              /* [B] fp32 */
  const float* y_t;                /* (B,3,H,W) */
  float* y_prev;                   /* (B,3,H,W); may alias y_t (in-place reverse loop) */
  int* t_state;                    /* [B] int32 device */
  int advance;                     /* != 0: t_state[b] = max(t-1, 0) and ++noise_ctr after use */
  unsigned long long* noise_ctr;   /* [2] device {offset, seed}, or NULL with z */
  uint64_t seed;                   /* != 0: Philox seed by value; 0: the device-resident noise_ctr[1] (replayed graphs) */
  const float* z;                  /* injected draw or NULL */
  int add_noise;                   /* 0 never, 1 always, 2 iff any t > 0 (decided on the device) */
  int clip_denoised, weighting;
  void* x0; int* img_sample; float* unet_out;      /* staging: [images*H*W, K0] act dtype, [images], [images*H*W, 8] fp32 */
  float* level; int* t_cur; vf_step_record* rec;   /* [B], [B], [1] device scratch */
  float* eps_out; float* weights_out; float* logits_out; int max_v;   /* optional outputs as in vf_compose_args */
  vf_schedule sched;
} vf_sample_step_args;
int vf_p_sample_step(vf_unet* u, const vf_sample_step_args* a, vf_stream stream);

int vf_compose_ddpm_step(const vf_compose_args* a, const vf_schedule* s, vf_stream stream);

/* Training tail (view_fusion.py:265-298): composition + MSE loss + gradient w.r.t. the UNet output.
 *   loss_acc : [1] fp32, must be zero on entry; receives mean((noise - eps_hat)^2)
 *   grad_out : [images*H*W, 8] fp32 or NULL: dL/d(unet_out) scaled by `grad_scale` */
int vf_compose_mse(const float* unet_out, const int* view_offset, const float* noise, int B, int H, int W,
                   int weighting, float* loss_acc, float* eps_out, float* grad_out, float grad_scale,
                   vf_stream stream);

/* ------------------------------------------------------------------------------------------------------
 * Stage-level operators (used by the plan above; exported for unit parity tests and ncu captures).
 * All activations [images*H*W, C] NHWC in `dtype`.
 * ---------------------------------------------------------------------------------------------------- */

/* PositionalEncoding + noise_level_mlp + all per-block Linear(inner->Cout) (unet.py:115-116,:147-157,:160-177).
 *   w0 [4*ic, ic], b0, w2 [ic, 4*ic], b2 : the MLP;  emb_w [E, ic], emb_b [E]: concatenated block linears
 *   out [rows, E] fp32 */
int vf_embed(const float* level, const float* angle, int rows, int inner_channel, const float* w0, const float* b0,
             const float* w2, const float* b2, const float* emb_w, const float* emb_b, int E, float* out,
             vf_stream stream);

/* GroupNorm statistics: per (image, channel) sum and sum of squares over the H*W pixels of the channel-concatenation
 * of src0 [.,C0] and src1 [.,C1] (src1 may be NULL), both PADDED.  stats [images, C0+C1, 2] fp32 must be zero on entry.
 * (The plan does not use this pass: convolution epilogues emit the same sums, vf_conv_args::stats.) */
int vf_gn_stats(const void* src0, int C0, const void* src1, int C1, int dtype, int images, int H, int W, float* stats,
                vf_stream stream);

/* GroupNorm(groups, eps=1e-5, affine) + optional Swish (unet.py:207-218,:254): PADDED sources -> PADDED dst [., C0+C1]
 * with exact zeros in the padding rows.
 * statsX: per (image, channel) {sum, sum of squares} of source X, rows of statsX_ld channels ([images, ld, 2]);
 * they come from vf_gn_stats or from the producing convolution's epilogue (vf_conv_args::stats). */
typedef struct vf_gn_shift {   /* source 0 is stored WITHOUT the per-(image, channel) constant  s[img][c] = bias[c] + emb[img_row[img]][c]  */
  const float* bias;           /* [C0] or NULL */
  const float* emb;            /* [rows, emb_ld] (column offset already applied) or NULL */
  const int* img_row;          /* [images], needed with emb */
  int emb_ld;
} vf_gn_shift;
/* `shift` (may be NULL): the normalisation is applied to src0 + s without s ever being added in memory — the statistics of
 * src0 (raw sums) are shifted in closed form, sum(x+s) = S1 + HW*s, sum((x+s)^2) = S2 + 2*s*S1 + HW*s^2, and s folds into the
 * per-channel FMA.  The plan uses it for ResnetBlock.block1's bias and the FeatureWiseAffine add (unet.py:242-243, :176),
 * whose only consumer is block2's GroupNorm: the convolution epilogue then adds nothing. */
int vf_gn_apply(const void* src0, int C0, const float* stats0, int stats0_ld, const void* src1, int C1,
                const float* stats1, int stats1_ld, int dtype, int images, int H, int W, int groups, const float* gamma,
                const float* beta, int swish, void* dst, const vf_gn_shift* shift, vf_stream stream);

/* Nearest-neighbour x2 up-sampling (unet.py:188): PADDED (H, W) -> PADDED (2H, 2W), zero padding rows. */
int vf_upsample2x(const void* src, int dtype, int images, int H, int W, int C, void* dst, vf_stream stream);

/* Writes zeros into the padding rows of a PADDED tensor (needed when a convolution OUTPUT is consumed directly by a
 * 3x3 convolution: Downsample, unet.py:195-201). */
int vf_zero_padding(void* dst, int dtype, int images, int H, int W, int C, vf_stream stream);

/* Row-order conversions (tests, taps): FLAT <-> PADDED for [., C] matrices in `dtype`. */
int vf_flat_to_padded(const void* src, int dtype, int images, int H, int W, int C, void* dst, vf_stream stream);
int vf_padded_to_flat(const void* src, int dtype, int images, int H, int W, int C, void* dst, vf_stream stream);

/* Implicit-GEMM convolution, out[m, n] = sum_seg sum_tap sum_c A_seg[pix(m, tap), c] * Wt[n, koff + tap*C_seg + c]
 * with the fused epilogue  + bias[n] + emb[img_row[img(m)]*emb_ld + n] + residual[m, n].
 * ksize 3 uses padding 1; stride 2 only with ksize 3 (Downsample, unet.py:195-201). */
typedef struct vf_conv_args {
  int dtype;                 /* VF_F32 -> CUDA-core kernel, VF_BF16 -> tcgen05 kernel */
  int images, H, W;          /* spatial size of the SOURCES (the output is H/2 x W/2 when stride == 2) */
  int in_padded;             /* row order of every source: 1 PADDED, 0 FLAT (FLAT only with ksize 1) */
  int out_padded;            /* row order of out / residual / stats rows.  1x1 layers may change the order: PADDED -> FLAT
                              * gathers the valid pixels through the source's tensor map, FLAT -> PADDED stores whole image
                              * lines through the output's; padding rows of a PADDED output are never written */
  int n_seg;                 /* 1..3 K-segments accumulated into the same output tile */
  const void* src[3];        /* [rows, src_c[i]] */
  int src_c[3];
  int ksize[3];              /* 1 or 3 (only segment 0 may be 3) */
  int stride;                /* 1, or 2 = Downsample (3x3, single segment).  The tcgen05 path computes the H/2 x W/2 output
                              * pixels only: TMA gathers the pixel phases of the PADDED source and supplies the zero halo
                              * by out-of-bounds fill (padding rows of the source are not read) when W/2 divides 64;
                              * otherwise, and on the CUDA-core path, it is computed at source resolution with the even
                              * pixels kept and the source's padding rows must hold zeros */
  const void* weight;        /* [Cout_pad, K_total] K-major in `dtype`; K_total = sum ksize^2 * src_c */
  int cout;                  /* logical output channels */
  int cout_pad;              /* rows of `weight` (multiple of 16) */
  const float* bias;         /* [cout] or NULL */
  const float* emb;          /* [rows, emb_ld] or NULL; column offset already applied */
  const int* img_row;        /* [images] */
  int emb_ld;
  const void* residual;      /* [out rows, cout] in `dtype` or NULL (same row order as out) */
  void* out;                 /* [out rows, out_ld] */
  int out_dtype;             /* `dtype`, or VF_F32 for the final layer */
  int out_ld;
  int qkv_split;             /* >0: C of an attention block; columns [2C,3C) are written transposed to out_vt */
  void* out_vt;              /* [images, C, H*W] (keys contiguous) when qkv_split (needs out_padded == 0) */
  float* stats;              /* optional [images, cout, 2]: GroupNorm partial sums of the OUTPUT (pre-zeroed) */
} vf_conv_args;

int vf_conv2d(const vf_conv_args* a, vf_stream stream);

/* Test hooks (host arithmetic only, usable without a device): the row splits per image the GroupNorm forward /
 * backward kernels would be launched with for `images` images of H x W x C (small layers are scheduled by
 * a latency model).  Return the split count (threads per CTA through threads_out), negative for unsupported shapes. */
int vf_debug_gn_splits(int images, int H, int W, int C, int dtype, int* threads_out);
/* Same for the tcgen05 convolution: the host-side plan for `a` (pointers are not dereferenced, nothing is launched).
 * out[16] = block_n, G (128-row accumulators per item), A stages, B stages, weights resident, dynamic smem bytes, TMEM
 * columns, work items, grid, staged TMA epilogue, gathered-source mode (W), line-map epilogue (W), stride-2 gather chunks
 * (Cin/64), GEMM rows, bias-table images, K. */
int vf_debug_conv_tiling(const vf_conv_args* a, int* out16);
int vf_debug_gn_bwd_splits(int images, int H, int W, int C, int dtype);
/* Same for the one-pass GroupNorm backward: channels per CTA (a slab of whole groups; 0 = the layer takes the two-pass kernels,
 * negative = unsupported shape), CTA threads and dynamic shared memory through the out parameters. */
int vf_debug_gn_bwd_slab(int C, int groups, int H, int W, int dtype, int* threads_out, int* smem_out);

/* Test hook: route VF_BF16 vf_conv2d / vf_attention through the CUDA-core kernels (same bf16 storage, fp32
 * accumulation) so the tcgen05 kernels can be cross-checked on the device.  Never enabled by the product path. */
void vf_debug_force_simt(int on);
/* Test hook.  Bits 0-3: ablations of the tcgen05 conv epilogue (timing studies only; results are wrong when set).  Bits with
 * correct results, used by the tests to reach the alternative code path: 0x100 print the plan, 0x200 / 0x1000 no resident / pinned
 * weight tiles, 0x400 no TMA line maps, 0x800 no closed-form bias column sums, 0x2000 two-pass GroupNorm backward even where the
 * one-pass kernel applies, 0x4000 64-row activation boxes only (no short last box); bits 16-27 force (G, block_n). */
void vf_debug_flags(int flags);
/* Test hook: device buffer [148*4] int64 receiving the tcgen05 conv's MMA-thread cycle counters (NULL = off). */
void vf_debug_counters(long long* dev_buf);

/* Single-head self-attention core (unet.py:267-274): O = softmax(Q K^T / sqrt(C)) V per image.
 *   qk  : [images*L, 3C] rows hold q in [0,C), k in [C,2C) (v columns unused when vt != NULL)
 *   vt  : [images, C, L] V transposed, or NULL: v is read row-major from columns [2C,3C) of qk (the tensor-core path
 *         then uses it as an MN-major operand; inference mode, nothing is written for the backward)
 *   out : [images*L, C]
 *   lse : optional [images*L] fp32, written by the tensor-core path for vf_attention_backward: log2 of the softmax
 *         denominator in the scaled domain, P = exp2(s * log2(e)/sqrt(C) - lse).  NULL when not training. */
int vf_attention(const void* qk, const void* vt, int dtype, int images, int L, int C, void* out, float* lse, vf_stream stream);

/* fp32 OIHW conv weight -> K-major [cout_pad][k_off + tap*cin + c] rows of length k_total in `dtype`. */
int vf_pack_conv_weight(const float* w_oihw, int cout, int cin, int ksize, int dtype, void* dst, int cout_pad,
                        int k_total, int k_off, vf_stream stream);

/* ------------------------------------------------------------------------------------------------------
 * Training: backward of the LAST vf_unet_forward on this plan (experiment.py:292 loss.backward()).
 *   packed_t       : transposed weight packs for the data gradients (vf_unet_packed_t_bytes, refreshed by
 *                    vf_unet_pack_weights_t after every optimizer step)
 *   grad_workspace : activation gradients + scratch (vf_unet_backward_workspace_bytes, valid after a forward); like the
 *                    forward workspace it must be zero-filled once per (plan, image count): padding rows are never
 *                    written and are read as zeros
 *   grad_out8      : [images*H*W, 8] fp32 gradient of vf_unet_forward's `out` (vf_compose_mse's grad_out)
 *   param_grads    : host array of vf_unet_num_params() device pointers (fp32, parameter shapes); ACCUMULATED into
 * ---------------------------------------------------------------------------------------------------- */
size_t vf_unet_packed_t_bytes(const vf_unet* u);
int vf_unet_pack_weights_t(vf_unet* u, void* packed_t, vf_stream stream);
size_t vf_unet_backward_workspace_bytes(vf_unet* u);
int vf_unet_backward(vf_unet* u, const void* packed_t, void* grad_workspace, size_t grad_workspace_bytes, const float* grad_out8,
                     float* const* param_grads_host, vf_stream stream);

/* ------------------------------------------------------------------------------------------------------
 * Backward operators (training: autograd through model/unet.py, experiment.py:292).  Gradients of activations
 * use the activation dtype and row order of the tensor they belong to; parameter gradients are fp32.
 * ---------------------------------------------------------------------------------------------------- */

/* Weight gradient of the convolution described by `fwd` (same struct as the forward call; weight/bias/out unused):
 *   dwp[n][koff(seg) + tap*C_seg + c] += sum_rows dy[out_row(m)][n] * src_seg[m + shift(tap)][c]
 * dwp: [cout_pad, K_total] fp32, the forward weight's K-major packed layout (accumulated, zero it first);
 * dy: gradient of the forward output, [out rows, dy_ld] in fwd->dtype with zero padding rows. */
int vf_conv2d_wgrad(const vf_conv_args* fwd, const void* dy, int dy_ld, float* dwp, vf_stream stream);

/* The same backward in n_phases + 1 calls (phase = 0 .. n_phases, in order, same arguments), for overlapping the data-parallel
 * gradient exchange with the rest of the backward (reference: DistributedDataParallel's bucketed all-reduce, experiment.py:104-107).
 * The reversed tape is cut into n_phases runs of roughly equal parameter bytes; after phase k returns, every gradient whose
 * parameter has param_phase[i] == k is completely enqueued on `stream`; the embedding MLP and the per-block Linear parameters
 * (gradients complete only at the very end) have param_phase == n_phases and are produced by the last call.
 * vf_unet_backward_plan (after a forward: it follows the recorded tape) fills param_phase[vf_unet_num_params()]. */
int vf_unet_backward_plan(vf_unet* u, int n_phases, int* param_phase_out);
int vf_unet_backward_phase(vf_unet* u, const void* packed_t, void* grad_workspace, size_t grad_workspace_bytes, const float* grad_out8,
                           float* const* param_grads_host, int phase, int n_phases, vf_stream stream);

/* dw_oihw[n][c_off + c][tap] += dwp[n][k_off + tap*cin + c]: packed gradient of one K-segment -> its channel slice of an
 * OIHW parameter gradient with cin_total input channels. */
int vf_unpack_conv_wgrad(const float* dwp, int cout, int cin, int ksize, int k_total, int k_off, float* dw_oihw, int cin_total,
                         int c_off, vf_stream stream);

/* Transposed, tap-flipped pack: dst[c][k_off + (taps-1-tap)*n_stride + n] = w[n][c][tap] (zero for c >= cin, n >= cout), so that
 * the DATA gradient of a convolution is a forward vf_conv2d over dy with this weight (rows = cin_pad, K = taps*n_stride). */
int vf_pack_conv_weight_t(const float* w_oihw, int cout, int cin, int ksize, int dtype, void* dst, int cin_pad, int k_total, int k_off,
                          int n_stride, vf_stream stream);

/* GroupNorm(+Swish) backward for dst = vf_gn_apply(src0 | src1): dy [images*P, C0+C1] PADDED.
 * dxK receives (accK == 0) or accumulates (accK != 0) the gradient of source K; dgamma/dbeta [C0+C1] fp32 are
 * accumulated; scratch: images*(C0+C1)*2 floats.  Two implementations behind this entry point: when an (image, slab of whole
 * groups) of x and dy fits in shared memory, ONE kernel reads x and dy once, keeps them on the SM between the reduction and
 * the apply phase and writes dx (3 tensor passes); otherwise a reduce and an apply kernel (6 passes).  With swish != 0 the
 * buffer behind dy MAY BE CONSUMED (the two-pass path overwrites it with dy * swish'(z) so that its second pass is a pure
 * FMA stream).  `shift` as in vf_gn_apply (dx0 is the gradient w.r.t. src0 + s, which equals the gradient w.r.t. src0). */
int vf_gn_backward(const void* src0, int C0, const float* stats0, int stats0_ld, const void* src1, int C1, const float* stats1,
                   int stats1_ld, int dtype, int images, int H, int W, int groups, const float* gamma, const float* beta, int swish,
                   const void* dy, float* scratch, float* dgamma, float* dbeta, void* dx0, int acc0, void* dx1, int acc1,
                   const vf_gn_shift* shift, vf_stream stream);

/* Backward of vf_attention: d_out [images*L, C] -> dqkv [images*L, 3C] (dq | dk | dv).
 *   out, lse : the forward results (tensor-core path: bf16, L in {128, 256}, C in {128, 192}; otherwise unused / NULL
 *              and the CUDA-core kernel recomputes the softmax)
 *   scratch  : images*L*2C*max(1, L/128) floats */
int vf_attention_backward(const void* qk, const void* vt, const void* out, const float* lse, const void* d_out, int dtype, int images,
                          int L, int C, float* scratch, void* dqkv, vf_stream stream);

/* Backward of vf_upsample2x: dx (PADDED H x W) = or += sum of the 2x2 block of dy (PADDED 2H x 2W). */
int vf_upsample2x_backward(const void* dy, int dtype, int images, int H, int W, int C, void* dx, int accumulate, vf_stream stream);
/* dst (PADDED 2H x 2W) = dy (PADDED H x W) at even pixels, zero elsewhere: turns the stride-2 Downsample gradients into
 * stride-1 problems at the source resolution. */
int vf_zero_insert2x(const void* dy, int dtype, int images, int H, int W, int C, void* dst, vf_stream stream);
int vf_add_inplace(void* dst, const void* src, int dtype, size_t n_elems, vf_stream stream);
/* [rows, 8] fp32 (vf_compose_mse's grad_out) -> [rows, ld] activation dtype, zero beyond channel 8. */
int vf_grad8_to_act(const float* g8, size_t rows, int dtype, int ld, void* dst, vf_stream stream);

/* Bias / embedding gradients of one convolution (autograd of unet.py:214 bias and :176 FeatureWiseAffine add): per-image column
 * sums of dy [images*rows_per_img, ld] (padding rows must hold zeros), first `cout` columns:
 *   db0[n], db1[n] += sum_rows dy[row][n]   (either may be NULL);   demb[img_row[img]][col + n] += sum_rows(img) dy[row][n]  (or NULL). */
int vf_colsum_bias(const void* dy, int dtype, int images, int rows_per_img, int ld, int cout, float* db0, float* db1, float* demb,
                   const int* img_row, int emb_ld, int col, vf_stream stream);

/* Backward of vf_embed (autograd of unet.py:115-116, 27-32, 165-176): demb [rows, E] = gradient of the embedding table.
 * dew [E, ic] and deb [E] are OVERWRITTEN with the gradients of the concatenated per-block Linear weights / biases; the
 * noise_level_mlp gradients dw0 [4ic, ic], db0 [4ic], dw2 [ic, 4ic], db2 [ic] are ACCUMULATED.  rowbuf: rows*11*ic floats. */
int vf_embed_backward(const float* level, const float* angle, int rows, int inner_channel, const float* w0, const float* b0,
                      const float* w2, const float* b2, const float* emb_w, int E, const float* demb, float* rowbuf, float* dew,
                      float* deb, float* dw0, float* db0, float* dw2, float* db2, vf_stream stream);

/* ---- optimizer step (reference: torch.optim.Adam of experiment.py:115-120, stepped at :293) -------------
 * One launch for all parameter tensors.  table_dev: one entry per tensor (fp32 device pointers); chunk_entry_dev /
 * chunk_start_dev: for every vf_adam_chunk_elems()-sized chunk the entry index and the first element.  `step` is the
 * 1-based step count used for the bias corrections (host scalar, as in torch's non-capturable Adam). */
typedef struct vf_adam_entry {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int numel;
  int reserved;
} vf_adam_entry;
int vf_adam_chunk_elems(void);
int vf_adam_step(const vf_adam_entry* table_dev, const int* chunk_entry_dev, const int* chunk_start_dev, int n_chunks, double lr,
                 double beta1, double beta2, double eps, double weight_decay, int step, vf_stream stream);

/* ------------------------------------------------------------------------------------------------------
 * Evaluation metrics on the device (utils/metrics.py:6-12, called from experiment.py:349-356): PSNR and SSIM (the
 * algorithm of pytorch_msssim.ssim(..., data_range=1.0, size_average=False): 11-tap Gaussian, sigma 1.5, VALID windows)
 * of generated vs target, both (B, C, H, W) fp32 NCHW.  psnr / ssim: [B] fp32, overwritten.  H, W >= 11; the two planes
 * and their five filtered copies must fit in shared memory (up to ~94 x 94 pixels).
 * ---------------------------------------------------------------------------------------------------- */
int vf_eval_metrics(const float* generated, const float* target, int B, int C, int H, int W, float* psnr, float* ssim,
                    vf_stream stream);

/* Input path (data/nmr_dataset.py:10-52, `process_sample`, for a whole batch): views (B, V, H, W, C) uint8 HWC as
 * decoded from the dataset, perm (B, V) int32 = the per-object view permutation (np.random.shuffle in the reference).
 *   target (B, C, H, W) = views[b, perm[b,0]] / 255;  cond (B, V-1, C, H, W) = views[b, perm[b,1:]] / 255;
 *   angle (B) = 2 pi / V * perm[b,0]        — all fp32, NCHW, bit-exact with the host computation. */
int vf_prepare_batch_u8(const uint8_t* views, const int* perm, int B, int V, int C, int H, int W, float* target, float* cond,
                        float* angle, vf_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* VIEWFUSION_B200_H_ */
