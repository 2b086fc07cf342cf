/* heatnet_b200.h -- C ABI of libheatnet_b200.so: the B200 (sm_100a) kernels behind HeatNet's
 * dense-segmentation hot path.
 *
 * The reference (jzuern/heatnet-pub) has no FFI: every op below is a PyTorch library call made from
 * Python modules.  Each entry point names the reference call site(s) it replaces (paths relative to the
 * reference root; "cm/" = models/confusion_maximization/).  The Python host in heatnet_pub_b200/ binds
 * these with ctypes; no torch types cross this boundary.
 *
 * Conventions
 *  - every function returns 0 on success and a negative code on failure; hn_last_error() returns a
 *    thread-local message for the last failure on the calling thread.
 *  - all device pointers are borrowed for the duration of the call; the library allocates no device
 *    memory.  Kernels are enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises.
 *  - activations are NHWC ("pixel-major") views: element (n,h,w,c) lives at
 *        ptr + (((n*H + h)*W + w) * ld + c) * sizeof(dtype)
 *    where `ld` >= C is the channel count of the underlying buffer, so a view can be a channel slice of
 *    a wider buffer (this is how torch.cat(dim=1) in cm/models/extractors.py:192-196 and
 *    cm/models/pspnet.py:24 becomes zero-copy).
 */
#ifndef HEATNET_B200_H
#define HEATNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { HN_F32 = 0, HN_BF16 = 1 };
enum { HN_ACT_NONE = 0, HN_ACT_RELU = 1, HN_ACT_LEAKY = 2 /* y = x>=0 ? x : slope*x (LeakyReLU, PReLU) */ };
enum {
    HN_OK = 0,
    HN_ERR_ARG = -1,      /* invalid argument / unsupported shape */
    HN_ERR_CUDA = -2,     /* CUDA runtime or driver error */
    HN_ERR_WORKSPACE = -3 /* workspace too small */
};

typedef struct hn_tensor {
    void *ptr;     /* device pointer to element (0,0,0,0) of the view */
    int32_t dtype; /* HN_F32 or HN_BF16 */
    int32_t n, h, w, c;
    int32_t ld;    /* channel stride of the underlying NHWC buffer (>= c) */
} hn_tensor;

/* Fused conv epilogue: y = act(acc * scale[k] + shift[k] + residual).  Replaces the separate
 * BatchNorm2d(eval) / bias / `out += residual` / ReLU / PReLU / LeakyReLU launches of
 * cm/models/extractors.py:85-101, cm/models/pspnet.py:25,32-34, cm/discriminator_model.py:51-59. */
typedef struct hn_epilogue {
    const float *scale;     /* [Cout] or NULL (1); the BF16 engine is fastest with scale == NULL (fold it: hn_pack_weight_scaled) */
    const float *shift;     /* [Cout] or NULL (0) */
    const void *residual;   /* NHWC view with the output's dtype/shape, or NULL */
    int32_t residual_ld;
    int32_t act;            /* HN_ACT_* */
    float slope;            /* used when slope_ptr == NULL */
    const float *slope_ptr; /* device scalar (PReLU weight), optional */
    int32_t out_nchw;       /* 1: y is written as NCHW FP32 (the reference's logits layout) */
    double *stat_sum;       /* optional [Cout] FP64, caller-zeroed: += sum over the output pixels of the stored value ... */
    double *stat_sqsum;     /* ... and of its square: BatchNorm2d batch statistics in train mode, fused into the conv epilogue
                             * (BF16 engine, 16-byte-aligned output view, Cout tile >= 32; FP32 or BF16 output -- BF16: statistics of
                             * the ROUNDED values, i.e. of the stored tensor, and no explicit scale; else the call fails) */
    int32_t per_image;      /* hn_affine_act only: scale (and shift, if given) are [N][C] -- Dropout2d channel masks */
    int32_t stat_groups;    /* 0 / 1: one set of statistics.  G > 1: the batch is G equal groups of consecutive images with separate
                             * statistics (a day and a night batch convolved as one): stat_sum / stat_sqsum are [G][Cout] */
} hn_epilogue;

typedef struct hn_conv {
    int32_t cout, r, s, stride, pad, dil;
} hn_conv;

const char *hn_last_error(void);
int hn_version(void);
/* SM count / cc of the current device; fails (HN_ERR_CUDA) without a usable sm_100 device. */
int hn_device_check(void);
/* tuning aid: per-role wait-cycle counters of the tcgen05 conv kernel; returns 0 (all zeros) unless the library was built
 * with -DHN_PROFILE_ROLES (scripts/profile_roles.py builds such a copy) */
int hn_prof_read(unsigned long long *out16, int reset);

/* ---- layout + parameter preparation (no reference counterpart: the reference keeps NCHW/OIHW) ---- */
/* NCHW FP32 (user tensors) -> NHWC view. */
int hn_nchw_to_nhwc(const float *src, const hn_tensor *dst, void *stream);
/* NHWC view -> dense NCHW FP32. */
int hn_nhwc_to_nchw(const hn_tensor *src, float *dst, void *stream);
/* OIHW FP32 conv weight -> K-major pack [cout_pad][kpad], k = (r*S + s)*Cin + c, zero padded.
 * dtype HN_BF16 (tensor-core path) or HN_F32 (FP32 parity path). */
int hn_pack_weight(const float *w_oihw, void *dst, int32_t dtype, int32_t cout, int32_t cin, int32_t r, int32_t s,
                   int32_t cout_pad, int32_t kpad, void *stream);
/* same, with every output-channel row multiplied by row_scale[o] first: BatchNorm2d(eval) folded into the filter, which
 * leaves only the per-channel shift for the conv epilogue (row_scale == NULL: plain pack) */
int hn_pack_weight_scaled(const float *w_oihw, const float *row_scale, void *dst, int32_t dtype, int32_t cout, int32_t cin,
                          int32_t r, int32_t s, int32_t cout_pad, int32_t kpad, void *stream);
/* BatchNorm2d(eval) + conv bias folded to per-channel scale/shift:
 * scale = gamma/sqrt(var+eps), shift = beta + (bias - mean)*scale.  gamma==NULL => plain bias. */
int hn_bn_fold(const float *gamma, const float *beta, const float *mean, const float *var, const float *bias,
               float eps, float *scale, float *shift, int32_t c, void *stream);

/* ---- convolution (replaces nn.Conv2d forward: cm/models/extractors.py:71-76,111-123,
 *      cm/models/pspnet.py:13,18,32,57, cm/discriminator_model.py:40-44) ---- */
/* rows of the packed weight matrix: FP32 pack -> multiple of 64; BF16 pack -> 16/32/64 or a multiple of 64 */
int32_t hn_conv_cout_pad(int32_t cout, int32_t dtype);
int32_t hn_conv_kpad(int32_t cin, int32_t r, int32_t s);
/* bytes of scratch hn_conv2d_fwd needs for this problem (0 for the implicit-GEMM cases) */
int64_t hn_conv2d_workspace_bytes(const hn_tensor *x, const hn_conv *cv);
/* BF16 x / BF16 packed weights -> tcgen05 implicit GEMM (FP32 accumulate in TMEM);
 * FP32 x / FP32 packed weights -> FP32 CUDA-core implicit GEMM (the 1e-4 parity path). */
int hn_conv2d_fwd(const hn_tensor *x, const void *w_packed, const hn_conv *cv, const hn_epilogue *ep,
                  const hn_tensor *y, void *workspace, int64_t workspace_bytes, void *stream);

/* PSPUpsample without the upsampled tensor (cm/models/pspnet.py:37-40): y = epilogue(conv3x3_pad1(bilinear_2x(x))).
 * The 2x-upsampled halo patch of every tile is computed in shared memory by producer warps (exact ATen bilinear rule)
 * and consumed by tcgen05 straight from there.  BF16, Cin % 64 == 0, w_packed = hn_pack_weight of the 3x3 filter. */
int hn_upconv3x3_fwd(const hn_tensor *x, const void *w_packed, const hn_conv *cv, const hn_epilogue *ep, const hn_tensor *y,
                     void *stream);

/* PSPUpsample's 3x3 conv and the PSPNet classifier in ONE kernel (cm/models/pspnet.py:72-75 in eval mode: up_3.conv ->
 * drop_2 (identity) -> final 1x1): logits[n][k][h][w] (NCHW FP32, the reference's layout) = head_b[k] + sum_c head_w[k][c] *
 * act(conv3x3(x)[c] + shift[c]).  The classifier runs in the epilogue on the FP32 activations; the 64-channel
 * full-resolution tensor is never written.  BF16 x, Cin % 64 == 0, Cout == 64, head_classes <= 16.
 * head_w_host [classes][64] / head_b_host [classes] (or NULL) are HOST pointers: the classifier travels as kernel parameters
 * (constant bank), so its FMAs need no loads; they are read during the call and need not outlive it. */
int hn_conv3x3_head_fwd(const hn_tensor *x, const void *w_packed, const hn_conv *cv, const hn_epilogue *ep, const float *head_w_host,
                        const float *head_b_host, int32_t head_classes, float *logits_nchw, void *stream);

/* The 7x7 stride-2 pad-3 stems (Cin <= 4; cm/models/extractors.py:111-123) without an im2col pass: hn_stem_pad writes the
 * input as a zero-bordered 4-channel BF16 image [N][>=2*Ho+6][>=2*Wo+6][4]; in it the 8x4 values a filter row needs for one
 * output pixel are contiguous and consecutive windows overlap, which a strided tensor map hands to TMA directly.
 * hn_pack_stem_weight: OIHW -> [64][7][8][4] (optionally scaled per output channel). */
int hn_stem_pad(const hn_tensor *x, const hn_tensor *xpad, void *stream);
/* hn_stem7x7s2_fwd feeds tcgen05 straight from dense rows of the padded image (the overlapping windows are expressed by the
 * operand descriptor's strides, hn_conv_stem.cu) and reads whole 2112-byte row segments: the allocation behind xpad must
 * extend at least this many bytes past the last row (contents irrelevant). */
int64_t hn_stem_pad_slack_bytes(void);
int hn_pack_stem_weight(const float *w_oihw, const float *row_scale, void *dst, int32_t cout, int32_t cin, void *stream);
int hn_stem7x7s2_fwd(const hn_tensor *xpad, const void *w_packed, int32_t cout, const hn_epilogue *ep, const hn_tensor *y, void *stream);

/* ---- bandwidth-bound ops ---- */
/* nn.MaxPool2d(3,2,1): cm/models/extractors.py:128 */
int hn_maxpool3x3s2_fwd(const hn_tensor *x, const hn_tensor *y, void *stream);
/* the four nn.AdaptiveAvgPool2d(s), s in sizes[0..nsizes), in one pass over x (cm/models/pspnet.py:17).
 * out (x's dtype): one dense NHWC block [N][s][s][C] per size, concatenated in the order of `sizes`,
 * i.e. size a starts at element N * C * sum_{b<a} s_b^2.  workspace: FP32 row partials. */
int64_t hn_pyramid_pool_workspace_bytes(const hn_tensor *x, const int32_t *sizes, int32_t nsizes);
int hn_pyramid_pool_fwd(const hn_tensor *x, const int32_t *sizes, int32_t nsizes, void *out, void *workspace,
                        int64_t workspace_bytes, void *stream);
/* F.upsample(mode='bilinear') == interpolate(align_corners=False): cm/models/pspnet.py:23,39,
 * cm/discriminator_model.py:47.  x and y may differ in dtype. */
int hn_bilinear_fwd(const hn_tensor *x, const hn_tensor *y, void *stream);
/* y = sum_i upsample(xs[i]) for up to 4 low-resolution sources, FP32 accumulation, one write: the PSP priors after the
 * bottleneck projection (a 1x1 conv commutes with bilinear upsampling -- cm/models/pspnet.py:23-25 restructured) */
int hn_bilinear_sum_fwd(const hn_tensor *xs, int32_t nsrc, const hn_tensor *y, void *stream);
/* per-channel affine + residual + activation over an NHWC view (BatchNorm2d apply in train mode; with
 * per_image=1 and shift=NULL the nn.Dropout2d channel mask of cm/models/pspnet.py:64-73):
 * y = act(x*scale[c] + shift[c] + residual) */
int hn_affine_act(const hn_tensor *x, const hn_epilogue *ep, const hn_tensor *y, void *stream);
/* nn.Dropout2d channel masks (cm/models/pspnet.py:49,55,64-73) as the [N*C] scale vector hn_affine_act(per_image=1) applies:
 * scale_out[i] = u_i >= p ? 1/(1-p) : 0 (all zeros for p == 1, like torch), u_i = 24-bit uniform from Philox4x32-10 keyed by
 * state_dev[0] (seed) with counter (state_dev[1], i).  The kernel increments state_dev[1] itself, so every call -- including every
 * replay of a captured CUDA graph -- draws a new mask, and a given seed reproduces the whole sequence. */
int hn_dropout2d_scale(uint64_t *state_dev, int64_t n, float p, float *scale_out, void *stream);
/* per-channel sum / sum of squares over all pixels (BatchNorm2d batch statistics), FP64 accumulators */
int hn_channel_stats(const hn_tensor *x, double *sum, double *sqsum, void *stream);
/* train-mode BatchNorm2d finalize (cm/models/extractors.py:72-77): from FP64 sums over `count` pixels
 * computes scale/shift for the apply pass and updates running_mean/var (momentum, unbiased var). */
int hn_bn_finalize(const double *sum, const double *sqsum, int64_t count, const float *gamma, const float *beta,
                   float eps, float momentum, float *running_mean, float *running_var, float *scale, float *shift,
                   float *save_mean, float *save_invstd, int32_t c, void *stream);
/* same, and `num_batches_tracked += 1` in the same launch (the tail of a conv with fused statistics) */
int hn_bn_finalize_tracked(const double *sum, const double *sqsum, int64_t count, const float *gamma, const float *beta,
                           float eps, float momentum, float *running_mean, float *running_var, int64_t *num_batches_tracked,
                           float *scale, float *shift, float *save_mean, float *save_invstd, int32_t c, void *stream);
/* hn_bn_finalize_tracked and the normalise pass hn_affine_act in ONE launch (cm/models/extractors.py:85-101 in train mode):
 * y = act(x * scale + shift + residual) with scale / shift formed from the FP64 sums in every thread's prologue; the vectors the
 * backward needs (scale, shift, save_mean, save_invstd: [C] each) are written and the running statistics / counter updated by
 * the threads of the first pixel chunk.  ep supplies residual / act / slope and stat_groups: with G > 1 groups of consecutive
 * images, sum / sqsum and the four output vectors are [G][C], count = pixels per group, and the running statistics receive G
 * momentum updates in group order (what G consecutive forward calls of the reference do).  Views must be 8-channel aligned. */
int hn_bn_apply_train(const hn_tensor *x, const double *sum, const double *sqsum, int64_t count, const float *gamma, const float *beta,
                      float eps, float momentum, float *running_mean, float *running_var, int64_t *num_batches_tracked,
                      const hn_epilogue *ep, const hn_tensor *y, float *scale, float *shift, float *save_mean, float *save_invstd,
                      void *stream);
/* train-mode BatchNorm2d statistics in ONE kernel (+ one memset): hn_channel_stats and hn_bn_finalize fused through a
 * last-CTA ticket, including `num_batches_tracked += 1` (cm/models/extractors.py:85-101: every nn.BatchNorm2d in train mode).
 * scratch: device memory of hn_bn_batch_stats_scratch_bytes(C) (FP64 sums + ticket), zeroed by the call. */
int64_t hn_bn_batch_stats_scratch_bytes(int32_t c);
int hn_bn_batch_stats(const hn_tensor *x, void *scratch, const float *gamma, const float *beta, float eps, float momentum,
                      float *running_mean, float *running_var, int64_t *num_batches_tracked, float *scale, float *shift,
                      float *save_mean, float *save_invstd, void *stream);

/* ---- input pipeline (cm/thermal_loader.py:649-659,715-728; cm/train_trgb_segnet_conf.py:82-86): the loaders' per-pixel
 * normalisation after the H2D copy instead of before it.  Both functions gather from host-built lookup tables (the reference's
 * own arithmetic evaluated once per possible input value), so the outputs are bit-identical to the CPU path.
 * rgb_nhwc: uint8 [N][H][W][3]; lut_3x256: device float [3][256] = ((v / 255) - mean[c]) / std[c]; dst: 3-channel NHWC view. */
int hn_prepare_rgb_u8(const uint8_t *rgb_nhwc, const float *lut_3x256, const hn_tensor *dst, void *stream);
/* ir_nhw: uint16 (src_bits 16) or int32 (src_bits 32) [N][H][W]; values are clipped to [minval, maxval];
 * lut: device float [maxval - minval + 1] = (((k / (maxval - minval)) - mean) / std); dst: 1-channel NHWC view. */
int hn_prepare_ir(const void *ir_nhw, int32_t src_bits, int32_t minval, int32_t maxval, const float *lut, const hn_tensor *dst, void *stream);
/* rectDropTensor: x[n, :, r0:r0+dh, c0:c0+dw] = 0, params_dev int32 [N][4] = (r0, c0, dh, dw), Python slice clamping. */
int hn_rect_drop(const hn_tensor *x, const int32_t *params_dev, void *stream);

/* IR intensity augmentation (cm/train_trgb_segnet_conf.py:101-110 smartAugment, :394-410 ir_scale_aug), in place:
 * factors_dev == NULL: x *= scale.  Otherwise x[n, :, p] *= factors_dev[labels_dev[n, p]] (labels int64 [N][H][W], factors float [k]);
 * labels outside [0, k) leave the pixel unchanged and set bit 0 of *flags_dev (optional). */
int hn_label_scale(const hn_tensor *x, const int64_t *labels_dev, const float *factors_dev, int32_t k, float scale, int32_t *flags_dev,
                   void *stream);

/* ---- backward (autograd of the ops above; the reference gets these from torch.autograd / cuDNN) ---- */
/* dz = dout * act'(out), the derivative taken through the saved output (conv + bias + activation layers) */
int hn_act_bwd(const hn_tensor *dout, const hn_tensor *out, int32_t act, float slope, const hn_tensor *dz, void *stream);
/* BatchNorm2d (train) backward fused with the activation / residual split:
 *   dz = dout*act'(out);  sums[0..C) = sum dz (= dbeta), sums[C..2C) = sum dz*xhat (= dgamma), sums[2C] = dPReLU slope
 *   draw = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat));  dres (+)= dz.   raw: pre-normalisation tensor (FP32, or BF16
 *   on the BF16 path).  dbeta / dgamma [C] and dslope [1] (each optional) receive the FP32 parameter gradients directly,
 *   added to their current contents when param_accumulate is set (a second backward pass into an existing .grad).
 *   fwd_scale / fwd_shift (both or neither): the scale/shift vectors the forward normalise pass used.  When given -- only valid
 *   for a layer WITHOUT a residual input -- `out` is not read: the pre-activation z = fma(raw, scale, shift) is recomputed
 *   bit-identically, which removes one tensor from both passes.
 *   groups > 1: the batch is `groups` equal groups of consecutive images normalised with SEPARATE statistics (hn_epilogue.stat_groups):
 *   mean / invstd / fwd_scale / fwd_shift are [groups][C], sums is [groups][2C+1]; parameter gradients are summed over the groups. */
int hn_bn_bwd(const hn_tensor *dout, const hn_tensor *out, const hn_tensor *raw, const float *mean, const float *invstd,
              const float *gamma, int32_t act, float slope, const float *slope_ptr, double *sums, const hn_tensor *draw,
              const hn_tensor *dres, int32_t dres_accumulate, int32_t want_prelu_grad, float *dbeta, float *dgamma, float *dslope,
              int32_t param_accumulate, const float *fwd_scale, const float *fwd_shift, int32_t groups, void *stream);
/* y = x (accumulate == 0) or y += x; dtypes may differ */
int hn_accumulate(const hn_tensor *x, const hn_tensor *y, int32_t accumulate, void *stream);
/* MaxPool2d(3,2,1) forward that also records the winning tap (uint8 [N][Ho][Wo][C]) and its backward */
int hn_maxpool3x3s2_fwd_idx(const hn_tensor *x, const hn_tensor *y, uint8_t *idx, void *stream);
int hn_maxpool3x3s2_bwd(const hn_tensor *dy, const uint8_t *idx, const hn_tensor *dx, int32_t accumulate, void *stream);
/* adjoints of hn_bilinear_fwd / hn_pyramid_pool_fwd (gather form, deterministic) */
/* adjoint of hn_bilinear_fwd: exact 2x (PSPUpsample) in 2x2 micro-tiles; large magnifications (PSP priors, the critics' x32
 * maps) by a separable two-pass reduction through an FP32 workspace of hn_bilinear_bwd_workspace_bytes() (0 = not needed;
 * without it the generic gather form runs); deterministic, no atomics */
int64_t hn_bilinear_bwd_workspace_bytes(const hn_tensor *dy, const hn_tensor *dx);
int hn_bilinear_bwd(const hn_tensor *dy, const hn_tensor *dx, int32_t accumulate, void *workspace, int64_t workspace_bytes, void *stream);
int hn_pyramid_pool_bwd(const void *dpool, const int32_t *sizes, int32_t nsizes, const hn_tensor *dx, int32_t accumulate, void *stream);
/* conv dgrad = hn_conv2d_fwd of the (zero-inserted, for stride > 1) output gradient with the flipped/transposed pack */
int hn_dilate(const hn_tensor *x, int32_t stride, const hn_tensor *up, void *stream);
int hn_pack_weight_dgrad(const float *w_oihw, void *dst, int32_t dtype, int32_t cout, int32_t cin, int32_t r, int32_t s,
                         int32_t cin_pad, int32_t kpad, void *stream);
/* Stride-2 dgrad WITHOUT the zero-inserted gradient (BF16 engine): the input gradient splits into the four parity phases
 * (rho_y, rho_x) = (phase >> 1, phase & 1) of dX; each is a stride-1 correlation of dY with the sub-filter of the taps that reach
 * that parity, written straight onto its sub-lattice of dX through a strided tensor map (k*k tap evaluations per output pixel quad
 * instead of 4*k*k).  hn_pack_weight_dgrad_phase builds a phase's [cin_pad][kpad] pack, kpad = hn_dgrad_s2_phase_kpad(...) (0: no
 * tap reaches the phase -- it is zero-filled, or skipped when accumulating).  cv = the FORWARD geometry (stride 2, dilation 1);
 * w_phase: 4 device pointers (NULL for empty phases).  hn_conv2d_dgrad_s2_ok: 1 when the views qualify, else use hn_dilate. */
int32_t hn_dgrad_s2_phase_kpad(int32_t cout, int32_t r, int32_t s, int32_t pad, int32_t phase);
int hn_pack_weight_dgrad_phase(const float *w_oihw, void *dst, int32_t dtype, int32_t cout, int32_t cin, int32_t r, int32_t s, int32_t pad,
                               int32_t phase, int32_t cin_pad, void *stream);
int hn_conv2d_dgrad_s2_ok(const hn_tensor *dy, const hn_conv *cv, const hn_tensor *dx);
int hn_conv2d_dgrad_s2(const hn_tensor *dy, const void *const *w_phase, const hn_conv *cv, const hn_tensor *dx, int32_t accumulate,
                       void *stream);
/* Every weight pack a training step needs, rebuilt from the FP32 masters in ONE launch (the optimizer invalidates ~170 packs per
 * step: forward, dgrad and stride-2 phase packs of every live convolution).  The caller fills one hn_pack_job per pack with
 * hn_pack_job_init (kind 0 = hn_pack_weight, 1 = hn_pack_weight_dgrad, 2 = hn_pack_weight_dgrad_phase layout), copies the array
 * to the device and passes a block map of (job index, chunk index) int32 pairs, chunk c covering destination elements
 * [c * hn_pack_chunk(), (c+1) * hn_pack_chunk()).  Results are bit-identical to the single-tensor pack functions. */
typedef struct hn_pack_job {
    const float *src; /* OIHW FP32 master (device) */
    void *dst;        /* packed destination (device) */
    int32_t cout, cin, r, s;
    int32_t rows_pad, kpad;
    int32_t kind, dtype;
    int32_t ty, tx, phiy, phix; /* kind 2: taps and first filter tap per axis of the phase */
} hn_pack_job;
int32_t hn_pack_chunk(void);
int hn_pack_job_init(hn_pack_job *job, const float *w_oihw, void *dst, int32_t dtype, int32_t kind, int32_t cout, int32_t cin, int32_t r,
                     int32_t s, int32_t pad, int32_t phase);
int hn_pack_weights_multi(const hn_pack_job *jobs_dev, const int32_t *block_map_dev, int32_t n_blocks, void *stream);
/* conv wgrad into the packed FP32 layout [cout_pad][kpad] (BF16: tcgen05 with MN-major operands, split over pixels;
 * FP32: CUDA cores), then packed -> OIHW parameter gradient */
int64_t hn_conv2d_wgrad_workspace_bytes(const hn_tensor *x, const hn_conv *cv);
int hn_conv2d_wgrad(const hn_tensor *x, const hn_tensor *dy, const hn_conv *cv, float *dw_packed, int32_t zero_init, void *workspace,
                    int64_t workspace_bytes, void *stream);
int hn_unpack_wgrad(const float *packed, float *grad_oihw, int32_t cout, int32_t cin, int32_t r, int32_t s, int32_t kpad,
                    int32_t accumulate, void *stream);
/* FP64 per-channel sums -> FP32 parameter gradient (bias / gamma / beta / PReLU slope) */
int hn_vec_to_grad(const double *src, float *grad, int32_t n, int32_t accumulate, void *stream);

/* ---- metric: scripts/iou_eval.py:53-88,154-159 (max(1) + np.bincount(pred + K*target) on the host) ---- */
/* pred_labels: int64 [n_pixels] class ids, or NULL when `scores` (NCHW FP32 [N][K][HW]) is given, in which
 * case the first-max argmax over K is fused.  conf: device int64 [K*K], row = target, col = predicted,
 * ACCUMULATED into.  flags: device int32 [1], bit0 set if any pred outside [0,K), bit1 if any target is
 * (the reference's range asserts, iou_eval.py:58-79); out-of-range pixels are not counted. */
int hn_confusion(const int64_t *pred_labels, const float *scores, int64_t n_images, int64_t hw,
                 const int64_t *target, int32_t k, int64_t *conf, int32_t *flags, void *stream);
/* argmax over classes of NCHW FP32 scores -> uint8 / int64 label map (validation_bdd_mf.py:331) */
int hn_argmax_labels(const float *scores, int64_t n_images, int64_t hw, int32_t k, uint8_t *out_u8,
                     int64_t *out_i64, void *stream);

/* ---- losses of the training step, forward + backward in one pass ----
 * cm/train_trgb_segnet_conf.py:244 `criterion_semseg = CrossEntropyLoss()` applied at :452; scripts/main.py:223
 * `CrossEntropyLoss(ignore_index=13)`; cm/train_trgb_segnet_conf.py:238-240 `BCEWithLogitsLoss()` / `MSELoss()` of a critic
 * map against `torch.full_like(c, 1 or 0)` at :437-446,529-546.  All reductions are 'mean'.
 * scratch: device memory of at least hn_loss_scratch_bytes() (FP64 per-block partials; deterministic fold). */
int64_t hn_loss_scratch_bytes(void);
/* logits NCHW FP32 [N][K][HW], labels int64 [N][HW].  loss_out: device float.  dlogits (optional, same layout) receives
 * grad_scale * d loss / d logits.  Labels equal to ignore_index are skipped (pass a value outside [0,K) for "none");
 * other labels outside [0,K) set bit 0 of *flags (torch raises a device assert there) and are skipped. */
int hn_ce_loss_fwd_bwd(const float *logits_nchw, const int64_t *labels, int64_t n_images, int32_t k, int64_t hw, int64_t ignore_index,
                       float grad_scale, float *loss_out, float *dlogits_nchw, int32_t *flags, void *scratch, int64_t scratch_bytes,
                       void *stream);
/* kind 0: mean((x - t)^2); kind 1: mean(BCE-with-logits(x, t)); t = target[i] or, when target == NULL, target_const (the
 * reference fills its target with a constant: torch.full_like(c, 1 / 0)).  dx (optional) = grad_scale * d loss / dx */
int hn_critic_loss_fwd_bwd(const float *x, const float *target, float target_const, int64_t n, int32_t kind, float grad_scale,
                           float *loss_out, float *dx, void *scratch, int64_t scratch_bytes, void *stream);
/* y *= *scalar_dev (chain rule with the upstream gradient of a loss term, kept on the device) */
int hn_scale_by_scalar(float *y, int64_t n, const float *scalar_dev, void *stream);

/* ---- optimiser step: cm/train_trgb_segnet_conf.py:270 `torch.optim.RMSprop(params, lr)`,
 * scripts/main.py:159 `Adam(params, lr)`, scripts/main.py:256-257 `clip_grad_norm(params, clip)` ----
 * One launch updates every tensor.  slots_dev: device array, one entry per parameter tensor (FP32, dense).
 * block_map_dev: device int32 pairs (slot index, chunk index); chunk c of a slot covers elements
 * [c*hn_optim_chunk(), (c+1)*hn_optim_chunk()).  The caller builds both tables (they change only when a tensor moves). */
typedef struct hn_param_slot {
    float *param;
    const float *grad;
    float *state1; /* RMSprop: square_avg.            Adam: exp_avg    */
    float *state2; /* RMSprop: momentum_buffer / NULL. Adam: exp_avg_sq */
    int64_t numel;
} hn_param_slot;
int32_t hn_optim_chunk(void);
/* sqnorm_dev[0] = sum over all slots of sum(grad^2) (FP64; deterministic: per-block partials in partial_dev[n_blocks]) */
int hn_grad_sqnorm(const hn_param_slot *slots_dev, const int32_t *block_map_dev, int32_t n_blocks, double *partial_dev,
                   double *sqnorm_dev, void *stream);
/* Every gradient is first multiplied by grad_scale (1/world after an all-reduce SUM) and, when max_norm > 0, by
 * min(1, max_norm / (|grad_scale| * sqrt(*sqnorm_dev) + 1e-6)) -- torch.nn.utils.clip_grad_norm_ without the extra pass. */
int hn_rmsprop_step(const hn_param_slot *slots_dev, const int32_t *block_map_dev, int32_t n_blocks, float lr, float alpha, float eps,
                    float weight_decay, float momentum, float grad_scale, float max_norm, const double *sqnorm_dev, void *stream);
/* step: 1-based step count (bias corrections 1 - beta^step) */
int hn_adam_step(const hn_param_slot *slots_dev, const int32_t *block_map_dev, int32_t n_blocks, float lr, float beta1, float beta2,
                 float eps, float weight_decay, int64_t step, float grad_scale, float max_norm, const double *sqnorm_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HEATNET_B200_H */
