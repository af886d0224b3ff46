/* terragan_b200.h — C ABI of libtg_b200.so, the sm_100a kernel library behind the TERRA-GAN
 * PConv hot path (PConv U-Net generator, conv discriminator, inpainting losses; fwd + bwd).
 *
 * The reference (FKGSOFTWARE/TERRA-GAN) has NO native/FFI layer: its hot path is the Python
 * nn.Module API of mvp_gan/src/models/{pconv,generator,discriminator}.py and
 * mvp_gan/src/utils/losses.py, which dispatches to ATen/cuDNN. Each entry point below therefore
 * names the reference *Python call sites* whose arithmetic it replaces (file:line relative to the
 * reference root). The Python mirror of those modules (terra-gan_b200/mvp_gan/...) binds these
 * symbols with ctypes (terra-gan_b200/tg_b200/_lib.py); INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer owned by the caller (PyTorch allocates everything); the
 *    library never allocates or frees user-visible memory and keeps no reference past the call.
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it, re-entrant,
 *    performs no host synchronisation and may be called from any host thread (autograd worker).
 *  - Return value: 0 on success, <0 on error; tg_last_error() returns the (thread-local) message.
 *    No C++ exception crosses the ABI.
 *  - Activations are bf16, channels-last, optionally parity-split: [B][P][H][W][C] with P = 1
 *    (plain NHWC) or P = 4 (plane 2*(h&1)+(w&1) holds pixels (h>>1, w>>1) of the full-res image;
 *    the layout stride-2 convolutions consume). Masks / window counts are uint8, one per pixel,
 *    in the same pixel order as the tensor they accompany. Parameters and gradients are fp32 in
 *    PyTorch's native layouts.
 */
#ifndef TERRAGAN_B200_H
#define TERRAGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TG_MAX_TAPS 64
#define TG_MAX_SUB 4
#define TG_ACT_NONE 0
#define TG_ACT_RELU 1
#define TG_ACT_LEAKY 2
/* Activation storage type. BF16 is the product path (bf16 storage, kind::f16 MMAs, fp32 accumulate). F32 is the
 * verification path of BASELINE.json's north star ("TF32 path <= 1e-3"): fp32 storage, kind::tf32 MMAs; the same
 * entry points with a `dtype` field, and `_f32` twins of the bandwidth kernels. */
#define TG_DTYPE_BF16 0
#define TG_DTYPE_F32 1

/* ---- library ------------------------------------------------------------------------------ */
int tg_version(void);                               /* ABI version (this header: 2) */
size_t tg_last_error(char* buf, size_t cap);        /* copies the last error text of this thread */
int tg_num_sms(void);                               /* SM count of the current device (<=0: error) */
/* Hit / miss counters of the calling thread's TMA tensor-map cache (descriptors are keyed by pointer, shape and
 * box, SURVEY.md §8b "TMA descriptor cache keyed by pointer/shape"). */
int tg_tmap_cache_stats(uint64_t* hits, uint64_t* misses);

/* ---- mask pyramid -------------------------------------------------------------------------
 * Integer restatement of mask_conv + (>0) (pconv.py:33-40) and of the decoder mask merge
 * max(nearest_up2(up_mask), skip_mask) (generator.py:50-54,68-74). Bit-exact by construction. */

/* sum[b][ho][wo] = sum of mask over the k x k window (stride s, zero padding pad), and
 * upd[...] = sum > 0. mask_in is [B][Hi][Wi] uint8 {0,1}. Outputs are [B][Ho][Wo] uint8; either
 * may be NULL. If split_out != NULL it also receives `upd` in parity-split order [B][4][Ho/2][Wo/2]
 * and if split_in_out != NULL the *input* mask in parity-split order [B][4][Hi/2][Wi/2]. */
int tg_mask_window_sum(const uint8_t* mask_in, int B, int Hi, int Wi, int k, int s, int pad,
                       uint8_t* sum, uint8_t* upd, uint8_t* upd_split, uint8_t* in_split,
                       void* stream);
/* out[b][h][w] = max(up[b][h/2][w/2], skip[b][h][w]);  up is [B][H/2][W/2], skip/out [B][H][W]. */
int tg_mask_merge_up(const uint8_t* up, const uint8_t* skip, int B, int H, int W, uint8_t* out,
                     void* stream);
/* mask_u8 = (mask_f32 > 0) — entry conversion of the [B,1,H,W] fp32 {0,1} mask (pconv.py:27). */
int tg_mask_from_f32(const float* mask, long n, uint8_t* out, void* stream);
int tg_mask_to_f32(const uint8_t* mask, long n, float* out, void* stream);

/* ---- implicit-GEMM convolution (tcgen05 + TMEM + TMA) --------------------------------------
 * Replaces nn.Conv2d forward/backward-data on the path: PConv2d.input_conv (pconv.py:30),
 * Discriminator convs idx 2,5,8 (discriminator.py:11), VGG16 features[:16] convs (losses.py:31-32)
 * and their autograd dgrad, with the PConv renormalisation (pconv.py:38-43), bias, eval-mode
 * BatchNorm (pconv.py:47) and ReLU/LeakyReLU (pconv.py:48, discriminator.py:14) fused in the
 * epilogue, and per-channel sum / sum-of-squares partials for train-mode BatchNorm.
 *
 *   out[pix][n] = act( ((sum_{tap,c} x[pix (+) tap][c] * w[n][k_off + tap*C + c]) + bias[n])
 *                      * lut[code[pix]] * scale[n] + shift[n] )
 *
 * A "tap" is (plane, dh, dw): it reads input plane `plane` at (h + dh, w + dw) of the output pixel
 * (h, w); out-of-range reads are zero (conv zero padding). Stride-2 forward convs are expressed on
 * the parity-split input (P = 4); stride-2 dgrad as num_sub = 4 sub-problems, one per input parity,
 * each with its own tap subset, weight slab (k_off) and output plane. */
typedef struct tg_conv_sub {
  int32_t tap_begin, tap_count, k_off, out_plane;
} tg_conv_sub;

typedef struct tg_conv_args {
  const void* x;            /* bf16 [B][P][H][W][C]; C % 64 == 0 */
  int32_t B, P, H, W, C;
  const void* w;            /* bf16 [N][Ktot] row-major (K contiguous); N % 64 == 0 */
  int32_t N, Ktot;
  int32_t num_sub;
  tg_conv_sub sub[TG_MAX_SUB];
  int32_t num_taps;
  int8_t tap_plane[TG_MAX_TAPS], tap_dh[TG_MAX_TAPS], tap_dw[TG_MAX_TAPS];
  void* out;                /* bf16 [B][Po][Ho][Wo][N] */
  int32_t Po, Ho, Wo;
  const uint8_t* code;      /* [B][Po][Ho][Wo] or NULL (row scale 1) */
  const float* lut;         /* host pointer, lut_len <= 64 entries; required iff code != NULL */
  int32_t lut_len;
  const float* bias;        /* device [N] or NULL */
  const float* scale;       /* device [N] or NULL */
  const float* shift;       /* device [N] or NULL */
  int32_t act;              /* TG_ACT_* */
  float slope;
  float* stats;             /* device [stats_rows_cap][2][N] or NULL */
  int32_t stats_rows_cap;   /* in: rows allocated (>= tg_num_sms()) */
  int32_t stats_rows_used;  /* out: rows written by this launch */
  const void* gate;         /* bf16, same shape as out, or NULL: out *= (gate > 0 ? 1 : gate_slope) —
                               ReLU/LeakyReLU derivative of the tensor a dgrad result flows into */
  float gate_slope;
  int32_t dtype;            /* TG_DTYPE_*: storage type of x, w, out and gate (F32: C % 32 == 0, generic kernel) */
  const float* addend;      /* F32 only, or NULL: fp32 tensor of out's shape added to the accumulator before bias / ratio /
                               statistics (tf32x3: the lo*hi + hi*lo cross terms computed by a first launch, so that the
                               long hi*hi accumulation chain is not lengthened by them) */
  void* pool_out;           /* bf16 [B][Ho/2][Wo/2][N] or NULL: also write the 2x2 max-pool of the (activated) output —
                               nn.MaxPool2d(2,2) of VGG16 features[4], [9] (losses.py:31-32) fused into the conv epilogue.
                               Only where tg_conv_pool_fusable() says so (the halo-reuse kernels) */
  int32_t skip_out;         /* with pool_out: do not store the full-resolution output (`out` may then be NULL) */
} tg_conv_args;
int tg_conv_igemm(tg_conv_args* args, void* stream);
/* 1 if tg_conv_igemm would accept `pool_out` for these arguments (shape / dtype / kernel family), else 0. */
int tg_conv_pool_fusable(const tg_conv_args* args);

/* Weight gradient of the same convolutions (autograd of pconv.py:30 / discriminator.py:11):
 *   partial[s][(tap, c)][n] = sum over the s-th share of output pixels of x[pix (+) tap][c] * g[pix][n]
 * followed by tg_wgrad_reduce which sums the shares in a fixed order and scatters into the fp32
 * NCHW-shaped .grad tensor dw[n][c][kh][kw] (accumulate != 0: dw += ...). */
typedef struct tg_wgrad_blk {  /* one 64-row block of dW^T: (tap, 64-channel block of x) */
  int8_t plane, dh, dw, pad;
  int32_t cb;                  /* channel block index (channels cb*64 .. cb*64+63) */
  int32_t row;                 /* first row in the packed [num_taps*C][N] gradient = tap*C + cb*64 */
} tg_wgrad_blk;

typedef struct tg_wgrad_args {
  const void* x;            /* bf16 [B][P][H][W][C] — the (masked) layer input saved by forward */
  int32_t B, P, H, W, C;
  const void* g;            /* bf16 [B][1][Ho][Wo][N] — gradient w.r.t. the conv output */
  int32_t Ho, Wo, N;
  int32_t num_taps;
  int8_t tap_plane[TG_MAX_TAPS], tap_dh[TG_MAX_TAPS], tap_dw[TG_MAX_TAPS];
  float* partial;           /* device [splits][num_taps*C][N] fp32 workspace */
  int64_t partial_cap;      /* in: floats available in `partial` */
  int32_t splits;           /* out: split-K factor chosen (pass to tg_wgrad_reduce) */
  const void* blks;         /* device table of num_taps*C/64 tg_wgrad_blk, tap-major then channel block */
  int32_t num_blk;
  int32_t dtype;            /* TG_DTYPE_*: storage type of x and g. F32: blks are 32-channel blocks (num_taps*C/32
                               entries, row = tap*C + cb*32), C % 32 == 0, N % 32 == 0 */
} tg_wgrad_args;
int tg_wgrad_igemm(tg_wgrad_args* args, void* stream);
/* dw[n][c][tap_of(kh,kw)] (+)= sum_s partial[s][tap*C + c][n];  tap order is the caller's tap table,
 * tap_perm[t] gives the kh*kw-flattened kernel position of tap t. */
int tg_wgrad_reduce(const float* partial, int splits, int num_taps, int C, int N,
                    const int32_t* tap_perm_dev, float* dw, int accumulate, void* stream);
/* How many fp32 the `partial` workspace of tg_wgrad_igemm needs for this problem. */
int64_t tg_wgrad_partial_floats(int B, int Ho, int Wo, int num_taps, int C, int N);
int64_t tg_wgrad_partial_floats_f32(int B, int Ho, int Wo, int num_taps, int C, int N);   /* dtype = TG_DTYPE_F32 */


/* ---- BatchNorm2d + activation (HBM-bound, 16-byte vectorised) -------------------------------
 * Replaces `self.bn` + `self.activation` of PConv2d (pconv.py:46-48) and BatchNorm2d + LeakyReLU of
 * the Discriminator blocks (discriminator.py:12-14), forward and backward. Train-mode statistics
 * come from the (sum, sumsq) partial rows written by the conv kernels. */

/* partial [rows][2][C] -> scale = gamma*invstd, shift = beta - mean*scale, saved mean / invstd and
 * (if running_mean != NULL) the running-stat update with `momentum` and the unbiased variance. */
int tg_bn_finalize(const float* partial, int rows, int C, double count, const float* gamma,
                   const float* beta, float eps, float momentum, float* running_mean,
                   float* running_var, float* scale, float* shift, float* mean, float* invstd,
                   void* stream);
/* eval mode: scale / shift from the running statistics. */
int tg_bn_eval_coeff(int C, const float* gamma, const float* beta, const float* running_mean,
                     const float* running_var, float eps, float* scale, float* shift, void* stream);
/* y = act(z*scale + shift). z is bf16 [B][H][W][C]. Writes y_nhwc (same layout) and/or y_split
 * (parity-split [B][4][H/2][W/2][C]); if mask_split, y_split is multiplied by [code[pixel] > 0]
 * (the layer's updated mask), i.e. it is the `input * mask` the next PConv2d consumes (pconv.py:27). */
int tg_bn_apply(const void* z, int B, int H, int W, int C, const float* scale, const float* shift,
                int act, float slope, const uint8_t* code, void* y_nhwc, void* y_split,
                int mask_split, void* stream);

/* A gradient arriving at a layer output: bf16, pixel p of the [B][H][W] grid at
 * ptr + pix(p)*pix_stride + chan_off, pix(p) = p (split = 0) or its parity-split index (split = 1). */
typedef struct tg_grad_src {
  const void* ptr;
  int64_t pix_stride;
  int32_t chan_off;
  int32_t split;
} tg_grad_src;
/* One pass over (g0 [+ g1], z): per-channel sums needed by BN backward and the conv-bias gradient.
 * partial is [rows_cap][5][C]; *rows_used rows are written. lut_dev: device LUT for `code`. */
int tg_bn_bwd_reduce(const tg_grad_src* g0, const tg_grad_src* g1, const void* z, int B, int H, int W,
                     int C, const float* scale, const float* shift, int act, float slope,
                     const uint8_t* code, const float* lut_dev, float* partial, int rows_cap,
                     int* rows_used, void* stream);
/* -> coeff [5][C] for tg_bn_bwd_apply, dgamma, dbeta, dbias (each may be NULL; accumulate: +=).
 * batch_stats = 1: train-mode BN (mean/var depend on the batch); 0: eval-mode BN / plain affine. */
int tg_bn_bwd_finalize(const float* partial, int rows, int C, double count, const float* scale,
                       const float* mean, const float* invstd, float* coeff, float* dgamma,
                       float* dbeta, float* dbias, int accumulate, int batch_stats, void* stream);
/* gz = lut[code] * scale * (g*act' - dbeta/M - zhat*dgamma/M): bf16 [B][H][W][C], the gradient w.r.t.
 * (conv + bias) that tg_conv_igemm (dgrad) and tg_wgrad_igemm consume. */
int tg_bn_bwd_apply(const tg_grad_src* g0, const tg_grad_src* g1, const void* z, int B, int H, int W,
                    int C, const float* shift, const float* coeff, int act, float slope,
                    const uint8_t* code, const float* lut_dev, void* gz, void* stream);

/* ---- decoder input assembly and pooling ----------------------------------------------------- */
/* out[B][2h][2w][Cu+Cs] = cat(bilinear_up2(up[B][h][w][Cu]), skip[B][2h][2w][Cs]) * merged_mask
 * (generator.py:66-76 / :50-55 + pconv.py:27). skip may be NULL with Cs = 0 (dec1). */
int tg_upsample_concat(const void* up, int B, int h, int w, int Cu, const void* skip, int Cs,
                       const uint8_t* merged_mask, void* out, void* stream);
/* d_up[B][h][w][Cu] = bilinear_up2^T applied to channels [0,Cu) of d_merged[B][2h][2w][Ctot]. */
int tg_upsample_concat_bwd(const void* d_merged, int B, int h, int w, int Cu, int Ctot, void* d_up,
                           void* stream);
/* nn.MaxPool2d(2,2) of VGG16 features[4], [9] (losses.py:31-32) on bf16 NHWC, and its backward
 * (gradient to the first maximum in scan order; relu_gate: also multiply by [x > 0]). */
int tg_maxpool2(const void* x, int B, int H, int W, int C, void* y, void* stream);
int tg_maxpool2_bwd(const void* x, const void* gy, int B, int H, int W, int C, int relu_gate, void* gx,
                    void* stream);

/* ---- bandwidth-bound convolutions: 1 input channel or 1 output channel ------------------------ */
/* x fp32 [B][H][W] (times xmask if given) -> bf16 [B][Ho][Wo][64] (out_split: parity-split order):
 *   out = act((conv_kxk/s(x) + bias) * lut[code]);  stats: per-CTA (sum, sumsq) rows of the pre-act value.
 * Supported (k, s): (7,2) enc1, (4,2) Discriminator model[0], (3,1) VGG conv0 with folded channels.
 * wgt is fp32 [64][k*k]. */
int tg_conv_c1_fwd(const float* x, const uint8_t* xmask, int B, int H, int W, int k, int s, int pad,
                   const float* wgt, const float* bias, const uint8_t* code, const float* lut_dev, int act,
                   float slope, void* out, int out_split, float* stats, int stats_rows_cap,
                   int* stats_rows_used, void* stream);
/* dw[64][k*k] (+)= sum_p g[p][co] * x[p (+) tap];  db[64] (+)= sum_p g[p][co]  (db may be NULL).
 * g is bf16 [B][Ho][Wo][64] (g_split: parity-split order). partial: >= rows_cap*64*(k*k+1) floats. */
int tg_conv_c1_wgrad(const float* x, const uint8_t* xmask, int B, int H, int W, int k, int s, int pad,
                     const void* g, int g_split, float* partial, int rows_cap, float* dw, float* db,
                     int accumulate, void* stream);
int tg_conv_c1_wgrad_rows(void);
/* C -> 1 gather convolution with tap classes (see direct_conv.cu): x bf16 [B][H][W][C] (x_split:
 * parity-split), wgt fp32 [ntaps][C], bias device scalar or NULL.
 * mode 0: out = acc + bias.  mode 1: sig = sigmoid(acc + bias); out = sig*(1-mask) + xin*mask
 * (generator.py:56-62; sig_out receives sig for backward). */
int tg_conv_to1_fwd(const void* x, int x_split, int B, int H, int W, int C, const float* wgt, int ncls,
                    const int* cls_count, const int8_t* tap_dh, const int8_t* tap_dw, const float* bias,
                    int Ho, int Wo, int mode, const uint8_t* mask, const float* xin, float* out,
                    float* sig_out, float* scratch, size_t scratch_floats, void* stream);
/* Floats of device scratch with which tg_conv_to1_fwd runs its C = 64 cases on the tensor cores (per-pixel tap
 * dot products, then the shifted sum); with scratch == NULL or too small it runs the CUDA-core kernels. */
size_t tg_conv_to1_fwd_scratch_floats(int B, int H, int W, int C, int ntaps);
/* Kernels tg_conv_to1_fwd launches for this shape when scratch is supplied: 1 = the one-kernel path (C = 64, taps inside a
 * 3x3 window, plain layout, Ho x Wo == H x W, H and W >= 16), which needs no scratch at all; 2 otherwise; 0 = bad tap table. */
int tg_conv_to1_fwd_kernels(int x_split, int H, int W, int C, int ncls, const int* cls_count, const int8_t* tap_dh,
                            const int8_t* tap_dw, int Ho, int Wo);
/* dx[B][H][W][C] (bf16) = sum_t g[b][h - dh_t][w - dw_t] * wgt[t][c];  g fp32 [B][Ho][Wo]. */
int tg_conv_to1_bwd_data(const float* g, int B, int Ho, int Wo, const float* wgt, int ntaps,
                         const int8_t* tap_dh, const int8_t* tap_dw, int H, int W, int C, void* dx,
                         void* stream);
/* dw[C][ntaps] (+)= sum_o g[o] * x[o + d_t][c];  db[1] (+)= sum_o g[o].  ntaps in {9, 16}. */
int tg_conv_to1_wgrad(const void* x, int B, int H, int W, int C, const float* g, int Ho, int Wo, int ntaps,
                      const int8_t* tap_dh, const int8_t* tap_dw, float* partial, float* partial_b,
                      int rows_cap, float* dw, float* db, int accumulate, void* stream);
int tg_conv_to1_wgrad_rows(void);
/* g_pre = g_out * (1 - mask) * sig * (1 - sig): backward of sigmoid + composite (generator.py:57-62). */
int tg_final_bwd_pre(const float* g_out, const float* sig, const uint8_t* mask, long n, float* g_pre,
                     void* stream);

/* ---- fused losses ---------------------------------------------------------------------------
 * terms[0] = mean|pred-target| (flags&1: weighted by [mask>0], the human-region L1 of losses.py:172)
 * terms[1] = total_variation_loss(pred*(1-mask)) (losses.py:118-127; skipped if flags&2)
 * terms[2] = BoundaryAwareLoss(pred, target, mask) (losses.py:406-423), terms[3] = boundary count.
 * pred / target / mask are fp32 [B][1][H][W]; partial: >= rows_cap*5 floats. No host sync. */
int tg_loss_rows(void);
int tg_inpaint_loss_fwd(const float* pred, const float* target, const float* mask, int B, int H, int W,
                        int flags, float eps, float* partial, int rows_cap, float* terms, void* stream);
/* grad_pred = grad_terms[0]*dterms0 + grad_terms[1]*dterms1 + grad_terms[2]*dterms2. */
int tg_inpaint_loss_bwd(const float* pred, const float* target, const float* mask, int B, int H, int W,
                        int flags, float eps, const float* terms, const float* grad_terms, float* grad_pred,
                        void* stream);
/* out[0] = mean|a - b| over n bf16 elements (perceptual L1 on VGG features, losses.py:86-89), and
 * ga = grad_out[0]*sign(a-b)/n (* [a > 0] if relu_gate). n % 8 == 0. */
int tg_l1_bf16_fwd(const void* a, const void* b, long n, float* partial, int rows_cap, float* out,
                   void* stream);
int tg_l1_bf16_bwd(const void* a, const void* b, long n, const float* grad_out, int relu_gate, void* ga,
                   void* stream);
/* nn.BCEWithLogitsLoss() (mean reduction) of the adversarial terms (train.py:115, used at :203, :215-216):
 *   out[0] = mean_i( max(x_i,0) - x_i*t_i + log1p(exp(-|x_i|)) ),   grad_logits_i = grad_out[0]*(sigmoid(x_i)-t_i)/n.
 * `target` is a device fp32 [n] or NULL, in which case every t_i = target_const (the loops pass
 * torch.ones_like / torch.zeros_like). Deterministic (fixed summation order), no workspace. */
int tg_bce_logits_fwd(const float* logits, const float* target, float target_const, long n, float* out,
                      void* stream);
int tg_bce_logits_bwd(const float* logits, const float* target, float target_const, long n,
                      const float* grad_out, float* grad_logits, void* stream);

/* ---- logging-interval image-quality metrics (metrics_kernels.cu) ------------------------
 * One fused reduction replacing ExperimentTracker._calculate_psnr/_calculate_ssim/_calculate_l1_l2
 * (utils/experiment_tracking.py:196-231, called from :678-695), PerformanceMetrics (mvp_gan/src/utils/metrics.py:12-46)
 * and calculate_boundary_quality (mvp_gan/src/evaluation/metrics.py:79-133), which the loops call every
 * log_interval batches (train.py:229-266) with ~40 ATen kernels and 8 .item() host syncs.
 * pred / target / mask: fp32 [B][1][H][W] (mask may be NULL: no boundary metrics). partial: rows_cap * 10 doubles.
 * out[9] (device) = psnr, ssim, l1_distance, l2_distance, mse, boundary_mse, boundary_psnr, boundary_gradient_diff,
 * boundary pixel count. No host synchronisation. */
int tg_quality_metrics_rows(void);
int tg_quality_metrics(const float* pred, const float* target, const float* mask, int B, int H, int W,
                       double* partial, int rows_cap, float* out, void* stream);

/* ---- batched inference I/O and DSM normalisation (image_io.cu) ----------------
 * The byte-level ends of the inference path, per batch on the device instead of per tile through PIL / numpy:
 *   tg_u8_prepare          uint8 tile + uint8 mask -> masked fp32 image (u8/255 * [mask>0]) and fp32 {0,1} mask
 *                          (mvp_gan/src/evaluate.py:28-33)
 *   tg_resize_bilinear_u8  PIL.Image.resize(..., BILINEAR) on 8-bit images, bit-exact with Pillow's Resample.c
 *                          (evaluate.py:58-59 512->500; the Resize((512,512)) of evaluate.py:21-25;
 *                          utils/data_extraction.py:106-107). src may be fp32, quantised on the fly as
 *                          (x*255).astype(uint8) (evaluate.py:54-55). tmp: B*Hin*Wout bytes.
 *   tg_resize_ksize / tg_resize_coeffs   HOST helpers building Pillow's fixed-point coefficient tables
 *                          (bounds [out][2], kk [out][ksize] int32) that the caller uploads once per (in, out) size
 *   tg_quantize_u8         (x * 255).astype(uint8)
 *   tg_dsm_normalize       NaN-aware per-tile min-max normalisation of float64 DSM tiles [B][H][W] to uint8
 *                          (utils/data_extraction.py:80-103); minmax: B*2 doubles of workspace / output. */
int tg_u8_prepare(const uint8_t* img, const uint8_t* mask, long n, float* masked, float* mask_out, void* stream);
int tg_resize_ksize(int in_size, int out_size);
int tg_resize_coeffs(int in_size, int out_size, int32_t* bounds, int32_t* kk, int ksize);
int tg_resize_bilinear_u8(const void* src, int src_is_f32, int B, int Hin, int Win, int Hout, int Wout,
                          const int32_t* bounds_w, const int32_t* kk_w, int ksize_w, const int32_t* bounds_h,
                          const int32_t* kk_h, int ksize_h, uint8_t* tmp, uint8_t* dst, void* stream);
int tg_quantize_u8(const float* src, long n, uint8_t* dst, void* stream);
int tg_dsm_normalize(const double* data, int B, int H, int W, double* minmax, uint8_t* out, void* stream);

/* ---- synthetic irregular hole masks (maskgen.cu) ---------------------------------------
 * The dense image operations of generate_dem_random_mask (random__annotation_mask_generator.py:33-148): scipy.ndimage
 * binary morphology with the default cross structuring element and border value 0, the separable float64 Gaussian
 * filter (mode 'reflect', scipy's summation order) and the threshold / distance tests, on [H][W] arrays. The random
 * draws stay on the host in the reference's order (tg_b200/maskgen.py): seeded runs reproduce the reference's masks.
 *   tg_morph_cross  one iteration of binary_dilation (erode = 0) or binary_erosion (erode = 1); out != in
 *   tg_gauss1d_f64  correlate1d of float64 data with a symmetric kernel weights[2*radius+1] along axis 0 / 1; out != in
 *   tg_mask_shape   mode 0: out = field > p0; 1: out |= dist <= p0 + field*p1; 2: out |= x^2/p0^2 + y^2/p1^2 <= 1;
 *                   3: out |= (x^2+y^2 <= p0^2) & (field > p1), with x = col - cx, y = row - cy */
int tg_morph_cross(const uint8_t* in, int H, int W, int erode, uint8_t* out, void* stream);
int tg_gauss1d_f64(const double* in, int H, int W, int axis, const double* weights, int radius, double* out,
                   void* stream);
int tg_mask_shape(const double* field, int H, int W, int mode, int cx, int cy, double p0, double p1, uint8_t* out,
                  void* stream);

/* ---- fp32-storage verification path ("TF32 path <= 1e-3", BASELINE.json north star) ---------------------------
 * Twins of the entry points above with fp32 activations instead of bf16 (same argument meaning; every `void*`
 * activation / gradient pointer is a float tensor). The implicit-GEMM entry points take `dtype = TG_DTYPE_F32` in
 * their argument block instead. These run the SAME kernel templates with the storage type swapped (BatchNorm,
 * resampling, pooling, L1) or the CUDA-core kernels of direct_conv.cu in fp32 (the 1<->64-channel convolutions);
 * they exist to prove the forward / backward formulation against the fp32 reference to a sharp bound. */
int tg_bn_apply_f32(const void* z, int B, int H, int W, int C, const float* scale, const float* shift,
                    int act, float slope, const uint8_t* code, void* y_nhwc, void* y_split,
                    int mask_split, void* stream);
int tg_bn_bwd_reduce_f32(const tg_grad_src* g0, const tg_grad_src* g1, const void* z, int B, int H, int W,
                         int C, const float* scale, const float* shift, int act, float slope,
                         const uint8_t* code, const float* lut_dev, float* partial, int rows_cap,
                         int* rows_used, void* stream);
int tg_bn_bwd_apply_f32(const tg_grad_src* g0, const tg_grad_src* g1, const void* z, int B, int H, int W,
                        int C, const float* shift, const float* coeff, int act, float slope,
                        const uint8_t* code, const float* lut_dev, void* gz, void* stream);
int tg_upsample_concat_f32(const void* up, int B, int h, int w, int Cu, const void* skip, int Cs,
                           const uint8_t* merged_mask, void* out, void* stream);
int tg_upsample_concat_bwd_f32(const void* d_merged, int B, int h, int w, int Cu, int Ctot, void* d_up,
                               void* stream);
int tg_maxpool2_f32(const void* x, int B, int H, int W, int C, void* y, void* stream);
int tg_maxpool2_bwd_f32(const void* x, const void* gy, int B, int H, int W, int C, int relu_gate, void* gx,
                        void* stream);
int tg_conv_c1_fwd_f32(const float* x, const uint8_t* xmask, int B, int H, int W, int k, int s, int pad,
                       const float* wgt, const float* bias, const uint8_t* code, const float* lut_dev, int act,
                       float slope, void* out, int out_split, float* stats, int stats_rows_cap,
                       int* stats_rows_used, void* stream);
int tg_conv_c1_wgrad_f32(const float* x, const uint8_t* xmask, int B, int H, int W, int k, int s, int pad,
                         const void* g, int g_split, float* partial, int rows_cap, float* dw, float* db,
                         int accumulate, void* stream);
int tg_conv_to1_fwd_f32(const void* x, int x_split, int B, int H, int W, int C, const float* wgt, int ncls,
                        const int* cls_count, const int8_t* tap_dh, const int8_t* tap_dw, const float* bias,
                        int Ho, int Wo, int mode, const uint8_t* mask, const float* xin, float* out,
                        float* sig_out, void* stream);
int tg_conv_to1_bwd_data_f32(const float* g, int B, int Ho, int Wo, const float* wgt, int ntaps,
                             const int8_t* tap_dh, const int8_t* tap_dw, int H, int W, int C, void* dx,
                             void* stream);
int tg_conv_to1_wgrad_f32(const void* x, int B, int H, int W, int C, const float* g, int Ho, int Wo, int ntaps,
                          const int8_t* tap_dh, const int8_t* tap_dw, float* partial, float* partial_b,
                          int rows_cap, float* dw, float* db, int accumulate, void* stream);
int tg_l1_f32_fwd(const void* a, const void* b, long n, float* partial, int rows_cap, float* out, void* stream);
int tg_l1_f32_bwd(const void* a, const void* b, long n, const float* grad_out, int relu_gate, void* ga,
                  void* stream);
/* Two-term TF32 split of an fp32 tensor [rows][C] along its last dimension: hi = tf32(x), lo = tf32(x - hi).
 * layout 0: out [rows][3C] = [hi | lo | hi]; 1: [hi | hi | lo]; 2: out [rows][2C] = [hi | lo]; 3: out [rows][C] = hi;
 * 4: out [rows][2C] = [lo | hi]. A contraction of a [lo | hi] operand with a [hi | lo] operand over the concatenated
 * axis gives the cross terms lo*hi + hi*lo, hi against hi the main term: together fp32-grade products from
 * kind::tf32 passes ("tf32x3" precision of tg_b200.precision). */
int tg_split_tf32(const float* x, long rows, int C, int layout, float* out, void* stream);

/* ---- fused Adam step + packed-weight refresh (adam_kernels.cu) ---------------------------------------
 * Replaces torch.optim.Adam.step() of the reference loops (mvp_gan/src/train.py:207,219,
 * training/human_guided_trainer.py:153; amsgrad off, weight_decay 0) for a list of fp32 parameter tensors and,
 * for conv weights, writes the bf16 fprop / dgrad operand matrices in the same pass: packed[dst[i]] = bf16(p[i]).
 * `tensors` is a HOST array; all pointers inside are device pointers. `step` is the 1-based step count. */
typedef struct tg_adam_tensor {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  void* packed_fprop;           /* bf16 or NULL */
  const int32_t* dst_fprop;     /* [n] or NULL */
  void* packed_dgrad;           /* bf16 or NULL */
  const int32_t* dst_dgrad;     /* [n] or NULL */
  int64_t n;
} tg_adam_tensor;
int tg_adam_repack(const tg_adam_tensor* tensors, int n_tensors, float lr, float beta1, float beta2, float eps,
                   int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TERRAGAN_B200_H */
