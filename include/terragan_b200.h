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

/* ---- library ------------------------------------------------------------------------------ */
int tg_version(void);                               /* ABI version (this header: 1) */
size_t tg_last_error(char* buf, size_t cap);        /* copies the last error text of this thread */
int tg_num_sms(void);                               /* SM count of the current device (<=0: error) */

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
} tg_conv_args;
int tg_conv_igemm(tg_conv_args* args, void* stream);

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
} tg_wgrad_args;
int tg_wgrad_igemm(tg_wgrad_args* args, void* stream);
/* dw[n][c][tap_of(kh,kw)] (+)= sum_s partial[s][tap*C + c][n];  tap order is the caller's tap table,
 * tap_perm[t] gives the kh*kw-flattened kernel position of tap t. */
int tg_wgrad_reduce(const float* partial, int splits, int num_taps, int C, int N,
                    const int32_t* tap_perm_dev, float* dw, int accumulate, void* stream);
/* How many fp32 the `partial` workspace of tg_wgrad_igemm needs for this problem. */
int64_t tg_wgrad_partial_floats(int B, int Ho, int Wo, int num_taps, int C, int N);

#ifdef __cplusplus
}
#endif
#endif /* TERRAGAN_B200_H */
