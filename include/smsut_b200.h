/*
 * smsut_b200.h -- C ABI of libsmsut_b200.so (hand-written sm_100a kernels for the SMSUT training step).
 *
 * The reference (Sue1347/SMSUT-MedicalImgSegmentation) has NO FFI layer: every op below replaces a
 * PyTorch aten call made by the reference's nn.Modules / trainer.  Each entry cites the reference
 * file:line whose arithmetic it implements.  Conventions:
 *   - all pointers are DEVICE pointers unless named host_*; activations are NHWC bf16 unless stated;
 *     parameters / gradients / statistics / losses are fp32; labels are int64 (torch.long) or u8.
 *   - every function takes the cudaStream_t to launch on, never synchronises, returns 0 on success and
 *     a negative code on failure; smsut_last_error() returns a thread-local message.
 *   - no global mutable state except per-process caches of driver entry points / device attributes.
 */
#ifndef SMSUT_B200_H_
#define SMSUT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* smsut_stream_t; /* == cudaStream_t */

const char* smsut_last_error(void);
int smsut_abi_version(void);
/* number of kernels launched by this library since process start (bench.py "gpu_launches") */
int64_t smsut_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (TMA-staged NHWC bf16 tiles, fp32 accum in TMEM).
 * Replaces aten::convolution / convolution_backward(input) for: conv3x3 s1 p1 and conv1x1
 * (network/blocks.py:10-16), nn.ConvTranspose2d k2 s2 (network/blocks.py:41), nn.Linear of netF
 * (network/ugan.py:295) and -- with the dgrad-packed weights -- their input gradients.
 * The two-source K loop replaces torch.cat([up, skip], 1) (network/blocks.py:50).
 * ---------------------------------------------------------------------------------------------- */
enum { SMSUT_TC_CONV = 0, SMSUT_TC_CONVT_FWD = 1, SMSUT_TC_CONVT_DGRAD = 2 };
enum { SMSUT_ACT_NONE = 0, SMSUT_ACT_RELU = 1, SMSUT_ACT_LRELU = 2 };

typedef struct smsut_conv_tc_args {
  int32_t kind;       /* SMSUT_TC_* */
  int32_t ksize;      /* 1, 3 or 5 (kind CONV); ignored otherwise */
  int32_t n, h, w;    /* GEMM-row space: output dims for CONV; the transposed conv's INPUT dims otherwise */
  int32_t nsrc;       /* 1 or 2 activation sources (channel-concatenated, src[0] first) */
  const void* src[2]; /* bf16 NHWC; CONVT_DGRAD: the (n, 2h, 2w, src_c[0]) output gradient */
  int32_t src_c[2];   /* channels consumed from each source (multiple of 16) */
  int32_t src_ld[2];  /* channel pitch of each source tensor in elements (>= src_c) */
  const void* wpack;  /* bf16 [ncols_pad][ntaps * (src_c[0]+src_c[1])], K index = tap*ctot + channel */
  int32_t ncols;      /* valid GEMM columns (Cout; 4*Cout for CONVT_FWD, column = tap*Cout + co) */
  int32_t ncols_pad;  /* rows of wpack, multiple of 16 */
  void* out0;         /* destination for columns [0, split) (or all) */
  int32_t out0_ld, out0_coff;
  void* out1;         /* optional destination for columns [split, ncols) */
  int32_t out1_ld, out1_coff, split; /* split <= 0: unused */
  const float* bias;  /* optional fp32 [ncols] */
  int32_t act;        /* SMSUT_ACT_* applied after bias */
  float slope;
  int32_t accumulate; /* out += result (read-modify-write) */
  int32_t out_f32;    /* store fp32 instead of bf16 */
  int32_t bn;         /* N tile, 0 = auto */
  float* stats;       /* optional fused InstanceNorm statistics (network/blocks.py:22-23): stats[n][2][ncols_pad] +=
                         {sum, sum of squares} over H*W of the STORED (bf16-rounded) outputs.  Only valid when
                         smsut_conv_tc_fuses_stats() returns 1 for these arguments; must be zeroed by the caller. */
} smsut_conv_tc_args;

int smsut_conv_tc(const smsut_conv_tc_args* a, smsut_stream_t stream);
/* 1 if smsut_conv_tc would honour a->stats for this shape (wide layers on the band kernel), else 0 */
int smsut_conv_tc_fuses_stats(const smsut_conv_tc_args* a);

/* Weight gradient on tcgen05 (MN-major operands, split-K over pixels, fp32 atomics into OIHW grads).
 * Replaces aten::convolution_backward(weight) for the same layer classes.
 *   dW[co][ci_off+ci][ky][kx] += sum_p dy[p][co] * x[p + tap][ci]                      (CONV)
 *   dW[ci_off+ci][co][ty][tx] += sum_p x[p][ci] * dy[2p + tap][co]                    (CONVT) */
typedef struct smsut_wgrad_tc_args {
  int32_t kind;  /* SMSUT_TC_CONV or SMSUT_TC_CONVT_FWD */
  int32_t ksize; /* 1, 3 or 5 */
  int32_t n, h, w;
  const void* x;  /* bf16 NHWC input activations (n,h,w,*) */
  int32_t x_c, x_ld;
  const void* dy; /* bf16 NHWC output gradient ((n,h,w,*) or (n,2h,2w,*)) */
  int32_t dy_c, dy_ld;
  float* dw;      /* fp32 OIHW (conv) / IOHW (convT) gradient, accumulated with atomics */
  int32_t cin_total, ci_off; /* full Cin of the weight and channel offset of this source */
  int32_t cout_total;
  int32_t c_valid;           /* CONV: channels of x that exist in the weight (x may be zero-padded); 0 = all */
  int32_t dw_layout;         /* 0: dw is OIHW / IOHW (scalar atomics, 36-byte stride between lanes);
                                1: dw is the tap-major scratch [tap][cout][cin] (conv) / [tap][cin][cout] (convT):
                                   a thread's 16 accumulator columns are 64 contiguous bytes -> red.global.v4.f32,
                                   a warp's lanes are contiguous channels in the band kernel; smsut_unpack_wgrads
                                   folds the scratch into the OIHW gradient once per optimizer step */
} smsut_wgrad_tc_args;

int smsut_wgrad_tc(const smsut_wgrad_tc_args* a, smsut_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Direct (CUDA-core) convolutions: the HBM-bound tiny-K layers and a generic reference path.
 *   - stems: 5x5 s1 p2 Cin in {1,5} (network/ugan.py:26), 4x4 s2 p1 Cin=1 + bias (network/ugan.py:202)
 *   - heads: 1x1 16->{1,5} + bias (+tanh) (network/ugan.py:70-83), conv_src 3x3 256->1, conv_cls 4x4 valid
 *     256->4 (network/ugan.py:213-215)
 * x: NHWC (bf16 or fp32), w: fp32 OIHW master weights, y: NHWC (bf16 or fp32).
 * ---------------------------------------------------------------------------------------------- */
typedef struct smsut_conv_direct_args {
  int32_t n, h, w, cin;        /* input dims */
  int32_t cout, kh, kw, stride, pad;
  int32_t ho, wo;              /* output dims */
  const void* x; int32_t x_ld; int32_t x_f32;
  const float* wt;             /* fp32 OIHW */
  const float* bias;           /* optional */
  void* y; int32_t y_ld; int32_t y_f32;
  int32_t act; float slope;    /* SMSUT_ACT_*; act==3: tanh */
  int32_t accumulate;
} smsut_conv_direct_args;

int smsut_conv_direct_fprop(const smsut_conv_direct_args* a, smsut_stream_t stream);
/* dx[n,h,w,ci] (+)= sum dy[n,ho,wo,co] * w[co,ci,ky,kx]; x/y fields name dx / dy here */
int smsut_conv_direct_dgrad(const smsut_conv_direct_args* a, smsut_stream_t stream);
/* dW[co,ci,ky,kx] += sum x * dy ; dbias[co] += sum dy (if dbias != NULL) */
int smsut_conv_direct_wgrad(const smsut_conv_direct_args* a, float* dw, float* dbias, smsut_stream_t stream);

/* Fused backward of the 1x1 heads (network/ugan.py:70-83: tsl/seg `fc`, network/blocks.py:166: U-Net `fc`):
 * g = dy * (1 - y^2) when y != NULL (tanh head); dx = g W (bf16, optional); dW += g^T x; dbias += sum g.
 * x (npix, 16) bf16, dy / y (npix, cout) fp32, w fp32 (cout, 16). */
int smsut_head1x1_bwd(const void* x, const float* dy, const float* y, const float* w, void* dx, float* dw, float* db,
                      int64_t npix, int32_t cin, int32_t cout, smsut_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * InstanceNorm2d(affine) + LeakyReLU + residual  (network/blocks.py:22-34, 66-80, 99-117)
 * ---------------------------------------------------------------------------------------------- */
/* stats[n][c] = {sum, sumsq} over H*W (fp32, must be zeroed by the caller); x bf16 NHWC */
int smsut_in_stats(const void* x, int32_t n, int32_t hw, int32_t c, float* stats, smsut_stream_t stream);
/* out = act( IN_a(xa) [+ IN_b(xb)] [+ res] ), biased variance, eps 1e-5.
 * c_params <= c: channels >= c_params are layout padding (gamma = beta = 0, no parameter gradients). */
int smsut_in_apply(const void* xa, const float* stats_a, const float* gamma_a, const float* beta_a,
                   const void* xb, const float* stats_b, const float* gamma_b, const float* beta_b,
                   const void* res, void* out, int32_t n, int32_t hw, int32_t c, int32_t c_params, int32_t act,
                   float slope, smsut_stream_t stream);
/* backward, pass 1: g = dout * act'(out);  red[n][c] = {sum g, sum g*xhat_a, sum g*xhat_b} (zeroed by caller).
 * out == NULL (with an activation): the sign of the activation's input is recomputed from xa / xb with the forward's
 * own expression (gamma / beta of every branch required) instead of read from `out` -- valid whenever the forward
 * had no residual input; saves one streamed tensor.  gamma / beta may be NULL otherwise. */
int smsut_in_bwd_reduce(const void* dout, const void* out, const void* xa, const float* stats_a, const float* gamma_a,
                        const float* beta_a, const void* xb, const float* stats_b, const float* gamma_b,
                        const float* beta_b, float* red, int32_t n, int32_t hw, int32_t c, int32_t c_params,
                        int32_t act, float slope, smsut_stream_t stream);
/* backward, pass 2: dxa = gamma_a*rstd_a*(g - mean g - xhat_a*mean(g xhat_a)); same for b; dres = g (optional);
 * dgamma/dbeta accumulated (atomics) into fp32 parameter gradients.  out == NULL as above (beta_a / beta_b required). */
int smsut_in_bwd_apply(const void* dout, const void* out, const void* xa, const float* stats_a, const float* gamma_a,
                       const float* beta_a, void* dxa, float* dgamma_a, float* dbeta_a, const void* xb,
                       const float* stats_b, const float* gamma_b, const float* beta_b, void* dxb, float* dgamma_b,
                       float* dbeta_b, void* dres, const float* red, int32_t n, int32_t hw, int32_t c,
                       int32_t c_params, int32_t act, float slope, smsut_stream_t stream);
/* both passes in ONE launch (reduce, a barrier over the CTAs of a sample, apply; the second pass re-reads its strip from
 * L2): same arguments as the pair above plus `counters`, n zeroed 32-bit words.  Falls back to the two kernels when
 * the grid cannot be co-resident (more samples than resident CTAs) or SMSUT_IN_FUSED=0.  red[n][3][c] zeroed by the
 * caller as for smsut_in_bwd_reduce. */
int smsut_in_bwd_fused(const void* dout, const void* out, const void* xa, const float* stats_a, const float* gamma_a,
                       const float* beta_a, void* dxa, float* dgamma_a, float* dbeta_a, const void* xb,
                       const float* stats_b, const float* gamma_b, const float* beta_b, void* dxb, float* dgamma_b,
                       float* dbeta_b, void* dres, float* red, void* counters, int32_t n, int32_t hw, int32_t c,
                       int32_t c_params, int32_t act, float slope, smsut_stream_t stream);
/* double backward of InstanceNorm (WGAN-GP, trainer/uganShp0Trainer.py:127-134):
 * given u = cotangent of dx, with dx = IN_bwd(dy; x, gamma):
 *   pass 1: red2[n][c] = {sum u, sum dy, sum u*xhat, sum dy*xhat, sum u*dy}
 *   pass 2: g_dy, g_x (bf16), dgamma += sum u*dx/gamma */
int smsut_in_bwd2_reduce(const void* u, const void* dy, const void* x, const float* stats, float* red2, int32_t n,
                         int32_t hw, int32_t c, smsut_stream_t stream);
int smsut_in_bwd2_apply(const void* u, const void* dy, const void* x, const float* stats, const float* gamma,
                        const float* red2, void* g_dy, void* g_x, float* dgamma, int32_t n, int32_t hw, int32_t c,
                        smsut_stream_t stream);
/* the double backward in ONE launch (same design as smsut_in_bwd_fused); counters: n zeroed 32-bit words */
int smsut_in_bwd2_fused(const void* u, const void* dy, const void* x, const float* stats, const float* gamma, float* red2,
                        void* counters, void* g_dy, void* g_x, float* dgamma, int32_t n, int32_t hw, int32_t c,
                        smsut_stream_t stream);
/* BatchNorm2d (network/blocks.py:19-26 get_norm('batch'); default norm of network/unet.py:14) on the kernels above:
 * out[j] = mean over the n rows of rows[n][k][c] for every j (batch statistics = pooled per-sample sums; the
 * backward pools `red` the same way between smsut_in_bwd_reduce and smsut_in_bwd_apply) */
int smsut_bn_pool(const float* rows, float* out, int32_t n, int32_t k, int32_t c, smsut_stream_t stream);
/* running_mean/var <- (1-momentum)*running + momentum*{batch mean, UNBIASED batch var} (torch.nn.BatchNorm2d) from a
 * pooled statistics table; only the first c_params channels carry parameters */
int smsut_bn_running_update(const float* pooled_stats, int32_t n, int32_t hw, int32_t c, int32_t c_params,
                            float momentum, float* running_mean, float* running_var, smsut_stream_t stream);
/* eval mode: stats[n][2][c] such that smsut_in_apply normalises with the running estimates */
int smsut_bn_eval_stats(const float* running_mean, const float* running_var, float* stats, int32_t n, int32_t hw,
                        int32_t c, int32_t c_params, smsut_stream_t stream);
/* elementwise LeakyReLU family on bf16: y = act(x) ; dx = dy*act'(ref) (+ add) ; plain add */
int smsut_act_fwd(const void* x, void* y, int64_t count, int32_t act, float slope, smsut_stream_t stream);
int smsut_act_bwd(const void* dy, const void* ref, const void* add, void* dx, int64_t count, int32_t act, float slope,
                  smsut_stream_t stream);
int smsut_add_bf16(const void* a, const void* b, void* out, int64_t count, smsut_stream_t stream);
/* out[c] += sum over rows of x[rows][c] (bf16 in, fp32 atomics): bias gradients of netF's Linear layers */
int smsut_colsum_bf16(const void* x, int32_t rows, int32_t c, float* out, smsut_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Pooling / resampling (network/ugan.py:31-37, network/blocks.py:44,101-112)
 * ---------------------------------------------------------------------------------------------- */
int smsut_maxpool2_fwd(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, smsut_stream_t stream);
/* dx = [add +] route(dy) to the first maximum of each 2x2 window (PyTorch index semantics) */
int smsut_maxpool2_bwd(const void* x, const void* dy, const void* add, void* dx, int32_t n, int32_t h, int32_t w,
                       int32_t c, smsut_stream_t stream);
int smsut_avgpool2_fwd(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, smsut_stream_t stream);
/* dx = [add +] 0.25 * dy (nearest upsample); also the double-backward partner of avgpool2_fwd */
int smsut_avgpool2_bwd(const void* dy, const void* add, void* dx, int32_t n, int32_t h, int32_t w, int32_t c,
                       smsut_stream_t stream);
/* bilinear x2, align_corners=False; (n,h,w,c) -> (n,2h,2w,c) and its adjoint */
int smsut_bilinear2_fwd(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, smsut_stream_t stream);
int smsut_bilinear2_bwd(const void* dy, void* dx, int32_t n, int32_t h, int32_t w, int32_t c, smsut_stream_t stream);

/* layout / dtype conversion at the module boundary */
int smsut_nchw_f32_to_nhwc_bf16(const float* x, void* y, int32_t n, int32_t c, int32_t h, int32_t w, int32_t c_pad,
                                smsut_stream_t stream);
int smsut_nhwc_bf16_to_nchw_f32(const void* x, float* y, int32_t n, int32_t c, int32_t h, int32_t w, int32_t x_ld,
                                smsut_stream_t stream);
/* tsl input: concat(x, modality planes) -> (n,h,w,c_pad) bf16; m is fp32 (n, n_modal)  (network/ugan.py:154-159) */
int smsut_build_tsl_input(const float* x, const float* m, void* y, int32_t n, int32_t hw, int32_t n_modal,
                          int32_t c_pad, smsut_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Losses (misc/loss.py:16-63, network/patchnce.py:13-51, trainer/uganConsisTrainer.py:129-177,
 *         trainer/uganShp0Trainer.py:127-134)
 * ---------------------------------------------------------------------------------------------- */
/* logits fp32 NHWC (npix, c<=8); labels int64 or (labels==NULL) argmax of label_logits (pseudo labels).
 * acc[3*c + 1] (zeroed by caller) = {tp[c], fp[c], fn[c], sum_ce}.  */
int smsut_dice_ce_fwd(const float* logits, const int64_t* labels, const float* label_logits, float* acc,
                      int64_t npix, int32_t c, smsut_stream_t stream);
/* loss = w_dc*(1 - mean_{c>=1} (2tp+s)/(2tp+fp+fn+s+1e-8)) + w_ce*sum_ce/npix_total; writes loss[0] */
int smsut_dice_ce_finish(const float* acc, float* loss, int64_t npix_total, int32_t c, float w_dc, float w_ce,
                         smsut_stream_t stream);
/* dlogits = gscale[0]*scale * dLoss/dlogits (fp32 NHWC) */
int smsut_dice_ce_bwd(const float* logits, const int64_t* labels, const float* label_logits, const float* acc,
                      const float* gscale, float scale, float* dlogits, int64_t npix, int64_t npix_total, int32_t c,
                      float w_dc, float w_ce, smsut_stream_t stream);
/* out[0] += mean over (pixels, classes) of (softmax(zs) - softmax(zt))^2; dzs = gscale * d/dzs (zt is constant):
 * the mean-teacher consistency term (trainer/meanTeacherTrainer.py:124-130).  fp32 (npix, c) logits. */
int smsut_softmax_mse_fwd(const float* zs, const float* zt, float* out, int64_t npix, int32_t c, smsut_stream_t stream);
int smsut_softmax_mse_bwd(const float* zs, const float* zt, const float* gscale, float* dzs, int64_t npix, int32_t c,
                          smsut_stream_t stream);
/* argmax over channels -> int64 (n,h,w)  (trainer/uganConsisTrainer.py:52, trainer/baseTrainer.py:230) */
int smsut_argmax_c(const float* logits, int64_t* out, int64_t npix, int32_t c, smsut_stream_t stream);
/* validation metric path (trainer/baseTrainer.py:207-252, trainer/uganShp0Trainer.py:250-287, misc/utils.py:180-203):
 * conf[label * c + argmax(logits)] += 1 over npix pixels (logits fp32 (npix, c) NHWC-flattened, labels int64; labels
 * outside [0, c) are ignored; ties go to the first maximum like torch.argmax); conf: c*c uint64 counters, accumulated */
int smsut_confusion_counts(const float* logits, const int64_t* labels, uint64_t* conf, int64_t npix, int32_t c,
                           smsut_stream_t stream);
/* out[0] += scale * sum|a-b| ; dA = gscale*scale*sign(a-b) */
int smsut_l1_fwd(const float* a, const float* b, float* out, int64_t count, float scale, smsut_stream_t stream);
int smsut_l1_bwd(const float* a, const float* b, const float* gscale, float scale, float* da, int64_t count,
                 smsut_stream_t stream);
/* out[0] += scale * sum(x) (fp32) */
int smsut_sum_f32(const float* x, float* out, int64_t count, float scale, smsut_stream_t stream);
int smsut_fill_f32(float* x, int64_t count, float value, smsut_stream_t stream);
/* x[i] = gscale[0]*scale: backward of the adversarial means -mean(D(x)) (trainer/uganConsisTrainer.py:130,136,154) */
int smsut_fill_scaled_f32(float* x, int64_t count, const float* gscale, float scale, smsut_stream_t stream);
/* dx = dy*(1 - y^2): backward of the tanh fused into the translation head (network/ugan.py:73,82) */
int smsut_tanh_bwd(const float* dy, const float* y, float* dx, int64_t count, smsut_stream_t stream);
/* out[r][i] = alpha[r]*x[r][i] + (1-alpha[r])*y[r][i]: x_hat of the gradient penalty (trainer/uganConsisTrainer.py:139) */
int smsut_lerp_rows_f32(const float* alpha, const float* x, const float* y, float* out, int32_t rows, int64_t per,
                        smsut_stream_t stream);
/* softmax cross-entropy on small (rows, c<=8) fp32 logits with int64 targets: out[0] += scale*mean CE;
 * dlogits = gscale*scale*(p - onehot)/rows */
int smsut_ce_rows_fwd(const float* logits, const int64_t* target, float* out, int32_t rows, int32_t c, float scale,
                      smsut_stream_t stream);
int smsut_ce_rows_bwd(const float* logits, const int64_t* target, const float* gscale, float scale, float* dlogits,
                      int32_t rows, int32_t c, smsut_stream_t stream);
/* WGAN-GP: norm[b] = ||g_b||_2 over `per` elements (fp32 g); out[0] += scale*mean((norm-1)^2);
 * norm2: b zeroed floats of scratch (the squared norms are accumulated by several blocks per sample);
 * backward: u = gscale*scale*2*(norm-1)/(B*norm) * g  (cotangent of g) */
int smsut_gp_fwd(const float* g, float* norm, float* norm2, float* out, int32_t b, int64_t per, float scale,
                 smsut_stream_t stream);
int smsut_gp_bwd(const float* g, const float* norm, const float* gscale, float scale, float* u, int32_t b, int64_t per,
                 smsut_stream_t stream);
/* PatchNCE: gather rows of a (n, hw, c) bf16 feature map at `ids` -> (n*nids, c) bf16 and its adjoint */
int smsut_gather_rows(const void* feat, const int64_t* ids, void* out, int32_t n, int32_t hw, int32_t c, int32_t nids,
                      smsut_stream_t stream);
int smsut_scatter_rows_add(const void* dout, const int64_t* ids, void* dfeat, int32_t n, int32_t hw, int32_t c,
                           int32_t nids, smsut_stream_t stream);
/* y = x / (||x||_2 + 1e-7) per row (fp32 in, fp32 out), and backward writing bf16 dx (network/networks.py:241-242) */
int smsut_l2norm_fwd(const float* x, float* y, float* norm, int32_t rows, int32_t c, smsut_stream_t stream);
int smsut_l2norm_bwd(const float* dy, const float* y, const float* norm, void* dx, int32_t rows, int32_t c,
                     smsut_stream_t stream);
/* PatchNCE logits + CE fused: q,k fp32 (groups*np, c) L2-normalised; loss_rows[r] and out[0] += scale*mean;
 * dq = gscale*scale/rows * dCE/dq  (k is detached)  (network/patchnce.py:13-51) */
int smsut_patchnce_fwd(const float* q, const float* k, float* loss_rows, float* out, int32_t groups, int32_t np,
                       int32_t c, float inv_t, float scale, smsut_stream_t stream);
int smsut_patchnce_bwd(const float* q, const float* k, const float* gscale, float scale, float* dq, int32_t groups,
                       int32_t np, int32_t c, float inv_t, smsut_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Optimisers over flat fp32 parameter buffers (trainer/uganShp0Trainer.py:72-74,
 * trainer/meanTeacherTrainer.py:63-69) and bf16 weight packing for the tensor-core convs.
 * lr is read from a DEVICE scalar so a captured CUDA graph follows the poly schedule.
 * ---------------------------------------------------------------------------------------------- */
int smsut_sgd_step(float* p, const float* g, float* mom, int64_t count, const float* lr, float momentum,
                   float weight_decay, float grad_scale, smsut_stream_t stream);
/* state[0] = step count (device, fp32) incremented inside */
int smsut_adam_step(float* p, const float* g, float* m, float* v, int64_t count, const float* lr, float beta1,
                    float beta2, float eps, float weight_decay, float* state, float grad_scale,
                    smsut_stream_t stream);
int smsut_ema_update(float* ema, const float* p, int64_t count, const float* alpha, smsut_stream_t stream);
/* lr_out[0] = base*(1 - max(iter-1,0)/max_iter)^power; iter_state[0] += 1  (trainer/uganConsisTrainer.py:198-203) */
int smsut_poly_lr_tick(float* iter_state, float* lr_out, float base_lr, float max_iter, float power,
                       smsut_stream_t stream);

typedef struct smsut_pack_entry {
  const float* w;   /* fp32 master: OIHW (conv) or IOHW (convT) */
  void* fprop;      /* bf16 [cout_pad][taps][cin] (conv) | [taps*cout][cin] (convT)        */
  void* dgrad;      /* bf16 [cin_pad][taps flipped][cout] (conv) | [cin][taps][cout] (convT); may be NULL */
  int32_t cout, cin, kh, kw;
  int32_t transposed; /* 1 = ConvTranspose2d weight */
  int32_t cout_pad, cin_pad;
} smsut_pack_entry;
/* `table` is a DEVICE array of n entries */
int smsut_pack_weights(const smsut_pack_entry* table, int32_t n, smsut_stream_t stream);

/* Tap-major weight-gradient scratch -> fp32 master-layout gradient (the mirror image of smsut_pack_weights):
 *   grad[(m * cols + c) * taps + t] += scratch[(t * rows + m) * cols + c]     rows x cols = Cout x Cin (conv), Cin x Cout (convT)
 * One launch per network and optimizer step; `table` is a DEVICE array of n entries. */
typedef struct smsut_unpack_entry {
  const float* scratch;
  float* grad;
  int32_t rows, cols, taps, pad;
} smsut_unpack_entry;
int smsut_unpack_wgrads(const smsut_unpack_entry* table, int32_t n, smsut_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Deterministic accumulation (SMSUT_DETERMINISTIC=1).  The reference's reductions (InstanceNorm sums over H*W,
 * network/blocks.py:22-23; aten::convolution_backward(weight); the batch-wide Dice statistics, misc/loss.py:52-63;
 * the scalar loss means, trainer/uganConsisTrainer.py:129-177) are cross-CTA sums; the library forms them with fp32
 * atomics, whose order changes run to run.  For every accumulator address inside a registered range the kernels add
 * into a 64-bit fixed-point (Q31.32) shadow instead -- integer addition is associative, so any arrival order, even
 * from concurrent streams, gives bit-identical totals -- and smsut_det_resolve folds the shadow into the fp32
 * destination (dst[i] += shadow[i] * 2^-32; shadow[i] = 0).  Unregistered destinations keep the fp32 atomics.
 * `shadow` holds bytes/4 int64 values (2x the bytes of the range) and must be zero before the first accumulation.
 * ---------------------------------------------------------------------------------------------- */
int smsut_det_register(const void* base, size_t bytes, void* shadow);
int smsut_det_unregister(const void* base);
int smsut_det_ranges(void);                       /* number of live registrations */
void* smsut_det_shadow(const void* p);            /* shadow address of an accumulator address, or NULL */
int smsut_det_resolve(float* dst, int64_t count, smsut_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Input pipeline: gather + joint augmentation + normalisation of one batch in one launch.
 * Replaces, per slice, the reference's host-side chain  JointRotate -> JointElasticDeform -> JointRandomResizedCrop
 * (data_loader/externalTransforms.py:46-101) -> [RandomGammaCorrection] -> ToTensor + Normalize(0.5, 0.5) /
 * MaskToTensor (data_loader/baseLoader.py:87-112), applied by BalanceDataset.__getitem__
 * (data_loader/balanceLoader.py:66-76) in DataLoader workers.
 *   images / labels : the resident u8 dataset, (slices, h, w) each
 *   index           : int64 [n] slice ids of the batch (the InTurn sampler's draw, data_loader/inTurnLoader.py:15-60)
 *   params          : fp32 [n][SMSUT_AUG_PARAM_FLOATS], per slice:
 *                     rot_on, m00, m01, m02, m10, m11, m12   inverse rotation, source = M (x+.5, y+.5, 1) - .5
 *                     ela_on                                  apply the elastic deformation
 *                     crop_on, top, left, height, width       crop box resized to (h, w)
 *                     gamma_on, gamma                         RandomGammaCorrection (image only)
 *                     points, coef[2][5][5]                   cubic B-spline coefficients of the control displacements
 *   x_out (n,1,h,w) fp32 in [-1, 1];  y_out (n,h,w) int64.   h * w <= 100 KiB (the plane lives in shared memory).
 * ---------------------------------------------------------------------------------------------- */
#define SMSUT_AUG_PARAM_FLOATS 66
int smsut_augment_batch(const uint8_t* images, const uint8_t* labels, const int64_t* index, const float* params,
                        float* x_out, int64_t* y_out, int32_t n, int32_t h, int32_t w, smsut_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * coraNet losses (trainer/coraNetTrainer.py).  All tensors fp32 NHWC-flattened (npix, c),
 * labels int64, accumulators zeroed by the caller.
 * ---------------------------------------------------------------------------------------------- */
/* the (1 + nheads*nlab)-channel output -> nheads stacked (npix, 1 + nlab) tensors, head h = [channel 0, channels
 * 1 + h*nlab .. (h+1)*nlab]: `torch.cat([out_back, out_h], dim=1)` of coraNetTrainer.py:279-297; the backward sums the
 * background gradients of the heads */
int smsut_heads_split_fwd(const float* z, float* heads, int64_t npix, int32_t nlab, int32_t nheads, smsut_stream_t stream);
int smsut_heads_split_bwd(const float* dheads, float* dz, int64_t npix, int32_t nlab, int32_t nheads, smsut_stream_t stream);
/* nn.CrossEntropyLoss(weight=cw[, reduction='none']) of coraNetTrainer.py:44-58: acc[0] += sum_p m_p cw[y_p] nll_p,
 * acc[1] += sum_p cw[y_p], acc[2] += sum_p m_p  (cw == NULL: ones; mask == NULL: ones).  'mean' loss = acc[0] / acc[1];
 * the masked certain-area loss `(CE_none * mask).sum() / (mask.sum() + 1e-16)` (:301-303) = acc[0] / (acc[2] + 1e-16).
 * bwd: dz = gscale[0] * d(loss)/dz with the denominator chosen by mask_den */
int smsut_wce_fwd(const float* z, const int64_t* y, const float* cw, const float* mask, float* acc, int64_t npix, int32_t c,
                  smsut_stream_t stream);
int smsut_wce_bwd(const float* z, const int64_t* y, const float* cw, const float* mask, const float* acc,
                  const float* gscale, int32_t mask_den, float* dz, int64_t npix, int32_t c, smsut_stream_t stream);
/* `(softmax_mse_loss(zs, zt) * m).sum() / (m.sum() + 1e-16)` of coraNetTrainer.py:137-149,331-337 with m = mask
 * (invert = 0) or 1 - mask (invert = 1), mask (npix) broadcast over the classes: acc[0] += sum_p m_p sum_c (ps - pt)^2,
 * acc[1] += sum_p m_p; bwd writes dzs (zt is a constant) */
int smsut_softmax_mse_masked_fwd(const float* zs, const float* zt, const float* mask, int32_t invert, float* acc,
                                 int64_t npix, int32_t c, smsut_stream_t stream);
int smsut_softmax_mse_masked_bwd(const float* zs, const float* zt, const float* mask, int32_t invert, const float* acc,
                                 const float* gscale, float* dzs, int64_t npix, int32_t c, smsut_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SMSUT_B200_H_ */
