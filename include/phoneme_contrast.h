/*
 * phoneme_contrast.h -- C ABI of libpc_b200.so, the B200 (sm_100a) implementation of the
 * phoneme_contrast data-parallel training hot path.
 *
 * The reference (brant01/phoneme_contrast) is pure Python: it has no FFI. Its "plugin API" for this
 * path is four Python seams; each group of entry points below states which
 * reference interface it sits behind (paths relative to the reference repo). The Python host side
 * (phoneme_contrast_b200/) mirrors those interfaces and calls these symbols through ctypes with raw
 * device pointers -- no torch types cross this boundary. INTEGRATION.md shows the binding stubs.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - tensors are dense fp32; activations are NHWC ([B,H,W,C]); conv weights cross the ABI in the
 *     reference's OIHW layout and are re-packed on the device by pc_pack_conv_weight;
 *   - labels are int64 (torch.long), as trainer.py:189,197 produces them;
 *   - every launcher is asynchronous on `stream` (the caller's current CUDA stream), allocates
 *     nothing, and returns 0 on success or a negative PcStatus; pc_last_error() gives the text.
 *     No exception crosses the ABI; the Python shim turns PC_EINVAL into the ValueError the reference
 *     raises at the same place (losses.py:44-45, features.py:168, registry.py:18,27).
 */
#ifndef PHONEME_CONTRAST_H_
#define PHONEME_CONTRAST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* pc_stream_t;

typedef enum PcStatus {
  PC_OK = 0,
  PC_EINVAL = -1,      /* bad shape / argument (maps to ValueError)            */
  PC_EUNSUPPORTED = -2, /* configuration outside what the kernels cover         */
  PC_ECUDA = -3        /* CUDA launch / runtime error (maps to RuntimeError)    */
} PcStatus;

const char* pc_last_error(void);
int pc_abi_version(void);
/* Number of kernels this library has launched since load (bench.py "gpu_launches"). */
unsigned long long pc_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * 1. MFCC / log-mel front end + view augmentation
 *    replaces: MFCCExtractor.forward            src/datasets/features.py:61-103
 *              MelSpectrogramExtractor.forward  src/datasets/features.py:134-153
 *              (torchaudio MFCC: _transforms.py:701-718, functional.py:123-144,390-405)
 *              TimeMask/FrequencyMask/GaussianNoise/Compose  src/datasets/transforms.py:38-97,139-144
 *              waveform gain                    src/datasets/dataset.py:165-167
 * ---------------------------------------------------------------------------------------------- */
typedef struct PcViewDesc {   /* one per output view; built on the host with the reference's RNG calls */
  float gain;                 /* waveform gain (1.0 = none)                      dataset.py:165-167   */
  int32_t t0, t1;             /* time-mask columns [t0,t1) zero-filled (0,0 = none)  transforms.py:45 */
  int32_t f0, f1;             /* freq-mask rows    [f0,f1)                           transforms.py:69 */
  float noise_level;          /* x + N(0,1)*level after the masks (0 = none)         transforms.py:93-96 */
  uint32_t noise_seed;        /* Philox key when no explicit noise tensor is passed                   */
  int32_t clip;               /* source clip index in `wave`                                          */
} PcViewDesc;

typedef struct PcMfccConsts { /* device pointers, built once by the host (features.py:25-59)          */
  const float* window;        /* [n_fft]  periodic Hann                                               */
  const int32_t* fb_start;    /* [n_mels] first non-zero FFT bin of each mel filter                   */
  const int32_t* fb_len;      /* [n_mels] number of non-zero bins (<= PC_FB_MAXW)                      */
  const float* fb_w;          /* [n_mels][PC_FB_MAXW] filter weights                                   */
  const float* dct;           /* [n_mels][n_mfcc] DCT-II ortho (unused for log-mel)                    */
  const float* tw;            /* twiddles: cos/sin tables, see mfcc.cu                                 */
  int32_t n_fft, hop, n_mels, n_mfcc;
  int32_t fb_wmax;            /* max over filters of fb_len (0 = unknown: assume PC_FB_MAXW)                  */
  float preemph;              /* optional pre-emphasis coefficient a: y[n] = x[n] - a x[n-1], y[0] = x[0], applied before
                               * the reflect padding. 0 = off = the reference's code path (north_star names the stage;
                               * src/datasets/features.py has none, publication/sections/02_methods.md:32 quotes 0.97) */
} PcMfccConsts;
#define PC_FB_MAXW 32

enum { PC_FE_MFCC = 0, PC_FE_LOGMEL = 1 };
enum { PC_CLAMP_PER_CLIP = 0, PC_CLAMP_NONE = 1, PC_CLAMP_GIVEN = 2 };

/* wave [n_clips, S] (row stride `wave_ld` floats) -> out [n_views, n_out, T] with T = 1 + S/hop and
 * n_out = n_mfcc (MFCC) or n_mels (log-mel). views == NULL means one un-augmented view per clip.
 * noise (optional) [n_views, n_out*T]: explicit N(0,1) draws (parity with the torch CPU generator);
 * otherwise Philox noise is generated on the device. clamp_mode PER_CLIP is the dataset's semantics
 * (one clip per call, dataset.py:90); GIVEN takes the floor reference from clamp_ref[0] (device),
 * which reproduces a batched MFCCExtractor call whose amax spans the batch (functional.py:396-399);
 * clip_max_out (optional) [n_views] receives each view's max dB. */
int pc_frontend_fwd(const float* wave, int n_clips, int S, int wave_ld, const PcMfccConsts* consts_host,
                    const PcViewDesc* views, int n_views, const float* noise, int kind, int clamp_mode,
                    float top_db, const float* clamp_ref, float* clip_max_out, float* out, pc_stream_t stream);
/* max over n floats -> out[0] (used for the whole-call clamp). */
int pc_reduce_max(const float* x, int n, float* out, pc_stream_t stream);
/* Apply masks + noise to existing features x [n_views, F, T] (BaseTransform.__call__ on a tensor,
 * transforms.py:12-22): out may alias x. */
int pc_augment_apply(const float* x, const PcViewDesc* views, int n_views, int F, int T, const float* noise,
                     float* out, pc_stream_t stream);
/* 5-tap regression deltas along T (features.py:82-95; torchaudio compute_deltas), rows = n*F. */
int pc_compute_deltas(const float* x, int rows, int T, float* out, pc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 2. Supervised contrastive loss
 *    replaces: SupervisedContrastiveLoss.forward + autograd backward  src/training/losses.py:26-86
 * ---------------------------------------------------------------------------------------------- */
/* Row block [row0,row0+nrows) of the N x N problem. F [N,D]; labels [N] (or NULL with mask [N,N] float,
 * losses.py:27,52). stats [nrows,4] = (row max m, den incl. 1e-6, n_pos, sum_pos(z-m)); row_loss [nrows]
 * = -(T/T_base) * mean_log_prob_pos (losses.py:76-79). The N x N logits never reach HBM. */
int pc_supcon_fwd(const float* F, const int64_t* labels, const float* mask, int N, int D, int row0, int nrows,
                  float temperature, float base_temperature, float* stats, float* row_loss, pc_stream_t stream);
/* Data-parallel exchange helpers (csrc/dp.cu). pack: [n][D] embeddings + [n] int64 labels -> [n][D+2] fp32 rows (label bits in the
 * last two columns) so that ONE all_gather carries both; unpack: gathered [N][D+2] -> contiguous F [N][D] and labels [N].
 * loss_from_stats: out[0] = scale * sum of the per-row losses implied by the gathered row statistics [N][4] (scale = 1/N for the
 * reference's mean over all rows, losses.py:81-82) -- every rank derives the identical global loss without a further collective. */
int pc_dp_pack(const float* emb, const int64_t* labels, int n, int D, float* packed, pc_stream_t stream);
int pc_dp_unpack(const float* packed, int N, int D, float* F, int64_t* labels, pc_stream_t stream);
int pc_supcon_loss_from_stats(const float* stats, int N, float temperature, float base_temperature, float scale, float* out,
                              pc_stream_t stream);
/* loss[0] = scale * sum(row_loss[0..n)) in a fixed order (scale = 1/N for 'mean', losses.py:81-84). */
int pc_sum_scaled(const float* x, int n, float scale, float* out, pc_stream_t stream);
/* dF[row0..row0+nrows) given the stats of ALL N rows (stats_all [N,4]); coef = (T/T_base)*grad_out/N
 * for 'mean'. grad_scale (optional, device scalar) multiplies coef (upstream gradient). */
int pc_supcon_bwd(const float* F, const int64_t* labels, const float* mask, int N, int D, int row0, int nrows,
                  float temperature, float coef, const float* grad_scale, const float* stats_all, float* dF,
                  pc_stream_t stream);

/* Tensor-core (tcgen05, FP16x2 operand split) forward of the label form: same outputs as pc_supcon_fwd. Eligible when
 * pc_supcon_tc_supported() != 0 (N >= 128, D = 64 | 128, row0 % 8 == 0); workspace >= pc_supcon_tc_workspace(N, D, nrows)
 * bytes, 128-byte aligned (packed fp16 hi/lo image of F + per-column-split partial row statistics). */
int pc_supcon_tc_supported(int N, int D, int row0, int nrows);
size_t pc_supcon_tc_workspace(int N, int D, int nrows);
int pc_supcon_fwd_tc(const float* F, const int64_t* labels, int N, int D, int row0, int nrows, float temperature,
                     float base_temperature, void* workspace, size_t workspace_bytes, float* stats, float* row_loss,
                     pc_stream_t stream);
/* Tensor-core backward of the label form: same dF as pc_supcon_bwd (S recomputed per tile in TMEM, W' = dL/dS turned into the
 * fp16 hi/lo A operand of a second MMA that accumulates dF in TMEM; column splits reduced in a fixed order). */
size_t pc_supcon_bwd_tc_workspace(int N, int D, int nrows);
int pc_supcon_bwd_tc(const float* F, const int64_t* labels, int N, int D, int row0, int nrows, float temperature, float coef,
                     const float* grad_scale, const float* stats_all, void* workspace, size_t workspace_bytes, float* dF,
                     pc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 3. CNN building blocks (PhonemeNet / PhonemeNetDeep forward + backward)
 *    replaces: nn.Conv2d / BatchNorm2d / ReLU / MaxPool2d / Dropout2d / SpatialAttention /
 *              AdaptiveAvgPool2d / Linear / BatchNorm1d / F.normalize as composed in
 *              src/models/phoneme_cnn.py:33-77,98-126 (PhonemeNet), :129-143, :146-184, :209-272,274-304
 * ---------------------------------------------------------------------------------------------- */
typedef struct PcConvGeom {
  int32_t B, H, W, Cin;       /* input  [B,H,W,Cin]   */
  int32_t Ho, Wo, Cout;       /* output [B,Ho,Wo,Cout] */
  int32_t R, S, stride, pad;
} PcConvGeom;

/* Per-channel input transform applied while a conv reads its input, so that the activation
 * a = drop * relu(scale*y + shift) of the previous BatchNorm never has to be materialised:
 * scale/shift [C] (NULL = identity), relu flag, drop [B,C] multipliers (NULL = none). */
typedef struct PcInXform {
  const float* scale;
  const float* shift;
  const float* drop;
  int32_t relu;
  int32_t presplit;   /* != 0: `x` is not fp32 but the fp16 hi | lo planes written by pc_bn_act_split (the transform is already
                         applied; scale/shift/drop must be NULL, relu 0). PC_PREC_FP16X2 tensor-core paths only. */
} PcInXform;

/* Train-mode BatchNorm finalisation folded into the kernel that first consumes the coefficients (pc_bn_act_split_fin,
 * pc_bn_add_relu_fwd_fin): stats = the [2][C] fp64 sums (sum y, sum y^2) a convolution accumulated over `count` elements per
 * channel. Every block derives scale / shift itself; block 0 writes scale, shift (required), mean, invstd (may be NULL), updates
 * running_mean / running_var (may be NULL, both or neither) with `momentum` and increments num_batches_tracked (may be NULL).
 * Same arithmetic as pc_bn_finalize(training = 1); saves one dependent launch per BatchNorm layer. */
typedef struct PcBnFinalize {
  const double* stats;
  double count;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  int64_t* num_batches_tracked;
  float momentum;
  float eps;
  float* scale;
  float* shift;
  float* mean;
  float* invstd;
} PcBnFinalize;

/* a = drop * relu(scale*y + shift) written ONCE in the tensor-core operand form: planes[0][n_pix*C] = fp16 hi, planes[1] = fp16
 * lo * 2^11 (see PC_PREC_FP16X2). A convolution that reads its input through `presplit` then only copies bytes instead of
 * redoing this arithmetic for every tap and output-channel tile. hw = pixels per sample (row of `drop` = pixel / hw). */
int pc_bn_act_split(const float* y, int64_t n_pix, int C, int hw, const float* scale, const float* shift, const float* drop,
                    int relu, void* planes, pc_stream_t stream);

/* FP16X2 activation planes are written unscaled; every plane writer (pc_bn_act_split, the `planes` outputs of pc_bn_act_fwd /
 * pc_bn_add_relu_fwd) raises a sticky device flag when a value exceeds fp16's range (|a| > 65504, or NaN). This call copies the
 * flag to *host_flag (synchronising `stream`) and clears it when reset != 0. A set flag means the step's convolutions saw inf:
 * re-run with PC_PREC_TF32X3, which has no range assumption (ContrastiveTrainer does this at its per-epoch read-back). */
int pc_f16_overflow_query(int reset, int* host_flag, pc_stream_t stream);

/* OIHW fp32 -> fwd layout Wf [(r,s,c)][o] and dgrad layout Wd [(r,s,o)][c] (either may be NULL). */
int pc_pack_conv_weight(const float* w_oihw, int O, int I, int R, int S, float* wf, float* wd, pc_stream_t stream);

/* Precision of the implicit-GEMM convolutions:
 *   PC_PREC_FP32   exact-fp32 SIMT kernels (operands packed by pc_pack_conv_weight);
 *   PC_PREC_TF32X3 tcgen05 tensor cores, fp32 operands split hi+lo, 3 kind::tf32 MMAs per k-step (fp32-level accuracy);
 *   PC_PREC_FP16X2 tcgen05 tensor cores, fp32 operands split x = hi + lo*2^-11 with hi, lo in fp16, 3 products per k-step at
 *                  the kind::f16 rate (same ~22-bit operand precision as TF32X3). Gradient operands are pre-scaled by a
 *                  power of two taken from the tensor's max magnitude (the `*_amax` arguments below) to fit fp16's range;
 *                  forward activations must stay below 65504 in magnitude (they follow a BatchNorm);
 *   PC_PREC_BF16   tcgen05 tensor cores, operands rounded to bf16 (1e-2 tolerance mode).
 * For the two tensor-core modes the weight operand is the pre-swizzled tile image written by
 * pc_pack_conv_weight_tc, and a layer is eligible when pc_conv_tc_supported() != 0 (gathered channels a multiple of
 * 32 / 64); ineligible layers must be run with PC_PREC_FP32. */
enum { PC_PREC_FP32 = 0, PC_PREC_TF32X3 = 1, PC_PREC_BF16 = 2, PC_PREC_FP16X2 = 3 };
int pc_conv_tc_supported(const PcConvGeom* g, int dgrad, int prec);
size_t pc_conv_tc_packed_bytes(int O, int I, int R, int S, int dgrad, int prec);
int pc_pack_conv_weight_tc(const float* w_oihw, int O, int I, int R, int S, int dgrad, int prec, void* out, pc_stream_t stream);
/* All layers' weight operands in ONE launch (they are re-packed after every optimiser step). jobs: DEVICE array; each job is
 * what one pc_pack_conv_weight_tc call would do; item_begin = running sum of pc_pack_conv_weight_tc_items() over the
 * preceding jobs, total_items = the sum over all jobs. */
typedef struct PcPackJob {
  const float* w_oihw;
  void* out;
  int32_t O, I, R, S, dgrad, prec;
  int64_t item_begin;
} PcPackJob;
int64_t pc_pack_conv_weight_tc_items(int O, int I, int R, int S, int dgrad, int prec);
int pc_pack_conv_weights_tc_batch(const PcPackJob* jobs, int n_jobs, int64_t total_items, pc_stream_t stream);
/* C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) through the same tcgen05 tile engine (unit check of the tensor-core path;
 * also usable as a stand-alone fp32-accurate GEMM). ws >= pc_tc_gemm_workspace(N, K, prec) bytes. */
size_t pc_tc_gemm_workspace(int N, int K, int prec);
int pc_tc_gemm(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int prec, void* ws,
               size_t ws_bytes, pc_stream_t stream);

/* y = conv(xform(x)) + bias; if stats != NULL accumulates per-channel sum / sum of squares of y into
 * stats [2][Cout] (fp64) for the following train-mode BatchNorm. Cin == 1 uses a direct kernel (wf = the OIHW
 * weight itself); otherwise Cin % 16 == 0 and the implicit-GEMM kernel of the requested precision runs. */
int pc_conv_fwd(const float* x, const float* wf, const float* bias, const PcConvGeom* g, const PcInXform* xf,
                float* y, double* stats, int prec, pc_stream_t stream);
/* dx (+)= conv_transpose(dy, w): accumulate != 0 adds into dx. dy_amax (may be NULL): device scalar holding max|dy|, written by
 * the BatchNorm-backward apply calls below; PC_PREC_FP16X2 derives its power-of-two operand scale from it (NULL: scale 1). */
int pc_conv_dgrad(const float* dy, const float* wd, const PcConvGeom* g, float* dx, int accumulate, int prec,
                  const float* dy_amax, int dy_presplit, pc_stream_t stream);
/* Halo-resident engine for stride-1 3x3 pad-1 layers on the FP16X2 planes (csrc/conv_halo.cu): persistent CTAs, the activation
 * tile + halo loaded once per 64-channel chunk by tiled TMA boxes (zero padding materialised by out-of-bounds fill) and shared
 * by all 9 taps through row-shifted UMMA descriptors, weight stages multicast across a cluster (PC_HALO_CLUSTER = 1 | 2 | 4),
 * double-buffered TMEM accumulators. pc_conv_fwd / pc_conv_dgrad route to it automatically when it covers the layer
 * (pc_conv_halo_supported; PC_CONV_HALO=0 disables); the weight operand is the one pc_pack_conv_weight_tc writes. Same
 * arithmetic (three fp16 products per operand pair) and the same outputs as the per-tap-gather kernels. */
int pc_conv_halo_supported(const PcConvGeom* g, int dgrad);
int pc_conv_fwd_halo(const void* x_planes, const void* wp, const float* bias, const PcConvGeom* g, float* y, double* stats,
                     pc_stream_t stream);
int pc_conv_dgrad_halo(const void* dy_planes, const void* wp, const PcConvGeom* g, float* dx, int accumulate,
                       const float* dy_amax, pc_stream_t stream);
/* dy_presplit != 0 (here and in pc_conv_wgrad): `dy` points to the fp16 hi | lo planes written by pc_bn_*_bwd_apply (dy_planes),
 * scaled by the power of two derived from *dy_amax; the gather copies bytes and the epilogue undoes the scale. */
/* dw (OIHW) and db from x (through xform) and dy. workspace >= pc_conv_wgrad_workspace(g) bytes. */
size_t pc_conv_wgrad_workspace(const PcConvGeom* g);
int pc_conv_wgrad(const float* x, const float* dy, const PcConvGeom* g, const PcInXform* xf, float* dw_oihw,
                  float* db, void* workspace, size_t workspace_bytes, int prec, const float* dy_amax, int dy_presplit,
                  pc_stream_t stream);

/* Halo weight-gradient engine (csrc/conv_halo_wgrad.cu) for stride-1 3x3 pad-1 layers with both operands as FP16X2 planes
 * (Cin % 64 == 0, Cout == 64 or Cout % 128 == 0): x and dy are loaded once per position tile by tiled TMA boxes (zero padding by
 * out-of-bounds fill) and the nine taps come from UMMA descriptors -- the three column taps as N groups one position apart, the
 * row taps as separate work units (or, for 64 output channels, as M groups one padded row apart). pc_conv_wgrad routes here when
 * prec == PC_PREC_FP16X2, xf->presplit, dy_presplit and db == NULL (PC_WGRAD_HALO=0 disables); dw only, no bias gradient. */
int pc_conv_wgrad_halo_supported(const PcConvGeom* g);
size_t pc_conv_wgrad_halo_workspace(const PcConvGeom* g);
int pc_conv_wgrad_halo(const void* x_planes, const void* dy_planes, const PcConvGeom* g, float* dw_oihw, void* workspace,
                       size_t workspace_bytes, const float* dy_amax, pc_stream_t stream);

/* Stem backward without the full-resolution tensors (csrc/stem_bwd.cu; reference init_conv, src/models/phoneme_cnn.py:211-216):
 * for Conv2d(1, 64, 7, pad 3) -> BatchNorm2d -> ReLU -> MaxPool2d(3, 2, 1) the gradient from the pool is non-zero only at each
 * window's argmax pixel, where the BatchNorm output equals the pooled output; the dense BatchNorm-projection terms of the weight
 * gradient are closed forms in the 49 x 49 Gram matrix G and the tap sums X1 of the input patches (y0 is linear in them). So
 * dw, db (= 0 exactly), dgamma, dbeta follow from dpool / p0 / argmax at POOLED resolution plus x; the pre-BatchNorm tensor
 * is never read. pc_stem_gram accumulates G [49*49] and X1 [49] (fp64; zero them first) from x [B][H][W]; pc_stem_bwd needs
 * zeroed `sums` [2][64] fp64 and `amax_slot` (float) and a workspace of pc_stem_bwd_workspace() bytes. */
int pc_stem_bwd_supported(int k, int Cout, int H, int W);
int pc_stem_gram(const float* x, int B, int H, int W, double* G, double* X1, pc_stream_t stream);
size_t pc_stem_bwd_workspace(void);
int pc_stem_bwd(const float* dpool, const float* p0, const uint8_t* argmax, const float* x, int B, int H, int W,
                const float* w_oihw, const float* bias, const float* gamma, const float* scale, const float* shift,
                const float* mean, const float* invstd, const double* G, const double* X1, double* sums, float* amax_slot,
                void* workspace, size_t workspace_bytes, float* dw, float* db, float* dgamma, float* dbeta, pc_stream_t stream);

/* Stem forward in ONE pass (csrc/stem_fwd.cu): the batch statistics of y0 = conv7x7(x) + b are closed forms in the Gram matrix /
 * tap sums of the input patches (pc_stem_gram), so BatchNorm's coefficients are known before the convolution runs and the kernel
 * -- tcgen05 implicit GEMM over 128-pixel row tiles, operand tiles built in shared memory from a staged input window -- applies
 * BatchNorm + ReLU + MaxPool2d(3, 2, 1) in its epilogue and stores only the pooled tensor (+ argmax, + optional fp16 hi | lo
 * planes). pc_stem_stats_from_gram: stats [2][64] fp64 = (sum y0, sum y0^2), the input of pc_bn_finalize. */
int pc_stem_fwd_supported(int k, int Cout, int H, int W);
int pc_stem_stats_from_gram(const double* G, const double* X1, const float* w_oihw, const float* bias, int B, int H, int W,
                            double* stats, const PcBnFinalize* fin /* may be NULL: also finalise the BatchNorm (train mode) */,
                            pc_stream_t stream);
int pc_stem_fwd(const float* x, const float* w_oihw, const float* bias, const float* scale, const float* shift, int B, int H, int W,
                float* p0, uint8_t* argmax, void* planes, pc_stream_t stream);

/* BatchNorm statistics -> per-channel coefficients.
 * training: mean/var from stats (count = elements per channel), running stats updated with `momentum`
 * (unbiased variance) exactly like nn.BatchNorm2d; eval: coefficients from the running stats.
 * Outputs: scale = gamma*invstd, shift = beta - mean*scale, mean, invstd (all [C]). */
int pc_bn_finalize(const double* stats, int C, double count, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum, float eps,
                   int training, float* scale, float* shift, float* mean, float* invstd, pc_stream_t stream);

/* Fused BatchNorm-apply + ReLU (+ max-pool) (+ Dropout2d multiplier) producing a materialised activation.
 * pool: 0 none, 2 = MaxPool2d(2,2) (phoneme_cnn.py:42,52), 3 = MaxPool2d(3,2,1) (:215; writes argmax u8). */
int pc_bn_act_fwd(const float* y, int B, int H, int W, int C, const float* scale, const float* shift,
                  const float* drop, int pool, float* out, uint8_t* argmax, void* planes, pc_stream_t stream);
/* `planes` (may be NULL; here and in pc_bn_add_relu_fwd): additionally write `out` as fp16 hi | lo planes (the layout of
 * pc_bn_act_split, numel(out) elements per plane) for a following convolution that reads its input with presplit != 0. */
/* Backward of the above w.r.t. y: two passes. pass 1 accumulates sums[0][C] = sum dz, sums[1][C] = sum dz*xhat (fp64);
 * pass 2 writes dy = scale*(dz - sum_dz/M - xhat*sum_dzxhat/M) and dgamma/dbeta. dout is the gradient w.r.t. `out`.
 * dy_amax (may be NULL): zero-initialised device scalar that receives max|dy| (atomic max). */
int pc_bn_act_bwd_reduce(const float* dout, const float* y, int B, int H, int W, int C, const float* scale,
                         const float* shift, const float* mean, const float* invstd, const float* drop, int pool,
                         const uint8_t* argmax, double* sums, float* maxes, pc_stream_t stream);
/* maxes (may be NULL): zero-initialised float[2] receiving max|dz| and max|xhat| (atomic max); needed by dy_planes below. */
int pc_bn_act_bwd_apply(const float* dout, const float* y, int B, int H, int W, int C, const float* scale,
                        const float* shift, const float* mean, const float* invstd, const float* drop, int pool,
                        const uint8_t* argmax, const double* sums, float* dy, float* dgamma, float* dbeta,
                        float* dy_amax, const float* maxes, void* dy_planes, const double* y_stats, float* db_conv,
                        pc_stream_t stream);
/* db_conv (may be NULL; with y_stats = the [2][C] fp64 sums of y the forward pass accumulated): gradient of the bias of the
 * convolution that produced y, i.e. sum_p dy, in closed form from the per-channel constants -- no pass over dy. The bias is
 * absorbed by the batch mean, so this is the fp32 round-off residual the reference reports there (csrc/bn_act.cu:
 * bias_grad_closed_form); pc_conv_wgrad can then be called with db = NULL and skips its column-sum kernel. */
/* dy_planes (may be NULL): write dy (also) in tensor-core operand form -- fp16 hi | lo planes (pc_bn_act_split layout) of
 * dy * 2^k, with 2^k derived from a bound of |dy| computed from `maxes` and `sums`; the bound is stored in dy_amax (which
 * pc_conv_dgrad / pc_conv_wgrad with dy_presplit != 0 read to undo the scale). `dy` itself may then be NULL. */

/* pc_bn_act_split / pc_bn_add_relu_fwd with the BatchNorm coefficients finalised inside the kernel (PcBnFinalize above).
 * fin_s == NULL: identity shortcut (ysc is added as it is). */
int pc_bn_act_split_fin(const float* y, int64_t n_pix, int C, int hw, const PcBnFinalize* fin, const float* drop, int relu,
                        void* planes, pc_stream_t stream);
int pc_bn_add_relu_fwd_fin(const float* y2, const PcBnFinalize* fin2, const float* ysc, const PcBnFinalize* fin_s, int64_t n_pix,
                           int C, float* out, void* planes, pc_stream_t stream);
/* pc_bn_act_fwd (BatchNorm + ReLU + pool + dropout, optional planes) with the coefficients finalised inside the kernel. */
int pc_bn_act_fwd_fin(const float* y, int B, int H, int W, int C, const PcBnFinalize* fin, const float* drop, int pool, float* out,
                      uint8_t* argmax, void* planes, pc_stream_t stream);

/* Residual tail: out = relu(bn2(y2) + (sc_scale ? bn_s(ysc) : ysc))   (phoneme_cnn.py:177-182). */
int pc_bn_add_relu_fwd(const float* y2, const float* scale2, const float* shift2, const float* ysc,
                       const float* sc_scale, const float* sc_shift, int64_t n_pix, int C, float* out,
                       void* planes, pc_stream_t stream);
/* g = dout * (out > 0); pass 1: sums2[2][C] over (g, y2) and sums_s[2][C] over (g, ysc) (sums_s NULL for identity);
 * pass 2: dy2, dysc (or, identity shortcut, dsc (+)= g into dx_identity), dgamma/dbeta for both norms. */
int pc_bn_add_relu_bwd_reduce(const float* dout, const float* out, const float* y2, const float* mean2,
                              const float* invstd2, const float* ysc, const float* mean_s, const float* invstd_s,
                              int64_t n_pix, int C, double* sums2, double* sums_s, float* maxes, pc_stream_t stream);
int pc_bn_add_relu_bwd_apply(const float* dout, const float* out, const float* y2, const float* scale2,
                             const float* mean2, const float* invstd2, const double* sums2, const float* ysc,
                             const float* sc_scale, const float* mean_s, const float* invstd_s, const double* sums_s,
                             int64_t n_pix, int C, float* dy2, float* dysc_or_dx, float* dgamma2, float* dbeta2,
                             float* dgamma_s, float* dbeta_s, float* dy2_amax, float* dysc_amax, const float* maxes,
                             void* dy2_planes, void* dysc_planes, const double* y2_stats, float* db2, const double* ysc_stats,
                             float* db_s, pc_stream_t stream);
/* maxes: float[3] (max|g|, max|xhat2|, max|xhat_s|) from the reduce pass; *_planes as dy_planes of pc_bn_act_bwd_apply;
 * db2 / db_s (may be NULL) with y2_stats / ysc_stats: bias gradients of conv2 / the shortcut convolution as db_conv there. */

/* SpatialAttention + AdaptiveAvgPool2d(1): pooled[b,c] = mean_p a[b,p,c] * sigmoid(w.a[b,p,:] + b0)
 * (phoneme_cnn.py:134-143,117-118). w == NULL: plain mean (use_attention False). gate [B,HW] saved for backward. */
int pc_attn_pool_fwd(const float* a, int B, int HW, int C, const float* w, const float* b0, float* gate,
                     float* pooled, pc_stream_t stream);
/* The same with a workspace that lets S = pc_attn_pool_splits(B) blocks share one sample's pixels (small batches leave SMs idle with
 * one block per sample): scratch >= B * S * C floats; counters >= B unsigned ints, ZERO before the first call (the kernel leaves them
 * zero). The S partial sums are combined in a fixed order by the block that finishes last, so the result is deterministic. */
int pc_attn_pool_splits(int B);
int pc_attn_pool_fwd_ws(const float* a, int B, int HW, int C, const float* w, const float* b0, float* gate, float* pooled,
                        float* scratch, unsigned int* counters, pc_stream_t stream);
int pc_attn_pool_bwd(const float* a, const float* gate, const float* dpooled, int B, int HW, int C, const float* w,
                     float* da, float* dw, float* db0, pc_stream_t stream);

/* Projection head: z = x W^T + b (Linear, [N,K] weight) ; zn = BatchNorm1d(z) ; e = zn / max(||zn||, 1e-12)
 * (phoneme_cnn.py:75-77,121-124). ws: >= pc_head_workspace(B,K,N) bytes, kept until the backward. */
size_t pc_head_workspace(int B, int K, int N);
int pc_head_fwd(const float* x, int B, int K, int N, const float* W, const float* bias, const float* gamma,
                const float* beta, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                float momentum, float eps, int training, float* emb, void* ws, pc_stream_t stream);
int pc_head_bwd(const float* demb, const float* x, int B, int K, int N, const float* W, const float* gamma,
                const float* beta, int training, void* ws, float* dx, float* dW, float* dbias, float* dgamma,
                float* dbeta, pc_stream_t stream);

/* Dropout2d multipliers drop[B,C] in {0, 1/(1-p)} from Philox(seed, offset) (phoneme_cnn.py:43,53,62,176).
 * step_dev (optional, device int64): its value << 24 is added to offset, so a CUDA-graph replay draws fresh masks. */
int pc_dropout2d_mask(float* drop, int B, int C, float p, uint64_t seed, uint64_t offset, const int64_t* step_dev,
                      pc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 4. Optimiser step on flat buffers
 *    replaces: clip_grad_norm_ (src/training/trainer.py:147-150) + torch.optim.Adam.step
 *              (scripts/train.py:129-133; L2 weight decay, betas .9/.999, eps 1e-8)
 * ---------------------------------------------------------------------------------------------- */
/* norm_sq[0] += sum g^2 (fp64; caller zeroes it). */
int pc_grad_sumsq(const float* g, int64_t n, double* norm_sq, pc_stream_t stream);
/* In-place Adam on n elements. grad is first multiplied by grad_prescale (e.g. 1/world_size) and by the
 * clip coefficient min(1, max_norm/(sqrt(norm_sq*prescale^2)+1e-6)) when max_norm > 0. `step` is 1-based. */
int pc_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                 float eps, float weight_decay, float max_norm, const double* norm_sq, float grad_prescale,
                 int64_t step, pc_stream_t stream);

/* Graph-capturable form: learning rate and 1-based step count are read from device memory (lr_dev[0], step_dev[0]),
 * so the launch can be captured once in a CUDA graph and replayed while the host updates those scalars. */
int pc_clip_adam_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev, float beta1, float beta2,
                     float eps, float weight_decay, float max_norm, const double* norm_sq, float grad_prescale,
                     const int64_t* step_dev, pc_stream_t stream);
/* counter[0] += inc on the stream (device-side step counters for captured steps). */
int pc_counter_add(int64_t* counter, int64_t inc, pc_stream_t stream);

/* Fused BatchNorm-backward reduce (csrc/conv_halo.cu): the stride-1 3x3 data gradient that produces dx = dL/d(drop * relu(bn(y)))
 * also accumulates, for the pixels it writes, sums [2][C] fp64 = (sum dz, sum dz xhat) and maxes [2] = (max |dz|, max |xhat|) with
 * dz = [bn(y) > 0] * dx * drop, xhat = (y - mean) * invstd -- the outputs of pc_bn_act_bwd_reduce(dx, y, ..., pool = 0), which is then
 * skipped (one read of dx and y less per layer). sums / maxes must be zeroed by the caller. PC_EUNSUPPORTED when the layer is not on
 * the halo engine (pc_conv_halo_supported). Reference: the autograd backward of ResidualBlock.forward, src/models/phoneme_cnn.py:173-184. */
typedef struct PcBnBwdReduce {
  const float* y;        /* [B,H,W,C] pre-BatchNorm tensor of the layer below                      */
  const float* scale;    /* [C] train-mode BatchNorm coefficients of that layer                    */
  const float* shift;
  const float* mean;
  const float* invstd;
  const float* drop;     /* [B,C] Dropout2d multipliers or NULL                                    */
  double* sums;          /* [2][C] out                                                             */
  float* maxes;          /* [2] out or NULL                                                        */
} PcBnBwdReduce;
int pc_conv_dgrad_halo_bnred(const void* dy_planes, const void* wp, const PcConvGeom* g, float* dx, const float* dy_amax,
                             const PcBnBwdReduce* red, pc_stream_t stream);

/* Two-phase forms of pc_head_fwd / pc_head_bwd for BatchNorm1d statistics synchronised across data-parallel ranks (train mode): the caller
 * exchanges bn_sums [2][N] fp64 between the phases. forward phase 1: Linear + (sum z, sum z^2) -> bn_sums; phase 2: mean / invstd /
 * running statistics from bn_sums over `count` rows (the GLOBAL batch) + normalise. backward phase 1: dL/d(bn output) + (sum d, sum d xhat)
 * -> bn_sums; phase 2: BatchNorm backward with bn_sums taken as the per-rank AVERAGE of the global sums (so that B stays the local row
 * count and dgamma / dbeta / dbias come out as this rank's share of the all-reduced gradient) + dW + dx. */
int pc_head_fwd_sync(const float* x, int B, int K, int N, const float* W, const float* bias, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum, float eps, float* emb, void* ws,
                     double* bn_sums, double count, int phase, pc_stream_t stream);
int pc_head_bwd_sync(const float* demb, const float* x, int B, int K, int N, const float* W, const float* gamma, const float* beta, void* ws,
                     float* dx, float* dW, float* dbias, float* dgamma, float* dbeta, double* bn_sums, int phase, pc_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * 8. Data-parallel exchanges over NVLink peer memory (csrc/peer.cu; SURVEY.md 8e C1 / C1' / C2 -- absent in the
 *    single-process reference: the hot loop they serve is src/training/trainer.py:126-164)
 *    Every rank owns one "peer region" of identical layout, exported with pc_peer_export and mapped by the other
 *    ranks of the node with pc_peer_open; `bases` is a HOST array of the R region addresses as seen from the calling
 *    process (own region at index `rank`). The exchanges are ordinary stream-ordered kernels (graph-capturable, no host
 *    participation, no NCCL).
 * ---------------------------------------------------------------------------------------------- */
/* Bytes of the flag block a region reserves at `flag_off` (zero it before the first barrier); most ranks supported. */
int pc_peer_flag_bytes(void);
int pc_peer_max_ranks(void);
/* IPC handle (64 bytes) of the cudaMalloc block containing ptr, and ptr's byte offset inside it. */
int pc_peer_export(const void* ptr, unsigned char* handle64, size_t* offset);
/* A zero-filled cudaMalloc block of the library's own, for callers whose allocator's memory cannot be exported (virtual-memory backed). */
int pc_peer_alloc(size_t bytes, void** ptr);
int pc_peer_free(void* ptr);
/* Map a peer's block into this process (peer access enabled lazily); *base = address of the block's first byte. */
int pc_peer_open(const unsigned char* handle64, void** base);
int pc_peer_close(void* base);
/* Flag barrier between the R ranks on `channel` (0 or 1; each has its own epoch). Writes of earlier kernels of the stream, local or
 * to peer regions, are visible to every rank's kernels that follow its own barrier. After timeout_ms without a peer's arrival the
 * region's sticky error word is set (1 + the missing rank) and this and all later barriers return without waiting. */
int pc_peer_barrier(const unsigned long long* bases, int R, int rank, size_t flag_off, int channel, int timeout_ms, pc_stream_t stream);
/* *out = the sticky error word (0 = none); synchronises the stream. */
int pc_peer_error(const unsigned long long* bases, int R, int rank, size_t flag_off, int reset, int* out, pc_stream_t stream);
/* The embedding / label all_gather as a store loop: emb [n][D] (D % 4 == 0) -> rows [row0, row0+n) of the [N][D] fp32 field at byte
 * offset f_off, labels [n] -> elements [row0, row0+n) of the [N] int64 field at y_off, of EVERY region. */
int pc_dp_gather_peer(const float* emb, const int64_t* labels, int n, int D, const unsigned long long* bases, int R, size_t f_off, size_t y_off,
                      int row0, pc_stream_t stream);
/* bytes (multiple of 16) from src to byte offset dst_off of every region. */
int pc_peer_bcast(const void* src, size_t bytes, const unsigned long long* bases, int R, size_t dst_off, pc_stream_t stream);
/* In-place sum over the ranks of `count` floats (multiple of 4) at byte offset off of every region: rank r reduces slice r in
 * rank order and stores it to all regions (results bit-identical on every rank). The caller brackets it with barriers. */
int pc_peer_allreduce(const unsigned long long* bases, int R, int rank, size_t off, long long count, int blocks, pc_stream_t stream);
/* One-shot all-reduce of a small fp64 block (synchronised BatchNorm statistics): after every rank has stored its [n]
 * block into slot `rank` of the [R][n] LOCAL slot array of every region (pc_peer_bcast) and a barrier, dst[i] = scale * sum_r slots[r][i]. */
int pc_peer_sum_slots(const double* slots, int R, int n, double scale, double* dst, pc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PHONEME_CONTRAST_H_ */
