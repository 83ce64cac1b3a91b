/*
 * sdt_b200.h -- C ABI of libsdt_b200.so: the B200 (sm_100a) kernels behind the Voice2Pose / Pose2Pose
 * training-step hot path of ShenhanQian/SpeechDrivesTemplates.
 *
 * Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in _host;
 *   - every entry point takes the CUDA stream it must launch on (cudaStream_t passed as void*), is
 *     re-entrant (backward runs on autograd worker threads), allocates no device memory and keeps no
 *     device state: workspaces are caller-owned;
 *   - returns 0 on success; non-zero = error, text in sdt_last_error() (thread-local). Nothing throws
 *     across the ABI;
 *   - activations are CHANNELS-LAST fp32: 2-D maps (B,H,W,C), 1-D sequences (B,L,C) == (B,1,L,C).
 *     The reference's NCHW/NCL tensors only exist at the Python boundary (weights, which keep the
 *     reference's (Cout,Cin,kh,kw) layout in HBM because they are the checkpoint/optimizer contract).
 *
 * The reference has no FFI: its arithmetic is torch/torchaudio library calls.  Each entry point below
 * cites the reference call site (file:line under /root/reference) whose computation it replaces.
 */
#ifndef SDT_B200_H
#define SDT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDT_OK 0
#define SDT_ERR_ARG 1
#define SDT_ERR_CUDA 2

/* ---- library ---------------------------------------------------------------------------------- */
const char* sdt_last_error(void);
int sdt_version(void);
/* math mode of the dense convolutions.  The PROCESS DEFAULT below applies to descriptors whose `math` field is 0; it starts at
 * 3 (this library runs on sm_100a only, and TF32 is the reference's own GPU default: torch.backends.cudnn.allow_tf32) unless the
 * environment variable SDT_CONV_MATH=0..4 says otherwise.  Engines pass their own mode per descriptor (sdt_conv_desc.math).
 *   0 = fp32 FFMA (SIMT);
 *   1 = tcgen05 tensor cores, TF32 operands, fp32 accumulation in TMEM, operand tiles built by producer warps
 *       (supports the loader transform) -- for shapes with C % 32 == 0 and N in {64,128,256}, FFMA otherwise;
 *   2 = as 1, plus TMA (cp.async.bulk.tensor) operand delivery for forward / data-gradient launches whose source is a
 *       plain tensor (no loader transform): the caller materialises activations for those layers;
 *   3 = as 2, plus operand reuse in shared memory for 2-D maps (csrc/tc_conv_ytap.cu): one TMA box serves all the
 *       vertical taps of a kernel column and one weight box serves several accumulators -- the L2 -> shared-memory
 *       traffic, which bounds mode 2, drops 2-3x.  Same arithmetic as mode 2 (same products, fp32 accumulation in a
 *       different order);
 *   4 = as 3, plus CTA pairs (tcgen05 cta_group::2, 256-row MMAs, each CTA supplying half of the weight tile) for 2-D
 *       convolutions with N % 128 == 0 (csrc/tc_conv_pair.cu) -- experimental. */
int sdt_set_conv_math(int mode);
int sdt_get_conv_math(void);
/* number of tcgen05 kernel launches made by this process so far (lets callers/tests verify which path ran) */
int64_t sdt_tc_launches(void);

/* ---- mel front end -----------------------------------------------------------------------------
 * torchaudio.transforms.MelSpectrogram(win_length=400, hop_length=160, n_fft=512, f_min=55,
 * f_max=7500, n_mels=80) as constructed at core/pipelines/voice2pose.py:27-30 (pose2pose.py:25-28)
 * and called at voice2pose.py:125 (pose2pose.py:48).  STFT (reflect pad 256, hann-400 centred in a
 * 512 frame) + |X|^2 + banded mel projection fused in one kernel: warp-per-frame-pair shared-memory
 * radix-2 FFT, the 257-bin spectrum never reaches HBM.  Power spectrogram, NO log.
 *   audio   (B, L) f32           window (400) f32   [state-dict buffer mel_transfm.spectrogram.window]
 *   fb_start/fb_count (80) i32, fb_weight (80, fb_stride) f32: band form of mel_transfm.mel_scale.fb
 *   mel     (B, 80, 1 + L/160) f32  (== channels-last (B,80,T,1))
 */
int sdt_mel_fwd(const float* audio, int B, int L, const float* window, const int32_t* fb_start,
                const int32_t* fb_count, const float* fb_weight, int fb_stride, float* mel, void* stream);

/* ---- convolution as implicit GEMM --------------------------------------------------------------
 * nn.Conv1d / nn.Conv2d inside ConvNormRelu (core/networks/building_blocks.py:15-22,31-36,49) and
 * their autograd backward.  One descriptor drives forward, data-gradient and weight-gradient:
 *
 *   rows   m = (b, gy, gx) over a GH x GW grid per image           (M = B*GH*GW)
 *   cols   n in [0, N)
 *   depth  k = (ty, tx, c), ty<TH, tx<TW, c<C                       (K = TH*TW*C)
 *   A(m,k) = xf( src[b, gy*y_mul + ty*ty_mul + y_off, gx*x_mul + tx*tx_mul + x_off, c] ), 0 outside src
 *   xf(v)  = act( v*xf_scale[b*xf_bstride + c] + xf_shift[...] ),  act = LeakyReLU(xf_slope)   (xf_scale != NULL)
 *            -- the PREVIOUS layer's normalisation + activation applied in the loader, so normalised
 *               activations are never materialised (building_blocks.py:50-54)
 *   conv:   dst[b, gy*dy_mul+dy_off, gx*dx_mul+dx_off, n] (+)= sum_k A(m,k) * wt[k*N + n] + bias[n]
 *           and, if stat_partial != NULL, per-row-tile column sums / sums of squares of the result
 *           (the input of sdt_norm_finalize) are written to stat_partial[tile][2][N]
 *   wgrad:  wpart[z][n][k] = sum_{m in split z} dy[m, n] * A(m,k)    (dy = (B,GH,GW,N) channels-last)
 */
typedef struct sdt_conv_desc {
    const float* src;          /* (B, SH, SW, C) */
    const float* wt;           /* conv: (K, N) row-major, from sdt_weight_prep mode 0/1 (FFMA path) */
    const float* wt_nk;        /* conv: (N, K) row-major (K-major rows), mode 2/3; enables the tcgen05 path, may be NULL */
    const float* bias;         /* (N) or NULL */
    const float* xf_scale;     /* loader transform, or NULL for identity */
    const float* xf_shift;
    float* dst;                /* conv: (B, DH, DW, N) */
    float* stat_partial;       /* conv: (row_tiles, 2, N) or NULL */
    const float* dy;           /* wgrad: (B, GH, GW, N) */
    float* wpart;              /* wgrad: (splits, N, K) */
    int32_t B, SH, SW, C;
    int32_t GH, GW, TH, TW;
    int32_t y_mul, ty_mul, y_off, x_mul, tx_mul, x_off;
    int32_t N;
    int32_t DH, DW, dy_mul, dy_off, dx_mul, dx_off;
    int32_t xf_bstride;        /* C for per-(b,c) statistics (InstanceNorm2d), 0 for per-channel (BatchNorm) */
    float xf_slope;            /* 0.2 LeakyReLU, 0 ReLU */
    int32_t accumulate;        /* conv: dst += result */
    int32_t per_image_tiles;   /* conv: row tiles never straddle images (needed for per-(b,c) statistics) */
    int32_t splits;            /* wgrad: number of K splits (== gridDim.z) */
    int32_t math;              /* 0 = the process default (sdt_get_conv_math()); m + 1 = math mode m for THIS problem, so every
                                * engine / model of a process carries its own mode and no call depends on shared mutable state */
    /* Fused channel-LayerNorm + activation epilogue (the reference's 1-D "IN" block: Conv1d -> InstanceNorm1d on the permuted
     * tensor -> LeakyReLU, building_blocks.py:31-54): when rn_act != NULL the convolution also writes
     * rn_act = act((dst - mean_row) * rstd_row) (same shape as dst) and the per-row statistics rn_mean / rn_rstd (B*DH*DW), i.e.
     * what sdt_rownorm_act_fwd would compute from dst in a second launch.  Only where sdt_conv_rownorm_ok() says so (the TMA
     * tensor-core kernel with the N = 256 output channels of a row spread over a cluster of four CTAs that exchange the row sums
     * through distributed shared memory); otherwise leave rn_act NULL and call sdt_rownorm_act_fwd. */
    float* rn_act;
    float* rn_mean;
    float* rn_rstd;
    float rn_eps, rn_slope;
    int32_t rn_out_tf32;
} sdt_conv_desc;

/* number of row tiles sdt_conv_gemm will use for this descriptor (size of stat_partial's first dim) */
int sdt_conv_row_tiles(const sdt_conv_desc* d);
int sdt_conv_gemm(const sdt_conv_desc* d, void* stream);
/* 1 if sdt_conv_gemm can run this problem with the fused row-norm epilogue (rn_* fields), else 0 (host only) */
int sdt_conv_rownorm_ok(const sdt_conv_desc* d);
/* which kernel sdt_conv_gemm would launch for this descriptor under the current math mode (host-only, no launch):
 * out10[0] = 0 fp32 FFMA, 1 tcgen05 (producer warps), 2 tcgen05 + TMA, 3 tcgen05 + TMA + shared-memory reuse, 4 CTA pairs; for 3/4 also
 * out10[1..9] = N tile, accumulators per CTA, patch rows | images per patch << 8, patch cols, box rows, A stages, B stages,
 * shared memory, tiles */
int sdt_conv_plan(const sdt_conv_desc* d, int32_t* out10);
/* n (1..16) independent problems in stream order -- the stride-parity classes of a strided layer's data gradient
 * (reference: one cuDNN dgrad call behind `loss.backward()`, core/networks/building_blocks.py:15-21).  Equivalent to n calls of
 * sdt_conv_gemm; in math mode 3, problems that share source, destination and shapes and differ in grid / offsets / weight
 * operand only run as ONE persistent launch.  *launched (may be NULL) receives the number of kernels launched. */
int sdt_conv_gemm_multi(const sdt_conv_desc* descs, int n, void* stream, int* launched);
/* wgrad: contractions with K = TH*TW*C <= 16 and N <= 64 (first encoder layer) use a streaming kernel whose CTA count
 * equals `splits`; otherwise split-K GEMM tiles (FFMA, or tcgen05 in math mode 1 when C % 32 == 0, N in {64,128,256}). */
int sdt_conv_wgrad(const sdt_conv_desc* d, void* stream);
/* sum the split-K partials in fixed order and store in the reference's parameter layout:
 * grad[(n*C + c)*T + t] (+)= sum_z wpart[z][n][(t*C + c)],  T = TH*TW  -> (Cout, Cin, kh, kw) */
int sdt_conv_wgrad_reduce(const float* wpart, int splits, int N, int C, int T, float* grad, int accumulate,
                          void* stream);
/* The same reduction for several layers in ONE launch (the fused trainers defer the reductions of a gradient bucket to its end:
 * 24 launches of ~10 us each per step become 2-3).  items_device: device array of n_items records, every item with C % 32 == 0
 * and T <= max_T <= 256; max_ctas = the largest N * (C / 32) among them. */
typedef struct sdt_reduce_item {
    const float* wpart;        /* (splits, N, T*C) partials of sdt_conv_wgrad */
    float* grad;               /* (N, C, T) reference-layout gradient */
    int32_t splits, N, C, T, accumulate;
    int32_t pad0;
} sdt_reduce_item;
int sdt_conv_wgrad_reduce_batch(const sdt_reduce_item* items_device, int n_items, int max_ctas, int max_T, void* stream);
/* reference-layout weight (Cout, Cin, KH, KW) -> GEMM operand (K, N):
 *   mode 0 forward : out[((ky*KW+kx)*Cin + ci)*Cout + co] = w[co,ci,ky,kx]
 *   mode 1 dgrad   : taps ky = ky0 + kstep*jy (jy < TH), kx likewise:
 *                    out[((jy*TW+jx)*Cout + co)*Cin + ci] = w[co,ci,ky0+kstep*jy,kx0+kstep*jx]
 *   mode 2 / 3     : the transposes of mode 0 / 1, i.e. (N, K) with K contiguous -- the K-major tensor-core operand:
 *                    mode 2 out[co*K + (ky*KW+kx)*Cin + ci], mode 3 out[ci*K' + (jy*TW+jx)*Cout + co];
 *                    values are rounded to nearest TF32 (see "out_tf32" below) because only the tcgen05 kernels read them */
int sdt_weight_prep(const float* w, int Cout, int Cin, int KH, int KW, int mode, int ky0, int kx0, int kstep,
                    int TH, int TW, float* out, void* stream);

/* All weight operands of a network in ONE launch: items_device is a device array of n_items records (pointers are
 * static across steps, so the table is uploaded once); max_elems = the largest TH*TW*Cin*Cout among them. */
typedef struct sdt_prep_item {
    const float* w;            /* (Cout, Cin, KH, KW) parameter */
    float* out;                /* operand buffer */
    int32_t Cout, Cin, KH, KW, mode, ky0, kx0, kstep, TH, TW;
    int32_t pad0;              /* mode 2 only: input channels of the operand, zero-padded (> Cin), e.g. 242 -> 256 so that a
                                * layer becomes tensor-core eligible; 0 = Cin */
    int32_t pad1;
} sdt_prep_item;
int sdt_weight_prep_batch(const sdt_prep_item* items_device, int n_items, long long max_elems, void* stream);

/* ---- "out_tf32" ------------------------------------------------------------------------------------
 * tcgen05.mma kind::tf32 reads fp32 operands and ignores their low 13 mantissa bits: it TRUNCATES, so every product is low by
 * ~7e-4 on average.  Normalisation on batch / instance statistics cancels that scale bias layer by layer; BatchNorm in eval
 * mode (validation, demo: voice2pose.py:333-410) does not, and 25 layers of it drift by 1.6e-2 (profiles/
 * r2_tf32_truncation_bias.txt).  Every entry point that PRODUCES a tensor-core operand (activations, input gradients, GEMM
 * weight copies) therefore takes `out_tf32`: non-zero = store the result rounded to the nearest TF32 value (cvt.rna.tf32.f32),
 * which makes the MMA's view of it exact and its rounding error unbiased -- what cuDNN's TF32 convolutions, the reference's own
 * GPU default, do.  Callers pass 0 in the fp32 math mode. */

/* ---- normalisation -----------------------------------------------------------------------------
 * nn.InstanceNorm2d / nn.BatchNorm{1,2}d inside ConvNormRelu (building_blocks.py:23-27,38-43,53) with
 * the LeakyReLU/ReLU at :46,54.  Statistics come from the conv epilogue partials; normalise+activate is
 * applied by the next consumer's loader through (scale, shift):  xhat*gamma+beta = x*scale + shift.
 *   partial (groups*tiles_per_group, 2, C); count = elements per statistic; groups = B (IN) or 1 (BN)
 *   outputs (groups, C): scale, shift, mean, rstd.  gamma/beta NULL -> no affine (InstanceNorm).
 *   running_mean/var != NULL (BatchNorm train): momentum update with the UNBIASED variance and
 *   num_batches_tracked += 1.
 */
int sdt_norm_finalize(const float* partial, int groups, int tiles_per_group, int C, double count, const float* gamma,
                      const float* beta, float eps, float* scale, float* shift, float* mean, float* rstd,
                      float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                      void* stream);
/* Column sums / sums of squares of the window [w0, w1) of a channels-last map x (B, H, W, C), in the partial layout
 * sdt_norm_finalize reads: partial (B, n_parts, 2, C), part p = positions [p*rows_per_part, (p+1)*rows_per_part) of the
 * window's H*(w1-w0) positions (parts past the end are written as zeros).  Time-tiled long-audio inference
 * (trainer.py:459-484 / voice2pose.py:386-410 run the whole utterance in one forward): InstanceNorm2d statistics span the
 * utterance, a tile contributes its OWNED columns only (the receptive-field halo is excluded). */
int sdt_chan_stats(const float* x, int B, int H, int W, int C, int w0, int w1, int rows_per_part, float* partial,
                   int n_parts, void* stream);
/* BatchNorm eval mode: scale/shift from running statistics (C). */
int sdt_bn_eval_scale_shift(const float* running_mean, const float* running_var, const float* gamma, const float* beta,
                            float eps, int C, float* scale, float* shift, void* stream);
/* Backward of [per-(g,c) normalise -> affine -> (Leaky)ReLU] over a channels-last map x (B, P, C), P = H*W:
 * pass 1 reduces  s1 = sum g', s2 = sum g'*xhat  (g' = g_act * act'(y)) into partial (B*tiles, 2, C);
 * finalize gives the means (groups, C) and, for BatchNorm, dgamma/dbeta; pass 2 writes
 * g_x = rstd*gamma*(g' - m1 - xhat*m2) in place over g. groups = B (IN) or 1 (BN). */
/* tiles_per_image in [1, P] is the caller's choice of parallelism for pass 1 (rows per tile = ceil(P / tiles)). */
int sdt_norm_bwd_reduce(const float* g, const float* x, const float* mean, const float* rstd, const float* gamma,
                        const float* beta, int B, int P, int C, int groups, float slope, float* partial,
                        int tiles_per_image, void* stream);
int sdt_norm_bwd_finalize(const float* partial, int groups, int tiles_per_group, int C, double count, float* m1,
                          float* m2, float* dgamma, float* dbeta, int accumulate, void* stream);
int sdt_norm_bwd_apply(float* g, const float* x, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, const float* m1, const float* m2, int B, int P, int C, int groups, float slope,
                       int out_tf32, void* stream);
/* InstanceNorm1d on the permuted tensor == LayerNorm over channels per (b,t), no affine
 * (building_blocks.py:50-51) + activation, on rows of a (R, C) channels-last matrix. */
int sdt_rownorm_act_fwd(const float* x, int R, int C, float eps, float slope, float* y, float* mean, float* rstd,
                        int out_tf32, void* stream);
int sdt_rownorm_act_bwd(const float* g_y, const float* x, const float* mean, const float* rstd, int R, int C,
                        float slope, float* g_x, int out_tf32, void* stream);
/* y = act(x*scale[g,c] + shift[g,c]) materialised (used at module boundaries only). */
int sdt_scale_shift_act(const float* x, const float* scale, const float* shift, int B, int P, int C, int bstride,
                        float slope, float* y, int out_tf32, void* stream);

/* ---- first block of the audio encoder, special-cased ------------------------------------------------
 * Conv2d(1 -> 64, 3x3, s1, p1, no bias) + InstanceNorm2d + LeakyReLU (generator.py:17 via building_blocks.py
 * ConvNormRelu): the largest map of the network and pure HBM traffic.  x (B,H,W) is the one-channel input (the mel
 * image), w (64,1,3,3) the reference weight.  With one input channel the InstanceNorm statistics follow in closed form
 * from the 9 tap means + 45 tap second moments of each image (`moments`, (B,54) f64), so the forward writes the
 * activated map act (B,H,W,64) in ONE pass and never stores the raw convolution; scale/shift (B,64) receive rstd and
 * -mean*rstd.  mom_partial: (B, sdt_first_layer_units(H,W), 54) f64 scratch. */
int sdt_first_layer_units(int H, int W);
int sdt_first_layer_fwd(const float* x, const float* w, int B, int H, int W, int C, float eps, float slope,
                        double* mom_partial, double* moments, float* scale, float* shift, float* act, int out_tf32,
                        void* stream);
/* The two halves separately, for the time-tiled long-audio forward (trainer.py:459-484 runs the whole utterance at once):
 * sdt_first_layer_fwd with act == NULL computes the statistics (moments, scale, shift) of the WHOLE image only -- they follow in
 * closed form from the input, no pass over the 64-channel map -- and sdt_first_layer_act writes the activated map of any tile of
 * the image (x = the tile, zero padding at its edges) with the given per-(image, channel) scale / shift. */
int sdt_first_layer_act(const float* x, const float* w, const float* scale, const float* shift, int B, int H, int W, int C,
                        float slope, float* act, int out_tf32, void* stream);
/* Weight gradient of that block from g_act = dLoss/d act in ONE pass over g_act: the InstanceNorm + LeakyReLU backward is
 * folded into per-(image, channel) sums and combined in closed form (f64) with the forward's moments; the block's input needs
 * no gradient (it is the mel spectrogram).  The pre-activation is recomputed from x with the forward's arithmetic, so `act`
 * is not read (it may be NULL; the parameter stays for ABI stability).  Requires slope > 0.
 * partial: (B, units, 11, 64) f32 scratch; dw (64,1,3,3) is overwritten. */
int sdt_first_layer_bwd(const float* g_act, const float* act, const float* x, const float* w, const double* moments,
                        const float* scale, const float* shift, int B, int H, int W, int C, float slope, float* partial,
                        float* dw, void* stream);

/* ---- resampling / concatenation ----------------------------------------------------------------
 * F.interpolate(x, (1, F), mode='bilinear') + squeeze (generator.py:41-42) fused with the code broadcast +
 * concat (generator.py:109-111): reads the last encoder block's raw output (B,H,W,C) through its
 * scale/shift/activation, writes the UNet input (B, F, C + D) channels-last; code (B, D) or NULL. */
int sdt_enc_to_seq_fwd(const float* x, const float* scale, const float* shift, int xf_bstride, float slope, int B, int H,
                       int W, int C, const float* code, int D, int F, float* out, int out_tf32, void* stream);
/* adjoint: g_out (B,F,C+D) -> g_act (B,H,W,C) (zero outside the sampled row) and g_code (B,D) = sum_t. */
int sdt_enc_to_seq_bwd(const float* g_out, int B, int H, int W, int C, int D, int F, float* g_act, float* g_code,
                       void* stream);
/* F.interpolate(x, Lout, mode='linear') + skip (generator.py:79-83; autoencoder.py:62-66 with skip NULL):
 * out (B,Lout,C) = lerp(x (B,Lin,C)) [+ skip]. */
int sdt_upsample_add_fwd(const float* x, const float* skip, int B, int Lin, int Lout, int C, float* out, int out_tf32,
                         void* stream);
/* adjoint of the lerp: g_x (B,Lin,C) (+)= A^T g_out. */
int sdt_upsample_bwd(const float* g_out, int B, int Lin, int Lout, int C, float* g_x, int accumulate, void* stream);

/* ---- losses -------------------------------------------------------------------------------------
 * L1: nn.L1Loss('none')(pred, gt) * lambda -> mean (voice2pose.py:141-142, pose2pose.py:71-72).
 *   loss_out[0] = value; g_pred = lambda*sign(pred-gt)/n (sign(0)=0) if g_pred != NULL. partial: >= 1024 floats. */
int sdt_l1_loss(const float* pred, const float* gt, int64_t n, float lambda, float* loss_out, float* g_pred,
                float* partial, void* stream);
/* Clip-code gather + batch-statistics KL (voice2pose.py:94,147-157): code (B,D) = table[idx];
 * out[0] = KL value (0 if skipped), out[1] = 1 if applied (all unbiased batch variances != 0) else 0;
 * g_code (B,D) = d KL / d code (0 when skipped). */
int sdt_code_gather_kl(const float* table, const int64_t* idx, int B, int D, float lambda, float* code, float* out,
                       float* g_code, void* stream);
/* index backward: g_table[idx[b]] += g_code_a[b] + g_code_b[b] (duplicates accumulate, fixed order).
 * g_table must be zeroed by the caller. */
int sdt_code_scatter_grad(const float* g_code_a, const float* g_code_b, const int64_t* idx, int B, int D,
                          float* g_table, void* stream);
/* Pose2Pose.train_step's buffer scatter (pose2pose.py:135-137): table_a[idx[b]] = src_a[b] (and table_b / src_b likewise when
 * given), rows of D floats.  Duplicate indices in a batch: the last occurrence wins (the sequential CPU semantics). */
int sdt_code_store_rows(const float* src_a, float* table_a, const float* src_b, float* table_b, const int64_t* idx, int B,
                        int D, void* stream);
/* column sums of a (R, C) matrix: bias gradient of the final Conv1d (generator.py:103). */
int sdt_colsum(const float* g, int R, int C, float* out, int accumulate, void* stream);
/* LSGAN terms, nn.MSELoss vs a constant target (voice2pose.py:82,195-202): out[0] = lambda*mean((s-target)^2),
 * g_s = lambda*2*(s-target)/n if g_s != NULL. */
int sdt_mse_const_loss(const float* s, int64_t n, float target, float lambda, float* out, float* g_s, void* stream);
/* motion = x[:,1:] - x[:,:-1] over a (B,T,C) sequence (voice2pose.py:187-189) and its adjoint. */
int sdt_motion_diff_fwd(const float* x, int B, int T, int C, float* out, void* stream);
int sdt_motion_diff_bwd(const float* g_out, int B, int T, int C, float* g_x, int accumulate, void* stream);

/* ---- pose VAE head ------------------------------------------------------------------------------
 * PoseSeqEncoder tail (autoencoder.py:31-35): take t=0 of the last block's raw output (B,L,2D) through its
 * scale/shift/activation, split even/odd channels -> mu, logvar (B,D). */
int sdt_pose_head_fwd(const float* x, const float* scale, const float* shift, float slope, int B, int L, int D2,
                      float* mu, float* logvar, void* stream);
/* VAE reparameterisation + KL (autoencoder.py:84-87, pose2pose.py:77): code = mu + exp(.5 logvar)*eps;
 * out[0] = lambda*0.5*mean(-logvar + mu^2 + exp(logvar) - 1). */
int sdt_vae_reparam_kl(const float* mu, const float* logvar, const float* eps, int n, float lambda, float* code,
                       float* out, void* stream);

/* adjoints of the two entry points above: g_act (B,L,2D) is zero except t = 0;
 * d mu = g_code + lambda*mu/n, d logvar = g_code*0.5*exp(0.5 logvar)*eps + lambda*0.5*(exp(logvar)-1)/n. */
int sdt_pose_head_bwd(const float* g_mu, const float* g_logvar, int B, int L, int D2, float* g_act, void* stream);
int sdt_vae_reparam_kl_bwd(const float* mu, const float* logvar, const float* eps, const float* g_code, int n, float lambda,
                           float* g_mu, float* g_logvar, void* stream);

/* ---- keypoint indexing / normalisation (bit-exact gates) ----------------------------------------
 * GestureDataset pose preprocessing (core/datasets/gesture_dataset.py:95-105,131-191):
 * raw (T,3,137) f32 -> gather 122 -> xy -= kp1 -> drop kp1 -> [parted] -> (x - f32(mean)) / f32(std) (IEEE div).
 * mean/std (242) f32.  out (T,2,121) f32. */
int sdt_pose_preprocess(const float* raw, int T, const float* mean, const float* std, int hierarchical, float* out,
                        void* stream);
/* GestureDataset.get_final_results (gesture_dataset.py:193-220): f64 x*std (rounded) + mean (rounded) ->
 * parted_to_global -> * scale.  poses (B,T,2,121) f32; mean/std (B,242) f64; scale (B) f64; out f64. */
int sdt_pose_final_results(const float* poses, int B, int T, const double* mean, const double* std, const double* scale,
                           int hierarchical, double* out, void* stream);
/* GestureDataset.transform_normalized_parted2global (gesture_dataset.py:221-234), fp32: poses (n_rows,2,121) normalised
 * in the parted (hierarchical) space -> normalised in the global space; statistics (242) f32 of one speaker. */
int sdt_pose_parted2global(const float* poses, int64_t n_rows, const float* mean_parted, const float* std_parted,
                           const float* mean_global, const float* std_global, float* out, void* stream);
/* Voice2Pose.evaluate_step (voice2pose.py:412-430) on final-result poses: out[0] = L2_dist, out[1] = lip_sync_error_n.
 * partial: >= 2*B doubles. */
int sdt_pose_metrics(const double* pred, const double* gt, int B, int T, double* partial, double* out, void* stream);

/* ---- optimizer ----------------------------------------------------------------------------------
 * torch.optim.Adam (voice2pose.py:249,263,274; pose2pose.py:114) over a flat fp32 buffer; weight_decay is Adam's L2 term
 * (grad += weight_decay * param, cfg.TRAIN.WD at voice2pose.py:250; 0 in every shipped config).
 * state_host-free: scalars (step count -> bias corrections) live in `scalars` (device, 4 floats + 1 i64 as 8 floats):
 *   sdt_adam_advance bumps the step and recomputes them on device so that a captured CUDA graph can replay
 *   (lr < 0: take the learning rate from scalars[3], so a schedule can change it between replays);
 *   grad_scale multiplies the gradient first (1/world_size after the NCCL sum). */
int sdt_adam_advance(float* scalars, float lr, double beta1, double beta2, void* stream);
int sdt_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const float* scalars,
                  double beta1, double beta2, double eps, float grad_scale, float weight_decay, void* stream);

/* ---- gradient exchange over peer memory ------------------------------------------------------------------------
 * All-reduce(sum) of a flat fp32 buffer of n elements (n % 4 == 0) that every rank of one NVSwitch node holds at the same place of
 * a symmetric allocation -- the DDP gradient all-reduce of voice2pose.py:222-223,298-309 / pose2pose.py:101-102 without NCCL.
 * peer_ptrs[world]: HOST array of the device addresses of the buffer on rank 0..world-1 (peer-mapped); multicast_ptr: the NVLS
 * multicast address of the same buffer, or 0 to use peer loads / stores.  Rank r reduces shard r (rank order 0..world-1) and
 * writes it into every rank's buffer, in place; all ranks end with bit-identical sums.  scal_*: scal_n (<= 512) doubles per rank
 * (the loss / metric scalars of trainer.py:323-327), read from every rank's scal_peer_ptrs[r], summed into the LOCAL scal_dst.
 * The caller issues a cross-GPU barrier before (all gradients final) and after (all shards written) this call. */
int sdt_p2p_allreduce(const uint64_t* peer_ptrs, uint64_t multicast_ptr, int rank, int world, long long n,
                      const uint64_t* scal_peer_ptrs, double* scal_dst, int scal_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDT_B200_H */
