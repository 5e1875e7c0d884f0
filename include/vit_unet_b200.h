/* vit_unet_b200.h -- C ABI of the B200-native ViT-UNet forward/backward kernels.
 *
 * Plain C: pointers, sizes, POD descriptors.  No torch types.  Every pointer is a DEVICE pointer unless
 * the name ends in _host.  Every launch goes to the cudaStream_t passed as `stream` (a torch
 * `current_stream().cuda_stream` value works); nothing here synchronises the device or allocates.
 * All functions return 0 on success; non-zero = VU_ERR_*; vu_last_error() returns the text.
 *
 * The reference has no native layer: each entry point below replaces the ATen op
 * sequence the reference's Python issues at the cited lines of /root/reference/vit_unet/torch/model.py.
 *
 * Layout vocabulary ("patch layout p"):  a (B, C, H, W) image stored as tokens (B, N, C*p*p) with
 * token r*(W/p)+c and feature ch*p*p+i*p+j  <->  pixel (ch, r*p+i, c*p+j)   [model.py:8-35].
 * p == 0 means plain NCHW.  Every level of the network is the same image in a different patch layout.
 */
#ifndef VIT_UNET_B200_H
#define VIT_UNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VU_OK 0
#define VU_ERR_ARG 1      /* bad shape / unsupported configuration */
#define VU_ERR_CUDA 2     /* CUDA runtime error (text in vu_last_error) */
#define VU_ERR_UNSUPPORTED 3

int vu_version(void);                 /* ABI version, bumped on any signature change */
const char* vu_last_error(void);      /* thread-local text of the last non-zero return */
int vu_device_sm_count(int device);   /* 148 on B200 */

/* ---------------------------------------------------------------- layout (model.py:8-53, :84-91) */
/* out[layout p_out] = in[layout p_in]; patch / unpatch / downsampling / upsampling in one permutation. */
int vu_repatch(const float* in, float* out, int B, int C, int H, int W, int p_in, int p_out, void* stream);
/* PatchEncoder: out[p_out] = in[p_in] + table[p_table] (table has no batch dim). model.py:84-91,
 * ViT_UNet.ipynb c16:L31-40 */
int vu_pe_fwd(const float* in, int p_in, const float* table, int p_table, float* out, int p_out,
              int B, int C, int H, int W, void* stream);
/* dtable[p_table] (+)= sum_b dout[b][p_out] */
int vu_pe_bwd_table(const float* dout, int p_out, float* dtable, int p_table,
                    int B, int C, int H, int W, int accumulate, void* stream);

/* dst[b][h][e][n] (bfloat16, row pitch ldn >= N, ldn % 8 == 0) = src[b][n][h*(D/h) + e]: per-head transposed bf16
 * copy of a (B,N,D) token tensor = K-major B operand of the bf16 attention GEMMs (head split of model.py:152). */
int vu_heads_transpose_bf16(const float* src, void* dst, int B, int N, int D, int h, int ldn, void* stream);

/* ---------------------------------------------------------------- 3x3 convs (model.py:137-139,152-154,428) */
/* nconv (1..3) bias-optional 3x3 C->C convs sharing one input.  Zero padding at the borders of
 * `border_p`-sized patches (border_p == 0: image borders).  Filters [C][C][3][3] per conv: w1 == NULL -> all convs
 * contiguous at w ([nconv][C][C][3][3]); else conv k reads its own block w / w1 / w2 (the q/k/v nn.Conv2d weights
 * where they lie, no concatenation pass).  bias: [nconv][C] or NULL.  x is read in layout p_x; out_k written in p_out. */
int vu_conv3x3_fwd(const float* x, int p_x, const float* w, const float* w1, const float* w2, const float* bias, int nconv,
                   float* out0, float* out1, float* out2, int p_out, int border_p,
                   int B, int C, int H, int W, int tf32, void* stream);
/* tf32 != 0 (the tensor-core precision class, all three entry points): the per-patch convs (source, destination and border
 * patch agree, patch >= 8, C <= 3) run as implicit GEMMs on TF32 warp MMAs (operands rounded to nearest TF32, fp32
 * accumulation); 0 = exact fp32 FMAs. */
/* dx[p_dx] (+)= sum_k conv_transpose(dy_k[p_dy], w_k) */
int vu_conv3x3_bwd_data(const float* dy0, const float* dy1, const float* dy2, int p_dy,
                        const float* w, const float* w1, const float* w2, int nconv, float* dx, int p_dx, int border_p,
                        int B, int C, int H, int W, int accumulate, int tf32, void* stream);
/* dw_k[C][C][3][3] += ..., dbias[nconv][C] += ... (atomic accumulation: caller zeroes).  dw1 == NULL: the blocks of all
 * convs are contiguous at dw; else conv k accumulates into dw / dw1 / dw2 (separate slots of a flat gradient buffer). */
int vu_conv3x3_bwd_weight(const float* x, int p_x, const float* dy0, const float* dy1, const float* dy2,
                          int p_dy, int nconv, float* dw, float* dw1, float* dw2, float* dbias, int border_p,
                          int B, int C, int H, int W, int tf32, void* stream);

/* ---------------------------------------------------------------- GEMM (model.py:155,161,162; :103,106) */
enum { VU_ACT_NONE = 0, VU_ACT_GELU = 1, VU_ACT_GELU_BWD = 2 };
enum { VU_PREC_FP32 = 0, VU_PREC_TF32 = 1 };   /* FP32: CUDA-core FMA; TF32: tcgen05 tensor cores */

typedef struct vu_gemm_desc {
  const float* A; const float* B; float* C;
  const float* bias;      /* [N] added to every row, or NULL */
  const float* residual;  /* same indexing as C (ldr), added after activation/dropout, or NULL */
  const float* aux_in;    /* VU_ACT_GELU_BWD: pre-activation, same indexing as C (ldaux) */
  float* aux_out;         /* VU_ACT_GELU: pre-activation written here (ldaux), or NULL */
  int M, N, K;
  int trans_a;            /* 0: A(m,k)=A[m*lda+k]   1: A(m,k)=A[k*lda+m] */
  int trans_b;            /* 0: B(k,n)=B[k*ldb+n]   1: B(k,n)=B[n*ldb+k]  (nn.Linear weight layout) */
  int64_t lda, ldb, ldc, ldr, ldaux;
  int batch_outer, batch_inner;               /* batch z = zo*batch_inner + zi */
  int64_t sAo, sAi, sBo, sBi, sCo, sCi;       /* element strides; residual/aux use the C strides */
  float alpha;            /* C = act(alpha*A.B + bias) [dropout] + residual */
  int act;
  int accumulate;         /* 1: C += result (no act/residual allowed with split_k>1) */
  int split_k;            /* >1: K is split, partials atomically added into C (caller zeroes C unless accumulate) */
  float drop_p;           /* >0: inverted dropout on act(...) keyed by (drop_seed, drop_stream, m*N+n) */
  uint64_t drop_seed; uint32_t drop_stream;
  int precision;          /* VU_PREC_* */
  /* element types (tensor-core path only): 0 = float32, 1 = bfloat16.  A and B must agree (either may be K- or
   * MN-major); a bf16 C takes every fused epilogue but cannot be accumulated into (no accumulate / split_k);
   * aux_bf16: aux_in / aux_out are bfloat16 (bf16 operands only).  bias and residual are always float32.
   * ld* and batch strides count elements of the tensor's own type. */
  int a_bf16, b_bf16, c_bf16, aux_bf16;
} vu_gemm_desc;

int vu_gemm(const vu_gemm_desc* d_host, void* stream);
/* out[n] (+)= sum_m X[m*ld+n]   (bias gradients); X float32, or bfloat16 when x_bf16 != 0 */
int vu_colsum(const void* X, int x_bf16, int64_t M, int N, int64_t ld, float* out, int accumulate, void* stream);

/* ---------------------------------------------------------------- Re-Attention core (model.py:155-161) */
/* maps are (B, h, N, ld) fp32 with ld >= N, ld % 4 == 0 */
/* in place: P[r, :N] = softmax(scale * S[r, :N]) for r in rows, pad columns zeroed. */
int vu_softmax_rows(float* S, int64_t rows, int N, int ld, float scale, void* stream);
/* train-mode moments of the dropped maps Pd_g = drop(P_g), centred at c = 1/N, over all (b,i,j):
 * sums[h + h*h] (double, caller zeroes) += { s'_g = sum(Pd_g - c),  G'_{gg'} = sum (Pd_g - c)(Pd_g' - c) }.
 * The BatchNorm batch statistics of M_h = sum_g W[h,g] Pd_g + b[h] follow in closed form (vu_reattn_bn_finalize). */
int vu_reattn_stats(const float* P, int B, int h, int N, int ld, float drop_p, uint64_t seed, uint32_t stream_id,
                    double* sums, void* stream);
/* Storage format flags of the attention maps (`map_fmt` below).  VU_MAP_BF16: the mixed map A and the gradient
 * map dA/dS are bfloat16 (ld % 8 == 0).  VU_MAP_P_CENTRED_BF16: the probabilities are stored as bfloat16 CENTRED
 * at the uniform row, Pc = P - 1/N (so the rounding is relative to what head mixing + BatchNorm actually see);
 * only together with VU_MAP_BF16 and where vu_reattn_tensor_core_path() != 0. */
enum { VU_MAP_BF16 = 1, VU_MAP_P_CENTRED_BF16 = 2, VU_MAP_TF32_MIX = 4 };
/* VU_MAP_TF32_MIX: fp32 maps, but the 8x8 head mixing may run on TF32 warp MMAs (the tensor-core precision class);
 * without it (and without VU_MAP_BF16) the map kernels use exact fp32 FMAs. */
/* 1 when the 8-head warp-MMA formulation of the map kernels applies (h == 8, ld == N, N % 4 == 0; VU_MAP_MMA=0 in
 * the environment switches it off): vu_softmax_stats(TF32), vu_reattn_mix / _mix_reduce / _bwd_rows with bf16 maps
 * or VU_MAP_TF32_MIX. */
int vu_reattn_tensor_core_path(int h, int N, int ld);
/* fused train-mode pass: softmax of every head + the moments above in ONE read of S / write of P.
 * Pc == NULL: P overwrites S in place (fp32).  Pc != NULL: centred bf16 probabilities are written to Pc (same
 * (B,h,N,ld) indexing), S is left untouched and the moments are those of the ROUNDED map.
 * precision = VU_PREC_TF32 lets the moments of 8-head maps without pad columns be accumulated by TF32 warp MMAs
 * (centred inputs, fp32 accumulation); VU_PREC_FP32 keeps the exact CUDA-core sums. */
int vu_softmax_stats(float* S, void* Pc, int B, int h, int N, int ld, float scale, float drop_p, uint64_t seed,
                     uint32_t stream_id, double* sums, int precision, void* stream);
/* fold conv1x1 + BatchNorm into one affine:  fold[h*h + h] = {alpha'[h][g], beta'[h]};
 * saved[2h] = {mean_h, invstd_h}.  train=1: batch statistics from `sums` (+ running-stat update,
 * momentum, unbiased variance, num_batches_tracked += 1); train=0: running statistics. model.py:136,159 */
int vu_reattn_bn_finalize(const double* sums, int64_t count, int h, int N, const float* W, const float* bconv,
                          const float* gamma, const float* beta, float* running_mean, float* running_var,
                          int64_t* num_batches_tracked, float eps, float momentum, int train,
                          float* fold, float* saved, void* stream);
/* A_h = sum_g alpha'[h,g]*drop(P_g) + beta'[h].  map_fmt: see VU_MAP_* (P fp32 or centred bf16; A fp32 or bf16). */
int vu_reattn_mix(const void* P, void* A, int map_fmt, const float* fold, int B, int h, int N, int ld,
                  float drop_p, uint64_t seed, uint32_t stream_id, void* stream);
/* backward reductions: red[h + h*h] (double, caller zeroes) += { s1_h = sum dA_h,  X'_{hg} = sum dA_h (Pd_g - c) } */
int vu_reattn_bwd_reduce(const float* P, const float* dA, int B, int h, int N, int ld, float drop_p, uint64_t seed,
                         uint32_t stream_id, double* red, void* stream);
/* fused backward pass: A = mix(P) (as vu_reattn_mix) AND the reductions of vu_reattn_bwd_reduce, one read of P, dA.
 * A == NULL (tensor-core map path only): the caller kept the forward map, only the reductions are computed. */
int vu_reattn_mix_reduce(const void* P, const void* dA, void* A, int map_fmt, const float* fold, int B, int h, int N,
                         int ld, float drop_p, uint64_t seed, uint32_t stream_id, double* red, void* stream);
/* closed-form parameter gradients from (red, sums): dW[h*h], dbconv[h], dgamma[h], dbeta[h] are ACCUMULATED
 * (atomic; caller zeroes); coef[2h] = BatchNorm-backward means {mean dA_h, mean dA_h*Ahat_h} for vu_reattn_bwd_rows.
 * sums may be NULL when train == 0. */
int vu_reattn_bwd_params(const double* red, const double* sums, int B, int h, int N, const float* W,
                         const float* bconv, const float* gamma, const float* saved, int train,
                         float* coef, float* dW, float* dbconv, float* dgamma, float* dbeta, void* stream);
/* in place dA -> dS (gradient of the pre-softmax scores) */
int vu_reattn_bwd_rows(const void* P, void* dA_dS, int map_fmt, int B, int h, int N, int ld, const float* W,
                       const float* bconv, const float* gamma, const float* saved, const float* coef,
                       int train, float scale, float drop_p, uint64_t seed, uint32_t stream_id, void* stream);

/* ---- streamed Re-Attention (vu_reattn_stream.cu): the same math as the chain softmax -> dropout -> mix/BN -> A.V
 * (model.py:155-161, :251-256) WITHOUT materialising the (B,h,N,N) maps, for the fine levels (many tokens, small
 * heads): (h, hd) in {(8,8), (8,24), (4,12), (4,48)}, N % 16 == 0.  q, k: (B,N,h*hd) fp32; vt: per-head
 * transposed bf16 values (B,h,hd,ldn) from vu_heads_transpose_bf16; o: (B,N,h*hd) fp32 (pre-projection output).
 * Scores run as TF32 warp MMAs, A.V as bf16 warp MMAs (the tensor-core precision class, VU_PREC_TF32).
 * mode 0: eval forward, one launch (fold = running-statistics affine from vu_reattn_bn_finalize(train=0)).
 * mode 1: train statistics: rowc[b,h,i] = log2-domain softmax constant of every row, sums += centred moments of the
 *         dropped maps (same definition as vu_softmax_stats); pc != NULL also writes the centred bf16 probabilities
 *         (B,h,N,N) consumed by the materialised backward kernels (VU_MAP_P_CENTRED_BF16).
 * mode 2: train apply: o = (fold . dropout(softmax)) v with the row constants of mode 1 and the batch-statistics fold;
 *         amap != NULL also writes the mixed map A as bf16 (B,h,N,N) for the backward product dV = A^T dO.
 * Dropout masks are those of the materialised kernels, element for element (same counter-hash keying).  mask
 * (optional, B * (N/16)^2 * 256 bytes = 64 bits per lane of every (image, 16-row unit, 16-key step)): mode 1 caches the keep-bits it generated, mode 2 reads them instead of
 * hashing again (the hash is ~40 % of the apply sweep's instructions); NULL = regenerate. */
int vu_reattn_stream_supported(int h, int hd, int N);
int vu_reattn_stream_fwd(int mode, const float* q, const float* k, const void* vt, float* o, const float* fold,
                         float* rowc, double* sums, void* pc, void* amap, void* mask, int B, int h, int N, int hd, int ldn,
                         float scale, float drop_p, uint64_t seed, uint32_t stream_id, void* stream);

/* streamed backward: pc = centred bf16 probabilities (B,h,N,N) written by vu_reattn_stream_fwd mode 1, mask = its
 * cached keep-bits (or NULL: re-hash), dO / v: (B,N,h*hd) fp32.  dA = dO v^T is formed on the fly, never stored.
 * _bwd_reduce: red[h + h*h] (double, caller zeroes) += { sum dA_h, sum dA_h (Pd_g - 1/N) }  (as vu_reattn_mix_reduce).
 * _bwd_ds: dS (bf16 (B,h,N,N), the gradient of the pre-softmax scores: softmax / dropout / mixing / BatchNorm
 * backward, as dA -> vu_reattn_bwd_rows) and dq = dS k (B,N,h*hd) with kt = per-head transposed bf16 keys. */
int vu_reattn_stream_bwd_reduce(const void* pc, const void* mask, const float* dO, const float* v, double* red,
                                int B, int h, int N, int hd, float drop_p, uint64_t seed, uint32_t stream_id, void* stream);
int vu_reattn_stream_bwd_ds(const void* pc, const void* mask, const float* dO, const float* v, const void* kt, void* ds,
                            float* dq, const float* W, const float* bconv, const float* gamma, const float* saved,
                            const float* coef, int train, int B, int h, int N, int hd, int ldn, float drop_p,
                            uint64_t seed, uint32_t stream_id, void* stream);

/* number of VU_PREC_TF32 GEMM requests that ran on the CUDA-core kernel because an operand was not TMA-addressable
 * (also logged to stderr the first few times; VU_LOG_FALLBACKS=1 logs all) */
int vu_gemm_tf32_fallbacks(void);

/* ---------------------------------------------------------------- LayerNorm over (N,D) (model.py:193-196,204,206) */
#define VU_LN_SPLIT 8      /* CTAs cooperating on one image's statistics */
/* stats[b] = {mean, rstd} over the n = N*D elements of image b; scratch: 2*VU_LN_SPLIT*B floats */
int vu_ln_stats(const float* x, int B, int64_t n, float eps, float* stats, float* scratch, void* stream);
/* out_bf16 (optional, NULL = none): a bfloat16 copy of the result written in the same pass -- the A operand of the next
 * tensor-core GEMM in the bf16 mode (the fp32 result stays the residual stream) */
int vu_ln_apply(const float* x, const float* stats, const float* w, const float* b, float* out, void* out_bf16,
                int B, int64_t n, void* stream);
/* dx = LN backward; dw/db += (atomic; caller zeroes or accumulates: the README variant shares one LN);
 * dx_bf16 (optional): bfloat16 copy of dx for the following data- / weight-gradient GEMMs */
int vu_ln_bwd(const float* g, const float* x, const float* stats, const float* w, float* dx, void* dx_bf16,
              float* dw, float* db, float* scratch /* (2 + 2*VU_LN_SPLIT)*B floats */, int B, int64_t n, void* stream);

/* ---------------------------------------------------------------- losses (run_denoising.py:80, README.md:91-101) */
enum { VU_LOSS_L1 = 0, VU_LOSS_MSE = 1, VU_LOSS_DICE = 2 };
/* sums[4] (double, zeroed here): L1 {sum|d|}, MSE {sum d^2}, DICE {sum xy, sum x, sum y}; loss[0] = the scalar */
int vu_loss_fwd(int kind, const float* pred, const float* target, int64_t n, double* sums, float* loss,
                void* stream);
/* loss[0] from sums alone (n = GLOBAL element count): lets data-parallel ranks all-reduce `sums` between the
 * reduction and the scalar, so soft-Dice stays the reference's whole-batch ratio (README.md:96-101). */
int vu_loss_finalize(int kind, int64_t n, const double* sums, float* loss, void* stream);
/* dpred = gscale[0] * dloss/dpred  (gscale: device scalar, the upstream gradient) */
int vu_loss_bwd(int kind, const float* pred, const float* target, int64_t n, const double* sums,
                const float* gscale, float* dpred, void* stream);

/* ---------------------------------------------------------------- callers either side of the path */
/* N2: per-image PSNR on the device (functions.py:7-19 + skimage.metrics.peak_signal_noise_ratio):
 * psnr[b] = 10 log10(range^2 / mean((pred_b - target_b)^2)), range = data_range if > 0 else 1 (min(target_b) >= 0) or 2.
 * scratch: B doubles + B floats. */
int vu_psnr(const float* pred, const float* target, int B, int64_t n, float data_range, double* scratch, float* psnr,
            void* stream);
/* N4: input pipeline step of DenoisingDataset (dataset.py:65-68) + Normalize (run_denoising.py:54) on the device:
 * uint8 HWC -> float32 CHW, dst = (src * scale - mean) / std */
int vu_u8hwc_to_chw(const uint8_t* src, float* dst, int B, int C, int H, int W, float scale, float mean, float std,
                    void* stream);

/* N4 continued: cv2.resize(INTER_LINEAR) of a uint8 HWC batch (dataset.py:59-60), and a fused affine warp (albumentations
 * ShiftScaleRotate with BORDER_CONSTANT, run_denoising.py:52-55) + Normalize + HWC->CHW float.  mats: B x 6 floats, the
 * map from OUTPUT to SOURCE pixel coordinates (NULL = identity); interp 1 bilinear / 0 nearest; round_u8 rounds the
 * sample to a grey level as cv2 does for uint8 images; dst = ((sample * scale) - mean) / std * post. */
int vu_resize_u8hwc(const uint8_t* src, uint8_t* dst, int B, int C, int Hs, int Ws, int Hd, int Wd, void* stream);
int vu_warp_u8hwc_to_chw(const uint8_t* src, float* dst, const float* mats, int B, int C, int Hs, int Ws, int Hd, int Wd,
                         int interp, float border, int round_u8, float scale, float mean, float std, float post,
                         void* stream);

/* ---------------------------------------------------------------- misc */
/* out = in * keep / (1-p), keep from the same counter-based stream the GEMM epilogue uses (index = flat element);
 * out is float32, or bfloat16 when out_bf16 != 0 (p == 0: a plain conversion) */
int vu_dropout(const float* in, void* out, int out_bf16, int64_t n, float p, uint64_t seed, uint32_t stream_id, void* stream);
/* bf16 mode: dst (bfloat16 [R][C], optional) and dst_t (bfloat16 [C][R], optional) = src (float32 [R][C]) -- the per-step
 * copies of an nn.Linear weight: W as the K-major B operand of y = x W^T, W^T as the K-major B operand of dx = dy W */
int vu_cast_bf16(const float* src, void* dst, void* dst_t, int R, int C, void* stream);
/* stream-ordered zero fill (cudaMemsetAsync) of `bytes` bytes: accumulators and gradient buffers */
int vu_zero(void* p, int64_t bytes, void* stream);
/* y = a*x + b*y elementwise */
int vu_axpby(const float* x, float* y, int64_t n, float a, float b, void* stream);
/* fused AdamW over one flat parameter buffer (torch.optim.AdamW semantics; run_denoising.py:81) */
int vu_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
             float eps, float weight_decay, int step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIT_UNET_B200_H */
