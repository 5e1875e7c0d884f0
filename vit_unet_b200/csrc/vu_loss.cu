// Fused scalar losses + their gradients (MSELoss: run_denoising.py:80; soft-Dice: README.md:91-101; L1 named by
// the benchmark configs), plus dropout regeneration, axpby and a fused AdamW step (run_denoising.py:81).
#include <cuda_bf16.h>

#include "vu_common.cuh"

namespace vu {

__global__ void __launch_bounds__(256)
loss_reduce_kernel(int kind, const float* __restrict__ p, const float* __restrict__ t, int64_t n, double* __restrict__ sums) {
  __shared__ double red[3 * 32];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float x = p[i], y = t[i];
    if (kind == VU_LOSS_L1) a0 += fabsf(x - y);
    else if (kind == VU_LOSS_MSE) { float d = x - y; a0 = fmaf(d, d, a0); }
    else { a0 = fmaf(x, y, a0); a1 += x; a2 += y; }
  }
  double v[3] = {a0, a1, a2};
  block_sum<3>(v, red);
  if (threadIdx.x == 0) { atomicAdd(sums, v[0]); if (kind == VU_LOSS_DICE) { atomicAdd(sums + 1, v[1]); atomicAdd(sums + 2, v[2]); } }
}

__global__ void loss_finalize_kernel(int kind, int64_t n, const double* __restrict__ sums, float* __restrict__ loss) {
  if (kind == VU_LOSS_DICE) loss[0] = (float)(1.0 - (2.0 * sums[0] + 1.0) / (sums[1] + sums[2] + 1.0));
  else loss[0] = (float)(sums[0] / (double)n);
}

__global__ void __launch_bounds__(256)
loss_bwd_kernel(int kind, const float* __restrict__ p, const float* __restrict__ t, int64_t n,
                const double* __restrict__ sums, const float* __restrict__ gscale, float* __restrict__ dp) {
  const float gs = gscale[0];
  float c0 = 0.f, c1 = 0.f;
  if (kind == VU_LOSS_DICE) {
    double den = sums[1] + sums[2] + 1.0, num = 2.0 * sums[0] + 1.0;
    c0 = (float)(-2.0 / den) * gs;          // coefficient of t_i
    c1 = (float)(num / (den * den)) * gs;   // constant term
  } else if (kind == VU_LOSS_L1) c0 = gs / (float)n;
  else c0 = 2.0f * gs / (float)n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float x = p[i], y = t[i], d = x - y;
    float r;
    if (kind == VU_LOSS_L1) r = d > 0.f ? c0 : (d < 0.f ? -c0 : 0.f);
    else if (kind == VU_LOSS_MSE) r = c0 * d;
    else r = fmaf(c0, y, c1);
    dp[i] = r;
  }
}

// TOut = float or __nv_bfloat16 (the bf16 mode hands the masked gradient straight to the tensor-core GEMMs; thresh == 0
// makes it a plain fp32 -> bf16 conversion)
template <typename TOut>
__global__ void __launch_bounds__(256)
dropout_kernel(const float* __restrict__ in, TOut* __restrict__ out, int64_t n, uint32_t thresh, float scale,
               uint64_t seed, uint32_t stream) {
  const bool vec = ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % (4 * sizeof(TOut)) == 0);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n; q += (int64_t)gridDim.x * blockDim.x) {
    uint4 r = thresh ? Philox::gen(seed, stream, (uint64_t)q) : make_uint4(0u, 0u, 0u, 0u);
    uint32_t rr[4] = {r.x, r.y, r.z, r.w};
    if (vec && q * 4 + 3 < n) {
      const float4 t = *reinterpret_cast<const float4*>(in + q * 4);
      float v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = rr[k] >= thresh ? v[k] * scale : 0.f;
      if constexpr (sizeof(TOut) == 4) *reinterpret_cast<float4*>(out + q * 4) = make_float4(v[0], v[1], v[2], v[3]);
      else {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
        uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(out + q * 4) = pk;
      }
      continue;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int64_t i = q * 4 + k;
      if (i < n) out[i] = (TOut)(rr[k] >= thresh ? in[i] * scale : 0.f);
    }
  }
}

// dst (bf16, [R][Cc]) and / or dstT (bf16, [Cc][R]) = src (fp32, [R][Cc]): the per-step bf16 copies of the Linear weights
// (W for the forward products, W^T as the K-major B operand of the data-gradient products).  32x32 tiles through smem.
__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, __nv_bfloat16* __restrict__ dstT, int R, int Cc) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i, c = c0 + tx;
    const float v = (r < R && c < Cc) ? src[(int64_t)r * Cc + c] : 0.f;
    tile[ty + i][tx] = v;
    if (dst && r < R && c < Cc) dst[(int64_t)r * Cc + c] = __float2bfloat16_rn(v);
  }
  if (!dstT) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, r = r0 + tx;
    if (c < Cc && r < R) dstT[(int64_t)c * R + r] = __float2bfloat16_rn(tile[tx][ty + i]);
  }
}

__global__ void __launch_bounds__(256)
axpby_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float a, float b) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = b == 0.f ? a * x[i] : fmaf(a, x[i], b * y[i]);
}

// torch.optim.AdamW (decoupled weight decay, bias-corrected, eps outside the sqrt of the corrected v)
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             int64_t n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float pi = p[i] * (1.0f - lr * wd);
    float mi = fmaf(b1, m[i], (1.0f - b1) * gi);
    float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// per-image sum of squared error: grid (VU_LN_SPLIT, B) partial sums merged with atomics (out zeroed by the caller)
__global__ void __launch_bounds__(512)
image_sse_kernel(const float* __restrict__ p, const float* __restrict__ t, int64_t n, int64_t len, double* __restrict__ out) {
  __shared__ double red[32];
  const int64_t beg = (int64_t)blockIdx.x * len, end = min(n, beg + len);
  const float* pb = p + (int64_t)blockIdx.y * n; const float* tb = t + (int64_t)blockIdx.y * n;
  float acc = 0.f;
  for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) { float d = pb[i] - tb[i]; acc = fmaf(d, d, acc); }
  double v[1] = {acc};
  block_sum<1>(v, red);
  if (threadIdx.x == 0) atomicAdd(out + blockIdx.y, v[0]);
}
// psnr[b] = 10 log10(range_b^2 / mse_b); range_b = data_range if > 0 else (min(target_b) >= 0 ? 1 : 2)  (skimage rule)
__global__ void psnr_finalize_kernel(const double* __restrict__ sse, const float* __restrict__ tmin, int B, int64_t n,
                                     float data_range, float* __restrict__ out) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float dr = data_range > 0.f ? data_range : (tmin[b] >= 0.f ? 1.f : 2.f);
  double mse = sse[b] / (double)n;
  out[b] = (float)(10.0 * log10((double)dr * dr / mse));
}
__global__ void __launch_bounds__(512)
image_min_kernel(const float* __restrict__ t, int64_t n, float* __restrict__ out) {
  __shared__ float red[32];
  const float* tb = t + (int64_t)blockIdx.x * n;
  float m = INFINITY;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = fminf(m, tb[i]);
  m = -warp_max(-m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : INFINITY;
    v = -warp_max(-v);
    if (threadIdx.x == 0) out[blockIdx.x] = v;
  }
}

// uint8 HWC (cv2.imread layout, any channel count C <= 4) -> float32 CHW, (v * scale - mean) / std per element:
// DenoisingDataset.__getitem__ (dataset.py:65-68: /255, transpose(2,0,1)) + albumentations.Normalize (run_denoising.py:54)
__global__ void __launch_bounds__(256)
u8hwc_to_chw_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t total, int C, int HW, float scale,
                    float mean, float inv_std) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = e / ((int64_t)C * HW); int r = (int)(e - b * C * HW);
    int c = r / HW, pix = r - c * HW;                      // enumerate in OUTPUT (CHW) order: coalesced stores
    dst[e] = ((float)src[(b * HW + pix) * C + c] * scale - mean) * inv_std;
  }
}

static int ew_grid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 256), (int64_t)sm_count() * 16)); }

}  // namespace vu

extern "C" int vu_loss_fwd(int kind, const float* pred, const float* target, int64_t n, double* sums, float* loss,
                           void* stream) {
  using namespace vu;
  const char* fn = "vu_loss_fwd";
  VU_REQUIRE(kind >= VU_LOSS_L1 && kind <= VU_LOSS_DICE, fn, "unknown loss kind");
  VU_REQUIRE(pred && target && sums && loss && n > 0, fn, "bad arguments");
  cudaStream_t s = as_stream(stream);
  if (cudaMemsetAsync(sums, 0, 4 * sizeof(double), s) != cudaSuccess) return check_launch(fn);
  int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 256 * 8), (int64_t)sm_count() * 4));
  loss_reduce_kernel<<<blocks, 256, 0, s>>>(kind, pred, target, n, sums);
  loss_finalize_kernel<<<1, 1, 0, s>>>(kind, n, sums, loss);
  return check_launch(fn);
}

// Second half of vu_loss_fwd on its own: the scalar from (possibly all-reduced) sums.  Data-parallel soft-Dice is a
// ratio of GLOBAL sums (README.md:96-101 flattens the whole batch), so the ranks add their three sums first.
extern "C" int vu_loss_finalize(int kind, int64_t n, const double* sums, float* loss, void* stream) {
  using namespace vu;
  const char* fn = "vu_loss_finalize";
  VU_REQUIRE(kind >= VU_LOSS_L1 && kind <= VU_LOSS_DICE, fn, "unknown loss kind");
  VU_REQUIRE(sums && loss && n > 0, fn, "bad arguments");
  loss_finalize_kernel<<<1, 1, 0, as_stream(stream)>>>(kind, n, sums, loss);
  return check_launch(fn);
}

extern "C" int vu_loss_bwd(int kind, const float* pred, const float* target, int64_t n, const double* sums,
                           const float* gscale, float* dpred, void* stream) {
  using namespace vu;
  const char* fn = "vu_loss_bwd";
  VU_REQUIRE(kind >= VU_LOSS_L1 && kind <= VU_LOSS_DICE, fn, "unknown loss kind");
  VU_REQUIRE(pred && target && sums && gscale && dpred && n > 0, fn, "bad arguments");
  loss_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(kind, pred, target, n, sums, gscale, dpred);
  return check_launch(fn);
}

extern "C" int vu_dropout(const float* in, void* out, int out_bf16, int64_t n, float p, uint64_t seed, uint32_t stream_id,
                          void* stream) {
  using namespace vu;
  const char* fn = "vu_dropout";
  VU_REQUIRE(in && out && n > 0 && p >= 0.f && p < 1.f, fn, "bad arguments");
  uint32_t th = p > 0.f ? drop_threshold(p) : 0u;
  if (out_bf16)
    dropout_kernel<__nv_bfloat16><<<ew_grid(cdiv(n, 4)), 256, 0, as_stream(stream)>>>(in, reinterpret_cast<__nv_bfloat16*>(out), n, th,
                                                                                  drop_keep_scale(p), seed, stream_id);
  else
    dropout_kernel<float><<<ew_grid(cdiv(n, 4)), 256, 0, as_stream(stream)>>>(in, reinterpret_cast<float*>(out), n, th,
                                                                          drop_keep_scale(p), seed, stream_id);
  return check_launch(fn);
}

extern "C" int vu_cast_bf16(const float* src, void* dst, void* dst_t, int R, int C, void* stream) {
  using namespace vu;
  const char* fn = "vu_cast_bf16";
  VU_REQUIRE(src && (dst || dst_t) && R > 0 && C > 0, fn, "bad arguments");
  dim3 grid((unsigned)cdiv(C, 32), (unsigned)cdiv(R, 32));
  cast_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, reinterpret_cast<__nv_bfloat16*>(dst),
                                                        reinterpret_cast<__nv_bfloat16*>(dst_t), R, C);
  return check_launch(fn);
}

extern "C" int vu_zero(void* p, int64_t bytes, void* stream) {
  using namespace vu;
  VU_REQUIRE(p && bytes >= 0, "vu_zero", "bad arguments");
  if (bytes && cudaMemsetAsync(p, 0, (size_t)bytes, as_stream(stream)) != cudaSuccess) return check_launch("vu_zero");
  return VU_OK;
}

extern "C" int vu_axpby(const float* x, float* y, int64_t n, float a, float b, void* stream) {
  using namespace vu;
  const char* fn = "vu_axpby";
  VU_REQUIRE(x && y && n > 0, fn, "bad arguments");
  axpby_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, y, n, a, b);
  return check_launch(fn);
}

extern "C" int vu_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                        float eps, float weight_decay, int step, float grad_scale, void* stream) {
  using namespace vu;
  const char* fn = "vu_adamw";
  VU_REQUIRE(p && g && m && v && n > 0 && step >= 1, fn, "bad arguments");
  float bc1 = 1.0f - powf(beta1, (float)step);
  float bc2s = sqrtf(1.0f - powf(beta2, (float)step));
  adamw_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2s, grad_scale);
  return check_launch(fn);
}

extern "C" int vu_psnr(const float* pred, const float* target, int B, int64_t n, float data_range, double* scratch,
                       float* psnr, void* stream) {
  using namespace vu;
  const char* fn = "vu_psnr";
  VU_REQUIRE(pred && target && scratch && psnr && B > 0 && n > 0, fn, "bad arguments");
  cudaStream_t s = as_stream(stream);
  // scratch: B doubles (sse) followed by B floats (per-image min of target)
  if (cudaMemsetAsync(scratch, 0, sizeof(double) * B, s) != cudaSuccess) return check_launch(fn);
  float* tmin = reinterpret_cast<float*>(scratch + B);
  const int S = VU_LN_SPLIT;
  const int64_t len = cdiv(n, S);
  image_sse_kernel<<<dim3(S, B), 512, 0, s>>>(pred, target, n, len, scratch);
  image_min_kernel<<<B, 512, 0, s>>>(target, n, tmin);
  psnr_finalize_kernel<<<(unsigned)cdiv(B, 128), 128, 0, s>>>(scratch, tmin, B, n, data_range, psnr);
  return check_launch(fn);
}

extern "C" int vu_u8hwc_to_chw(const uint8_t* src, float* dst, int B, int C, int H, int W, float scale, float mean,
                               float std, void* stream) {
  using namespace vu;
  const char* fn = "vu_u8hwc_to_chw";
  VU_REQUIRE(src && dst && B > 0 && C > 0 && H > 0 && W > 0 && std != 0.f, fn, "bad arguments");
  int64_t total = (int64_t)B * C * H * W;
  u8hwc_to_chw_kernel<<<ew_grid(total), 256, 0, as_stream(stream)>>>(src, dst, total, C, H * W, scale, mean, 1.0f / std);
  return check_launch(fn);
}
