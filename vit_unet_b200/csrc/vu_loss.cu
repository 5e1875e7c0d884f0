// Fused scalar losses + their gradients (MSELoss: run_denoising.py:80; soft-Dice: README.md:91-101; L1 named by
// the benchmark configs), plus dropout regeneration, axpby and a fused AdamW step (run_denoising.py:81).
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

__global__ void __launch_bounds__(256)
dropout_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, uint32_t thresh, float scale,
               uint64_t seed, uint32_t stream) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n; q += (int64_t)gridDim.x * blockDim.x) {
    uint4 r = Philox::gen(seed, stream, (uint64_t)q);
    uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int64_t i = q * 4 + k;
      if (i < n) out[i] = rr[k] >= thresh ? in[i] * scale : 0.f;
    }
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

extern "C" int vu_loss_bwd(int kind, const float* pred, const float* target, int64_t n, const double* sums,
                           const float* gscale, float* dpred, void* stream) {
  using namespace vu;
  const char* fn = "vu_loss_bwd";
  VU_REQUIRE(kind >= VU_LOSS_L1 && kind <= VU_LOSS_DICE, fn, "unknown loss kind");
  VU_REQUIRE(pred && target && sums && gscale && dpred && n > 0, fn, "bad arguments");
  loss_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(kind, pred, target, n, sums, gscale, dpred);
  return check_launch(fn);
}

extern "C" int vu_dropout(const float* in, float* out, int64_t n, float p, uint64_t seed, uint32_t stream_id, void* stream) {
  using namespace vu;
  const char* fn = "vu_dropout";
  VU_REQUIRE(in && out && n > 0 && p >= 0.f && p < 1.f, fn, "bad arguments");
  uint32_t th = p > 0.f ? drop_threshold(p) : 0u;
  dropout_kernel<<<ew_grid(cdiv(n, 4)), 256, 0, as_stream(stream)>>>(in, out, n, th, drop_keep_scale(p), seed, stream_id);
  return check_launch(fn);
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
