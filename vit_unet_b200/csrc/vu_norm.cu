// LayerNorm over the whole (N, D) token matrix of an image (normalized_shape=(num_patches, projection_dim),
// model.py:193-196,204,206): a per-image reduction over n = C*H*W elements with an (N,D)-shaped affine.
#include <cuda_bf16.h>

#include <algorithm>

#include "vu_common.cuh"

namespace vu {

__device__ __forceinline__ void store4_bf16(__nv_bfloat16* p, float a, float b, float c, float d) {   // p 8-byte aligned
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = pk;
}

// grid (VU_LN_SPLIT, B): every CTA reduces one slice of an image with two passes (mean, then centred sum of
// squares -- the slice is L1/L2 resident for the second pass); slices are merged exactly (Chan et al.) by a tiny
// finalize kernel.  fp32-exact statistics like ATen's, but B*VU_LN_SPLIT CTAs instead of B.
__global__ void __launch_bounds__(512)
ln_stats_partial_kernel(const float* __restrict__ x, int64_t n, int64_t len, float* __restrict__ part) {
  __shared__ double red[32];
  __shared__ float s_mean;
  const int64_t beg = (int64_t)blockIdx.x * len, end = min(n, beg + len);
  const float* xb = x + (int64_t)blockIdx.y * n;
  const int64_t cnt = max((int64_t)0, end - beg);
  const bool vec = (beg % 4 == 0) && (cnt % 4 == 0) && ((uintptr_t)xb % 16 == 0);
  double v[1]; float acc = 0.f;
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(xb + beg);
    for (int64_t i = threadIdx.x; i < cnt / 4; i += blockDim.x) { float4 t = x4[i]; acc += (t.x + t.y) + (t.z + t.w); }
  } else {
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) acc += xb[i];
  }
  v[0] = acc; block_sum<1>(v, red);
  if (threadIdx.x == 0) s_mean = cnt > 0 ? (float)(v[0] / (double)cnt) : 0.f;
  __syncthreads();
  const float mean = s_mean;
  acc = 0.f;
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(xb + beg);
    for (int64_t i = threadIdx.x; i < cnt / 4; i += blockDim.x) {
      float4 t = x4[i];
      float a = t.x - mean, b = t.y - mean, c = t.z - mean, d = t.w - mean;
      acc += (a * a + b * b) + (c * c + d * d);
    }
  } else {
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) { float a = xb[i] - mean; acc = fmaf(a, a, acc); }
  }
  v[0] = acc; block_sum<1>(v, red);
  if (threadIdx.x == 0) {
    float* o = part + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 2;
    o[0] = mean; o[1] = (float)v[0];
  }
}

__global__ void ln_stats_finalize_kernel(const float* __restrict__ part, int B, int S, int64_t n, int64_t len, float eps,
                                         float* __restrict__ stats) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double tot = 0, mean = 0;
  for (int s = 0; s < S; ++s) {
    double c = (double)max((int64_t)0, min(n, (int64_t)(s + 1) * len) - (int64_t)s * len);
    mean += c * (double)part[((int64_t)b * S + s) * 2]; tot += c;
  }
  mean /= tot;
  double m2 = 0;
  for (int s = 0; s < S; ++s) {
    double c = (double)max((int64_t)0, min(n, (int64_t)(s + 1) * len) - (int64_t)s * len);
    double d = (double)part[((int64_t)b * S + s) * 2] - mean;
    m2 += (double)part[((int64_t)b * S + s) * 2 + 1] + c * d * d;
  }
  stats[2 * b] = (float)mean;
  stats[2 * b + 1] = 1.0f / sqrtf((float)(m2 / tot) + eps);
}

template <int V>
__global__ void __launch_bounds__(256)
ln_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ w,
                const float* __restrict__ bias, float* __restrict__ out, __nv_bfloat16* __restrict__ out16, int64_t n,
                int64_t total) {
  for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V; e < total; e += (int64_t)gridDim.x * blockDim.x * V) {
    int64_t b = e / n; int64_t r = e - b * n;
    float mean = __ldg(stats + 2 * b), rstd = __ldg(stats + 2 * b + 1);
    if (V == 4) {
      float4 t = *reinterpret_cast<const float4*>(x + e);
      float4 ww = *reinterpret_cast<const float4*>(w + r);
      float4 bb = *reinterpret_cast<const float4*>(bias + r);
      float4 o;
      o.x = fmaf((t.x - mean) * rstd, ww.x, bb.x); o.y = fmaf((t.y - mean) * rstd, ww.y, bb.y);
      o.z = fmaf((t.z - mean) * rstd, ww.z, bb.z); o.w = fmaf((t.w - mean) * rstd, ww.w, bb.w);
      *reinterpret_cast<float4*>(out + e) = o;
      if (out16) store4_bf16(out16 + e, o.x, o.y, o.z, o.w);
    } else {
      const float o = fmaf((x[e] - mean) * rstd, w[r], bias[r]);
      out[e] = o;
      if (out16) out16[e] = __float2bfloat16_rn(o);
    }
  }
}

// part[(b*S+s)*2 + {0,1}] = slice sums of g*w and g*w*xhat   (grid (S, B)); merged by ln_bwd_finalize_kernel
__global__ void __launch_bounds__(512)
ln_bwd_stats_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ stats,
                    const float* __restrict__ w, int64_t n, int64_t len, float* __restrict__ part) {
  __shared__ double red[64];
  const int64_t b = blockIdx.y;
  const int64_t beg = (int64_t)blockIdx.x * len, end = min(n, beg + len);
  const float mean = stats[2 * b], rstd = stats[2 * b + 1];
  const float* gb = g + b * n; const float* xb = x + b * n;
  float a1 = 0.f, a2 = 0.f;
  const int64_t cnt = max((int64_t)0, end - beg);
  if ((beg % 4 == 0) && (cnt % 4 == 0) && ((uintptr_t)gb % 16 == 0) && ((uintptr_t)xb % 16 == 0) && ((uintptr_t)w % 16 == 0)) {
    const float4* g4 = reinterpret_cast<const float4*>(gb + beg);
    const float4* x4 = reinterpret_cast<const float4*>(xb + beg);
    const float4* w4 = reinterpret_cast<const float4*>(w + beg);
    for (int64_t i = threadIdx.x; i < cnt / 4; i += blockDim.x) {
      const float4 gg = g4[i], xx = x4[i], ww = __ldg(w4 + i);
      const float g0 = gg.x * ww.x, g1 = gg.y * ww.y, g2 = gg.z * ww.z, g3 = gg.w * ww.w;
      a1 += (g0 + g1) + (g2 + g3);
      a2 = fmaf(g0, (xx.x - mean) * rstd, a2); a2 = fmaf(g1, (xx.y - mean) * rstd, a2);
      a2 = fmaf(g2, (xx.z - mean) * rstd, a2); a2 = fmaf(g3, (xx.w - mean) * rstd, a2);
    }
  } else {
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
      float gw = gb[i] * w[i];
      a1 += gw; a2 = fmaf(gw, (xb[i] - mean) * rstd, a2);
    }
  }
  double v[2] = {a1, a2};
  block_sum<2>(v, red);
  if (threadIdx.x == 0) {
    float* o = part + (b * gridDim.x + blockIdx.x) * 2;
    o[0] = (float)v[0]; o[1] = (float)v[1];
  }
}

__global__ void ln_bwd_finalize_kernel(const float* __restrict__ part, int B, int S, float* __restrict__ out) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double a1 = 0, a2 = 0;
  for (int s = 0; s < S; ++s) { a1 += (double)part[((int64_t)b * S + s) * 2]; a2 += (double)part[((int64_t)b * S + s) * 2 + 1]; }
  out[2 * b] = (float)a1; out[2 * b + 1] = (float)a2;
}

// dx = rstd * (g*w - a1/n - xhat * a2/n);  dw += sum_b g*xhat;  db += sum_b g
// grid (ceil(n/V/256), batch chunks); each CTA owns a slice of elements for a slice of the batch.
template <int V>
__global__ void __launch_bounds__(256)
ln_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ stats,
                    const float* __restrict__ w, const float* __restrict__ scratch, float* __restrict__ dx,
                    __nv_bfloat16* __restrict__ dx16, float* __restrict__ dw, float* __restrict__ db, int B, int64_t n,
                    int b_per_chunk) {
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (e >= n) return;
  const int b0 = blockIdx.y * b_per_chunk, b1 = min(B, b0 + b_per_chunk);
  const float inv_n = 1.0f / (float)n;
  float ww[V], aw[V], ab[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { ww[k] = w[e + k]; aw[k] = 0.f; ab[k] = 0.f; }
  for (int b = b0; b < b1; ++b) {
    const float mean = __ldg(stats + 2 * b), rstd = __ldg(stats + 2 * b + 1);
    const float c1 = __ldg(scratch + 2 * b) * inv_n, c2 = __ldg(scratch + 2 * b + 1) * inv_n;
    float gv[V], xv[V], o[V];
    if (V == 4) {
      float4 t = *reinterpret_cast<const float4*>(g + (int64_t)b * n + e);
      float4 u = *reinterpret_cast<const float4*>(x + (int64_t)b * n + e);
      gv[0] = t.x; gv[1] = t.y; gv[2] = t.z; gv[3] = t.w; xv[0] = u.x; xv[1] = u.y; xv[2] = u.z; xv[3] = u.w;
    } else { gv[0] = g[(int64_t)b * n + e]; xv[0] = x[(int64_t)b * n + e]; }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float xh = (xv[k] - mean) * rstd;
      o[k] = rstd * (gv[k] * ww[k] - c1 - xh * c2);
      aw[k] = fmaf(gv[k], xh, aw[k]); ab[k] += gv[k];
    }
    if (V == 4) {
      *reinterpret_cast<float4*>(dx + (int64_t)b * n + e) = make_float4(o[0], o[1], o[2], o[3]);
      if (dx16) store4_bf16(dx16 + (int64_t)b * n + e, o[0], o[1], o[2], o[3]);
    } else {
      dx[(int64_t)b * n + e] = o[0];
      if (dx16) dx16[(int64_t)b * n + e] = __float2bfloat16_rn(o[0]);
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) { atomicAdd(dw + e + k, aw[k]); atomicAdd(db + e + k, ab[k]); }
}

}  // namespace vu

extern "C" int vu_ln_stats(const float* x, int B, int64_t n, float eps, float* stats, float* scratch, void* stream) {
  using namespace vu;
  const char* fn = "vu_ln_stats";
  VU_REQUIRE(x && stats && scratch && B > 0 && n > 0, fn, "bad arguments");
  const int S = VU_LN_SPLIT;
  const int64_t len = cdiv(cdiv(n, S), 4) * 4;
  cudaStream_t s = as_stream(stream);
  ln_stats_partial_kernel<<<dim3(S, B), 512, 0, s>>>(x, n, len, scratch);
  ln_stats_finalize_kernel<<<(unsigned)cdiv(B, 128), 128, 0, s>>>(scratch, B, S, n, len, eps, stats);
  return check_launch(fn);
}

extern "C" int vu_ln_apply(const float* x, const float* stats, const float* w, const float* b, float* out,
                           void* out_bf16, int B, int64_t n, void* stream) {
  using namespace vu;
  const char* fn = "vu_ln_apply";
  VU_REQUIRE(x && stats && w && b && out && B > 0 && n > 0, fn, "bad arguments");
  int64_t total = n * B;
  __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  bool vec = (n % 4 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)w % 16 == 0) &&
             ((uintptr_t)b % 16 == 0) && ((uintptr_t)o16 % 8 == 0);
  int64_t work = vec ? total / 4 : total;
  int blocks = (int)std::min<int64_t>(cdiv(work, 256), (int64_t)sm_count() * 16);
  if (vec) ln_apply_kernel<4><<<blocks, 256, 0, as_stream(stream)>>>(x, stats, w, b, out, o16, n, total);
  else ln_apply_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(x, stats, w, b, out, o16, n, total);
  return check_launch(fn);
}

extern "C" int vu_ln_bwd(const float* g, const float* x, const float* stats, const float* w, float* dx, void* dx_bf16,
                         float* dw, float* db, float* scratch, int B, int64_t n, void* stream) {
  using namespace vu;
  const char* fn = "vu_ln_bwd";
  VU_REQUIRE(g && x && stats && w && dx && dw && db && scratch && B > 0 && n > 0, fn, "bad arguments");
  cudaStream_t s = as_stream(stream);
  const int S = VU_LN_SPLIT;
  const int64_t len = cdiv(cdiv(n, S), 4) * 4;
  float* part = scratch + 2 * (int64_t)B;          // scratch: [2B merged | 2*S*B partials]
  ln_bwd_stats_kernel<<<dim3(S, B), 512, 0, s>>>(g, x, stats, w, n, len, part);
  ln_bwd_finalize_kernel<<<(unsigned)cdiv(B, 128), 128, 0, s>>>(part, B, S, scratch);
  int rc = check_launch(fn); if (rc) return rc;
  __nv_bfloat16* dx16 = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
  bool vec = (n % 4 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)dx % 16 == 0) &&
             ((uintptr_t)dx16 % 8 == 0);
  int64_t work = vec ? n / 4 : n;
  int gx = (int)cdiv(work, 256);
  int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(B, cdiv((int64_t)sm_count() * 8, gx)));
  int bpc = (int)cdiv(B, chunks);
  dim3 grid(gx, (unsigned)cdiv(B, bpc));
  if (vec) ln_bwd_apply_kernel<4><<<grid, 256, 0, s>>>(g, x, stats, w, scratch, dx, dx16, dw, db, B, n, bpc);
  else ln_bwd_apply_kernel<1><<<grid, 256, 0, s>>>(g, x, stats, w, scratch, dx, dx16, dw, db, B, n, bpc);
  return check_launch(fn);
}
