// LayerNorm over the whole (N, D) token matrix of an image (normalized_shape=(num_patches, projection_dim),
// model.py:193-196,204,206): a per-image reduction over n = C*H*W elements with an (N,D)-shaped affine.
#include "vu_common.cuh"

namespace vu {

// one CTA per image; two passes (mean, then centred variance) like ATen's CPU kernel -> fp32-exact statistics
__global__ void __launch_bounds__(1024)
ln_stats_kernel(const float* __restrict__ x, int64_t n, float eps, float* __restrict__ stats) {
  __shared__ double red[32];
  __shared__ float s_mean;
  const float* xb = x + (int64_t)blockIdx.x * n;
  const bool vec = (n % 4 == 0) && ((uintptr_t)xb % 16 == 0);
  double v[1]; float acc = 0.f;
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(xb);
    for (int64_t i = threadIdx.x; i < n / 4; i += blockDim.x) { float4 t = x4[i]; acc += (t.x + t.y) + (t.z + t.w); }
  } else {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += xb[i];
  }
  v[0] = acc; block_sum<1>(v, red);
  if (threadIdx.x == 0) s_mean = (float)(v[0] / (double)n);
  __syncthreads();
  const float mean = s_mean;
  acc = 0.f;
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(xb);
    for (int64_t i = threadIdx.x; i < n / 4; i += blockDim.x) {
      float4 t = x4[i];
      float a = t.x - mean, b = t.y - mean, c = t.z - mean, d = t.w - mean;
      acc += (a * a + b * b) + (c * c + d * d);
    }
  } else {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { float a = xb[i] - mean; acc = fmaf(a, a, acc); }
  }
  v[0] = acc; block_sum<1>(v, red);
  if (threadIdx.x == 0) {
    float var = (float)(v[0] / (double)n);
    stats[2 * blockIdx.x] = mean;
    stats[2 * blockIdx.x + 1] = 1.0f / sqrtf(var + eps);
  }
}

template <int V>
__global__ void __launch_bounds__(256)
ln_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ w,
                const float* __restrict__ bias, float* __restrict__ out, int64_t n, int64_t total) {
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
    } else {
      out[e] = fmaf((x[e] - mean) * rstd, w[r], bias[r]);
    }
  }
}

// scratch[2b] = sum g*w ; scratch[2b+1] = sum g*w*xhat   (one CTA per image)
__global__ void __launch_bounds__(1024)
ln_bwd_stats_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ stats,
                    const float* __restrict__ w, int64_t n, float* __restrict__ scratch) {
  __shared__ double red[64];
  const int64_t b = blockIdx.x;
  const float mean = stats[2 * b], rstd = stats[2 * b + 1];
  const float* gb = g + b * n; const float* xb = x + b * n;
  float a1 = 0.f, a2 = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    float gw = gb[i] * w[i];
    a1 += gw; a2 = fmaf(gw, (xb[i] - mean) * rstd, a2);
  }
  double v[2] = {a1, a2};
  block_sum<2>(v, red);
  if (threadIdx.x == 0) { scratch[2 * b] = (float)v[0]; scratch[2 * b + 1] = (float)v[1]; }
}

// dx = rstd * (g*w - a1/n - xhat * a2/n);  dw += sum_b g*xhat;  db += sum_b g
// grid (ceil(n/V/256), batch chunks); each CTA owns a slice of elements for a slice of the batch.
template <int V>
__global__ void __launch_bounds__(256)
ln_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ stats,
                    const float* __restrict__ w, const float* __restrict__ scratch, float* __restrict__ dx,
                    float* __restrict__ dw, float* __restrict__ db, int B, int64_t n, int b_per_chunk) {
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
    if (V == 4) *reinterpret_cast<float4*>(dx + (int64_t)b * n + e) = make_float4(o[0], o[1], o[2], o[3]);
    else dx[(int64_t)b * n + e] = o[0];
  }
#pragma unroll
  for (int k = 0; k < V; ++k) { atomicAdd(dw + e + k, aw[k]); atomicAdd(db + e + k, ab[k]); }
}

}  // namespace vu

extern "C" int vu_ln_stats(const float* x, int B, int64_t n, float eps, float* stats, void* stream) {
  using namespace vu;
  const char* fn = "vu_ln_stats";
  VU_REQUIRE(x && stats && B > 0 && n > 0, fn, "bad arguments");
  ln_stats_kernel<<<B, 1024, 0, as_stream(stream)>>>(x, n, eps, stats);
  return check_launch(fn);
}

extern "C" int vu_ln_apply(const float* x, const float* stats, const float* w, const float* b, float* out,
                           int B, int64_t n, void* stream) {
  using namespace vu;
  const char* fn = "vu_ln_apply";
  VU_REQUIRE(x && stats && w && b && out && B > 0 && n > 0, fn, "bad arguments");
  int64_t total = n * B;
  bool vec = (n % 4 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)w % 16 == 0) && ((uintptr_t)b % 16 == 0);
  int64_t work = vec ? total / 4 : total;
  int blocks = (int)std::min<int64_t>(cdiv(work, 256), (int64_t)sm_count() * 16);
  if (vec) ln_apply_kernel<4><<<blocks, 256, 0, as_stream(stream)>>>(x, stats, w, b, out, n, total);
  else ln_apply_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(x, stats, w, b, out, n, total);
  return check_launch(fn);
}

extern "C" int vu_ln_bwd(const float* g, const float* x, const float* stats, const float* w, float* dx,
                         float* dw, float* db, float* scratch, int B, int64_t n, void* stream) {
  using namespace vu;
  const char* fn = "vu_ln_bwd";
  VU_REQUIRE(g && x && stats && w && dx && dw && db && scratch && B > 0 && n > 0, fn, "bad arguments");
  cudaStream_t s = as_stream(stream);
  ln_bwd_stats_kernel<<<B, 1024, 0, s>>>(g, x, stats, w, n, scratch);
  int rc = check_launch(fn); if (rc) return rc;
  bool vec = (n % 4 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)dx % 16 == 0);
  int64_t work = vec ? n / 4 : n;
  int gx = (int)cdiv(work, 256);
  int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(B, cdiv((int64_t)sm_count() * 8, gx)));
  int bpc = (int)cdiv(B, chunks);
  dim3 grid(gx, (unsigned)cdiv(B, bpc));
  if (vec) ln_bwd_apply_kernel<4><<<grid, 256, 0, s>>>(g, x, stats, w, scratch, dx, dw, db, B, n, bpc);
  else ln_bwd_apply_kernel<1><<<grid, 256, 0, s>>>(g, x, stats, w, scratch, dx, dw, db, B, n, bpc);
  return check_launch(fn);
}
