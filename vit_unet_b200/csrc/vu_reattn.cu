// Re-Attention map kernels (DeepViT-style head mixing + BatchNorm over the attention maps):
//   attn = softmax(q k^T * scale); attn = dropout(attn); attn = BN_h(Conv1x1_{h->h}(attn))   model.py:155-159
// Maps are (B, h, N, ld) fp32, ld % 4 == 0.  The 1x1 conv + BatchNorm are folded into ONE h x h affine per position
// (SURVEY.md F5).  Everything that is a reduction over all (b,i,j) positions is expressed through CENTRED moments
// of the dropped maps Pd_g (centre c = 1/N, the exact mean of a softmax row):
//     forward  (train):  s'_g = sum (Pd_g - c)            G'_{gg'} = sum (Pd_g - c)(Pd_g' - c)
//     backward        :  s1_h = sum dA_h                  X'_{hg}  = sum dA_h (Pd_g - c)
// from which the BatchNorm batch statistics, the BatchNorm backward means and ALL parameter gradients (mixing
// matrix, conv bias, BN gamma/beta) follow in closed form (reattn_bwd_params_kernel) -- the big kernels only stream.
// Every streaming kernel handles 4 consecutive keys per thread (float4) so that one counter-hash call yields the
// dropout mask of the whole quad; masks are regenerated, never stored.
#include <cuda_bf16.h>
#include <algorithm>
#include <cstdlib>

#include "vu_common.cuh"

namespace vu {

// The mixed map A and the gradient map dA/dS may be stored as bfloat16 (half the HBM bytes; consumed by the bf16
// tensor-core GEMMs); the probabilities P are fp32 (or, optionally, centred bf16 on the tensor-core map path, see
// vu_reattn_mma.cuh).  The kernels in THIS file are the exact fp32 FMA versions for any head count 1..8.
__device__ __forceinline__ float4 map_ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 map_ld(const __nv_bfloat16* p) {
  // bf16 -> fp32 is a 16-bit shift: one SHL / one AND per element (the cuda_bf16.h conversions compile to PRMT + shift pairs)
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                     __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ void map_st(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void map_st(__nv_bfloat16* p, const float4& v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 u; u.x = *reinterpret_cast<uint32_t*>(&lo); u.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
}

// ------------------------------------------------------------------ row softmax (one warp per row)
__global__ void __launch_bounds__(256)
softmax_rows_kernel(float* __restrict__ S, int64_t rows, int N, int ld, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float sl2 = scale * 1.4426950408889634f;       // exp(x) = exp2(x*log2e)
  for (int64_t r = wid; r < rows; r += nw) {
    float* row = S + r * ld;
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, row[j] * sl2);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) sum += exp2f(fmaf(row[j], sl2, -mx));
    sum = warp_sum(sum);
    float inv = 1.0f / sum;
    for (int j = lane; j < N; j += 32) row[j] = exp2f(fmaf(row[j], sl2, -mx)) * inv;
    for (int j = N + lane; j < ld; j += 32) row[j] = 0.f;
  }
}

// ------------------------------------------------------------------ quad loader
// P quad of head g at (b, i, j..j+3) with dropout applied and the centre subtracted; invalid (pad) lanes -> 0.
struct QuadCtx {
  uint32_t thresh; float dscale; uint64_t seed; uint32_t stream; float c; int N;
  uint32_t key;          // Philox::key(seed, stream), computed on the host
};
__device__ __forceinline__ float4 load_pd(const float* __restrict__ p, uint64_t flat_idx, const QuadCtx& q) {
  float4 v = *reinterpret_cast<const float4*>(p);
  if (q.thresh) {
    uint4 rr = Philox::gen_k(q.key, (uint32_t)(flat_idx >> 2));
    v.x = rr.x >= q.thresh ? v.x * q.dscale : 0.f; v.y = rr.y >= q.thresh ? v.y * q.dscale : 0.f;
    v.z = rr.z >= q.thresh ? v.z * q.dscale : 0.f; v.w = rr.w >= q.thresh ? v.w * q.dscale : 0.f;
  }
  return v;
}
__device__ __forceinline__ void centre(float4& v, int j, const QuadCtx& q) {
  v.x = j + 0 < q.N ? v.x - q.c : 0.f; v.y = j + 1 < q.N ? v.y - q.c : 0.f;
  v.z = j + 2 < q.N ? v.z - q.c : 0.f; v.w = j + 3 < q.N ? v.w - q.c : 0.f;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float sum4(const float4& a) { return (a.x + a.y) + (a.z + a.w); }

// ------------------------------------------------------------------ forward statistics (train)
// sums[g] += sum (Pd_g - c);  sums[H + g*H + g'] += sum (Pd_g - c)(Pd_g' - c)   (g' >= g filled, mirrored later)
template <int H>
__global__ void __launch_bounds__(256)
reattn_stats_kernel(const float* __restrict__ P, int B, int N, int ld, QuadCtx q, double* __restrict__ sums) {
  constexpr int NV = H + H * (H + 1) / 2;
  __shared__ double red[NV * 32];
  const int ld4 = ld >> 2;
  const int64_t head_stride = (int64_t)N * ld, img_stride = head_stride * H;
  const int64_t per_img4 = (int64_t)N * ld4, total = per_img4 * B;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = t / per_img4; int64_t r = t - b * per_img4;
    int j = (int)(r % ld4) * 4;
    int64_t off = b * img_stride + r * 4;
    float4 p[H];
#pragma unroll
    for (int g = 0; g < H; ++g) { p[g] = load_pd(P + off + g * head_stride, (uint64_t)(off + g * head_stride), q); centre(p[g], j, q); }
    int k = H;
#pragma unroll
    for (int g = 0; g < H; ++g) {
      acc[g] += sum4(p[g]);
#pragma unroll
      for (int g2 = g; g2 < H; ++g2) { acc[k] += dot4(p[g], p[g2]); ++k; }
    }
  }
  double v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = acc[i];
  block_sum<NV>(v, red);
  if (threadIdx.x == 0) {
    int k = H;
    for (int g = 0; g < H; ++g) {
      atomicAdd(sums + g, v[g]);
      for (int g2 = g; g2 < H; ++g2) {
        atomicAdd(sums + H + g * H + g2, v[k]);
        if (g2 != g) atomicAdd(sums + H + g2 * H + g, v[k]);
        ++k;
      }
    }
  }
}

// Fused train-mode pass: one warp owns the rows (b, i) of ALL heads: softmax of every head in place, then the
// centred moments of the freshly written probabilities (re-read from L1/L2) -- one HBM read of S, one write of P.
template <int H>
__global__ void __launch_bounds__(256)
softmax_stats_kernel(float* __restrict__ S, int B, int N, int ld, float scale, QuadCtx q, double* __restrict__ sums) {
  constexpr int NV = H + H * (H + 1) / 2;
  __shared__ double red[NV * 32];
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float sl2 = scale * 1.4426950408889634f;
  const int64_t head_stride = (int64_t)N * ld, img_stride = head_stride * H;
  const int64_t rows = (int64_t)B * N;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  for (int64_t r = wid; r < rows; r += nw) {
    const int64_t b = r / N; const int i = (int)(r - b * N);
    const int64_t row_off = b * img_stride + (int64_t)i * ld;
#pragma unroll 1
    for (int g = 0; g < H; ++g) {
      float* row = S + row_off + g * head_stride;
      // sweep 1 (float4): row maximum;  sweep 2: sum of exp2;  sweep 3: write exp2 * 1/sum (pads -> 0).
      // sweeps 2 and 3 recompute exp2 instead of storing the un-normalised value: 3 reads (L1 after the first) + 1 write
      float mx = -INFINITY;
      for (int j = lane * 4; j < ld; j += 128) {
        const float4 t = *reinterpret_cast<const float4*>(row + j);
        if (j + 0 < N) mx = fmaxf(mx, t.x * sl2); if (j + 1 < N) mx = fmaxf(mx, t.y * sl2);
        if (j + 2 < N) mx = fmaxf(mx, t.z * sl2); if (j + 3 < N) mx = fmaxf(mx, t.w * sl2);
      }
      mx = warp_max(mx);
      float sum = 0.f;
      for (int j = lane * 4; j < ld; j += 128) {
        const float4 t = *reinterpret_cast<const float4*>(row + j);
        if (j + 0 < N) sum += exp2f(fmaf(t.x, sl2, -mx)); if (j + 1 < N) sum += exp2f(fmaf(t.y, sl2, -mx));
        if (j + 2 < N) sum += exp2f(fmaf(t.z, sl2, -mx)); if (j + 3 < N) sum += exp2f(fmaf(t.w, sl2, -mx));
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      for (int j = lane * 4; j < ld; j += 128) {
        const float4 t = *reinterpret_cast<const float4*>(row + j);
        float4 o;
        o.x = j + 0 < N ? exp2f(fmaf(t.x, sl2, -mx)) * inv : 0.f; o.y = j + 1 < N ? exp2f(fmaf(t.y, sl2, -mx)) * inv : 0.f;
        o.z = j + 2 < N ? exp2f(fmaf(t.z, sl2, -mx)) * inv : 0.f; o.w = j + 3 < N ? exp2f(fmaf(t.w, sl2, -mx)) * inv : 0.f;
        *reinterpret_cast<float4*>(row + j) = o;
      }
    }
    __syncwarp();
    for (int j = lane * 4; j < ld; j += 128) {
      float4 p[H];
#pragma unroll
      for (int g = 0; g < H; ++g) {
        const int64_t off = row_off + g * head_stride + j;
        p[g] = load_pd(S + off, (uint64_t)off, q); centre(p[g], j, q);
      }
      int k = H;
#pragma unroll
      for (int g = 0; g < H; ++g) {
        acc[g] += sum4(p[g]);
#pragma unroll
        for (int g2 = g; g2 < H; ++g2) { acc[k] += dot4(p[g], p[g2]); ++k; }
      }
    }
  }
  double v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = acc[i];
  block_sum<NV>(v, red);
  if (threadIdx.x == 0) {
    int k = H;
    for (int g = 0; g < H; ++g) {
      atomicAdd(sums + g, v[g]);
      for (int g2 = g; g2 < H; ++g2) {
        atomicAdd(sums + H + g * H + g2, v[k]);
        if (g2 != g) atomicAdd(sums + H + g2 * H + g, v[k]);
        ++k;
      }
    }
  }
}

// one block: statistics -> folded affine + saved (mean, invstd) + running-stat update
__global__ void reattn_bn_finalize_kernel(const double* __restrict__ sums, double count, int H, int N,
                                          const float* __restrict__ W, const float* __restrict__ bconv,
                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                          float* __restrict__ rmean, float* __restrict__ rvar,
                                          int64_t* __restrict__ nbt, float eps, float momentum, int train,
                                          float* __restrict__ fold, float* __restrict__ saved) {
  int h = threadIdx.x;
  if (h < H) {
    float mean, var;
    if (train) {
      // M_h - c_h = sum_g W_hg (Pd_g - c),  c_h = b_h + c * sum_g W_hg
      double rs = 0, m1 = 0, m2 = 0;
      for (int g = 0; g < H; ++g) {
        rs += (double)W[h * H + g];
        m1 += (double)W[h * H + g] * sums[g];
        for (int g2 = 0; g2 < H; ++g2) m2 += (double)W[h * H + g] * (double)W[h * H + g2] * sums[H + g * H + g2];
      }
      m1 /= count; m2 /= count;
      double c_h = (double)bconv[h] + rs / (double)N;
      double dmean = c_h + m1, dvar = m2 - m1 * m1;
      if (dvar < 0) dvar = 0;
      mean = (float)dmean; var = (float)dvar;
      double unbiased = count > 1 ? dvar * (count / (count - 1.0)) : dvar;
      rmean[h] = (1.f - momentum) * rmean[h] + momentum * mean;
      rvar[h] = (1.f - momentum) * rvar[h] + momentum * (float)unbiased;
    } else {
      mean = rmean[h]; var = rvar[h];
    }
    float invstd = 1.0f / sqrtf(var + eps);
    float a = gamma[h] * invstd;
    for (int g = 0; g < H; ++g) fold[h * H + g] = a * W[h * H + g];
    fold[H * H + h] = a * (bconv[h] - mean) + beta[h];
    saved[h] = mean; saved[H + h] = invstd;
  }
  if (train && threadIdx.x == 0 && nbt) *nbt += 1;
}

// A_h = sum_g fold[h][g] * Pd_g + fold[H*H + h]; 4 positions per thread along j (float4)
template <int H, typename MT>
__global__ void __launch_bounds__(256, 3)
reattn_mix_kernel(const float* __restrict__ P, MT* __restrict__ A, const float* __restrict__ fold,
                  int B, int N, int ld, QuadCtx q) {
  __shared__ float sF[H * H + H];
  for (int i = threadIdx.x; i < H * H + H; i += blockDim.x) sF[i] = fold[i];
  __syncthreads();
  const int ld4 = ld >> 2;
  const int64_t head_stride = (int64_t)N * ld, img_stride = head_stride * H;
  const int64_t per_img4 = (int64_t)N * ld4, total = per_img4 * B;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = t / per_img4; int64_t r = t - b * per_img4;      // r = i*ld4 + j4
    int j = (int)(r % ld4) * 4;
    int64_t off = b * img_stride + r * 4;
    float4 p[H];
#pragma unroll
    for (int g = 0; g < H; ++g) p[g] = load_pd(P + off + g * head_stride, (uint64_t)(off + g * head_stride), q);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float bb = sF[H * H + h];
      float4 a = make_float4(bb, bb, bb, bb);
#pragma unroll
      for (int g = 0; g < H; ++g) {
        float w = sF[h * H + g];
        a.x = fmaf(w, p[g].x, a.x); a.y = fmaf(w, p[g].y, a.y); a.z = fmaf(w, p[g].z, a.z); a.w = fmaf(w, p[g].w, a.w);
      }
      if (j + 3 >= N) {      // keep the pad columns at zero
        if (j + 0 >= N) a.x = 0.f; if (j + 1 >= N) a.y = 0.f; if (j + 2 >= N) a.z = 0.f; if (j + 3 >= N) a.w = 0.f;
      }
      map_st(A + off + h * head_stride, a);
    }
  }
}

// Backward fusion: one read of P and dA gives both the recomputed mixed map A (needed for dV = A^T dO) and the
// backward reductions red[h] += sum dA_h, red[H + h*H + g] += sum dA_h (Pd_g - c).
template <int H, typename MT>
__global__ void __launch_bounds__(256)
reattn_mix_reduce_kernel(const float* __restrict__ P, const MT* __restrict__ dA, MT* __restrict__ A,
                         const float* __restrict__ fold, int B, int N, int ld, QuadCtx q, double* __restrict__ out) {
  constexpr int NV = H + H * H;
  __shared__ double red[NV * 32];
  __shared__ float sF[H * H + H];
  for (int i = threadIdx.x; i < H * H + H; i += blockDim.x) sF[i] = fold[i];
  __syncthreads();
  const int ld4 = ld >> 2;
  const int64_t head_stride = (int64_t)N * ld, img_stride = head_stride * H;
  const int64_t per_img4 = (int64_t)N * ld4, total = per_img4 * B;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = t / per_img4; int64_t r = t - b * per_img4;
    int j = (int)(r % ld4) * 4;
    int64_t off = b * img_stride + r * 4;
    float4 p[H];
#pragma unroll
    for (int g = 0; g < H; ++g) p[g] = load_pd(P + off + g * head_stride, (uint64_t)(off + g * head_stride), q);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float bb = sF[H * H + h];
      float4 a = make_float4(bb, bb, bb, bb);
#pragma unroll
      for (int g = 0; g < H; ++g) {
        float w = sF[h * H + g];
        a.x = fmaf(w, p[g].x, a.x); a.y = fmaf(w, p[g].y, a.y); a.z = fmaf(w, p[g].z, a.z); a.w = fmaf(w, p[g].w, a.w);
      }
      if (j + 3 >= N) { if (j + 0 >= N) a.x = 0.f; if (j + 1 >= N) a.y = 0.f; if (j + 2 >= N) a.z = 0.f; if (j + 3 >= N) a.w = 0.f; }
      map_st(A + off + h * head_stride, a);
    }
#pragma unroll
    for (int g = 0; g < H; ++g) centre(p[g], j, q);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float4 d = map_ld(dA + off + h * head_stride);
      if (j + 3 >= N) { if (j + 0 >= N) d.x = 0.f; if (j + 1 >= N) d.y = 0.f; if (j + 2 >= N) d.z = 0.f; if (j + 3 >= N) d.w = 0.f; }
      acc[h] += sum4(d);
#pragma unroll
      for (int g = 0; g < H; ++g) acc[H + h * H + g] += dot4(d, p[g]);
    }
  }
  double v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = acc[i];
  block_sum<NV>(v, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) atomicAdd(out + i, v[i]);
  }
}

// ------------------------------------------------------------------ backward reductions
// red[h] += sum dA_h ;  red[H + h*H + g] += sum dA_h (Pd_g - c)
template <int H>
__global__ void __launch_bounds__(256)
reattn_bwd_reduce_kernel(const float* __restrict__ P, const float* __restrict__ dA, int B, int N, int ld, QuadCtx q,
                         double* __restrict__ out) {
  constexpr int NV = H + H * H;
  __shared__ double red[NV * 32];
  const int ld4 = ld >> 2;
  const int64_t head_stride = (int64_t)N * ld, img_stride = head_stride * H;
  const int64_t per_img4 = (int64_t)N * ld4, total = per_img4 * B;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = t / per_img4; int64_t r = t - b * per_img4;
    int j = (int)(r % ld4) * 4;
    int64_t off = b * img_stride + r * 4;
    float4 p[H];
#pragma unroll
    for (int g = 0; g < H; ++g) { p[g] = load_pd(P + off + g * head_stride, (uint64_t)(off + g * head_stride), q); centre(p[g], j, q); }
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float4 d = *reinterpret_cast<const float4*>(dA + off + h * head_stride);
      if (j + 3 >= N) { if (j + 0 >= N) d.x = 0.f; if (j + 1 >= N) d.y = 0.f; if (j + 2 >= N) d.z = 0.f; if (j + 3 >= N) d.w = 0.f; }
      acc[h] += sum4(d);
#pragma unroll
      for (int g = 0; g < H; ++g) acc[H + h * H + g] += dot4(d, p[g]);
    }
  }
  double v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = acc[i];
  block_sum<NV>(v, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) atomicAdd(out + i, v[i]);
  }
}

// One block.  From the backward reductions (red) and the forward centred moments (sums) derive
//   coef[h] = m1_h = mean(dA_h), coef[H+h] = m2_h = mean(dA_h * Ahat_h)      (train; zero in eval)
// and accumulate the parameter gradients dW[h][g], dbconv[h], dgamma[h], dbeta[h].
__global__ void reattn_bwd_params_kernel(const double* __restrict__ red, const double* __restrict__ sums, double count,
                                         int H, int N, const float* __restrict__ W, const float* __restrict__ bconv,
                                         const float* __restrict__ gamma, const float* __restrict__ saved, int train,
                                         float* __restrict__ coef, float* __restrict__ dW, float* __restrict__ dbconv,
                                         float* __restrict__ dgamma, float* __restrict__ dbeta) {
  int h = threadIdx.x;
  if (h >= H) return;
  const double c = 1.0 / (double)N;
  const double mean = saved[h], invstd = saved[H + h], k_h = (double)gamma[h] * invstd;
  const double s1 = red[h];
  double rs = 0, wx = 0;
  for (int g = 0; g < H; ++g) { rs += (double)W[h * H + g]; wx += (double)W[h * H + g] * red[H + h * H + g]; }
  const double c_h = (double)bconv[h] + c * rs;         // value of M_h when every Pd_g equals its centre
  double delta = mean - c_h;                            // M_h - mean = sum_g W_hg (Pd_g - c) - delta
  if (train) {                                          // batch statistics: delta is the centred mean, exactly
    delta = 0;
    for (int g = 0; g < H; ++g) delta += (double)W[h * H + g] * sums[g];
    delta /= count;
  }
  // s2 = sum dA_h * Ahat_h = invstd * ( sum_g W_hg X'_hg - delta * s1 )
  const double s2 = invstd * (wx - delta * s1);
  atomicAdd(dgamma + h, (float)s2);
  atomicAdd(dbeta + h, (float)s1);
  if (train) {
    const double m1 = s1 / count, m2 = s2 / count;
    coef[h] = (float)m1; coef[H + h] = (float)m2;
    // dW_hg = sum dM_h Pd_g = sum dM_h (Pd_g - c)      (sum dM_h == 0 under batch statistics)
    //       = k_h [ X'_hg - m1 s'_g - m2 invstd ( sum_g' W_hg' G'_g'g - delta s'_g ) ]
    for (int g = 0; g < H; ++g) {
      double wg = 0;
      for (int g2 = 0; g2 < H; ++g2) wg += (double)W[h * H + g2] * sums[H + g2 * H + g];
      double v = k_h * (red[H + h * H + g] - m1 * sums[g] - m2 * invstd * (wg - delta * sums[g]));
      atomicAdd(dW + h * H + g, (float)v);
    }
    // dbconv_h = sum dM_h = 0 exactly: nothing to add
  } else {
    coef[h] = 0.f; coef[H + h] = 0.f;
    // eval: dM_h = k_h dA_h  ->  dW_hg = k_h (X'_hg + c s1),  dbconv_h = k_h s1
    for (int g = 0; g < H; ++g) atomicAdd(dW + h * H + g, (float)(k_h * (red[H + h * H + g] + c * s1)));
    if (dbconv) atomicAdd(dbconv + h, (float)(k_h * s1));
  }
}

// One warp walks whole rows (b, i) for all heads, 4 keys per lane:
//   dM_h  = k_h (dA_h - m1_h - Ahat_h m2_h)   (train)   |   k_h dA_h   (eval)
//   dPd_g = sum_h W[h][g] dM_h ;  dP_g = keep_g dPd_g / (1-p)
//   r_g   = sum_j dP_g P_g ;       dS_g = scale * P_g (dP_g - r_g)           (written over dA)
template <int H, typename MT>
__global__ void __launch_bounds__(128)
reattn_bwd_rows_kernel(const float* __restrict__ P, MT* __restrict__ dA, int B, int N, int ld,
                       const float* __restrict__ W, const float* __restrict__ bconv, const float* __restrict__ gamma,
                       const float* __restrict__ saved, const float* __restrict__ coef, int train, float scale,
                       QuadCtx q) {
  __shared__ float sW[H * H];
  __shared__ float sOff[H], sInv[H], sK[H], sM1[H], sM2[H];
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) sW[i] = W[i];
  if (threadIdx.x < H) {
    int h = threadIdx.x;
    sOff[h] = bconv[h] - saved[h]; sInv[h] = saved[H + h]; sK[h] = gamma[h] * saved[H + h];
    sM1[h] = train ? coef[h] : 0.f; sM2[h] = train ? coef[H + h] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t head_stride = (int64_t)N * ld, img_stride = head_stride * H;
  const int64_t rows = (int64_t)B * N;
  for (int64_t r = wid; r < rows; r += nw) {
    int64_t b = r / N; int i = (int)(r - b * N);
    const int64_t row_off = b * img_stride + (int64_t)i * ld;
    float rg[H];
#pragma unroll
    for (int g = 0; g < H; ++g) rg[g] = 0.f;
    for (int j = lane * 4; j < ld; j += 128) {
      float4 p[H], pd[H], dm[H];
      uint4 keep[H];
#pragma unroll
      for (int g = 0; g < H; ++g) {
        p[g] = *reinterpret_cast<const float4*>(P + row_off + g * head_stride + j);
        pd[g] = p[g];
        keep[g] = make_uint4(1, 1, 1, 1);
        if (q.thresh) {
          uint4 rr = Philox::gen_k(q.key, (uint32_t)((uint64_t)(row_off + g * head_stride + j) >> 2));
          keep[g] = make_uint4(rr.x >= q.thresh, rr.y >= q.thresh, rr.z >= q.thresh, rr.w >= q.thresh);
          pd[g].x = keep[g].x ? p[g].x * q.dscale : 0.f; pd[g].y = keep[g].y ? p[g].y * q.dscale : 0.f;
          pd[g].z = keep[g].z ? p[g].z * q.dscale : 0.f; pd[g].w = keep[g].w ? p[g].w * q.dscale : 0.f;
        }
      }
#pragma unroll
      for (int h = 0; h < H; ++h) {
        float4 d = map_ld(dA + row_off + h * head_stride + j);
        float4 t = d;
        if (train) {
          float4 m = make_float4(sOff[h], sOff[h], sOff[h], sOff[h]);
#pragma unroll
          for (int g = 0; g < H; ++g) {
            float w = sW[h * H + g];
            m.x = fmaf(w, pd[g].x, m.x); m.y = fmaf(w, pd[g].y, m.y); m.z = fmaf(w, pd[g].z, m.z); m.w = fmaf(w, pd[g].w, m.w);
          }
          const float a2 = sInv[h] * sM2[h], a1 = sM1[h];
          t.x = d.x - a1 - m.x * a2; t.y = d.y - a1 - m.y * a2; t.z = d.z - a1 - m.z * a2; t.w = d.w - a1 - m.w * a2;
        }
        const float kh = sK[h];
        dm[h] = make_float4(kh * t.x, kh * t.y, kh * t.z, kh * t.w);
      }
#pragma unroll
      for (int g = 0; g < H; ++g) {
        float4 dp = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int h = 0; h < H; ++h) {
          float w = sW[h * H + g];
          dp.x = fmaf(w, dm[h].x, dp.x); dp.y = fmaf(w, dm[h].y, dp.y); dp.z = fmaf(w, dm[h].z, dp.z); dp.w = fmaf(w, dm[h].w, dp.w);
        }
        dp.x = (keep[g].x && j + 0 < N) ? dp.x * q.dscale : 0.f; dp.y = (keep[g].y && j + 1 < N) ? dp.y * q.dscale : 0.f;
        dp.z = (keep[g].z && j + 2 < N) ? dp.z * q.dscale : 0.f; dp.w = (keep[g].w && j + 3 < N) ? dp.w * q.dscale : 0.f;
        rg[g] += dot4(dp, p[g]);
        map_st(dA + row_off + g * head_stride + j, dp);
      }
    }
#pragma unroll
    for (int g = 0; g < H; ++g) rg[g] = warp_sum(rg[g]);
    for (int j = lane * 4; j < ld; j += 128) {
#pragma unroll
      for (int g = 0; g < H; ++g) {
        float4 pv = *reinterpret_cast<const float4*>(P + row_off + g * head_stride + j);
        float4 dp = map_ld(dA + row_off + g * head_stride + j);
        float4 o;
        o.x = scale * pv.x * (dp.x - rg[g]); o.y = scale * pv.y * (dp.y - rg[g]);
        o.z = scale * pv.z * (dp.z - rg[g]); o.w = scale * pv.w * (dp.w - rg[g]);   // pad columns: P == 0 -> 0
        map_st(dA + row_off + g * head_stride + j, o);
      }
    }
  }
}

}  // namespace vu
#include "vu_reattn_mma.cuh"
namespace vu {

// 8-head tensor-core formulation (vu_reattn_mma.cuh): TF32 path, no pad columns.  VU_MAP_MMA=0 disables it.
static bool mma_path(int h, int N, int ld) {
  static const bool on = []() { const char* e = getenv("VU_MAP_MMA"); return !(e && e[0] == '0'); }();
  return on && h == 8 && ld == N && N % 4 == 0 && N <= 8192;   // 32-bit offsets inside one image
}
// persistent grid: as many CTAs as are resident (occupancy query, cached per kernel), capped by the work
template <typename K>
static int resident_grid(K kernel, int threads, int64_t warps_of_work, size_t dyn_smem = 0) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem);
  if (per_sm < 1) per_sm = 1;
  const int64_t need = cdiv(warps_of_work, threads / 32);
  return (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)sm_count() * per_sm));
}

static int grid_for(int64_t work_items, int threads, int per_sm) {
  int64_t b = cdiv(work_items, threads);
  int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)std::max<int64_t>(1, std::min(b, cap));
}
static QuadCtx make_ctx(float drop_p, uint64_t seed, uint32_t stream_id, int N) {
  QuadCtx q;
  q.thresh = drop_p > 0.f ? drop_threshold(drop_p) : 0u;
  q.dscale = drop_keep_scale(drop_p); q.seed = seed; q.stream = stream_id; q.c = 1.0f / (float)N; q.N = N;
  q.key = Philox::key(seed, stream_id);
  return q;
}

}  // namespace vu

#define VU_DISPATCH_H(h, fn, ...)                                   \
  switch (h) {                                                      \
    case 1: { constexpr int HH = 1; __VA_ARGS__; } break;           \
    case 2: { constexpr int HH = 2; __VA_ARGS__; } break;           \
    case 3: { constexpr int HH = 3; __VA_ARGS__; } break;           \
    case 4: { constexpr int HH = 4; __VA_ARGS__; } break;           \
    case 5: { constexpr int HH = 5; __VA_ARGS__; } break;           \
    case 6: { constexpr int HH = 6; __VA_ARGS__; } break;           \
    case 7: { constexpr int HH = 7; __VA_ARGS__; } break;           \
    case 8: { constexpr int HH = 8; __VA_ARGS__; } break;           \
    default: return vu::fail_arg(fn, "num_heads must be in 1..8");  \
  }
#define VU_MAP_ARGS_OK(P) ((P) && B > 0 && N > 0 && ld >= N && ld % 4 == 0 && ((uintptr_t)(P) % 16 == 0))

extern "C" int vu_reattn_tensor_core_path(int h, int N, int ld) { return vu::mma_path(h, N, ld) ? 1 : 0; }

extern "C" int vu_softmax_rows(float* S, int64_t rows, int N, int ld, float scale, void* stream) {
  using namespace vu;
  const char* fn = "vu_softmax_rows";
  VU_REQUIRE(S && rows > 0 && N > 0 && ld >= N, fn, "bad arguments");
  int blocks = grid_for(rows * 32, 256, 16);
  softmax_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(S, rows, N, ld, scale);
  return check_launch(fn);
}

extern "C" int vu_reattn_stats(const float* P, int B, int h, int N, int ld, float drop_p, uint64_t seed,
                               uint32_t stream_id, double* sums, void* stream) {
  using namespace vu;
  const char* fn = "vu_reattn_stats";
  VU_REQUIRE(VU_MAP_ARGS_OK(P) && sums, fn, "bad arguments (maps need ld % 4 == 0 and 16-byte alignment)");
  VU_REQUIRE(drop_p >= 0.f && drop_p < 1.f, fn, "drop_p must be in [0,1)");
  QuadCtx q = make_ctx(drop_p, seed, stream_id, N);
  int blocks = grid_for((int64_t)B * N * (ld / 4), 256 * 2, 4);
  VU_DISPATCH_H(h, fn, reattn_stats_kernel<HH><<<blocks, 256, 0, as_stream(stream)>>>(P, B, N, ld, q, sums));
  return check_launch(fn);
}

extern "C" int vu_reattn_bn_finalize(const double* sums, int64_t count, int h, int N, const float* W,
                                     const float* bconv, const float* gamma, const float* beta,
                                     float* running_mean, float* running_var, int64_t* num_batches_tracked,
                                     float eps, float momentum, int train, float* fold, float* saved, void* stream) {
  using namespace vu;
  const char* fn = "vu_reattn_bn_finalize";
  VU_REQUIRE(W && bconv && gamma && beta && running_mean && running_var && fold && saved, fn, "null pointer");
  VU_REQUIRE(h >= 1 && h <= 32 && N > 0, fn, "bad head count");
  VU_REQUIRE(!train || (sums && count > 0), fn, "train mode needs sums and count");
  reattn_bn_finalize_kernel<<<1, 32, 0, as_stream(stream)>>>(sums, (double)count, h, N, W, bconv, gamma, beta,
                                                             running_mean, running_var, num_batches_tracked,
                                                             eps, momentum, train, fold, saved);
  return check_launch(fn);
}

extern "C" int vu_reattn_mix(const void* Pv, void* A, int map_fmt, const float* fold, int B, int h, int N, int ld,
                             float drop_p, uint64_t seed, uint32_t stream_id, void* stream) {
  using namespace vu;
  const char* fn = "vu_reattn_mix";
  const int map_bf16 = map_fmt & VU_MAP_BF16, p_bf16 = map_fmt & VU_MAP_P_CENTRED_BF16;
  const float* P = (const float*)Pv;
  VU_REQUIRE(VU_MAP_ARGS_OK(P) && A && fold && ((uintptr_t)A % 16 == 0), fn, "bad arguments");
  VU_REQUIRE(!map_bf16 || ld % 8 == 0, fn, "bf16 maps need ld % 8 == 0");
  VU_REQUIRE(drop_p >= 0.f && drop_p < 1.f, fn, "drop_p must be in [0,1)");
  VU_REQUIRE(!p_bf16 || (map_bf16 && mma_path(h, N, ld)), fn, "centred bf16 probabilities need bf16 maps, h == 8, ld == N, N % 8 == 0");
  QuadCtx q = make_ctx(drop_p, seed, stream_id, N);
  if ((map_bf16 || (map_fmt & VU_MAP_TF32_MIX)) && mma_path(h, N, ld)) {
    VU_REQUIRE(B <= 65535, fn, "at most 65535 images per call on the tensor-core map path");
    const int64_t tiles = cdiv((int64_t)N * N / 4, 8);          // per image; grid = (x, B)
    const dim3 grid((unsigned)std::max<int64_t>(1, cdiv(tiles, 8 * 2 * 8)), B);   // 8 warps x 2 tiles x ~8 iterations per CTA
    cudaStream_t st = as_stream(stream);
    if (p_bf16) mma::reattn_mix_mma_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)Pv, (__nv_bfloat16*)A, fold, N, q);
    else if (map_bf16) mma::reattn_mix_mma_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(P, (__nv_bfloat16*)A, fold, N, q);
    else mma::reattn_mix_mma_kernel<float, float><<<grid, 256, 0, st>>>(P, (float*)A, fold, N, q);
    return check_launch(fn);
  }
  int blocks = grid_for((int64_t)B * N * (ld / 4), 256, 16);
  if (map_bf16) { VU_DISPATCH_H(h, fn, reattn_mix_kernel<HH, __nv_bfloat16><<<blocks, 256, 0, as_stream(stream)>>>(P, (__nv_bfloat16*)A, fold, B, N, ld, q)); }
  else { VU_DISPATCH_H(h, fn, reattn_mix_kernel<HH, float><<<blocks, 256, 0, as_stream(stream)>>>(P, (float*)A, fold, B, N, ld, q)); }
  return check_launch(fn);
}

extern "C" int vu_reattn_bwd_reduce(const float* P, const float* dA, int B, int h, int N, int ld, float drop_p,
                                    uint64_t seed, uint32_t stream_id, double* red, void* stream) {
  using namespace vu;
  const char* fn = "vu_reattn_bwd_reduce";
  VU_REQUIRE(VU_MAP_ARGS_OK(P) && dA && red && ((uintptr_t)dA % 16 == 0), fn, "bad arguments");
  QuadCtx q = make_ctx(drop_p, seed, stream_id, N);
  int blocks = grid_for((int64_t)B * N * (ld / 4), 256 * 2, 4);
  VU_DISPATCH_H(h, fn, reattn_bwd_reduce_kernel<HH><<<blocks, 256, 0, as_stream(stream)>>>(P, dA, B, N, ld, q, red));
  return check_launch(fn);
}

extern "C" int vu_reattn_bwd_params(const double* red, const double* sums, int B, int h, int N, const float* W,
                                    const float* bconv, const float* gamma, const float* saved, int train,
                                    float* coef, float* dW, float* dbconv, float* dgamma, float* dbeta, void* stream) {
  using namespace vu;
  const char* fn = "vu_reattn_bwd_params";
  VU_REQUIRE(red && W && bconv && gamma && saved && coef && dW && dgamma && dbeta, fn, "null pointer");
  VU_REQUIRE(!train || sums, fn, "train mode needs the forward moments");
  VU_REQUIRE(h >= 1 && h <= 32 && B > 0 && N > 0, fn, "bad shape");
  reattn_bwd_params_kernel<<<1, 32, 0, as_stream(stream)>>>(red, sums, (double)B * N * N, h, N, W, bconv, gamma, saved,
                                                            train, coef, dW, dbconv, dgamma, dbeta);
  return check_launch(fn);
}

extern "C" int vu_reattn_bwd_rows(const void* Pv, void* dA_dS, int map_fmt, int B, int h, int N, int ld, const float* W,
                                  const float* bconv, const float* gamma, const float* saved, const float* coef,
                                  int train, float scale, float drop_p, uint64_t seed, uint32_t stream_id, void* stream) {
  using namespace vu;
  const char* fn = "vu_reattn_bwd_rows";
  const int map_bf16 = map_fmt & VU_MAP_BF16, p_bf16 = map_fmt & VU_MAP_P_CENTRED_BF16;
  const float* P = (const float*)Pv;
  VU_REQUIRE(VU_MAP_ARGS_OK(P) && dA_dS && W && bconv && gamma && saved && ((uintptr_t)dA_dS % 16 == 0), fn, "bad arguments");
  VU_REQUIRE(!train || coef, fn, "train mode needs the BN-backward coefficients");
  QuadCtx q = make_ctx(drop_p, seed, stream_id, N);
  VU_REQUIRE(!map_bf16 || ld % 8 == 0, fn, "bf16 maps need ld % 8 == 0");
  VU_REQUIRE(!p_bf16 || (map_bf16 && mma_path(h, N, ld)), fn, "centred bf16 probabilities need bf16 maps, h == 8, ld == N, N % 8 == 0");
  if ((map_bf16 || (map_fmt & VU_MAP_TF32_MIX)) && mma_path(h, N, ld)) {
    cudaStream_t st = as_stream(stream);
    __nv_bfloat16* d = (__nv_bfloat16*)dA_dS;
    const __nv_bfloat16* Pb = (const __nv_bfloat16*)Pv;
    if (map_bf16 && N > 256 && N <= 1024) {         // long rows: one CTA per row, row kept in registers between the sweeps
      // (warps, tiles per warp): 5 x 5 when the 32-key tiles divide by five (N = 784: 25 tiles; 10.1 vs 11.5 ms per step
      // against 8 x 4, where one warp does four rounds and seven do three), else 8 x 4
      static const int nw_env = []() { const char* e = getenv("VU_ROWS_NW"); return e ? atoi(e) : 0; }();
      const int ntiles = (N / 4 + 7) / 8;
      int nw = (ntiles <= 25 && ntiles % 5 == 0) ? 5 : 8;
      if (ntiles <= 28 && (nw_env == 7)) nw = 7;
      if (nw_env == 8) nw = 8;
      if (nw_env == 5 && ntiles <= 25) nw = 5;
#define VU_ROWS(TPWV, NWV, PTV, PPTR)                                                                               \
      do {                                                                                                          \
        const int grid = resident_grid(mma::reattn_bwd_rows_mma_cta_kernel<TPWV, NWV, PTV>, NWV * 32, (int64_t)B * N * NWV); \
        mma::reattn_bwd_rows_mma_cta_kernel<TPWV, NWV, PTV><<<grid, NWV * 32, 0, st>>>(PPTR, d, B, N, W, bconv, gamma, saved, coef, train, scale, q); \
      } while (0)
      // centred bf16 probabilities, 5 x 5: four resident CTAs (96 registers, 8 bytes of spill) instead of three at 128
      // registers -- the kernel is latency-bound: 8.90 vs 9.21 ms per Base step (five CTAs at 72 registers spill: 11.45 ms; 7 warps x 4
      // tiles: 10.87).  VU_ROWS_MINB=3 restores three.
      static const bool minb3 = []() { const char* e = getenv("VU_ROWS_MINB"); return e && atoi(e) == 3; }();
      if (p_bf16 && nw == 5 && !minb3) {
        const int grid = resident_grid(mma::reattn_bwd_rows_mma_cta_kernel<5, 5, __nv_bfloat16, 4>, 5 * 32, (int64_t)B * N * 5);
        mma::reattn_bwd_rows_mma_cta_kernel<5, 5, __nv_bfloat16, 4><<<grid, 5 * 32, 0, st>>>(Pb, d, B, N, W, bconv, gamma, saved, coef, train, scale, q);
        return check_launch(fn);
      }
      if (p_bf16) { if (nw == 5) VU_ROWS(5, 5, __nv_bfloat16, Pb); else if (nw == 7) VU_ROWS(4, 7, __nv_bfloat16, Pb); else VU_ROWS(4, 8, __nv_bfloat16, Pb); }
      else { if (nw == 5) VU_ROWS(5, 5, float, P); else if (nw == 7) VU_ROWS(4, 7, float, P); else VU_ROWS(4, 8, float, P); }
#undef VU_ROWS
      return check_launch(fn);
    }
    if (p_bf16) {
      const int grid = resident_grid(mma::reattn_bwd_rows_mma_kernel<__nv_bfloat16, __nv_bfloat16>, 256, (int64_t)B * N);
      mma::reattn_bwd_rows_mma_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(Pb, d, B, N, W, bconv, gamma, saved, coef, train, scale, q);
    } else if (map_bf16) {
      const int grid = resident_grid(mma::reattn_bwd_rows_mma_kernel<float, __nv_bfloat16>, 256, (int64_t)B * N);
      mma::reattn_bwd_rows_mma_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(P, d, B, N, W, bconv, gamma, saved, coef, train, scale, q);
    } else {
      const int grid = resident_grid(mma::reattn_bwd_rows_mma_kernel<float, float>, 256, (int64_t)B * N);
      mma::reattn_bwd_rows_mma_kernel<float, float><<<grid, 256, 0, st>>>(P, (float*)dA_dS, B, N, W, bconv, gamma, saved, coef, train, scale, q);
    }
    return check_launch(fn);
  }
  int blocks = grid_for((int64_t)B * N * 32, 128, 12);
  if (map_bf16) { VU_DISPATCH_H(h, fn, reattn_bwd_rows_kernel<HH, __nv_bfloat16><<<blocks, 128, 0, as_stream(stream)>>>(
      P, (__nv_bfloat16*)dA_dS, B, N, ld, W, bconv, gamma, saved, coef, train, scale, q)); }
  else { VU_DISPATCH_H(h, fn, reattn_bwd_rows_kernel<HH, float><<<blocks, 128, 0, as_stream(stream)>>>(
      P, (float*)dA_dS, B, N, ld, W, bconv, gamma, saved, coef, train, scale, q)); }
  return check_launch(fn);
}

extern "C" int vu_softmax_stats(float* S, void* Pc, int B, int h, int N, int ld, float scale, float drop_p, uint64_t seed,
                                uint32_t stream_id, double* sums, int precision, void* stream) {
  using namespace vu;
  const char* fn = "vu_softmax_stats";
  VU_REQUIRE(VU_MAP_ARGS_OK(S) && sums, fn, "bad arguments (maps need ld % 4 == 0 and 16-byte alignment)");
  VU_REQUIRE(drop_p >= 0.f && drop_p < 1.f, fn, "drop_p must be in [0,1)");
  QuadCtx q = make_ctx(drop_p, seed, stream_id, N);
  VU_REQUIRE(precision == VU_PREC_FP32 || precision == VU_PREC_TF32, fn, "precision must be VU_PREC_FP32 or VU_PREC_TF32");
  const bool use_mma = precision == VU_PREC_TF32 && mma_path(h, N, ld);
  VU_REQUIRE(!Pc || (use_mma && (uintptr_t)Pc % 16 == 0), fn,
             "centred bf16 output needs VU_PREC_TF32, h == 8, ld == N, N % 8 == 0 and 16-byte alignment");
  if (use_mma) {
    if (N > 256 && N <= 1024) {          // asynchronous row pipeline (cp.async.bulk ring in shared memory)
      constexpr int ST = 2;
      const size_t smem = mma::bulk_smem_bytes(ST, 8 * N);
      // consumer warps: 8 (measured at N = 784: 8.9 ms per step with 8 warps, 9.0 with 7, 9.7 with the evenly
      // dividing 5 -- the shared-memory sweeps want the extra warps more than the balance; nine warps (three even rounds)
      // at 48 registers / four CTAs per SM: 9.5 vs 8.6 ms); VU_SOFTMAX_NW overrides
      static const int nw_env = []() { const char* e = getenv("VU_SOFTMAX_NW"); return e ? atoi(e) : 0; }();
      int nw = 8;
      if (nw_env >= 5 && nw_env <= 8) nw = nw_env;
      cudaStream_t st = as_stream(stream);
#define VU_SMB(NWV)                                                                                                       \
      do {                                                                                                                \
        static uint64_t seen = 0;                                                                                         \
        if (first_use_on_device(seen)) cudaFuncSetAttribute(mma::softmax_stats_mma_bulk_kernel<ST, NWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
        const int grid = resident_grid(mma::softmax_stats_mma_bulk_kernel<ST, NWV>, (NWV + 1) * 32, (int64_t)B * N * (NWV + 1), smem); \
        mma::softmax_stats_mma_bulk_kernel<ST, NWV><<<grid, (NWV + 1) * 32, smem, st>>>(S, (__nv_bfloat16*)Pc, B, N, scale, q, sums); \
      } while (0)
      if (nw == 5) VU_SMB(5); else if (nw == 6) VU_SMB(6); else if (nw == 7) VU_SMB(7); else VU_SMB(8);
#undef VU_SMB
      return check_launch(fn);
    }
    const int grid = resident_grid(mma::softmax_stats_mma_kernel, 256, (int64_t)B * N);
    mma::softmax_stats_mma_kernel<<<grid, 256, 0, as_stream(stream)>>>(S, (__nv_bfloat16*)Pc, B, N, scale, q, sums);
    return check_launch(fn);
  }
  int blocks = grid_for((int64_t)B * N * 32, 256, 8);
  VU_DISPATCH_H(h, fn, softmax_stats_kernel<HH><<<blocks, 256, 0, as_stream(stream)>>>(S, B, N, ld, scale, q, sums));
  return check_launch(fn);
}

extern "C" int vu_reattn_mix_reduce(const void* Pv, const void* dA, void* A, int map_fmt, const float* fold, int B, int h,
                                    int N, int ld, float drop_p, uint64_t seed, uint32_t stream_id, double* red,
                                    void* stream) {
  using namespace vu;
  const char* fn = "vu_reattn_mix_reduce";
  const int map_bf16 = map_fmt & VU_MAP_BF16, p_bf16 = map_fmt & VU_MAP_P_CENTRED_BF16;
  const float* P = (const float*)Pv;
  VU_REQUIRE(VU_MAP_ARGS_OK(P) && dA && fold && red && ((uintptr_t)dA % 16 == 0) && ((uintptr_t)A % 16 == 0), fn,
             "bad arguments");
  VU_REQUIRE(A || ((map_bf16 || (map_fmt & VU_MAP_TF32_MIX)) && mma_path(h, N, ld)), fn,
             "A == NULL (reductions only) is available on the tensor-core map path only");
  VU_REQUIRE(!map_bf16 || ld % 8 == 0, fn, "bf16 maps need ld % 8 == 0");
  VU_REQUIRE(A != dA, fn, "A and dA must be distinct buffers");
  VU_REQUIRE(!p_bf16 || (map_bf16 && mma_path(h, N, ld)), fn, "centred bf16 probabilities need bf16 maps, h == 8, ld == N, N % 8 == 0");
  QuadCtx q = make_ctx(drop_p, seed, stream_id, N);
  if ((map_bf16 || (map_fmt & VU_MAP_TF32_MIX)) && mma_path(h, N, ld)) {
    VU_REQUIRE(B <= 65535, fn, "at most 65535 images per call on the tensor-core map path");
    const int64_t tiles = cdiv((int64_t)N * N / 4, 8);          // per image; grid = (x, B)
    const dim3 grid((unsigned)std::max<int64_t>(1, cdiv(tiles, 8 * 16)), B);      // 8 warps x ~16 iterations per CTA
    cudaStream_t st = as_stream(stream);
#define VU_MR(PTV, MTV, MIXV) mma::reattn_mix_reduce_mma_kernel<PTV, MTV, MIXV><<<grid, 256, 0, st>>>( \
        (const PTV*)Pv, (const MTV*)dA, (MTV*)A, fold, N, q, red)
    if (p_bf16) { if (A) VU_MR(__nv_bfloat16, __nv_bfloat16, true); else VU_MR(__nv_bfloat16, __nv_bfloat16, false); }
    else if (map_bf16) { if (A) VU_MR(float, __nv_bfloat16, true); else VU_MR(float, __nv_bfloat16, false); }
    else { if (A) VU_MR(float, float, true); else VU_MR(float, float, false); }
#undef VU_MR
    return check_launch(fn);
  }
  int blocks = grid_for((int64_t)B * N * (ld / 4), 256 * 2, 4);
  if (map_bf16) { VU_DISPATCH_H(h, fn, reattn_mix_reduce_kernel<HH, __nv_bfloat16><<<blocks, 256, 0, as_stream(stream)>>>(P, (const __nv_bfloat16*)dA, (__nv_bfloat16*)A, fold, B, N, ld, q, red)); }
  else { VU_DISPATCH_H(h, fn, reattn_mix_reduce_kernel<HH, float><<<blocks, 256, 0, as_stream(stream)>>>(P, (const float*)dA, (float*)A, fold, B, N, ld, q, red)); }
  return check_launch(fn);
}
