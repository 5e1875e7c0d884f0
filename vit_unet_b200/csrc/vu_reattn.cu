// Re-Attention map kernels (DeepViT-style head mixing + BatchNorm over the attention maps):
//   attn = softmax(q k^T * scale); attn = dropout(attn); attn = BN_h(Conv1x1_{h->h}(attn))   model.py:155-159
// Maps are (B, h, N, ld) fp32.  The 1x1 conv + BatchNorm are folded into ONE h x h affine per position
// (SURVEY.md F5); train-mode batch statistics come from a streamed reduction over all (b,i,j).
// Attention dropout masks are regenerated from Philox in every kernel that needs them (never stored).
#include "vu_common.cuh"

namespace vu {

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
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    // scale may be negative in principle; softmax(scale*s): shift by max of scale*s
    float mxs = mx * sl2;
    if (scale < 0.f) {
      float mn = INFINITY;
      for (int j = lane; j < N; j += 32) mn = fminf(mn, row[j]);
      mn = -warp_max(-mn);
      mxs = mn * sl2;
    }
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
      float e = exp2f(fmaf(row[j], sl2, -mxs));
      row[j] = e; sum += e;
    }
    sum = warp_sum(sum);
    float inv = 1.0f / sum;
    for (int j = lane; j < N; j += 32) row[j] *= inv;
    for (int j = N + lane; j < ld; j += 32) row[j] = 0.f;
  }
}

// ------------------------------------------------------------------ per-position head vector helpers
template <int H>
struct HeadMix {
  // load P_g(b,i,j) for all heads with the dropout mask applied
  __device__ __forceinline__ static void load(const float* __restrict__ P, int64_t head_stride, int64_t off,
                                              uint32_t thresh, float dscale, uint64_t seed, uint32_t stream,
                                              int64_t flat_base, float (&p)[H]) {
#pragma unroll
    for (int g = 0; g < H; ++g) {
      float v = __ldg(P + g * head_stride + off);
      if (thresh) v = Philox::keep(seed, stream, (uint64_t)(flat_base + g * head_stride + off), thresh) ? v * dscale : 0.f;
      p[g] = v;
    }
  }
};

// sums[h] += sum (M_h - c_h), sums[H+h] += sum (M_h - c_h)^2,  M_h = sum_g W[h][g] Pd_g + b_h
template <int H>
__global__ void __launch_bounds__(256)
reattn_stats_kernel(const float* __restrict__ P, int B, int N, int ld, const float* __restrict__ W,
                    const float* __restrict__ bconv, uint32_t thresh, float dscale, uint64_t seed, uint32_t stream,
                    double* __restrict__ sums) {
  __shared__ float sW[H * H];
  __shared__ float sc[H];
  __shared__ double red[2 * H * 32];
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) sW[i] = W[i];
  __syncthreads();
  if (threadIdx.x < H) {
    float rs = 0.f;
    for (int g = 0; g < H; ++g) rs += sW[threadIdx.x * H + g];
    sc[threadIdx.x] = rs / (float)N;       // M_h - b_h - rowsum/N : shift keeps the sums well conditioned
  }
  __syncthreads();
  const int64_t head_stride = (int64_t)N * ld, img_stride = head_stride * H;
  const int64_t per_img = (int64_t)N * N, total = per_img * B;
  float s1[H], s2[H];
#pragma unroll
  for (int h = 0; h < H; ++h) { s1[h] = 0.f; s2[h] = 0.f; }
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = t / per_img; int64_t r = t - b * per_img;
    int i = (int)(r / N), j = (int)(r - (int64_t)i * N);
    int64_t off = (int64_t)i * ld + j;
    float p[H];
    HeadMix<H>::load(P + b * img_stride, head_stride, off, thresh, dscale, seed, stream, b * img_stride, p);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float m = -sc[h];
#pragma unroll
      for (int g = 0; g < H; ++g) m = fmaf(sW[h * H + g], p[g], m);
      s1[h] += m; s2[h] = fmaf(m, m, s2[h]);
    }
  }
  double v[2 * H];
#pragma unroll
  for (int h = 0; h < H; ++h) { v[h] = s1[h]; v[H + h] = s2[h]; }
  block_sum<2 * H>(v, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 2 * H; ++i) atomicAdd(sums + i, v[i]);
  }
  (void)bconv;
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
      float rs = 0.f;
      for (int g = 0; g < H; ++g) rs += W[h * H + g];
      double c = (double)bconv[h] + (double)(rs / (float)N);
      double m1 = sums[h] / count, m2 = sums[H + h] / count;
      double dmean = c + m1, dvar = m2 - m1 * m1;
      if (dvar < 0) dvar = 0;
      mean = (float)dmean; var = (float)dvar;
      double unbiased = count > 1 ? dvar * (count / (count - 1.0)) : dvar;
      rmean[h] = (1.f - momentum) * rmean[h] + momentum * mean;
      rvar[h] = (1.f - momentum) * rvar[h] + momentum * (float)unbiased;
    } else {
      mean = rmean[h]; var = rvar[h];
    }
    float invstd = rsqrtf(var + eps);
    // torch computes 1/sqrt in fp32 as well; use the correctly rounded form for parity
    invstd = 1.0f / sqrtf(var + eps);
    float a = gamma[h] * invstd;
    for (int g = 0; g < H; ++g) fold[h * H + g] = a * W[h * H + g];
    fold[H * H + h] = a * (bconv[h] - mean) + beta[h];
    saved[h] = mean; saved[H + h] = invstd;
  }
  if (train && threadIdx.x == 0 && nbt) *nbt += 1;
}

// A_h = sum_g fold[h][g] * Pd_g + fold[H*H + h]; 4 positions per thread along j (float4)
template <int H>
__global__ void __launch_bounds__(256)
reattn_mix_kernel(const float* __restrict__ P, float* __restrict__ A, const float* __restrict__ fold,
                  int B, int N, int ld, uint32_t thresh, float dscale, uint64_t seed, uint32_t stream) {
  __shared__ float sF[H * H + H];
  for (int i = threadIdx.x; i < H * H + H; i += blockDim.x) sF[i] = fold[i];
  __syncthreads();
  const int ld4 = ld >> 2;
  const int64_t head_stride = (int64_t)N * ld, img_stride = head_stride * H;
  const int64_t per_img4 = (int64_t)N * ld4, total = per_img4 * B;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = t / per_img4; int64_t r = t - b * per_img4;      // r = i*ld4 + j4
    int j = (int)(r % ld4) * 4;
    int64_t off = r * 4;
    float4 p[H];
#pragma unroll
    for (int g = 0; g < H; ++g) {
      float4 v = *reinterpret_cast<const float4*>(P + b * img_stride + g * head_stride + off);
      if (thresh) {
        // element index is a multiple of 4 -> one Philox call covers the quad
        uint64_t idx = (uint64_t)(b * img_stride + g * head_stride + off);
        uint4 rr = Philox::gen(seed, stream, idx >> 2);
        v.x = rr.x >= thresh ? v.x * dscale : 0.f; v.y = rr.y >= thresh ? v.y * dscale : 0.f;
        v.z = rr.z >= thresh ? v.z * dscale : 0.f; v.w = rr.w >= thresh ? v.w * dscale : 0.f;
      }
      p[g] = v;
    }
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
      *reinterpret_cast<float4*>(A + b * img_stride + h * head_stride + off) = a;
    }
  }
}

// red[h] += sum dA_h ; red[H+h] += sum dA_h * Ahat_h ;  Ahat_h = (M_h - mean_h) * invstd_h
template <int H>
__global__ void __launch_bounds__(256)
reattn_bwd_reduce_kernel(const float* __restrict__ P, const float* __restrict__ dA, int B, int N, int ld,
                         const float* __restrict__ W, const float* __restrict__ bconv, const float* __restrict__ saved,
                         uint32_t thresh, float dscale, uint64_t seed, uint32_t stream, double* __restrict__ out) {
  __shared__ float sW[H * H];
  __shared__ float sOff[H], sInv[H];
  __shared__ double red[2 * H * 32];
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) sW[i] = W[i];
  if (threadIdx.x < H) { sOff[threadIdx.x] = bconv[threadIdx.x] - saved[threadIdx.x]; sInv[threadIdx.x] = saved[H + threadIdx.x]; }
  __syncthreads();
  const int64_t head_stride = (int64_t)N * ld, img_stride = head_stride * H;
  const int64_t per_img = (int64_t)N * N, total = per_img * B;
  float s1[H], s2[H];
#pragma unroll
  for (int h = 0; h < H; ++h) { s1[h] = 0.f; s2[h] = 0.f; }
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = t / per_img; int64_t r = t - b * per_img;
    int i = (int)(r / N), j = (int)(r - (int64_t)i * N);
    int64_t off = (int64_t)i * ld + j;
    float p[H];
    HeadMix<H>::load(P + b * img_stride, head_stride, off, thresh, dscale, seed, stream, b * img_stride, p);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float m = sOff[h];
#pragma unroll
      for (int g = 0; g < H; ++g) m = fmaf(sW[h * H + g], p[g], m);
      float ah = m * sInv[h];
      float d = __ldg(dA + b * img_stride + h * head_stride + off);
      s1[h] += d; s2[h] = fmaf(d, ah, s2[h]);
    }
  }
  double v[2 * H];
#pragma unroll
  for (int h = 0; h < H; ++h) { v[h] = s1[h]; v[H + h] = s2[h]; }
  block_sum<2 * H>(v, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 2 * H; ++i) atomicAdd(out + i, v[i]);
  }
}

// One warp walks whole rows (b, i) for all heads:
//   dM_h  = k_h (dA_h - m1_h - Ahat_h m2_h)          (train)   |   k_h dA_h   (eval),  k_h = gamma_h invstd_h
//   dPd_g = sum_h W[h][g] dM_h ;  dP_g = keep_g dPd_g / (1-p)
//   r_g   = sum_j dP_g P_g ;       dS_g = scale * P_g (dP_g - r_g)           (written over dA)
//   dW[h][g] += dM_h Pd_g ; dbconv[h] += dM_h ; dgamma[h] = sum dA_h Ahat_h ; dbeta[h] = sum dA_h
template <int H>
__global__ void __launch_bounds__(128)
reattn_bwd_rows_kernel(const float* __restrict__ P, float* __restrict__ dA, int B, int N, int ld,
                       const float* __restrict__ W, const float* __restrict__ bconv, const float* __restrict__ gamma,
                       const float* __restrict__ saved, const double* __restrict__ red, double count, int train,
                       float scale, uint32_t thresh, float dscale, uint64_t seed, uint32_t stream,
                       float* __restrict__ dW, float* __restrict__ dbconv, float* __restrict__ dgamma,
                       float* __restrict__ dbeta) {
  __shared__ float sW[H * H];
  __shared__ float sOff[H], sInv[H], sK[H], sM1[H], sM2[H];
  __shared__ float sAcc[H * H + H];
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) sW[i] = W[i];
  for (int i = threadIdx.x; i < H * H + H; i += blockDim.x) sAcc[i] = 0.f;
  if (threadIdx.x < H) {
    int h = threadIdx.x;
    sOff[h] = bconv[h] - saved[h]; sInv[h] = saved[H + h]; sK[h] = gamma[h] * saved[H + h];
    sM1[h] = (train && red) ? (float)(red[h] / count) : 0.f;
    sM2[h] = (train && red) ? (float)(red[H + h] / count) : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t head_stride = (int64_t)N * ld, img_stride = head_stride * H;
  const int64_t rows = (int64_t)B * N;
  float aW[H][H], aB[H];
#pragma unroll
  for (int h = 0; h < H; ++h) { aB[h] = 0.f;
#pragma unroll
    for (int g = 0; g < H; ++g) aW[h][g] = 0.f; }

  for (int64_t r = wid; r < rows; r += nw) {
    int64_t b = r / N; int i = (int)(r - b * N);
    const float* Pb = P + b * img_stride + (int64_t)i * ld;
    float* Db = dA + b * img_stride + (int64_t)i * ld;
    float rg[H];
#pragma unroll
    for (int g = 0; g < H; ++g) rg[g] = 0.f;
    for (int j = lane; j < N; j += 32) {
      float p[H], pd[H], dm[H];
      bool keep[H];
#pragma unroll
      for (int g = 0; g < H; ++g) {
        p[g] = Pb[g * head_stride + j];
        keep[g] = thresh ? Philox::keep(seed, stream, (uint64_t)(b * img_stride + g * head_stride + (int64_t)i * ld + j), thresh) : true;
        pd[g] = keep[g] ? p[g] * dscale : 0.f;
      }
#pragma unroll
      for (int h = 0; h < H; ++h) {
        float d = Db[h * head_stride + j];
        float t = d;
        if (train) {
          float m = sOff[h];
#pragma unroll
          for (int g = 0; g < H; ++g) m = fmaf(sW[h * H + g], pd[g], m);
          t = d - sM1[h] - (m * sInv[h]) * sM2[h];
        }
        dm[h] = sK[h] * t;
        aB[h] += dm[h];
#pragma unroll
        for (int g = 0; g < H; ++g) aW[h][g] = fmaf(dm[h], pd[g], aW[h][g]);
      }
#pragma unroll
      for (int g = 0; g < H; ++g) {
        float dpd = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) dpd = fmaf(sW[h * H + g], dm[h], dpd);
        float dp = keep[g] ? dpd * dscale : 0.f;
        rg[g] = fmaf(dp, p[g], rg[g]);
        Db[g * head_stride + j] = dp;
      }
    }
#pragma unroll
    for (int g = 0; g < H; ++g) rg[g] = warp_sum(rg[g]);
    for (int j = lane; j < N; j += 32) {
#pragma unroll
      for (int g = 0; g < H; ++g) {
        float pv = Pb[g * head_stride + j];
        float dp = Db[g * head_stride + j];
        Db[g * head_stride + j] = scale * pv * (dp - rg[g]);
      }
    }
    for (int j = N + lane; j < ld; j += 32) {
#pragma unroll
      for (int g = 0; g < H; ++g) Db[g * head_stride + j] = 0.f;
    }
  }
  // parameter gradients: warp -> block (smem atomics) -> global atomics
#pragma unroll
  for (int h = 0; h < H; ++h) {
#pragma unroll
    for (int g = 0; g < H; ++g) {
      float v = warp_sum(aW[h][g]);
      if (lane == 0) atomicAdd(&sAcc[h * H + g], v);
    }
    float v = warp_sum(aB[h]);
    if (lane == 0) atomicAdd(&sAcc[H * H + h], v);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * H + H; i += blockDim.x) {
    if (i < H * H) atomicAdd(dW + i, sAcc[i]);
    else if (dbconv) atomicAdd(dbconv + (i - H * H), sAcc[i]);
  }
  if (blockIdx.x == 0 && threadIdx.x < H) {
    int h = threadIdx.x;
    if (red) {     // BN affine gradients come straight from the two reductions
      atomicAdd(dgamma + h, (float)red[H + h]);
      atomicAdd(dbeta + h, (float)red[h]);
    }
  }
}

static int grid_for(int64_t work_items, int threads, int per_sm) {
  int64_t b = cdiv(work_items, threads);
  int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)std::max<int64_t>(1, std::min(b, cap));
}

}  // namespace vu

#define VU_DISPATCH_H(h, fn, ...)                                   \
  switch (h) {                                                      \
    case 1: { constexpr int HH = 1; __VA_ARGS__; } break;           \
    case 2: { constexpr int HH = 2; __VA_ARGS__; } break;           \
    case 4: { constexpr int HH = 4; __VA_ARGS__; } break;           \
    case 8: { constexpr int HH = 8; __VA_ARGS__; } break;           \
    default: return vu::fail_arg(fn, "num_heads must be 1, 2, 4 or 8"); \
  }

extern "C" int vu_softmax_rows(float* S, int64_t rows, int N, int ld, float scale, void* stream) {
  using namespace vu;
  const char* fn = "vu_softmax_rows";
  VU_REQUIRE(S && rows > 0 && N > 0 && ld >= N, fn, "bad arguments");
  int blocks = grid_for(rows * 32, 256, 16);
  softmax_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(S, rows, N, ld, scale);
  return check_launch(fn);
}

extern "C" int vu_reattn_stats(const float* P, int B, int h, int N, int ld, const float* W, const float* bconv,
                               float drop_p, uint64_t seed, uint32_t stream_id, double* sums, void* stream) {
  using namespace vu;
  const char* fn = "vu_reattn_stats";
  VU_REQUIRE(P && W && bconv && sums && B > 0 && N > 0 && ld >= N && ld % 4 == 0, fn, "bad arguments");
  VU_REQUIRE(drop_p >= 0.f && drop_p < 1.f, fn, "drop_p must be in [0,1)");
  uint32_t th = drop_p > 0.f ? drop_threshold(drop_p) : 0u; float ds = 1.f / (1.f - drop_p);
  int blocks = grid_for((int64_t)B * N * N, 256 * 4, 8);
  VU_DISPATCH_H(h, fn, reattn_stats_kernel<HH><<<blocks, 256, 0, as_stream(stream)>>>(P, B, N, ld, W, bconv, th, ds, seed, stream_id, sums));
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

extern "C" int vu_reattn_mix(const float* P, float* A, const float* fold, int B, int h, int N, int ld,
                             float drop_p, uint64_t seed, uint32_t stream_id, void* stream) {
  using namespace vu;
  const char* fn = "vu_reattn_mix";
  VU_REQUIRE(P && A && fold && B > 0 && N > 0 && ld >= N && ld % 4 == 0, fn, "bad arguments");
  VU_REQUIRE(((uintptr_t)P % 16 == 0) && ((uintptr_t)A % 16 == 0), fn, "maps must be 16-byte aligned");
  VU_REQUIRE(drop_p >= 0.f && drop_p < 1.f, fn, "drop_p must be in [0,1)");
  uint32_t th = drop_p > 0.f ? drop_threshold(drop_p) : 0u; float ds = 1.f / (1.f - drop_p);
  int blocks = grid_for((int64_t)B * N * (ld / 4), 256, 16);
  VU_DISPATCH_H(h, fn, reattn_mix_kernel<HH><<<blocks, 256, 0, as_stream(stream)>>>(P, A, fold, B, N, ld, th, ds, seed, stream_id));
  return check_launch(fn);
}

extern "C" int vu_reattn_bwd_reduce(const float* P, const float* dA, int B, int h, int N, int ld, const float* W,
                                    const float* bconv, const float* saved, float drop_p, uint64_t seed,
                                    uint32_t stream_id, double* red, void* stream) {
  using namespace vu;
  const char* fn = "vu_reattn_bwd_reduce";
  VU_REQUIRE(P && dA && W && bconv && saved && red && B > 0 && N > 0 && ld >= N, fn, "bad arguments");
  uint32_t th = drop_p > 0.f ? drop_threshold(drop_p) : 0u; float ds = 1.f / (1.f - drop_p);
  int blocks = grid_for((int64_t)B * N * N, 256 * 4, 8);
  VU_DISPATCH_H(h, fn, reattn_bwd_reduce_kernel<HH><<<blocks, 256, 0, as_stream(stream)>>>(P, dA, B, N, ld, W, bconv, saved, th, ds, seed, stream_id, red));
  return check_launch(fn);
}

extern "C" int vu_reattn_bwd_rows(const float* P, float* dA_dS, int B, int h, int N, int ld, const float* W,
                                  const float* bconv, const float* gamma, const float* saved, const double* red,
                                  int train, float scale, float drop_p, uint64_t seed, uint32_t stream_id,
                                  float* dW, float* dbconv, float* dgamma, float* dbeta, void* stream) {
  using namespace vu;
  const char* fn = "vu_reattn_bwd_rows";
  VU_REQUIRE(P && dA_dS && W && bconv && gamma && saved && dW && dgamma && dbeta, fn, "null pointer");
  VU_REQUIRE(B > 0 && N > 0 && ld >= N, fn, "bad shape");
  VU_REQUIRE(!train || red, fn, "train mode needs the BN reductions");
  uint32_t th = drop_p > 0.f ? drop_threshold(drop_p) : 0u; float ds = 1.f / (1.f - drop_p);
  int blocks = grid_for((int64_t)B * N * 32, 128, 8);
  double count = (double)B * N * N;
  VU_DISPATCH_H(h, fn, reattn_bwd_rows_kernel<HH><<<blocks, 128, 0, as_stream(stream)>>>(
      P, dA_dS, B, N, ld, W, bconv, gamma, saved, red, count, train, scale, th, ds, seed, stream_id, dW, dbconv, dgamma, dbeta));
  return check_launch(fn);
}
