// FP32 CUDA-core batched GEMM with fused epilogues -- the exact (1e-5 parity) path for every contraction
// of the ViT-UNet block:  QK^T / PV (model.py:155,161), proj (:162), FeedForward (:103,106) and their
// data / weight gradients.  The tensor-core (tcgen05) path lives in vu_gemm_tc.cu and shares vu_gemm_desc.
//
// Tiling: BMxBN output tile per CTA, BK-deep smem stages (double buffered through registers), each thread owns
// a TMxTN micro-tile split into 4-wide column/row groups so that every shared-memory read is a conflict-free
// LDS.128.  Both operands may be transposed in memory; batch index z = (zo, zi) with independent strides so that
// head-sliced views of (B,N,h,hd) tensors need no copies.
#include <atomic>
#include <cstdio>
#include <cstdlib>

#include <cuda_bf16.h>

#include "vu_common.cuh"

namespace vu {

struct GemmArgs {
  const float* A; const float* B; float* C;
  const float* bias; const float* residual; const float* aux_in; float* aux_out;
  int M, N, K;
  int64_t lda, ldb, ldc, ldr, ldaux;
  int batch_inner;
  int64_t sAo, sAi, sBo, sBi, sCo, sCi;
  float alpha; int act; int accumulate; int split_k; int k_per_split;
  float drop_scale; uint32_t drop_thresh; uint64_t drop_seed; uint32_t drop_stream;
};

// Load a (ROWS x BK) operand tile into registers.  `contig_k` : the k index is the contiguous one in memory.
//   element(r, k) = contig_k ? P[r*ld + k] : P[k*ld + r]
// Thread mapping keeps global reads as wide and coalesced as alignment allows (float4 when legal).
template <int ROWS, int BK, int NT, bool CONTIG_K>
struct TileLoader {
  static constexpr int VEC = 4;
  static constexpr int NVEC = ROWS * BK / VEC;          // float4s per tile
  static constexpr int PER_T = (NVEC + NT - 1) / NT;    // float4s per thread
  float4 reg[PER_T];

  __device__ __forceinline__ void load(const float* __restrict__ P, int64_t ld, int r0, int k0, int rmax, int kmax,
                                       bool aligned, int tid) {
#pragma unroll
    for (int it = 0; it < PER_T; ++it) {
      int v = tid + it * NT;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (NVEC % NT == 0 || v < NVEC) {
        if (CONTIG_K) {
          int r = v / (BK / VEC), kq = (v % (BK / VEC)) * VEC;
          int gr = r0 + r, gk = k0 + kq;
          if (gr < rmax) {
            const float* src = P + (int64_t)gr * ld + gk;
            if (aligned && gk + 3 < kmax) val = *reinterpret_cast<const float4*>(src);
            else {
              if (gk < kmax) val.x = src[0];
              if (gk + 1 < kmax) val.y = src[1];
              if (gk + 2 < kmax) val.z = src[2];
              if (gk + 3 < kmax) val.w = src[3];
            }
          }
        } else {
          int k = v / (ROWS / VEC), rq = (v % (ROWS / VEC)) * VEC;
          int gr = r0 + rq, gk = k0 + k;
          if (gk < kmax) {
            const float* src = P + (int64_t)gk * ld + gr;
            if (aligned && gr + 3 < rmax) val = *reinterpret_cast<const float4*>(src);
            else {
              if (gr < rmax) val.x = src[0];
              if (gr + 1 < rmax) val.y = src[1];
              if (gr + 2 < rmax) val.z = src[2];
              if (gr + 3 < rmax) val.w = src[3];
            }
          }
        }
      }
      reg[it] = val;
    }
  }
  // smem tile is [BK][ROWS + PAD] (k-major) so the compute loop reads rows of one k as float4s.
  template <int LDS>
  __device__ __forceinline__ void store(float* __restrict__ S, int tid) const {
#pragma unroll
    for (int it = 0; it < PER_T; ++it) {
      int v = tid + it * NT;
      if (NVEC % NT == 0 || v < NVEC) {
        if (CONTIG_K) {
          int r = v / (BK / VEC), kq = (v % (BK / VEC)) * VEC;
          S[(kq + 0) * LDS + r] = reg[it].x; S[(kq + 1) * LDS + r] = reg[it].y;
          S[(kq + 2) * LDS + r] = reg[it].z; S[(kq + 3) * LDS + r] = reg[it].w;
        } else {
          int k = v / (ROWS / VEC), rq = (v % (ROWS / VEC)) * VEC;
          *reinterpret_cast<float4*>(S + k * LDS + rq) = reg[it];
        }
      }
    }
  }
};

template <int BM, int BN, int BK, int TM, int TN, bool TA, bool TB>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_f32_kernel(GemmArgs g) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
  constexpr int GM = TM / 4, GN = TN / 4;           // 4-wide groups per thread
  __shared__ __align__(16) float As[2][BK * LDA_S];
  __shared__ __align__(16) float Bs[2][BK * LDB_S];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int z = blockIdx.z, ks = 0;
  if (g.split_k > 1) { ks = z % g.split_k; z /= g.split_k; }
  const int zo = z / g.batch_inner, zi = z % g.batch_inner;
  const float* A = g.A + zo * g.sAo + zi * g.sAi;
  const float* B = g.B + zo * g.sBo + zi * g.sBi;
  const int64_t coff = zo * g.sCo + zi * g.sCi;
  const int kbeg = ks * g.k_per_split;
  const int kend = min(g.K, kbeg + g.k_per_split);

  // A(m,k): !TA -> k contiguous.  B(k,n): TB -> k contiguous (nn.Linear weight), !TB -> n contiguous.
  TileLoader<BM, BK, NT, !TA> la;
  TileLoader<BN, BK, NT, TB> lb;
  const bool a_al = ((uintptr_t)A % 16 == 0) && (g.lda % 4 == 0) && (!TA ? (kbeg % 4 == 0) : true);
  const bool b_al = ((uintptr_t)B % 16 == 0) && (g.ldb % 4 == 0) && (TB ? (kbeg % 4 == 0) : true);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (kend - kbeg + BK - 1) / BK;
  if (nk > 0) {
    la.load(A, g.lda, m0, kbeg, g.M, kend, a_al, tid);
    lb.load(B, g.ldb, n0, kbeg, g.N, kend, b_al, tid);
    la.template store<LDA_S>(As[0], tid);
    lb.template store<LDB_S>(Bs[0], tid);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      la.load(A, g.lda, m0, kbeg + (kt + 1) * BK, g.M, kend, a_al, tid);
      lb.load(B, g.ldb, n0, kbeg + (kt + 1) * BK, g.N, kend, b_al, tid);
    }
    const float* as = As[cur];
    const float* bs = Bs[cur];
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int gi = 0; gi < GM; ++gi) {
        float4 t = *reinterpret_cast<const float4*>(as + k * LDA_S + gi * (BM / GM) + ty * 4);
        a[gi * 4 + 0] = t.x; a[gi * 4 + 1] = t.y; a[gi * 4 + 2] = t.z; a[gi * 4 + 3] = t.w;
      }
#pragma unroll
      for (int gj = 0; gj < GN; ++gj) {
        float4 t = *reinterpret_cast<const float4*>(bs + k * LDB_S + gj * (BN / GN) + tx * 4);
        b[gj * 4 + 0] = t.x; b[gj * 4 + 1] = t.y; b[gj * 4 + 2] = t.z; b[gj * 4 + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      la.template store<LDA_S>(As[cur ^ 1], tid);
      lb.template store<LDB_S>(Bs[cur ^ 1], tid);
    }
    __syncthreads();
  }

  // ---------------------------------------------------------------- epilogue
  float* C = g.C + coff;
  const bool c_vec = ((uintptr_t)C % 16 == 0) && (g.ldc % 4 == 0) && g.split_k <= 1;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (i / 4) * (BM / GM) + ty * 4 + (i % 4);
    if (m >= g.M) continue;
#pragma unroll
    for (int gj = 0; gj < GN; ++gj) {
      const int n = n0 + gj * (BN / GN) + tx * 4;
      if (n >= g.N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[i][gj * 4 + j] * g.alpha;
      const int nv = min(4, g.N - n);
      if (g.split_k > 1) {
        // partial sums: bias/residual contributed once by split 0; act/dropout are rejected on the host
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < nv) {
            float t = v[j];
            if (ks == 0) {
              if (g.bias) t += g.bias[n + j];
              if (g.residual) t += g.residual[coff + (int64_t)m * g.ldr + n + j];
            }
            atomicAdd(C + (int64_t)m * g.ldc + n + j, t);
          }
        }
        continue;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < nv) {
          float t = v[j];
          if (g.bias) t += g.bias[n + j];
          if (g.act == VU_ACT_GELU) {
            if (g.aux_out) g.aux_out[coff + (int64_t)m * g.ldaux + n + j] = t;
            t = gelu_exact(t);
          } else if (g.act == VU_ACT_GELU_BWD) {
            t *= gelu_exact_grad(g.aux_in[coff + (int64_t)m * g.ldaux + n + j]);
          }
          if (g.drop_thresh) {
            uint64_t idx = (uint64_t)z * g.M * g.N + (uint64_t)m * g.N + (n + j);
            t = Philox::keep(g.drop_seed, g.drop_stream, idx, g.drop_thresh) ? t * g.drop_scale : 0.f;
          }
          if (g.residual) t += g.residual[coff + (int64_t)m * g.ldr + n + j];
          v[j] = t;
        }
      }
      float* dst = C + (int64_t)m * g.ldc + n;
      if (c_vec && nv == 4) {
        float4 o = make_float4(v[0], v[1], v[2], v[3]);
        if (g.accumulate) { float4 c = *reinterpret_cast<float4*>(dst); o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
        *reinterpret_cast<float4*>(dst) = o;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nv) dst[j] = g.accumulate ? dst[j] + v[j] : v[j];
      }
    }
  }
}

template <int BM, int BN, int BK, int TM, int TN>
static void launch_cfg(const GemmArgs& g, int trans_a, int trans_b, int nbatch, cudaStream_t s) {
  dim3 grid((unsigned)cdiv(g.N, BN), (unsigned)cdiv(g.M, BM), (unsigned)(nbatch * (g.split_k > 1 ? g.split_k : 1)));
  dim3 block((BM / TM) * (BN / TN));
  if (!trans_a && trans_b) gemm_f32_kernel<BM, BN, BK, TM, TN, false, true><<<grid, block, 0, s>>>(g);
  else if (!trans_a && !trans_b) gemm_f32_kernel<BM, BN, BK, TM, TN, false, false><<<grid, block, 0, s>>>(g);
  else if (trans_a && !trans_b) gemm_f32_kernel<BM, BN, BK, TM, TN, true, false><<<grid, block, 0, s>>>(g);
  else gemm_f32_kernel<BM, BN, BK, TM, TN, true, true><<<grid, block, 0, s>>>(g);
}

int gemm_simt(const vu_gemm_desc& d, cudaStream_t s) {
  GemmArgs g;
  g.A = d.A; g.B = d.B; g.C = d.C; g.bias = d.bias; g.residual = d.residual; g.aux_in = d.aux_in; g.aux_out = d.aux_out;
  g.M = d.M; g.N = d.N; g.K = d.K; g.lda = d.lda; g.ldb = d.ldb; g.ldc = d.ldc; g.ldr = d.ldr; g.ldaux = d.ldaux;
  g.batch_inner = d.batch_inner > 0 ? d.batch_inner : 1;
  g.sAo = d.sAo; g.sAi = d.sAi; g.sBo = d.sBo; g.sBi = d.sBi; g.sCo = d.sCo; g.sCi = d.sCi;
  g.alpha = d.alpha; g.act = d.act; g.accumulate = d.accumulate;
  g.split_k = d.split_k > 1 ? d.split_k : 1;
  g.drop_thresh = d.drop_p > 0.f ? drop_threshold(d.drop_p) : 0u;
  g.drop_scale = drop_keep_scale(d.drop_p);
  g.drop_seed = d.drop_seed; g.drop_stream = d.drop_stream;
  const int nbatch = (d.batch_outer > 0 ? d.batch_outer : 1) * g.batch_inner;
  // split boundaries on multiples of 16 so vector loads stay aligned
  int kps = (int)cdiv(cdiv(g.K, g.split_k), 16) * 16;
  g.k_per_split = kps;
  g.split_k = (int)cdiv(g.K, kps);

  const int64_t tiles128 = cdiv(g.M, 128) * cdiv(g.N, 128) * nbatch * g.split_k;
  if (g.N <= 32) launch_cfg<128, 32, 16, 4, 4>(g, d.trans_a, d.trans_b, nbatch, s);
  else if (g.M <= 64 || g.N <= 64 || tiles128 < sm_count()) launch_cfg<64, 64, 16, 4, 4>(g, d.trans_a, d.trans_b, nbatch, s);
  else launch_cfg<128, 128, 8, 8, 8>(g, d.trans_a, d.trans_b, nbatch, s);
  return check_launch("vu_gemm");
}

static std::atomic<long> g_tf32_fallbacks{0};
int gemm_tc(const vu_gemm_desc& d, cudaStream_t s, bool* handled);   // vu_gemm_tc.cu
int gemm_scores(const vu_gemm_desc& d, cudaStream_t s, bool* handled);   // vu_gemm_scores.cu

// column sums: out[n] (+)= sum_m X[m*ld + n].  grid (N/32, chunks of M); smem transpose-free: each warp
// owns 32 columns, threads stride over rows, partials combined with atomics.
// column sums: out[n] (+)= sum_m X[m*ld + n].  A CTA covers 128 columns x a chunk of rows: every lane owns 4 consecutive
// columns (one 16-byte fp32 / 8-byte bf16 load per row), the 8 warps take interleaved rows with 4 loads in flight each;
// partials are combined through shared memory and one atomicAdd per column and CTA.
template <typename TX>
__device__ __forceinline__ float4 colsum_load4(const TX* p);
template <>
__device__ __forceinline__ float4 colsum_load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 colsum_load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 pk = *reinterpret_cast<const uint2*>(p);
  const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
  const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

template <typename TX, bool VEC>
__global__ void __launch_bounds__(256)
colsum_kernel(const TX* __restrict__ X, int64_t M, int N, int64_t ld, float* __restrict__ out, int64_t rows_per_block) {
  __shared__ float red[8][132];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 128 + lane * 4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (VEC) {
    if (n < N) {             // N % 4 == 0: whole quads
      int64_t r = r0 + warp;
      for (; r + 24 < r1; r += 32) {
        const float4 v0 = colsum_load4<TX>(X + r * ld + n), v1 = colsum_load4<TX>(X + (r + 8) * ld + n),
                     v2 = colsum_load4<TX>(X + (r + 16) * ld + n), v3 = colsum_load4<TX>(X + (r + 24) * ld + n);
        a0 += (v0.x + v1.x) + (v2.x + v3.x); a1 += (v0.y + v1.y) + (v2.y + v3.y);
        a2 += (v0.z + v1.z) + (v2.z + v3.z); a3 += (v0.w + v1.w) + (v2.w + v3.w);
      }
      for (; r < r1; r += 8) {
        const float4 v = colsum_load4<TX>(X + r * ld + n);
        a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
      }
    }
  } else {
    for (int64_t r = r0 + warp; r < r1; r += 8) {
      if (n < N) a0 += (float)X[r * ld + n];
      if (n + 1 < N) a1 += (float)X[r * ld + n + 1];
      if (n + 2 < N) a2 += (float)X[r * ld + n + 2];
      if (n + 3 < N) a3 += (float)X[r * ld + n + 3];
    }
  }
  *reinterpret_cast<float4*>(&red[warp][lane * 4]) = make_float4(a0, a1, a2, a3);
  __syncthreads();
  if (threadIdx.x < 128) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c < N) atomicAdd(out + c, s);
  }
}

}  // namespace vu

extern "C" int vu_gemm(const vu_gemm_desc* d, void* stream) {
  using namespace vu;
  const char* fn = "vu_gemm";
  VU_REQUIRE(d != nullptr, fn, "null descriptor");
  VU_REQUIRE(d->A && d->B && d->C, fn, "null operand pointer");
  VU_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, fn, "M, N, K must be positive");
  VU_REQUIRE(d->act >= VU_ACT_NONE && d->act <= VU_ACT_GELU_BWD, fn, "unknown activation");
  VU_REQUIRE(d->act != VU_ACT_GELU_BWD || d->aux_in, fn, "GELU_BWD needs aux_in");
  VU_REQUIRE(d->drop_p >= 0.f && d->drop_p < 1.f, fn, "drop_p must be in [0,1)");
  if (d->split_k > 1)
    VU_REQUIRE(d->act == VU_ACT_NONE && d->drop_p == 0.f, fn, "split_k cannot be combined with activation/dropout");
  VU_REQUIRE(d->drop_p == 0.f || d->ldc == d->N, fn, "dropout epilogue needs a dense C (ldc == N)");
  cudaStream_t s = as_stream(stream);
  const bool any_bf16 = d->a_bf16 || d->b_bf16 || d->c_bf16;
  VU_REQUIRE(!any_bf16 || d->precision == VU_PREC_TF32, fn, "bfloat16 operands need the tensor-core path (precision = VU_PREC_TF32)");
  if (d->precision == VU_PREC_TF32) {
    bool handled = false;
    int rc = gemm_scores(*d, s, &handled);        // K <= 32, map-sized output: the HBM write stream kernel
    if (rc != VU_OK || handled) return rc;
    rc = gemm_tc(*d, s, &handled);
    if (rc != VU_OK || handled) return rc;
    VU_REQUIRE(!any_bf16, fn, "bfloat16 GEMM could not be mapped onto the tensor-core kernel");
    // shapes the tensor-core kernel does not cover (an operand TMA cannot address: base or stride not 16-byte aligned)
    // fall through to the CUDA-core kernel (same numerics class or better) -- a ~7x slower kernel choice inside the CUDA
    // path, not a CPU fallback.  It is COUNTED (vu_gemm_tf32_fallbacks) and the first occurrence of each shape class is
    // logged to stderr, so a performance cliff cannot hide.
    const long n = ++g_tf32_fallbacks;
    if (n <= 4 || getenv("VU_LOG_FALLBACKS"))
      fprintf(stderr, "[vit_unet_b200] vu_gemm: TF32 request M=%d N=%d K=%d (lda=%lld ldb=%lld, trans %d/%d) is not TMA-addressable; "
              "running the CUDA-core kernel (%ld so far)\n", d->M, d->N, d->K, (long long)d->lda, (long long)d->ldb,
              d->trans_a, d->trans_b, n);
  } else {
    VU_REQUIRE(d->precision == VU_PREC_FP32, fn, "unknown precision");
  }
  return gemm_simt(*d, s);
}

extern "C" int vu_gemm_tf32_fallbacks(void) { return (int)vu::g_tf32_fallbacks.load(); }

extern "C" int vu_colsum(const void* X, int x_bf16, int64_t M, int N, int64_t ld, float* out, int accumulate, void* stream) {
  using namespace vu;
  const char* fn = "vu_colsum";
  VU_REQUIRE(X && out && M > 0 && N > 0 && ld >= N, fn, "bad arguments");
  cudaStream_t s = as_stream(stream);
  if (!accumulate) {
    if (cudaMemsetAsync(out, 0, sizeof(float) * N, s) != cudaSuccess) return check_launch(fn);
  }
  int64_t chunks = std::min<int64_t>(cdiv(M, 64), std::max<int64_t>(1, (int64_t)sm_count() * 8 / cdiv(N, 128)));
  int64_t rpb = cdiv(M, chunks);
  dim3 grid((unsigned)cdiv(N, 128), (unsigned)cdiv(M, rpb));
  const int esz = x_bf16 ? 2 : 4;
  const bool vec = N % 4 == 0 && ld % 4 == 0 && (uintptr_t)X % (4 * esz) == 0;
  if (x_bf16) {
    const __nv_bfloat16* Xb = reinterpret_cast<const __nv_bfloat16*>(X);
    if (vec) colsum_kernel<__nv_bfloat16, true><<<grid, 256, 0, s>>>(Xb, M, N, ld, out, rpb);
    else colsum_kernel<__nv_bfloat16, false><<<grid, 256, 0, s>>>(Xb, M, N, ld, out, rpb);
  } else {
    const float* Xf = reinterpret_cast<const float*>(X);
    if (vec) colsum_kernel<float, true><<<grid, 256, 0, s>>>(Xf, M, N, ld, out, rpb);
    else colsum_kernel<float, false><<<grid, 256, 0, s>>>(Xf, M, N, ld, out, rpb);
  }
  return check_launch(fn);
}
