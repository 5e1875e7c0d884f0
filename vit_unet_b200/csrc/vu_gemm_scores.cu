// "Scores" GEMM of the attention path on the fine levels:  C[z] (M x N) = alpha * A[z] (M x K) . B[z]^T (N x K),
// K = head_dim <= 128 (B operand must fit shared memory), M = N = tokens (784 at the Base bottleneck), batched over (image, head):
//   S = Q K^T (model.py:155)  and  dA = dO V^T (its backward).
// The contraction is 24 long but the output is the N x N attention map -- 5 GB per launch at 256 images -- so the
// kernel is an HBM WRITE stream: the general tcgen05 tile kernel (one 128x64 output tile per CTA, TMEM round trip,
// staged epilogue) spends its time in per-CTA set-up and reaches 2.4 TB/s.  Here one CTA owns a whole (image, head):
// the B operand (N x K, 75 KB) is staged once in shared memory as TF32, every warp walks 16-row strips with
// m16n8k8 warp MMAs whose accumulator fragments go straight from registers to global memory (full 32-byte sectors,
// four consecutive n-tiles complete a 128-byte line), no TMEM, no epilogue hand-off: 11 instructions per 128 outputs.
// Negative result (round 2, not in the tree): two row strips per work item, so that every B fragment (one LDS.64 per MMA) feeds
// two MMAs -- 128 registers, and the Base step's scores launches went from 10.15 to 11.66 ms per step: the kernel is not bound by
// the B-fragment loads; the second strip's accumulators only lengthen the dependent store phase of each item.
// Second negative result: the staged 16-row tiles leaving shared memory as bulk (TMA) row copies (cp.async.bulk.global.shared::cta,
// 16 rows of 128 / 256 bytes per instruction) instead of LDS.128 + STG.128 -- correct, but 12.47 vs 10.10 ms per step: 128-byte
// bulk requests are bound by the copy engine's request rate, not by bytes.
// Third: 12 instead of 8 warps per CTA (24 instead of 16 resident warps per SM): 10.62 vs 10.20 ms -- not latency either; what is left
// is the L1 / shared-memory pipe itself (B-fragment LDS.64 + the staging round trip, ~72 wavefronts per 512 outputs).
#include <cuda_bf16.h>

#include <algorithm>

#include "vu_common.cuh"

namespace vu {

__device__ __forceinline__ uint32_t tf32_rn(float x) { return __float_as_uint(x) + 0x1000u; }   // round on the dropped bits
__device__ __forceinline__ void mma_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct ScoresArgs {
  const float* A; const float* B; void* C;
  int M, N, K;
  int64_t lda, ldb, ldc;
  int batch_inner;
  int64_t sAo, sAi, sBo, sBi, sCo, sCi;
  float alpha;
  int vec4;              // B rows can be staged with 16-byte loads
  int chunks;            // CTAs per batch entry (each takes a contiguous range of 16-row strips)
  int staged;            // output rows are 16-byte aligned: stage 16x32 (fp32) / 16x64 (bf16) tiles through shared memory
};

// shared-memory layout of the B operand: row n at n * PITCH; inside a row the k index is permuted so that the two
// values of one k-step a lane needs (k = tig, tig + 4) are adjacent (one LDS.64): pos(k) = 8*(k/8) + 2*(k%4) + (k/4)%2.
// PITCH % 32 in {8, 24} keeps the eight rows of a quarter-warp on disjoint banks.
template <int KS>
struct ScoresTile {
  static constexpr int KP = KS * 8;
  static constexpr int PITCH = (KP % 16 == 0) ? KP + 8 : KP;
};

__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b), r = __floats2bfloat162_rn(c, d);
  uint2 u; u.x = *reinterpret_cast<uint32_t*>(&p); u.y = *reinterpret_cast<uint32_t*>(&r);
  return u;
}

template <int KS, bool BF16>
__global__ void __launch_bounds__(256)
scores_mma_kernel(ScoresArgs g) {
  using T = ScoresTile<KS>;
  extern __shared__ __align__(16) uint32_t Bs[];               // N8 * PITCH tf32 words
  const int z = blockIdx.x / g.chunks, chunk = blockIdx.x - z * g.chunks;     // CTA = (batch entry, chunk of row strips)
  const int zo = z / g.batch_inner, zi = z - zo * g.batch_inner;
  const float* __restrict__ A = g.A + zo * g.sAo + zi * g.sAi;
  const float* __restrict__ B = g.B + zo * g.sBo + zi * g.sBi;
  const int N8 = (g.N + 7) & ~7;
  // ---- stage B: (n, k) -> Bs[n * PITCH + pos(k)], zero padded in n and k.  Four independent loads in flight per
  // thread; 16-byte loads when the rows allow it (K % 4 == 0, 16-byte aligned rows).
  if (g.vec4) {
    constexpr int Q = T::KP / 4;                               // float4 slots per (padded) row
    const int total = N8 * Q;
    for (int base = threadIdx.x; base < total; base += 4 * blockDim.x) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = base + u * blockDim.x, n = idx / Q, k = 4 * (idx - n * Q);
        v[u] = (idx < total && n < g.N && k < g.K) ? __ldg(reinterpret_cast<const float4*>(B + (int64_t)n * g.ldb + k))
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = base + u * blockDim.x, n = idx / Q, k = 4 * (idx - n * Q);
        if (idx < total) {       // k % 4 == 0: the quad k..k+3 lands on pos = 8*(k/8) + {0,2,4,6} + (k/4)%2
          uint32_t* dst = Bs + n * T::PITCH + 8 * (k >> 3) + ((k >> 2) & 1);
          dst[0] = tf32_rn(v[u].x); dst[2] = tf32_rn(v[u].y); dst[4] = tf32_rn(v[u].z); dst[6] = tf32_rn(v[u].w);
        }
      }
    }
  } else {
    const int total = N8 * T::KP;
    for (int base = threadIdx.x; base < total; base += 4 * blockDim.x) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = base + u * blockDim.x, n = idx / T::KP, k = idx - n * T::KP;
        v[u] = (idx < total && n < g.N && k < g.K) ? __ldg(B + (int64_t)n * g.ldb + k) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = base + u * blockDim.x, n = idx / T::KP, k = idx - n * T::KP;
        if (idx < total) Bs[n * T::PITCH + 8 * (k >> 3) + 2 * (k & 3) + ((k >> 2) & 1)] = tf32_rn(v[u]);
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
  const int strips = (g.M + 15) >> 4, ntiles = N8 >> 3, jfull = g.N >> 3;     // tiles < jfull have all 8 columns
  const int halves = ntiles >= 32 ? 2 : 1;                     // work item = (strip, column half)
  const int tph = (ntiles + halves - 1) / halves;
  const int spc = (strips + g.chunks - 1) / g.chunks;          // strips per chunk
  const int it0 = chunk * spc * halves, items = min(strips, (chunk + 1) * spc) * halves;
  const int odd = tig & 1;
  for (int it = it0 + w; it < items; it += 8) {
    const int strip = it / halves, half = it - strip * halves;
    const int r0 = strip * 16 + gid, r1 = r0 + 8;
    const bool full_rows = strip * 16 + 16 <= g.M;             // warp-uniform
    uint32_t a[KS][4];                                         // alpha folded into the A fragments
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const int k0 = 8 * s + tig, k1 = k0 + 4;
      a[s][0] = (r0 < g.M && k0 < g.K) ? tf32_rn(g.alpha * __ldg(A + (int64_t)r0 * g.lda + k0)) : 0u;
      a[s][1] = (r1 < g.M && k0 < g.K) ? tf32_rn(g.alpha * __ldg(A + (int64_t)r1 * g.lda + k0)) : 0u;
      a[s][2] = (r0 < g.M && k1 < g.K) ? tf32_rn(g.alpha * __ldg(A + (int64_t)r0 * g.lda + k1)) : 0u;
      a[s][3] = (r1 < g.M && k1 < g.K) ? tf32_rn(g.alpha * __ldg(A + (int64_t)r1 * g.lda + k1)) : 0u;
    }
    const int j0 = half * tph, j1 = min(ntiles, j0 + tph);
    const int jfast = full_rows ? min(j1, jfull) : j0;         // [j0, jfast): no predicates needed
    const uint32_t* bp = Bs + (8 * j0 + gid) * T::PITCH + 2 * tig;
    int j = j0;
    if (!BF16) {
      float* __restrict__ C = reinterpret_cast<float*>(g.C) + zo * g.sCo + zi * g.sCi;
      float* c0p = C + (int64_t)r0 * g.ldc + 2 * tig + 8 * j0;
      float* c1p = c0p + 8 * g.ldc;
      if (g.staged) {
        // four n-tiles (32 columns = one 128-byte line per row) at a time through a per-warp staging tile, so every
        // global store instruction writes 4 full lines (512 contiguous-per-row bytes) instead of 8 x 32-byte sectors
        float* stg = reinterpret_cast<float*>(Bs + N8 * T::PITCH) + w * (16 * 40);
        float* crow = C + (int64_t)(strip * 16 + (lane >> 3)) * g.ldc + 4 * (lane & 7);
        for (; j + 3 < jfast; j += 4, bp += 32 * T::PITCH, c0p += 32, c1p += 32) {
          float c[4][4];
#pragma unroll
          for (int t = 0; t < 4; ++t) { c[t][0] = 0.f; c[t][1] = 0.f; c[t][2] = 0.f; c[t][3] = 0.f; }
#pragma unroll
          for (int s = 0; s < KS; ++s)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint2 b = *reinterpret_cast<const uint2*>(bp + 8 * t * T::PITCH + 8 * s);
              mma_16x8x8(c[t], a[s], b.x, b.y);
            }
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            *reinterpret_cast<float2*>(stg + gid * 40 + 8 * t + 2 * tig) = make_float2(c[t][0], c[t][1]);
            *reinterpret_cast<float2*>(stg + (gid + 8) * 40 + 8 * t + 2 * tig) = make_float2(c[t][2], c[t][3]);
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(stg + (4 * i + (lane >> 3)) * 40 + 4 * (lane & 7));
            *reinterpret_cast<float4*>(crow + (int64_t)(4 * i) * g.ldc + 8 * j) = v;
          }
          __syncwarp();
        }
      }
      for (; j + 1 < jfast; j += 2, bp += 16 * T::PITCH, c0p += 16, c1p += 16) {        // two independent n-tiles
        float c[4] = {0.f, 0.f, 0.f, 0.f}, d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const uint2 b = *reinterpret_cast<const uint2*>(bp + 8 * s);
          const uint2 e = *reinterpret_cast<const uint2*>(bp + 8 * T::PITCH + 8 * s);
          mma_16x8x8(c, a[s], b.x, b.y);
          mma_16x8x8(d, a[s], e.x, e.y);
        }
        *reinterpret_cast<float2*>(c0p) = make_float2(c[0], c[1]);
        *reinterpret_cast<float2*>(c1p) = make_float2(c[2], c[3]);
        *reinterpret_cast<float2*>(c0p + 8) = make_float2(d[0], d[1]);
        *reinterpret_cast<float2*>(c1p + 8) = make_float2(d[2], d[3]);
      }
      for (; j < j1; ++j, bp += 8 * T::PITCH, c0p += 8, c1p += 8) {                      // remainder, fully predicated
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const uint2 b = *reinterpret_cast<const uint2*>(bp + 8 * s);
          mma_16x8x8(c, a[s], b.x, b.y);
        }
        const int col = 8 * j + 2 * tig;
        if (col + 1 < g.N) {
          if (r0 < g.M) *reinterpret_cast<float2*>(c0p) = make_float2(c[0], c[1]);
          if (r1 < g.M) *reinterpret_cast<float2*>(c1p) = make_float2(c[2], c[3]);
        } else if (col < g.N) {
          if (r0 < g.M) c0p[0] = c[0];
          if (r1 < g.M) c1p[0] = c[2];
        }
      }
    } else {
      // bf16 output: two n-tiles per step; even lanes of a pair keep tile j (4 consecutive columns after one exchange
      // with the odd neighbour), odd lanes keep tile j + 1 -> one 8-byte store per row, full 32-byte sectors per row
      __nv_bfloat16* __restrict__ C = reinterpret_cast<__nv_bfloat16*>(g.C) + zo * g.sCo + zi * g.sCi;
      __nv_bfloat16* c0p = C + (int64_t)r0 * g.ldc + 8 * (j0 + odd) + 2 * (tig - odd);
      __nv_bfloat16* c1p = c0p + 8 * g.ldc;
      if (g.staged) {
        // eight n-tiles (64 columns = one 128-byte line of bf16 per row) at a time through the per-warp staging tile:
        // packed bf16 pairs go to shared memory (conflict-free: row pitch 36 words), come back as 16-byte chunks and
        // leave as full-line stores -- no lane exchange, 4 store instructions per 1024 outputs instead of 16
        uint32_t* stg = Bs + N8 * T::PITCH + w * (16 * 40);
        __nv_bfloat16* crow = C + (int64_t)(strip * 16 + (lane >> 3)) * g.ldc + 8 * (lane & 7);
        for (; j + 7 < jfast; j += 8, bp += 64 * T::PITCH, c0p += 64, c1p += 64) {
          float c[8][4];
#pragma unroll
          for (int t = 0; t < 8; ++t) { c[t][0] = 0.f; c[t][1] = 0.f; c[t][2] = 0.f; c[t][3] = 0.f; }
#pragma unroll
          for (int s = 0; s < KS; ++s)
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const uint2 b = *reinterpret_cast<const uint2*>(bp + 8 * t * T::PITCH + 8 * s);
              mma_16x8x8(c[t], a[s], b.x, b.y);
            }
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(c[t][0], c[t][1]), hi = __floats2bfloat162_rn(c[t][2], c[t][3]);
            stg[gid * 36 + 4 * t + tig] = *reinterpret_cast<uint32_t*>(&lo);
            stg[(gid + 8) * 36 + 4 * t + tig] = *reinterpret_cast<uint32_t*>(&hi);
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 v = *reinterpret_cast<const uint4*>(stg + (4 * i + (lane >> 3)) * 36 + 4 * (lane & 7));
            *reinterpret_cast<uint4*>(crow + (int64_t)(4 * i) * g.ldc + 8 * j) = v;
          }
          __syncwarp();
        }
      }
      for (; j < j1; j += 2, bp += 16 * T::PITCH, c0p += 16, c1p += 16) {
        float c[4] = {0.f, 0.f, 0.f, 0.f}, d[4] = {0.f, 0.f, 0.f, 0.f};
        const bool second = j + 1 < j1;                        // warp-uniform
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const uint2 b = *reinterpret_cast<const uint2*>(bp + 8 * s);
          mma_16x8x8(c, a[s], b.x, b.y);
          if (second) {
            const uint2 e = *reinterpret_cast<const uint2*>(bp + 8 * T::PITCH + 8 * s);
            mma_16x8x8(d, a[s], e.x, e.y);
          }
        }
        // even lane sends its tile j+1 values, odd lane sends its tile j values
        float s0 = odd ? c[0] : d[0], s1 = odd ? c[1] : d[1], s2 = odd ? c[2] : d[2], s3 = odd ? c[3] : d[3];
        s0 = __shfl_xor_sync(0xffffffffu, s0, 1); s1 = __shfl_xor_sync(0xffffffffu, s1, 1);
        s2 = __shfl_xor_sync(0xffffffffu, s2, 1); s3 = __shfl_xor_sync(0xffffffffu, s3, 1);
        // even: cols [8j + 2tig, +3] = own c then the partner's c; odd: tile j+1, the partner's d first then own d
        const uint2 lo = odd ? pack4_bf16(s0, s1, d[0], d[1]) : pack4_bf16(c[0], c[1], s0, s1);      // row r0
        const uint2 hi = odd ? pack4_bf16(s2, s3, d[2], d[3]) : pack4_bf16(c[2], c[3], s2, s3);      // row r1
        if (j + 1 < jfast) {                                   // warp-uniform fast path
          *reinterpret_cast<uint2*>(c0p) = lo;
          *reinterpret_cast<uint2*>(c1p) = hi;
        } else {
          const int col = 8 * (j + odd) + 2 * (tig - odd);
          if ((!odd || second) && col < g.N) {
            if (col + 3 < g.N) {
              if (r0 < g.M) *reinterpret_cast<uint2*>(c0p) = lo;
              if (r1 < g.M) *reinterpret_cast<uint2*>(c1p) = hi;
            } else {
              const __nv_bfloat16* l4 = reinterpret_cast<const __nv_bfloat16*>(&lo);
              const __nv_bfloat16* h4 = reinterpret_cast<const __nv_bfloat16*>(&hi);
              for (int t = 0; t < 4 && col + t < g.N; ++t) {
                if (r0 < g.M) c0p[t] = l4[t];
                if (r1 < g.M) c1p[t] = h4[t];
              }
            }
          }
        }
      }
    }
  }
}

template <int KS, bool BF16>
static int launch_scores(const ScoresArgs& g, int batch, size_t smem, cudaStream_t s) {
  static uint64_t seen = 0;
  if (first_use_on_device(seen))
    cudaFuncSetAttribute(scores_mma_kernel<KS, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  scores_mma_kernel<KS, BF16><<<(unsigned)(batch * g.chunks), 256, smem, s>>>(g);
  return check_launch("vu_gemm");
}

// Called from vu_gemm for TF32 requests; sets *handled when the shape is a scores GEMM.  VU_GEMM_SCORES=0 disables.
int gemm_scores(const vu_gemm_desc& d, cudaStream_t s, bool* handled) {
  *handled = false;
  static const bool on = []() { const char* e = getenv("VU_GEMM_SCORES"); return !(e && e[0] == '0'); }();
  if (!on) return VU_OK;
  if (d.trans_a || !d.trans_b || d.K > 128 || d.M < 64 || d.N < 64) return VU_OK;
  if (d.a_bf16 || d.b_bf16 || d.bias || d.residual || d.aux_in || d.aux_out || d.act != VU_ACT_NONE || d.accumulate ||
      d.split_k > 1 || d.drop_p > 0.f)
    return VU_OK;
  const int64_t batch = (int64_t)std::max(1, d.batch_outer) * std::max(1, d.batch_inner);
  if (batch > 0x7fffffff || d.N > 16384) return VU_OK;
  // vector stores: fp32 pairs need 8-byte, bf16 quads 8-byte alignment of every row start
  const int esz = d.c_bf16 ? 2 : 4;
  const int need = d.c_bf16 ? 4 : 2;
  if ((uintptr_t)d.C % 8 || d.ldc % need || d.sCo % need || d.sCi % need) return VU_OK;
  (void)esz;
  int KS = (d.K + 7) / 8;                                      // k-steps, rounded up to an instantiated count
  KS = KS <= 4 ? KS : (KS <= 6 ? 6 : (KS <= 8 ? 8 : (KS <= 12 ? 12 : 16)));
  const int N8 = (d.N + 7) & ~7;
  const int pitch = ((KS * 8) % 16 == 0) ? KS * 8 + 8 : KS * 8;
  static const bool staged_on = []() { const char* e = getenv("VU_SCORES_STAGED"); return !(e && e[0] == '0'); }();
  const int va = d.c_bf16 ? 8 : 4;                             // elements per 16-byte store
  const bool staged = staged_on && d.ldc % va == 0 && d.sCo % va == 0 && d.sCi % va == 0 && (uintptr_t)d.C % 16 == 0;
  const size_t smem = (size_t)N8 * pitch * 4 + (staged ? 8 * 16 * 40 * 4 : 0);
  if (smem > 200 * 1024) return VU_OK;
  ScoresArgs g;
  g.A = d.A; g.B = d.B; g.C = d.C; g.M = d.M; g.N = d.N; g.K = d.K;
  g.lda = d.lda; g.ldb = d.ldb; g.ldc = d.ldc; g.batch_inner = std::max(1, d.batch_inner);
  g.sAo = d.sAo; g.sAi = d.sAi; g.sBo = d.sBo; g.sBi = d.sBi; g.sCo = d.sCo; g.sCi = d.sCi;
  g.alpha = d.alpha;
  g.staged = staged ? 1 : 0;
  g.vec4 = (d.K % 4 == 0 && d.ldb % 4 == 0 && d.sBo % 4 == 0 && d.sBi % 4 == 0 && (uintptr_t)d.B % 16 == 0) ? 1 : 0;
  // enough CTAs to fill the machine, but at least 32 strips (512 rows) per CTA to amortise staging B
  const int64_t strips = cdiv(d.M, 16);
  const int64_t want = cdiv((int64_t)sm_count() * 4, batch);
  g.chunks = (int)std::max<int64_t>(1, std::min<int64_t>(want, strips / 32));
  if (batch * g.chunks > 0x7fffffff) return VU_OK;
  *handled = true;
#define VU_SC(KSV) (d.c_bf16 ? launch_scores<KSV, true>(g, (int)batch, smem, s) : launch_scores<KSV, false>(g, (int)batch, smem, s))
  switch (KS) {
    case 1: return VU_SC(1);
    case 2: return VU_SC(2);
    case 3: return VU_SC(3);
    case 4: return VU_SC(4);
    case 6: return VU_SC(6);
    case 8: return VU_SC(8);
    case 12: return VU_SC(12);
    default: return VU_SC(16);
  }
#undef VU_SC
}

}  // namespace vu
