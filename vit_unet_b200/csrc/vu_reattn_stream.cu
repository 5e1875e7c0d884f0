// Streamed Re-Attention forward for the fine levels (many tokens, small heads): the (B, h, N, N) attention maps are
// never written.  Reference math (ReAttention.forward, model.py:155-161; SkipConnection.forward :251-256):
//     S_g = q_g k_g^T * hd^-1/2 ;  P_g = softmax(S_g) ;  Pd_g = dropout(P_g) ;
//     A_h = BN_h( sum_g W[h,g] Pd_g + b_h )  = sum_g alpha[h,g] Pd_g + beta[h]   (1x1 conv + BatchNorm folded)
//     O_h = A_h v_h
// Head mixing acts on NORMALISED probabilities of all heads at the same (query, key) position, so flash-attention's
// running rescale does not apply; instead the scores are recomputed in each sweep over the keys (the contraction is
// only hd = 8..48 long, a few warp MMAs per 16x16 tile) and nothing but per-row softmax constants is kept:
//   sweep A  row maximum m and sum l of every (head, query row)              -> c = m + log2 l  (log2 domain)
//   sweep B  (train) Pd = dropout(exp2(s - c)); centred moments s'_g, G'_gg' -> BatchNorm batch statistics
//            [vu_reattn_bn_finalize runs between the two launches: the statistics couple the whole batch]
//   sweep C  Pd again, A = alpha Pd + beta (8x8 mix, thread-local: every lane holds all heads of its positions),
//            O += A v on the tensor cores (A fragments are the mixed accumulators, re-used as the MMA A operand)
// One CTA = (image, 7 x 16 query rows); one warp owns 16 query rows and ALL heads.  K tiles (fp32, read as TF32 by
// m16n8k8) and V^T tiles (bf16, m16n8k16) of 32 keys are double-buffered in shared memory with cp.async; Q stays
// resident.  Within a 16-key step the key index is permuted so that a lane owns FOUR CONSECUTIVE keys of two rows:
// one counter-hash call per quad gives the dropout mask (identical to the mask of the materialised kernels in
// vu_reattn.cu, element for element), and V^T fragments are single 8-byte shared loads.
#include <cuda_bf16.h>

#include <algorithm>

#include "vu_common.cuh"

#ifndef VU_RS_FFMA2
#define VU_RS_FFMA2 1
#endif

namespace vu {
namespace rs {

constexpr int KT = 32;            // keys per shared-memory tile (two 16-key steps)
constexpr int WARPS = 7;          // 7 x 16 = 112 query rows per CTA: 784 = 7 x 112, 3136 = 28 x 112
constexpr int VP = 96;            // bytes per V^T row in shared memory (64 data + 32: conflict-free 8-byte fragment loads)

enum { MODE_EVAL = 0, MODE_STATS = 1, MODE_APPLY = 2 };

struct Args {
  const float* q; const float* k;           // (B, N, D) fp32, head h = columns [h*hd, (h+1)*hd)
  const __nv_bfloat16* vt;                  // (B, H, hd, ldn) bf16: per-head transposed values
  float* o;                                 // (B, N, D)
  const float* fold;                        // H*H alpha[h][g] then H beta[h]
  float* rowc;                              // (B, H, N) softmax constants c = m + log2 l (written by STATS, read by APPLY)
  double* sums;                             // H + H*H centred moments (STATS)
  __nv_bfloat16* pc;                        // optional (B, H, N, N) centred probabilities P - 1/N for the backward pass
  __nv_bfloat16* amap;                      // optional (B, H, N, N) mixed map A written by APPLY (backward: dV = A^T dO)
  uint2* mask;                              // dropout keep-bits cached by STATS for APPLY: (B, N/16, N/16, 32 lanes) x 64 bits
  int N, D, ldn;
  float sl2;                                // hd^-1/2 * log2(e)
  uint32_t thresh; float dscale; uint32_t key;     // dropout: 16-bit threshold, 1/(1-p), hash key of (seed, stream)
  float cN;                                 // 1/N
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = valid ? 16 : 0;             // src-size 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {             // MUFU.EX2 (2^-22 relative), flushes denormal results to 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32 FMA (sm_100: FFMA2, two lanes per issue slot): (d0, d1) += (a0, a1) * w
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float w) {
#if VU_RS_FFMA2
  asm("{ .reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %4}; mov.b64 rd, {%0, %1};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rd;\n\t"
      "mov.b64 {%0, %1}, rd; }"
      : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1), "f"(w));
#else
  d0 = fmaf(a0, w, d0); d1 = fmaf(a1, w, d1);
#endif
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&p);
}

// Shared-memory plan (bytes), QP = D + 4 floats per Q / K row: (QP * 4) % 128 is an odd multiple of 16 when D % 8 == 0,
// so the eight 16-byte rows of every ldmatrix 8x8 block fall on disjoint bank groups.
template <int H, int HD>
struct Plan {
  static constexpr int D = H * HD, QP = D + 4;
  static constexpr int KS = (HD + 7) / 8;                 // k-steps of the score MMAs = n-tiles of the O accumulators
  static constexpr size_t q_bytes = (size_t)WARPS * 16 * QP * 4;
  static constexpr size_t k_bytes = (size_t)KT * QP * 4;  // one stage
  static constexpr size_t v_bytes = (size_t)(D + 8) * VP; // one stage (+8 rows: the last head's padded n-tile)
  static constexpr size_t w_bytes = (size_t)(H * H + H) * 4;
  static constexpr size_t total(bool with_v) { return q_bytes + 2 * k_bytes + (with_v ? 2 * v_bytes : 0) + w_bytes + 16; }
};

// Scores of one 16-key step for all heads: s[g][u][0..3] (u = n-tile of the pair).  Lane (gid, tig) ends up with rows
// gid / gid + 8 and keys 4 tig .. 4 tig + 3 of the step: (u, c) -> key 4 tig + 2 u + c.
template <int H, int HD>
__device__ __forceinline__ void scores_step(float (&s)[H][2][4], const float* Qw, const float* Kst, int lane) {
  using P = Plan<H, HD>;
  const int m = lane >> 3, r = lane & 7;
  // A (queries): matrix m -> rows (m & 1) * 8 + r, k offset (m >> 1) * 4
  const float* qa = Qw + ((m & 1) * 8 + r) * P::QP + (m >> 1) * 4;
  // B (keys), two n-tiles: matrix m -> tile u = m >> 1, k half m & 1; n index r -> key 4 (r / 2) + 2 u + (r % 2)
  const float* ka = Kst + (4 * (r >> 1) + 2 * (m >> 1) + (r & 1)) * P::QP + (m & 1) * 4;
#pragma unroll
  for (int g = 0; g < H; ++g) {
#pragma unroll
    for (int u = 0; u < 2; ++u) { s[g][u][0] = 0.f; s[g][u][1] = 0.f; s[g][u][2] = 0.f; s[g][u][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < P::KS; ++ks) {
      uint32_t a[4], b[4];
      ldmatrix_x4(a, qa + g * HD + ks * 8);
      ldmatrix_x4(b, ka + g * HD + ks * 8);
      if (ks * 8 + 4 >= HD) { a[2] = 0u; a[3] = 0u; }      // head dims that are not a multiple of 8 (12): k >= hd reads the next head
      mma_tf32(s[g][0], a, b[0], b[1]);
      mma_tf32(s[g][1], a, b[2], b[3]);
    }
  }
}

template <int H, int HD, int MODE, bool WRITE_PC>
__global__ void __launch_bounds__(WARPS * 32, 1)
stream_fwd_kernel(const Args g) {
  using P = Plan<H, HD>;
  constexpr int D = P::D, QP = P::QP, KS = P::KS;
  constexpr bool HAS_A = MODE != MODE_APPLY, HAS_B = MODE == MODE_STATS, HAS_C = MODE != MODE_STATS;
  extern __shared__ __align__(128) unsigned char smem[];
  float* Qs = reinterpret_cast<float*>(smem);
  float* Ks = Qs + WARPS * 16 * QP;
  unsigned char* Vs = reinterpret_cast<unsigned char*>(Ks + 2 * KT * QP);
  float* Ws = reinterpret_cast<float*>(Vs + (HAS_C ? 2 * P::v_bytes : 0));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int N = g.N, b = blockIdx.y, row0 = blockIdx.x * (WARPS * 16);
  const float* __restrict__ qb = g.q + (size_t)b * N * D;
  const float* __restrict__ kb = g.k + (size_t)b * N * D;
  const __nv_bfloat16* __restrict__ vb = HAS_C ? g.vt + (size_t)b * D * g.ldn : nullptr;
  const int ntiles = (N + KT - 1) / KT;

  auto load_k = [&](int tile, int stage) {
    float* dst = Ks + stage * KT * QP;
    const int key0 = tile * KT;
    for (int c = tid; c < KT * (D / 4); c += WARPS * 32) {
      const int kr = c / (D / 4), q4 = c - kr * (D / 4);
      const bool ok = key0 + kr < N;
      cp_async16(dst + kr * QP + 4 * q4, ok ? kb + (size_t)(key0 + kr) * D + 4 * q4 : kb, ok);
    }
  };
  auto load_v = [&](int tile, int stage) {
    unsigned char* dst = Vs + stage * P::v_bytes;
    const int key0 = tile * KT;
    for (int c = tid; c < D * (KT / 8); c += WARPS * 32) {
      const int row = c / (KT / 8), q8 = c - row * (KT / 8);
      const bool ok = key0 + 8 * q8 < N;        // ldn >= N rounded up to 8 and N % 16 == 0: whole 16-byte chunks
      cp_async16(dst + row * VP + 16 * q8, ok ? (const void*)(vb + (size_t)row * g.ldn + key0 + 8 * q8) : (const void*)vb, ok);
    }
  };

  // ---- prologue: resident Q rows of the CTA, the fold, zeroed pads; first K (and V) tile in flight
  for (int c = tid; c < WARPS * 16 * (D / 4); c += WARPS * 32) {
    const int r = c / (D / 4), q4 = c - r * (D / 4);
    const bool ok = row0 + r < N;
    cp_async16(Qs + r * QP + 4 * q4, ok ? qb + (size_t)(row0 + r) * D + 4 * q4 : qb, ok);
  }
  constexpr int FIRST_HAS_V = (MODE == MODE_APPLY);
  load_k(0, 0);
  if (FIRST_HAS_V) load_v(0, 0);
  cp_async_commit();
  for (int r = tid; r < WARPS * 16; r += WARPS * 32) *reinterpret_cast<float4*>(Qs + r * QP + D) = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = tid; r < 2 * KT; r += WARPS * 32) *reinterpret_cast<float4*>(Ks + r * QP + D) = make_float4(0.f, 0.f, 0.f, 0.f);
  if (HAS_C) {
    for (int i = tid; i < 2 * 8 * (VP / 4); i += WARPS * 32) {       // the 8 spare rows of both V stages
      const int st = i / (8 * (VP / 4)), w = i - st * (8 * (VP / 4));
      reinterpret_cast<uint32_t*>(Vs + st * P::v_bytes + D * VP)[w] = 0u;
    }
    // fold for sweep C: alpha (dropout keep-scale folded in: the mix sees keep ? p : 0) and beta
    for (int i = tid; i < H * H + H; i += WARPS * 32) Ws[i] = i < H * H ? g.fold[i] * g.dscale : g.fold[i];
  }

  const bool active = row0 + warp * 16 < N;                  // warp-uniform (N % 16 == 0: a unit is whole or absent)
  const float* Qw = Qs + warp * 16 * QP;
  const int rowa = row0 + warp * 16 + gid;                   // this lane's rows: rowa and rowa + 8
  float cst[H][2];                                           // softmax constants c = m + log2 l of (head, row)

  int item = 0;                                              // running index over (sweep, tile) -> stage = item & 1
  auto next_tile = [&](int t, int sweep_has_v, bool more, int next_t, int next_has_v) {
    // wait for tile `item`, make it visible, then prefetch the following one into the other stage
    cp_async_wait<0>();
    __syncthreads();
    if (more) {
      load_k(next_t, (item + 1) & 1);
      if (next_has_v) load_v(next_t, (item + 1) & 1);
      cp_async_commit();
    }
    (void)t; (void)sweep_has_v;
  };

  // =================================================================== sweep A: row maxima and sums (log2 domain)
  if (HAS_A) {
    float mx[H][2], l[H][2];
#pragma unroll
    for (int h = 0; h < H; ++h) { mx[h][0] = -INFINITY; mx[h][1] = -INFINITY; l[h][0] = 0.f; l[h][1] = 0.f; }
    for (int t = 0; t < ntiles; ++t, ++item) {
      const bool last = t + 1 == ntiles;
      next_tile(t, 0, true, last ? 0 : t + 1, last ? (MODE == MODE_EVAL) : 0);       // after A always comes B or C from tile 0
      if (!active) continue;
      const float* Kst = Ks + (item & 1) * KT * QP;
      const int steps = min(KT / 16, (N - t * KT) / 16);
      for (int st = 0; st < steps; ++st) {
        float s[H][2][4];
        scores_step<H, HD>(s, Qw, Kst + st * 16 * QP, lane);
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const float a0 = s[h][0][2 * r], a1 = s[h][0][2 * r + 1], a2 = s[h][1][2 * r], a3 = s[h][1][2 * r + 1];
            const float tm = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)) * g.sl2;
            const float mn = fmaxf(mx[h][r], tm);
            const float e = ex2(fmaf(a0, g.sl2, -mn)) + ex2(fmaf(a1, g.sl2, -mn)) + ex2(fmaf(a2, g.sl2, -mn)) +
                            ex2(fmaf(a3, g.sl2, -mn));
            l[h][r] = fmaf(l[h][r], ex2(mx[h][r] - mn), e);
            mx[h][r] = mn;
          }
      }
    }
    // combine the four lanes of a row (they hold disjoint key quads)
#pragma unroll
    for (int h = 0; h < H; ++h)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float m = mx[h][r], ll = l[h][r];
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
          const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, ll, o);
          const float mn = fmaxf(m, m2);
          ll = ll * ex2(m - mn) + l2 * ex2(m2 - mn);
          m = mn;
        }
        cst[h][r] = m + log2f(ll);
      }
    if (MODE == MODE_STATS && active && tig == 0) {
#pragma unroll
      for (int h = 0; h < H; ++h) {
        g.rowc[((size_t)b * H + h) * N + rowa] = cst[h][0];
        g.rowc[((size_t)b * H + h) * N + rowa + 8] = cst[h][1];
      }
    }
  } else if (active) {
#pragma unroll
    for (int h = 0; h < H; ++h) {
      cst[h][0] = __ldg(g.rowc + ((size_t)b * H + h) * N + rowa);
      cst[h][1] = __ldg(g.rowc + ((size_t)b * H + h) * N + rowa + 8);
    }
  }

  // dropout counters: flat element index of (b, g, row, key) in a (B, H, N, N) map, divided by 4 (one hash per quad)
  const uint32_t nn4 = (uint32_t)(((size_t)N * N) >> 2);
  const uint32_t ctr_r0 = (uint32_t)((((size_t)b * H * N + rowa) * N) >> 2) + tig;
  const uint32_t ctr_r1 = ctr_r0 + 8u * (uint32_t)(N >> 2);

  // =================================================================== sweep B: centred moments of the dropped maps
  if (HAS_B) {
    constexpr int NV = H + H * (H + 1) / 2;
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    for (int t = 0; t < ntiles; ++t, ++item) {
      const bool last = t + 1 == ntiles;
      next_tile(t, 0, !last, t + 1, 0);
      if (!active) continue;
      const float* Kst = Ks + (item & 1) * KT * QP;
      const int steps = min(KT / 16, (N - t * KT) / 16);
      for (int st = 0; st < steps; ++st) {
        float s[H][2][4];
        scores_step<H, HD>(s, Qw, Kst + st * 16 * QP, lane);
        const uint32_t j4 = (uint32_t)((t * KT + st * 16) >> 2);
        uint32_t bits[2] = {0u, 0u};              // keep-bits of this lane's 2 rows x H heads x 4 keys (bit 4h + key)
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            float p0 = ex2(fmaf(s[h][0][2 * r], g.sl2, -cst[h][r])), p1 = ex2(fmaf(s[h][0][2 * r + 1], g.sl2, -cst[h][r]));
            float p2 = ex2(fmaf(s[h][1][2 * r], g.sl2, -cst[h][r])), p3 = ex2(fmaf(s[h][1][2 * r + 1], g.sl2, -cst[h][r]));
            if (WRITE_PC) {
              uint2 w; w.x = pack_bf16(p0 - g.cN, p1 - g.cN); w.y = pack_bf16(p2 - g.cN, p3 - g.cN);
              *reinterpret_cast<uint2*>(g.pc + (((size_t)b * H + h) * N + rowa + 8 * r) * N + t * KT + st * 16 + 4 * tig) = w;
            }
            if (g.thresh) {
              const uint4 rr = Philox::gen_k(g.key, (r ? ctr_r1 : ctr_r0) + (uint32_t)h * nn4 + j4);
              const bool k0 = rr.x >= g.thresh, k1 = rr.y >= g.thresh, k2 = rr.z >= g.thresh, k3 = rr.w >= g.thresh;
              bits[r] |= ((uint32_t)k0 | ((uint32_t)k1 << 1) | ((uint32_t)k2 << 2) | ((uint32_t)k3 << 3)) << (4 * h);
              p0 = k0 ? p0 * g.dscale : 0.f; p1 = k1 ? p1 * g.dscale : 0.f;
              p2 = k2 ? p2 * g.dscale : 0.f; p3 = k3 ? p3 * g.dscale : 0.f;
            }
            s[h][0][2 * r] = p0 - g.cN; s[h][0][2 * r + 1] = p1 - g.cN; s[h][1][2 * r] = p2 - g.cN; s[h][1][2 * r + 1] = p3 - g.cN;
          }
        if (g.thresh && g.mask)
          g.mask[(((size_t)b * (N >> 4) + (rowa >> 4)) * (N >> 4) + (t * (KT / 16) + st)) * 32 + lane] = make_uint2(bits[0], bits[1]);
        int kk = H;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          float sm = 0.f;
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int e = 0; e < 4; ++e) sm += s[h][u][e];
          acc[h] += sm;
#pragma unroll
          for (int h2 = h; h2 < H; ++h2) {
            float d = 0.f;
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
              for (int e = 0; e < 4; ++e) d = fmaf(s[h][u][e], s[h2][u][e], d);
            acc[kk] += d; ++kk;
          }
        }
      }
    }
    // CTA reduction (the shared-memory tiles are dead after the barrier) -> one set of double atomics per CTA
    __syncthreads();
    double* red = reinterpret_cast<double*>(smem);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float v = warp_sum(acc[i]);
      if (lane == 0) red[i * WARPS + warp] = (double)v;
    }
    __syncthreads();
    if (tid < NV) {
      double v = 0.0;
      for (int w = 0; w < WARPS; ++w) v += red[tid * WARPS + w];
      // tid -> (g, g2): first H entries are s'_g, then the upper triangle row by row
      if (tid < H) atomicAdd(g.sums + tid, v);
      else {
        int idx = tid - H, gg = 0;
        while (idx >= H - gg) { idx -= H - gg; ++gg; }
        const int g2 = gg + idx;
        atomicAdd(g.sums + H + gg * H + g2, v);
        if (g2 != gg) atomicAdd(g.sums + H + g2 * H + gg, v);
      }
    }
    return;
  }

  // =================================================================== sweep C: mix + A.V
  if (HAS_C) {
    float o[H][KS][4];
#pragma unroll
    for (int h = 0; h < H; ++h)
#pragma unroll
      for (int t = 0; t < KS; ++t) { o[h][t][0] = 0.f; o[h][t][1] = 0.f; o[h][t][2] = 0.f; o[h][t][3] = 0.f; }
    for (int t = 0; t < ntiles; ++t, ++item) {
      const bool last = t + 1 == ntiles;
      next_tile(t, 1, !last, t + 1, 1);
      if (!active) continue;
      const float* Kst = Ks + (item & 1) * KT * QP;
      const unsigned char* Vst = Vs + (item & 1) * P::v_bytes;
      const int steps = min(KT / 16, (N - t * KT) / 16);
      for (int st = 0; st < steps; ++st) {
        float s[H][2][4];
        scores_step<H, HD>(s, Qw, Kst + st * 16 * QP, lane);
        const uint32_t j4 = (uint32_t)((t * KT + st * 16) >> 2);
        uint2 bits = make_uint2(0xffffffffu, 0xffffffffu);
        const bool cached = MODE == MODE_APPLY && g.thresh && g.mask;         // keep-bits cached by the statistics launch
        if (cached) bits = __ldg(g.mask + (((size_t)b * (N >> 4) + (rowa >> 4)) * (N >> 4) + (t * (KT / 16) + st)) * 32 + lane);
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            float p0 = ex2(fmaf(s[h][0][2 * r], g.sl2, -cst[h][r])), p1 = ex2(fmaf(s[h][0][2 * r + 1], g.sl2, -cst[h][r]));
            float p2 = ex2(fmaf(s[h][1][2 * r], g.sl2, -cst[h][r])), p3 = ex2(fmaf(s[h][1][2 * r + 1], g.sl2, -cst[h][r]));
            if (cached) {
              const uint32_t w = (r ? bits.y : bits.x) >> (4 * h);
              p0 = (w & 1u) ? p0 : 0.f; p1 = (w & 2u) ? p1 : 0.f; p2 = (w & 4u) ? p2 : 0.f; p3 = (w & 8u) ? p3 : 0.f;
            } else if (MODE == MODE_APPLY && g.thresh) {
              const uint4 rr = Philox::gen_k(g.key, (r ? ctr_r1 : ctr_r0) + (uint32_t)h * nn4 + j4);
              p0 = rr.x >= g.thresh ? p0 : 0.f; p1 = rr.y >= g.thresh ? p1 : 0.f;
              p2 = rr.z >= g.thresh ? p2 : 0.f; p3 = rr.w >= g.thresh ? p3 : 0.f;
            }
            s[h][0][2 * r] = p0; s[h][0][2 * r + 1] = p1; s[h][1][2 * r] = p2; s[h][1][2 * r + 1] = p3;
          }
        // head mixing, one target head at a time; the mixed values become the bf16 A fragment of the A.V MMAs:
        // reg0 = (row gid, keys 4tig, 4tig+1) = MMA k 2tig, 2tig+1; reg1 = row gid+8; reg2/3 = keys 4tig+2, +3 = MMA k 2tig+8, +9
        const unsigned char* vrow = Vst + gid * VP + st * 32 + tig * 8;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float bias = Ws[H * H + h];
          float a[2][4];
#pragma unroll
          for (int u = 0; u < 2; ++u) { a[u][0] = bias; a[u][1] = bias; a[u][2] = bias; a[u][3] = bias; }
#pragma unroll
          for (int q4 = 0; q4 < H / 4; ++q4) {
            const float4 w = *reinterpret_cast<const float4*>(Ws + h * H + 4 * q4);
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int gi = 0; gi < 4; ++gi)
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                ffma2(a[u][0], a[u][1], s[4 * q4 + gi][u][0], s[4 * q4 + gi][u][1], wv[gi]);
                ffma2(a[u][2], a[u][3], s[4 * q4 + gi][u][2], s[4 * q4 + gi][u][3], wv[gi]);
              }
          }
          uint32_t af[4];
          af[0] = pack_bf16(a[0][0], a[0][1]); af[1] = pack_bf16(a[0][2], a[0][3]);
          af[2] = pack_bf16(a[1][0], a[1][1]); af[3] = pack_bf16(a[1][2], a[1][3]);
          if (MODE == MODE_APPLY && g.amap) {      // keys 4tig..4tig+3 of rows gid / gid+8: one 8-byte store each
            __nv_bfloat16* ap = g.amap + (((size_t)b * H + h) * N + rowa) * N + t * KT + st * 16 + 4 * tig;
            *reinterpret_cast<uint2*>(ap) = make_uint2(af[0], af[2]);
            *reinterpret_cast<uint2*>(ap + (size_t)8 * N) = make_uint2(af[1], af[3]);
          }
#pragma unroll
          for (int nt = 0; nt < KS; ++nt) {
            const uint2 bv = *reinterpret_cast<const uint2*>(vrow + (h * HD + nt * 8) * VP);
            mma_bf16(o[h][nt], af, bv.x, bv.y);
          }
        }
      }
    }
    if (active) {
      float* ob = g.o + ((size_t)b * N + rowa) * D;
#pragma unroll
      for (int h = 0; h < H; ++h)
#pragma unroll
        for (int nt = 0; nt < KS; ++nt) {
          const int e = nt * 8 + 2 * tig;
          if (e < HD) {
            *reinterpret_cast<float2*>(ob + h * HD + e) = make_float2(o[h][nt][0], o[h][nt][1]);
            *reinterpret_cast<float2*>(ob + (size_t)8 * D + h * HD + e) = make_float2(o[h][nt][2], o[h][nt][3]);
          }
        }
    }
  }
}

template <int H, int HD, int MODE, bool WRITE_PC>
static int launch_fwd(const Args& a, int B, cudaStream_t st, const char* fn) {
  using P = Plan<H, HD>;
  const size_t smem = P::total(MODE != MODE_STATS);
  static uint64_t seen = 0;
  if (first_use_on_device(seen))
    cudaFuncSetAttribute(stream_fwd_kernel<H, HD, MODE, WRITE_PC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const dim3 grid((unsigned)cdiv(a.N, WARPS * 16), (unsigned)B);
  stream_fwd_kernel<H, HD, MODE, WRITE_PC><<<grid, WARPS * 32, smem, st>>>(a);
  return check_launch(fn);
}

static bool supported(int h, int hd, int N) {
  const bool shape = (h == 8 && (hd == 8 || hd == 24)) || (h == 4 && (hd == 12 || hd == 48));
  return shape && N % 16 == 0 && N >= 64 && N <= 8192;
}

}  // namespace rs
}  // namespace vu

#define VU_RS_DISPATCH(h, hd, CALL)                                                 \
  if (h == 8 && hd == 24) { constexpr int HH = 8, HDD = 24; CALL; }                 \
  else if (h == 8 && hd == 8) { constexpr int HH = 8, HDD = 8; CALL; }              \
  else if (h == 4 && hd == 12) { constexpr int HH = 4, HDD = 12; CALL; }            \
  else { constexpr int HH = 4, HDD = 48; CALL; }

extern "C" int vu_reattn_stream_supported(int h, int hd, int N) { return vu::rs::supported(h, hd, N) ? 1 : 0; }

// mode: 0 = eval forward in one launch (sweeps A + C with the fold of the running statistics; no dropout);
//       1 = train statistics (sweeps A + B): writes rowc (B,h,N), accumulates the centred moments into sums,
//           optionally writes the centred bf16 probabilities pc (B,h,N,N) that the backward pass consumes;
//       2 = train apply (sweep C) with the fold of the batch statistics and the row constants of mode 1.
extern "C" int vu_reattn_stream_fwd(int mode, const float* q, const float* k, const void* vt, float* o, const float* fold,
                                    float* rowc, double* sums, void* pc, void* amap, void* mask, int B, int h, int N, int hd, int ldn,
                                    float scale, float drop_p, uint64_t seed, uint32_t stream_id, void* stream) {
  using namespace vu;
  const char* fn = "vu_reattn_stream_fwd";
  VU_REQUIRE(mode >= 0 && mode <= 2, fn, "mode must be 0 (eval), 1 (train statistics) or 2 (train apply)");
  VU_REQUIRE(q && k && B > 0 && B <= 65535, fn, "null pointer or bad batch");
  VU_REQUIRE(rs::supported(h, hd, N), fn, "unsupported (heads, head_dim, tokens): see vu_reattn_stream_supported");
  VU_REQUIRE(drop_p >= 0.f && drop_p < 1.f, fn, "drop_p must be in [0,1)");
  VU_REQUIRE(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0), fn, "q / k must be 16-byte aligned");
  if (mode != 1) VU_REQUIRE(vt && o && fold && ldn % 8 == 0 && ldn >= N && ((uintptr_t)vt % 16 == 0) && ((uintptr_t)o % 8 == 0), fn,
                            "apply needs vt (ldn % 8 == 0), o and fold");
  if (mode != 0) VU_REQUIRE(rowc, fn, "train modes need the row-constant buffer");
  if (mode == 1) VU_REQUIRE(sums && (!pc || (uintptr_t)pc % 8 == 0), fn, "statistics need sums; pc must be 8-byte aligned");
  VU_REQUIRE((uintptr_t)mask % 8 == 0 && (uintptr_t)amap % 8 == 0, fn, "mask / amap must be 8-byte aligned");
  rs::Args a;
  a.q = q; a.k = k; a.vt = (const __nv_bfloat16*)vt; a.o = o; a.fold = fold; a.rowc = rowc; a.sums = sums;
  a.pc = (__nv_bfloat16*)pc; a.amap = (__nv_bfloat16*)amap; a.mask = (uint2*)mask; a.N = N; a.D = h * hd; a.ldn = ldn;
  a.sl2 = scale * 1.4426950408889634f;
  const bool drop = mode != 0 && drop_p > 0.f;
  a.thresh = drop ? drop_threshold(drop_p) : 0u;
  a.dscale = drop ? drop_keep_scale(drop_p) : 1.0f;
  a.key = Philox::key(seed, stream_id);
  a.cN = 1.0f / (float)N;
  cudaStream_t st = as_stream(stream);
  if (mode == 0) { VU_RS_DISPATCH(h, hd, return (rs::launch_fwd<HH, HDD, rs::MODE_EVAL, false>(a, B, st, fn))); }
  else if (mode == 1) {
    if (pc) { VU_RS_DISPATCH(h, hd, return (rs::launch_fwd<HH, HDD, rs::MODE_STATS, true>(a, B, st, fn))); }
    else { VU_RS_DISPATCH(h, hd, return (rs::launch_fwd<HH, HDD, rs::MODE_STATS, false>(a, B, st, fn))); }
  } else { VU_RS_DISPATCH(h, hd, return (rs::launch_fwd<HH, HDD, rs::MODE_APPLY, false>(a, B, st, fn))); }
  return VU_OK;
}
