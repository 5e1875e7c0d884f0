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
//
// Status (B200, DESIGN.md section 5, profiles/r02_stream.md): parity-green against the reference goldens in eval and
// train mode, forward and backward.  The no-grad forward is the default wherever the shape is covered (Lite inference
// 3.9x, Base inference 1.4x over the materialised chain).  For TRAINING steps the kernels are opt-in: a lane holds
// all heads of its 8 positions (64 registers) plus 96 output / 72 reduction accumulators -> 240-255 registers, one
// 7-warp CTA per SM, 15-42 % issue utilisation (latency-bound), and the materialised map kernels (23-68 % occupancy,
// 4.5-6.7 TB/s) are faster despite moving 15x the bytes.
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
// mma.sync reads fp32 bit patterns as TF32 by TRUNCATION (a -2^-11 relative bias on every operand).  Sums with heavy
// cancellation (sum dA over a map, the BatchNorm moments) turn that bias into O(1) relative errors, so every fp32 tile
// is rounded to nearest in shared memory right after it lands: one integer add per element (the low 13 bits that
// remain are ignored by the tensor core).  Each thread rounds exactly the 16-byte chunks it copied itself, after its own
// cp.async group completed and before the barrier that publishes the tile.
__device__ __forceinline__ void round_tf32_chunk(float* p) {
  uint4 v = *reinterpret_cast<uint4*>(p);
  v.x += 0x1000u; v.y += 0x1000u; v.z += 0x1000u; v.w += 0x1000u;
  *reinterpret_cast<uint4*>(p) = v;
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

// Shared-memory row of key kr (0 .. KT-1) of a tile: inside every 16-key block the keys are stored in MMA order -- rows
// 0..7 = n-tile 0 (keys 0,1,4,5,8,9,12,13), rows 8..15 = n-tile 1 (keys 2,3,6,7,10,11,14,15) -- so that each ldmatrix
// 8x8 block reads EIGHT CONSECUTIVE rows (conflict-free with the padded pitch); with keys in natural order the blocks
// would pick rows {0,1,4,5,8,9,12,13}, and rows 8 apart share their banks (2-way conflicts, measured 78 M per launch).
__device__ __forceinline__ int key_row(int kr) {
  const int k = kr & 15;
  return (kr & ~15) + (((k >> 1) & 1) << 3) + ((k >> 2) << 1) + (k & 1);
}

// Scores of one 16-key step for all heads: s[g][u][0..3] (u = n-tile of the pair).  Lane (gid, tig) ends up with rows
// gid / gid + 8 and keys 4 tig .. 4 tig + 3 of the step: (u, c) -> key 4 tig + 2 u + c.
template <int H, int HD>
__device__ __forceinline__ void scores_step(float (&s)[H][2][4], const float* Qw, const float* Kst, int lane) {
  using P = Plan<H, HD>;
  const int m = lane >> 3, r = lane & 7;
  // A (queries): matrix m -> rows (m & 1) * 8 + r, k offset (m >> 1) * 4
  const float* qa = Qw + ((m & 1) * 8 + r) * P::QP + (m >> 1) * 4;
  // B (keys), two n-tiles: matrix m -> tile u = m >> 1, k half m & 1; n index r -> key 4 (r / 2) + 2 u + (r % 2), which
  // key_row() stored at row 8 u + r of the 16-key block
  const float* ka = Kst + ((m >> 1) * 8 + r) * P::QP + (m & 1) * 4;
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

// (4 heads of 12): 140-153 registers without a cap -- just above the 146 that let TWO 7-warp CTAs share an SM; capped.
template <int H, int HD> constexpr int kMinCtas = (H == 4 && HD == 12) ? 2 : 1;

template <int H, int HD, int MODE, bool WRITE_PC>
__global__ void __launch_bounds__(WARPS * 32, kMinCtas<H, HD>)
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
      cp_async16(dst + key_row(kr) * QP + 4 * q4, ok ? kb + (size_t)(key0 + kr) * D + 4 * q4 : kb, ok);
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

  auto round_k = [&](int stage) {          // same chunk ownership as load_k
    float* dst = Ks + stage * KT * QP;
    for (int c = tid; c < KT * (D / 4); c += WARPS * 32) {
      const int kr = c / (D / 4), q4 = c - kr * (D / 4);
      round_tf32_chunk(dst + key_row(kr) * QP + 4 * q4);
    }
  };
  // ---- prologue: resident Q rows of the CTA, the fold, zeroed pads; first K (and V) tile in flight
  for (int c = tid; c < WARPS * 16 * (D / 4); c += WARPS * 32) {
    const int r = c / (D / 4), q4 = c - r * (D / 4);
    const bool ok = row0 + r < N;
    cp_async16(Qs + r * QP + 4 * q4, ok ? qb + (size_t)(row0 + r) * D + 4 * q4 : qb, ok);
  }
  cp_async_commit();                        // group 1: the Q rows (rounded below while the first K tile is in flight)
  constexpr int FIRST_HAS_V = (MODE == MODE_APPLY);
  load_k(0, 0);
  if (FIRST_HAS_V) load_v(0, 0);
  cp_async_commit();
  cp_async_wait<1>();
  for (int c = tid; c < WARPS * 16 * (D / 4); c += WARPS * 32) {
    const int r = c / (D / 4), q4 = c - r * (D / 4);
    round_tf32_chunk(Qs + r * QP + 4 * q4);
  }
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
    // wait for tile `item`, round it to TF32, make it visible, then prefetch the following one into the other stage
    cp_async_wait<0>();
    round_k(item & 1);
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

// ======================================================================================= backward (streamed)
// The backward pass reads the centred bf16 probabilities Pc = P - 1/N that the statistics launch wrote (2 bytes per
// map element, no Q / K tiles and no exp needed) and recomputes everything else on the fly:
//   stream_bwd_reduce_kernel   dA = dO v^T (TF32 MMAs, never stored);  red[h] += sum dA_h;  red[H + h H + g] += sum dA_h (Pd_g - 1/N)
//                              -> vu_reattn_bwd_params: BatchNorm-backward means + all parameter gradients of the mixing stage
//   stream_bwd_ds_kernel       sweep 1: dA again, dM_h = k_h (dA_h - m1_h - Mhat_h m2_h), dPd_g = sum_h W_hg dM_h,
//                              dP_g = keep dPd_g / (1-p) -> bf16 into the dS buffer, delta_g = sum_j dP_g P_g per row;
//                              sweep 2: dS_g = scale P_g (dP_g - delta_g) -> bf16 map (dK = dS^T q runs as a tcgen05 GEMM)
//                              and dq += dS k on the tensor cores (dS fragments are the MMA A operand, k^T tiles in smem).
// Same CTA shape and key permutation as the forward kernel (a lane owns 4 consecutive keys of two rows, all heads).
struct BwdArgs {
  const __nv_bfloat16* pc;                  // (B, H, N, N) centred probabilities
  const uint2* mask;                        // keep-bits cached by the forward statistics launch (NULL: re-hash)
  const float* dO; const float* v;          // (B, N, D) fp32
  const __nv_bfloat16* kt;                  // (B, H, hd, ldn) bf16 transposed keys (ds kernel)
  __nv_bfloat16* ds;                        // (B, H, N, N) bf16: dP after sweep 1, dS after sweep 2
  float* dq;                                // (B, N, D)
  double* red;                              // H + H*H (reduce kernel)
  const float* W; const float* bconv; const float* gamma; const float* saved; const float* coef;   // ds kernel
  int N, D, ldn, train;
  float scale;                              // hd^-1/2
  uint32_t thresh; float dscale; uint32_t key; float cN;
};

// keep-bits of this lane's (2 rows x H heads x 4 keys) for one 16-key step: cached, or regenerated with the same hash
template <int H>
__device__ __forceinline__ uint2 step_keep_bits(const BwdArgs& g, int b, int rowa, int lane, int step, uint32_t ctr_r0,
                                                uint32_t ctr_r1, uint32_t nn4) {
  if (!g.thresh) return make_uint2(0xffffffffu, 0xffffffffu);
  if (g.mask) return __ldg(g.mask + (((size_t)b * (g.N >> 4) + (rowa >> 4)) * (g.N >> 4) + step) * 32 + lane);
  uint32_t bits[2] = {0u, 0u};
  const uint32_t j4 = (uint32_t)(step * 4);
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const uint4 rr = Philox::gen_k(g.key, (r ? ctr_r1 : ctr_r0) + (uint32_t)h * nn4 + j4);
      bits[r] |= ((uint32_t)(rr.x >= g.thresh) | ((uint32_t)(rr.y >= g.thresh) << 1) | ((uint32_t)(rr.z >= g.thresh) << 2) |
                  ((uint32_t)(rr.w >= g.thresh) << 3)) << (4 * h);
    }
  return make_uint2(bits[0], bits[1]);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

template <int H, int HD>
struct BwdPlan {
  static constexpr int D = H * HD, QP = D + 4, KS = (HD + 7) / 8;
  static constexpr size_t do_bytes = (size_t)WARPS * 16 * QP * 4;
  static constexpr size_t v_bytes = (size_t)KT * QP * 4;           // one stage of fp32 value rows (B operand of dA = dO v^T)
  static constexpr size_t kt_bytes = (size_t)(D + 8) * VP;         // one stage of bf16 k^T rows (B operand of dq += dS k)
  static constexpr size_t reduce_total = do_bytes + 2 * v_bytes + 16;
  static constexpr size_t ds_total = do_bytes + 2 * (v_bytes > kt_bytes ? v_bytes : kt_bytes) + (size_t)(3 * H * H + 4 * H) * 4 + 16;
};

template <int H, int HD>
__global__ void __launch_bounds__(WARPS * 32, 1)
stream_bwd_reduce_kernel(const BwdArgs g) {
  using P = BwdPlan<H, HD>;
  constexpr int D = P::D, QP = P::QP;
  extern __shared__ __align__(128) unsigned char smem[];
  float* Gs = reinterpret_cast<float*>(smem);                 // resident dO rows of the CTA
  float* Vs = Gs + WARPS * 16 * QP;                           // 2 stages of value rows
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int N = g.N, b = blockIdx.y, row0 = blockIdx.x * (WARPS * 16);
  const float* __restrict__ gb = g.dO + (size_t)b * N * D;
  const float* __restrict__ vb = g.v + (size_t)b * N * D;
  const int ntiles = (N + KT - 1) / KT;
  auto load_v = [&](int tile, int stage) {
    float* dst = Vs + stage * KT * QP;
    const int key0 = tile * KT;
    for (int c = tid; c < KT * (D / 4); c += WARPS * 32) {
      const int kr = c / (D / 4), q4 = c - kr * (D / 4);
      const bool ok = key0 + kr < N;
      cp_async16(dst + key_row(kr) * QP + 4 * q4, ok ? vb + (size_t)(key0 + kr) * D + 4 * q4 : vb, ok);
    }
  };
  auto round_v = [&](int stage) {
    float* dst = Vs + stage * KT * QP;
    for (int c = tid; c < KT * (D / 4); c += WARPS * 32) {
      const int kr = c / (D / 4), q4 = c - kr * (D / 4);
      round_tf32_chunk(dst + key_row(kr) * QP + 4 * q4);
    }
  };
  for (int c = tid; c < WARPS * 16 * (D / 4); c += WARPS * 32) {
    const int r = c / (D / 4), q4 = c - r * (D / 4);
    const bool ok = row0 + r < N;
    cp_async16(Gs + r * QP + 4 * q4, ok ? gb + (size_t)(row0 + r) * D + 4 * q4 : gb, ok);
  }
  cp_async_commit();
  load_v(0, 0);
  cp_async_commit();
  cp_async_wait<1>();
  for (int c = tid; c < WARPS * 16 * (D / 4); c += WARPS * 32) {
    const int r = c / (D / 4), q4 = c - r * (D / 4);
    round_tf32_chunk(Gs + r * QP + 4 * q4);
  }
  for (int r = tid; r < WARPS * 16; r += WARPS * 32) *reinterpret_cast<float4*>(Gs + r * QP + D) = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = tid; r < 2 * KT; r += WARPS * 32) *reinterpret_cast<float4*>(Vs + r * QP + D) = make_float4(0.f, 0.f, 0.f, 0.f);

  const bool active = row0 + warp * 16 < N;
  const float* Gw = Gs + warp * 16 * QP;
  const int rowa = row0 + warp * 16 + gid;
  const uint32_t nn4 = (uint32_t)(((size_t)N * N) >> 2);
  const uint32_t ctr_r0 = (uint32_t)((((size_t)b * H * N + rowa) * N) >> 2) + tig;
  const uint32_t ctr_r1 = ctr_r0 + 8u * (uint32_t)(N >> 2);
  const float keep_off = g.cN * g.dscale - g.cN;              // kept: (pc + cN) dscale - cN = pc dscale + keep_off; dropped: -cN
  float x[H][H], s1[H];                                       // x[h][g] = sum dA_h (Pd_g - cN)
#pragma unroll
  for (int h = 0; h < H; ++h) {
    s1[h] = 0.f;
#pragma unroll
    for (int gg = 0; gg < H; ++gg) x[h][gg] = 0.f;
  }
  for (int t = 0; t < ntiles; ++t) {
    cp_async_wait<0>();
    round_v(t & 1);
    __syncthreads();
    if (t + 1 < ntiles) { load_v(t + 1, (t + 1) & 1); cp_async_commit(); }
    if (!active) continue;
    const float* Vst = Vs + (t & 1) * KT * QP;
    const int steps = min(KT / 16, (N - t * KT) / 16);
    for (int st = 0; st < steps; ++st) {
      const int step = t * (KT / 16) + st;
      uint2 pw[H][2];                                         // centred probabilities of (head, row): 4 keys as 4 bf16
#pragma unroll
      for (int h = 0; h < H; ++h)
#pragma unroll
        for (int r = 0; r < 2; ++r)
          pw[h][r] = __ldg(reinterpret_cast<const uint2*>(g.pc + (((size_t)b * H + h) * N + rowa + 8 * r) * N + step * 16 + 4 * tig));
      const uint2 bits = step_keep_bits<H>(g, b, rowa, lane, step, ctr_r0, ctr_r1, nn4);
      float dA[H][2][4];
      scores_step<H, HD>(dA, Gw, Vst + st * 16 * QP, lane);
      // positions e of a lane: (u, c) -> row (c >> 1), key 4 tig + 2 u + (c & 1); pdc[g][u][c] in the same order as dA
#pragma unroll
      for (int gg = 0; gg < H; ++gg) {
        float pdc[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t w = (r ? bits.y : bits.x) >> (4 * gg);
          const float c0 = bf16_lo(pw[gg][r].x), c1 = bf16_hi(pw[gg][r].x), c2 = bf16_lo(pw[gg][r].y), c3 = bf16_hi(pw[gg][r].y);
          pdc[0][2 * r] = (w & 1u) ? fmaf(c0, g.dscale, keep_off) : -g.cN;
          pdc[0][2 * r + 1] = (w & 2u) ? fmaf(c1, g.dscale, keep_off) : -g.cN;
          pdc[1][2 * r] = (w & 4u) ? fmaf(c2, g.dscale, keep_off) : -g.cN;
          pdc[1][2 * r + 1] = (w & 8u) ? fmaf(c3, g.dscale, keep_off) : -g.cN;
        }
#pragma unroll
        for (int h = 0; h < H; h += 2)
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int e = 0; e < 4; ++e) ffma2(x[h][gg], x[h + 1][gg], dA[h][u][e], dA[h + 1][u][e], pdc[u][e]);
      }
#pragma unroll
      for (int h = 0; h < H; ++h)
#pragma unroll
        for (int u = 0; u < 2; ++u) s1[h] += (dA[h][u][0] + dA[h][u][1]) + (dA[h][u][2] + dA[h][u][3]);
    }
  }
  __syncthreads();
  constexpr int NV = H + H * H;
  double* red = reinterpret_cast<double*>(smem);
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const float v = warp_sum(s1[h]);
    if (lane == 0) red[h * WARPS + warp] = (double)v;
#pragma unroll
    for (int gg = 0; gg < H; ++gg) {
      const float w = warp_sum(x[h][gg]);
      if (lane == 0) red[(H + h * H + gg) * WARPS + warp] = (double)w;
    }
  }
  __syncthreads();
  if (tid < NV) {
    double v = 0.0;
    for (int w = 0; w < WARPS; ++w) v += red[tid * WARPS + w];
    atomicAdd(g.red + tid, v);
  }
}

template <int H, int HD, bool TRAIN>
__global__ void __launch_bounds__(WARPS * 32, 1)
stream_bwd_ds_kernel(const BwdArgs g) {
  using P = BwdPlan<H, HD>;
  constexpr int D = P::D, QP = P::QP, KS = P::KS;
  constexpr size_t stage_bytes = P::v_bytes > P::kt_bytes ? P::v_bytes : P::kt_bytes;
  extern __shared__ __align__(128) unsigned char smem[];
  float* Gs = reinterpret_cast<float*>(smem);
  unsigned char* St = reinterpret_cast<unsigned char*>(Gs + WARPS * 16 * QP);       // 2 stages: fp32 value rows (sweep 1) / bf16 k^T rows (sweep 2)
  float* We = reinterpret_cast<float*>(St + 2 * stage_bytes);   // e[h][g] (H*H), Wb[g][h] (H*H, mix-back, transposed), kh[H], c0[H]
  float* Wb = We + H * H;
  float* Kh = Wb + H * H;
  float* C0 = Kh + H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int N = g.N, b = blockIdx.y, row0 = blockIdx.x * (WARPS * 16);
  const float* __restrict__ gb = g.dO + (size_t)b * N * D;
  const float* __restrict__ vb = g.v + (size_t)b * N * D;
  const __nv_bfloat16* __restrict__ ktb = g.kt + (size_t)b * D * g.ldn;
  const int ntiles = (N + KT - 1) / KT;
  auto load_v = [&](int tile, int stage) {
    float* dst = reinterpret_cast<float*>(St + stage * stage_bytes);
    const int key0 = tile * KT;
    for (int c = tid; c < KT * (D / 4); c += WARPS * 32) {
      const int kr = c / (D / 4), q4 = c - kr * (D / 4);
      const bool ok = key0 + kr < N;
      cp_async16(dst + key_row(kr) * QP + 4 * q4, ok ? vb + (size_t)(key0 + kr) * D + 4 * q4 : vb, ok);
    }
  };
  auto load_kt = [&](int tile, int stage) {
    unsigned char* dst = St + stage * stage_bytes;
    const int key0 = tile * KT;
    for (int c = tid; c < D * (KT / 8); c += WARPS * 32) {
      const int row = c / (KT / 8), q8 = c - row * (KT / 8);
      const bool ok = key0 + 8 * q8 < N;
      cp_async16(dst + row * VP + 16 * q8, ok ? (const void*)(ktb + (size_t)row * g.ldn + key0 + 8 * q8) : (const void*)ktb, ok);
    }
  };
  auto round_v = [&](int stage) {
    float* dst = reinterpret_cast<float*>(St + stage * stage_bytes);
    for (int c = tid; c < KT * (D / 4); c += WARPS * 32) {
      const int kr = c / (D / 4), q4 = c - kr * (D / 4);
      round_tf32_chunk(dst + key_row(kr) * QP + 4 * q4);
    }
  };
  for (int c = tid; c < WARPS * 16 * (D / 4); c += WARPS * 32) {
    const int r = c / (D / 4), q4 = c - r * (D / 4);
    const bool ok = row0 + r < N;
    cp_async16(Gs + r * QP + 4 * q4, ok ? gb + (size_t)(row0 + r) * D + 4 * q4 : gb, ok);
  }
  cp_async_commit();
  load_v(0, 0);
  cp_async_commit();
  cp_async_wait<1>();
  for (int c = tid; c < WARPS * 16 * (D / 4); c += WARPS * 32) {
    const int r = c / (D / 4), q4 = c - r * (D / 4);
    round_tf32_chunk(Gs + r * QP + 4 * q4);
  }
  for (int r = tid; r < WARPS * 16; r += WARPS * 32) *reinterpret_cast<float4*>(Gs + r * QP + D) = make_float4(0.f, 0.f, 0.f, 0.f);
  // folded coefficients (one thread per (h, g)):
  //   dM_h = kh_h dA_h + c0_h + sum_g e_hg pk_g          pk_g = keep ? p_g : 0
  //   kh = gamma invstd; e_hg = -kh m2 invstd W_hg dscale; c0_h = -kh (m1 + m2 invstd (b_h - mean_h))     (train)
  //   x_g = sum_h Wb[g][h] dM_h = dPd_g / (1-p),  Wb[g][h] = W[h][g] dscale
  for (int i = tid; i < H * H; i += WARPS * 32) {
    const int h = i / H, gg = i - h * H;
    const float invstd = g.saved[H + h], kh = g.gamma[h] * invstd;
    const float m2 = TRAIN ? g.coef[H + h] : 0.f;
    We[h * H + gg] = -kh * m2 * invstd * g.W[h * H + gg] * g.dscale;
    Wb[gg * H + h] = g.W[h * H + gg] * g.dscale;
    if (gg == 0) {
      Kh[h] = kh;
      C0[h] = TRAIN ? -kh * (g.coef[h] + m2 * invstd * (g.bconv[h] - g.saved[h])) : 0.f;
    }
  }
  const bool active = row0 + warp * 16 < N;
  const float* Gw = Gs + warp * 16 * QP;
  const int rowa = row0 + warp * 16 + gid;
  const uint32_t nn4 = (uint32_t)(((size_t)N * N) >> 2);
  const uint32_t ctr_r0 = (uint32_t)((((size_t)b * H * N + rowa) * N) >> 2) + tig;
  const uint32_t ctr_r1 = ctr_r0 + 8u * (uint32_t)(N >> 2);
  float delta[H][2];
#pragma unroll
  for (int h = 0; h < H; ++h) { delta[h][0] = 0.f; delta[h][1] = 0.f; }

  // ---------------------------------------------------------------- sweep 1: dP (bf16 -> ds buffer) and the row dots
  for (int t = 0; t < ntiles; ++t) {
    cp_async_wait<0>();
    round_v(t & 1);
    __syncthreads();
    // the first k^T tile of sweep 2 is prefetched behind the last value tile
    if (t + 1 < ntiles) load_v(t + 1, (t + 1) & 1); else load_kt(0, (t + 1) & 1);
    cp_async_commit();
    if (!active) continue;
    const float* Vst = reinterpret_cast<const float*>(St + (t & 1) * stage_bytes);
    const int steps = min(KT / 16, (N - t * KT) / 16);
    for (int st = 0; st < steps; ++st) {
      const int step = t * (KT / 16) + st;
      uint2 pw[H][2];
#pragma unroll
      for (int h = 0; h < H; ++h)
#pragma unroll
        for (int r = 0; r < 2; ++r)
          pw[h][r] = __ldg(reinterpret_cast<const uint2*>(g.pc + (((size_t)b * H + h) * N + rowa + 8 * r) * N + step * 16 + 4 * tig));
      const uint2 bits = step_keep_bits<H>(g, b, rowa, lane, step, ctr_r0, ctr_r1, nn4);
      float dM[H][2][4];
      scores_step<H, HD>(dM, Gw, Vst + st * 16 * QP, lane);           // dA
      float pk[H][2][4];                                              // keep ? p : 0
#pragma unroll
      for (int gg = 0; gg < H; ++gg)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t w = (r ? bits.y : bits.x) >> (4 * gg);
          pk[gg][0][2 * r] = (w & 1u) ? bf16_lo(pw[gg][r].x) + g.cN : 0.f;
          pk[gg][0][2 * r + 1] = (w & 2u) ? bf16_hi(pw[gg][r].x) + g.cN : 0.f;
          pk[gg][1][2 * r] = (w & 4u) ? bf16_lo(pw[gg][r].y) + g.cN : 0.f;
          pk[gg][1][2 * r + 1] = (w & 8u) ? bf16_hi(pw[gg][r].y) + g.cN : 0.f;
        }
      // dM_h = kh dA_h + c0_h + sum_g e_hg pk_g   (in place over dA)
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float kh = Kh[h], c0 = C0[h];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int e = 0; e < 4; ++e) dM[h][u][e] = fmaf(kh, dM[h][u][e], c0);
        if (TRAIN) {
#pragma unroll
          for (int q4 = 0; q4 < H / 4; ++q4) {
            const float4 w = *reinterpret_cast<const float4*>(We + h * H + 4 * q4);
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int gi = 0; gi < 4; ++gi)
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                ffma2(dM[h][u][0], dM[h][u][1], pk[4 * q4 + gi][u][0], pk[4 * q4 + gi][u][1], wv[gi]);
                ffma2(dM[h][u][2], dM[h][u][3], pk[4 * q4 + gi][u][2], pk[4 * q4 + gi][u][3], wv[gi]);
              }
          }
        }
      }
      // x_g = sum_h Wb[g][h] dM_h;  dP_g = keep ? x_g : 0 (bf16 -> ds);  delta_g += x_g pk_g
#pragma unroll
      for (int gg = 0; gg < H; ++gg) {
        float xg[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) { xg[u][0] = 0.f; xg[u][1] = 0.f; xg[u][2] = 0.f; xg[u][3] = 0.f; }
#pragma unroll
        for (int q4 = 0; q4 < H / 4; ++q4) {
          const float4 w = *reinterpret_cast<const float4*>(Wb + gg * H + 4 * q4);
          const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int hi = 0; hi < 4; ++hi)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              ffma2(xg[u][0], xg[u][1], dM[4 * q4 + hi][u][0], dM[4 * q4 + hi][u][1], wv[hi]);
              ffma2(xg[u][2], xg[u][3], dM[4 * q4 + hi][u][2], dM[4 * q4 + hi][u][3], wv[hi]);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t w = (r ? bits.y : bits.x) >> (4 * gg);
          delta[gg][r] = fmaf(xg[0][2 * r], pk[gg][0][2 * r], delta[gg][r]);
          delta[gg][r] = fmaf(xg[0][2 * r + 1], pk[gg][0][2 * r + 1], delta[gg][r]);
          delta[gg][r] = fmaf(xg[1][2 * r], pk[gg][1][2 * r], delta[gg][r]);
          delta[gg][r] = fmaf(xg[1][2 * r + 1], pk[gg][1][2 * r + 1], delta[gg][r]);
          uint2 o;
          o.x = pack_bf16((w & 1u) ? xg[0][2 * r] : 0.f, (w & 2u) ? xg[0][2 * r + 1] : 0.f);
          o.y = pack_bf16((w & 4u) ? xg[1][2 * r] : 0.f, (w & 8u) ? xg[1][2 * r + 1] : 0.f);
          *reinterpret_cast<uint2*>(g.ds + (((size_t)b * H + gg) * N + rowa + 8 * r) * N + step * 16 + 4 * tig) = o;
        }
      }
    }
  }
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float d = delta[h][r];
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      delta[h][r] = d;
    }

  // ---------------------------------------------------------------- sweep 2: dS = scale P (dP - delta) -> map, dq += dS k
  float dq[H][KS][4];
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int nt = 0; nt < KS; ++nt) { dq[h][nt][0] = 0.f; dq[h][nt][1] = 0.f; dq[h][nt][2] = 0.f; dq[h][nt][3] = 0.f; }
  for (int t = 0; t < ntiles; ++t) {
    const int item = ntiles + t;
    cp_async_wait<0>();
    __syncthreads();
    if (t + 1 < ntiles) { load_kt(t + 1, (item + 1) & 1); cp_async_commit(); }
    if (!active) continue;
    const unsigned char* Kst = St + (item & 1) * stage_bytes;
    const int steps = min(KT / 16, (N - t * KT) / 16);
    for (int st = 0; st < steps; ++st) {
      const int step = t * (KT / 16) + st;
      const unsigned char* krow = Kst + gid * VP + st * 32 + tig * 8;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        uint32_t af[4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const size_t off = (((size_t)b * H + h) * N + rowa + 8 * r) * N + step * 16 + 4 * tig;
          const uint2 pw = __ldg(reinterpret_cast<const uint2*>(g.pc + off));
          const uint2 dw = *reinterpret_cast<const uint2*>(g.ds + off);          // written by this very lane in sweep 1
          const float d = delta[h][r];
          const float s0 = g.scale * (bf16_lo(pw.x) + g.cN) * (bf16_lo(dw.x) - d), s1 = g.scale * (bf16_hi(pw.x) + g.cN) * (bf16_hi(dw.x) - d);
          const float s2 = g.scale * (bf16_lo(pw.y) + g.cN) * (bf16_lo(dw.y) - d), s3 = g.scale * (bf16_hi(pw.y) + g.cN) * (bf16_hi(dw.y) - d);
          const uint32_t lo = pack_bf16(s0, s1), hi = pack_bf16(s2, s3);
          *reinterpret_cast<uint2*>(g.ds + off) = make_uint2(lo, hi);
          af[r] = lo; af[2 + r] = hi;            // reg0/1: keys 4tig, +1 of rows gid / gid+8; reg2/3: keys 4tig+2, +3
        }
#pragma unroll
        for (int nt = 0; nt < KS; ++nt) {
          const uint2 bv = *reinterpret_cast<const uint2*>(krow + (h * HD + nt * 8) * VP);
          mma_bf16(dq[h][nt], af, bv.x, bv.y);
        }
      }
    }
  }
  if (active) {
    float* ob = g.dq + ((size_t)b * N + rowa) * D;
#pragma unroll
    for (int h = 0; h < H; ++h)
#pragma unroll
      for (int nt = 0; nt < KS; ++nt) {
        const int e = nt * 8 + 2 * tig;
        if (e < HD) {
          *reinterpret_cast<float2*>(ob + h * HD + e) = make_float2(dq[h][nt][0], dq[h][nt][1]);
          *reinterpret_cast<float2*>(ob + (size_t)8 * D + h * HD + e) = make_float2(dq[h][nt][2], dq[h][nt][3]);
        }
      }
  }
}

template <int H, int HD>
static int launch_bwd_reduce(const BwdArgs& a, int B, cudaStream_t st, const char* fn) {
  const size_t smem = BwdPlan<H, HD>::reduce_total;
  static uint64_t seen = 0;
  if (first_use_on_device(seen))
    cudaFuncSetAttribute(stream_bwd_reduce_kernel<H, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const dim3 grid((unsigned)cdiv(a.N, WARPS * 16), (unsigned)B);
  stream_bwd_reduce_kernel<H, HD><<<grid, WARPS * 32, smem, st>>>(a);
  return check_launch(fn);
}
template <int H, int HD, bool TRAIN>
static int launch_bwd_ds(const BwdArgs& a, int B, cudaStream_t st, const char* fn) {
  const size_t smem = BwdPlan<H, HD>::ds_total;
  static uint64_t seen = 0;
  if (first_use_on_device(seen))
    cudaFuncSetAttribute(stream_bwd_ds_kernel<H, HD, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const dim3 grid((unsigned)cdiv(a.N, WARPS * 16), (unsigned)B);
  stream_bwd_ds_kernel<H, HD, TRAIN><<<grid, WARPS * 32, smem, st>>>(a);
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

static void fill_bwd_args(vu::rs::BwdArgs& a, const void* pc, const void* mask, const float* dO, const float* v, int h, int N,
                          int hd, float drop_p, uint64_t seed, uint32_t stream_id) {
  using namespace vu;
  a = rs::BwdArgs();
  a.pc = (const __nv_bfloat16*)pc; a.mask = (const uint2*)mask; a.dO = dO; a.v = v;
  a.N = N; a.D = h * hd; a.scale = 1.0f / sqrtf((float)hd);
  a.thresh = drop_p > 0.f ? drop_threshold(drop_p) : 0u;
  a.dscale = drop_p > 0.f ? drop_keep_scale(drop_p) : 1.0f;
  a.key = Philox::key(seed, stream_id);
  a.cN = 1.0f / (float)N;
}

// red[h + h*h] (double, caller zeroes) += { sum dA_h , sum dA_h (Pd_g - 1/N) } with dA = dO v^T formed on the fly
// (the streamed twin of vu_reattn_mix_reduce / vu_reattn_bwd_reduce: no dA map is read or written).
extern "C" int vu_reattn_stream_bwd_reduce(const void* pc, const void* mask, const float* dO, const float* v, double* red,
                                           int B, int h, int N, int hd, float drop_p, uint64_t seed, uint32_t stream_id,
                                           void* stream) {
  using namespace vu;
  const char* fn = "vu_reattn_stream_bwd_reduce";
  VU_REQUIRE(pc && dO && v && red && B > 0 && B <= 65535, fn, "null pointer or bad batch");
  VU_REQUIRE(rs::supported(h, hd, N), fn, "unsupported (heads, head_dim, tokens): see vu_reattn_stream_supported");
  VU_REQUIRE(drop_p >= 0.f && drop_p < 1.f, fn, "drop_p must be in [0,1)");
  VU_REQUIRE((uintptr_t)pc % 8 == 0 && (uintptr_t)mask % 8 == 0 && (uintptr_t)dO % 16 == 0 && (uintptr_t)v % 16 == 0, fn, "misaligned pointer");
  rs::BwdArgs a; fill_bwd_args(a, pc, mask, dO, v, h, N, hd, drop_p, seed, stream_id);
  a.red = red;
  cudaStream_t st = as_stream(stream);
  VU_RS_DISPATCH(h, hd, return (rs::launch_bwd_reduce<HH, HDD>(a, B, st, fn)));
  return VU_OK;
}

// dS (bf16 map, (B,h,N,N)) and dq (B,N,h*hd) from the centred probabilities, dO and v: the streamed twin of
// dA = dO v^T  +  vu_reattn_bwd_rows  +  dq = dS k.   kt: per-head transposed bf16 keys (vu_heads_transpose_bf16).
// coef: BatchNorm-backward means from vu_reattn_bwd_params (train); saved: {mean, invstd} from vu_reattn_bn_finalize.
extern "C" int vu_reattn_stream_bwd_ds(const void* pc, const void* mask, const float* dO, const float* v, const void* kt,
                                       void* ds, float* dq, const float* W, const float* bconv, const float* gamma,
                                       const float* saved, const float* coef, int train, int B, int h, int N, int hd,
                                       int ldn, float drop_p, uint64_t seed, uint32_t stream_id, void* stream) {
  using namespace vu;
  const char* fn = "vu_reattn_stream_bwd_ds";
  VU_REQUIRE(pc && dO && v && kt && ds && dq && W && bconv && gamma && saved && B > 0 && B <= 65535, fn, "null pointer or bad batch");
  VU_REQUIRE(!train || coef, fn, "train mode needs the BN-backward coefficients");
  VU_REQUIRE(rs::supported(h, hd, N), fn, "unsupported (heads, head_dim, tokens): see vu_reattn_stream_supported");
  VU_REQUIRE(drop_p >= 0.f && drop_p < 1.f, fn, "drop_p must be in [0,1)");
  VU_REQUIRE(ldn % 8 == 0 && ldn >= N, fn, "kt needs ldn % 8 == 0");
  VU_REQUIRE((uintptr_t)pc % 8 == 0 && (uintptr_t)mask % 8 == 0 && (uintptr_t)dO % 16 == 0 && (uintptr_t)v % 16 == 0 &&
             (uintptr_t)kt % 16 == 0 && (uintptr_t)ds % 8 == 0 && (uintptr_t)dq % 8 == 0, fn, "misaligned pointer");
  rs::BwdArgs a; fill_bwd_args(a, pc, mask, dO, v, h, N, hd, drop_p, seed, stream_id);
  a.kt = (const __nv_bfloat16*)kt; a.ds = (__nv_bfloat16*)ds; a.dq = dq; a.ldn = ldn; a.train = train;
  a.W = W; a.bconv = bconv; a.gamma = gamma; a.saved = saved; a.coef = coef;
  cudaStream_t st = as_stream(stream);
  if (train) { VU_RS_DISPATCH(h, hd, return (rs::launch_bwd_ds<HH, HDD, true>(a, B, st, fn))); }
  else { VU_RS_DISPATCH(h, hd, return (rs::launch_bwd_ds<HH, HDD, false>(a, B, st, fn))); }
  return VU_OK;
}
