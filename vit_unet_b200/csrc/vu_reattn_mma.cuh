// Re-Attention map kernels, 8-head tensor-core formulation (TF32 path with bf16 mixed / gradient maps).
//
// The head mixing  y_h = sum_g F[h][g] x_g  at one map position is an 8x8 matrix applied to an 8-vector; at 614k
// positions per image and five sweeps per training step it is the instruction-issue bottleneck of the CUDA-core
// kernels in vu_reattn.cu (64 FFMA per position and ~160 live registers per thread, 12-25% occupancy).  Here a warp
// treats 32 consecutive keys x 8 heads as two m16n8k8 TF32 warp MMAs (rows = positions, K = source head, N = target
// head): a lane then only ever holds TWO heads of one key quad ("pair layout"), so the elementwise work (dropout
// mask, centring, BatchNorm-backward terms, bf16 packing) runs on 8 values per lane and the register footprint
// drops ~4x.  The MMA output fragment lands in the very same layout (columns are permuted by sigma() when the
// weight fragment is built), so mixing can be chained (rows kernel: M = W Pd, then dPd = W^T dM) without shuffles.
// Reductions over positions (G' = sum Pdc Pdc^T, X' = sum dA Pdc^T) are MMAs with K = positions in the "stats
// layout" (lane = one head, two key quads).  Inputs are centred (Pd - 1/N) before the TF32 rounding, so the rounding
// error is relative to the deviation from the row mean -- the quantity BatchNorm normalises.
//
// Preconditions (checked by the dispatchers): h == 8, ld == N (no pad columns), N % 8 == 0.
#pragma once

namespace vu {
namespace mma {

constexpr int H = 8;

__device__ __forceinline__ uint32_t tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ uint32_t bits(float x) { return __float_as_uint(x); }
// round-to-nearest (half away from zero) on the 13 mantissa bits the tensor core drops: one integer add.  Used for
// the statistics, where a truncation bias (-0.7 * 2^-10 on every second moment) would shift the BatchNorm variance.
__device__ __forceinline__ uint32_t rnd(float x) { return __float_as_uint(x) + 0x1000u; }
__device__ __forceinline__ void mma8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                     uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// output column n of the MMA carries head sigma(n): lane (e, k4) reads heads k4 and k4+4 from columns 2*k4, 2*k4+1
__device__ __forceinline__ int sigma(int n) { return (n >> 1) + ((n & 1) << 2); }

// Pair layout: x0 / x1 = key quad of heads k4 / k4+4.  y_{k4}, y_{k4+4} = sum_g Bm[g][.] x_g, with (b0, b1) the
// lane's fragment of the 8x8 weight matrix (see frag_fwd / frag_bwd).
__device__ __forceinline__ void mix_pair(const float4& x0, const float4& x1, uint32_t b0, uint32_t b1,
                                         float4& y0, float4& y1, float i0 = 0.f, float i1 = 0.f) {
  // data operands go in as raw fp32 bits: the tensor core reads the top 19 bits (truncation, 2^-11 mean relative
  // shrink of CENTRED values -- far below the bf16 rounding of the stored result); only the weights are rounded (rna).
  // (i0, i1): per-head constants of the two output heads, added by starting the accumulators from them
  float c[4] = {i0, i1, i0, i1}, d[4] = {i0, i1, i0, i1};
  mma8(c, bits(x0.x), bits(x0.y), bits(x1.x), bits(x1.y), b0, b1);    // rows e / e+8 = components 0 / 1
  mma8(d, bits(x0.z), bits(x0.w), bits(x1.z), bits(x1.w), b0, b1);    // rows e / e+8 = components 2 / 3
  y0 = make_float4(c[0], c[2], d[0], d[2]);
  y1 = make_float4(c[1], c[3], d[1], d[3]);
}
// y_h = sum_g F[h][g] x_g  (F row-major h x g)
__device__ __forceinline__ void frag_fwd(const float* __restrict__ F, int e, int k4, uint32_t& b0, uint32_t& b1) {
  b0 = tf32(F[sigma(e) * H + k4]); b1 = tf32(F[sigma(e) * H + k4 + 4]);
}
// y_g = sum_h F[h][g] x_h
__device__ __forceinline__ void frag_bwd(const float* __restrict__ F, int e, int k4, uint32_t& b0, uint32_t& b1) {
  b0 = tf32(F[k4 * H + sigma(e)]); b1 = tf32(F[(k4 + 4) * H + sigma(e)]);
}

// dropout of one quad in place (ctr = low 32 bits of element index / 4); returns the keep bits (bit t = component t)
__device__ __forceinline__ uint32_t drop_quad(float4& v, uint32_t ctr, const QuadCtx& q) {
  if (!q.thresh) return 0xFu;
  const uint4 rr = Philox::gen_k(q.key, ctr);
  const uint32_t m = (rr.x >= q.thresh ? 1u : 0u) | (rr.y >= q.thresh ? 2u : 0u) | (rr.z >= q.thresh ? 4u : 0u) |
                     (rr.w >= q.thresh ? 8u : 0u);
  v.x = (m & 1u) ? v.x * q.dscale : 0.f; v.y = (m & 2u) ? v.y * q.dscale : 0.f;
  v.z = (m & 4u) ? v.z * q.dscale : 0.f; v.w = (m & 8u) ? v.w * q.dscale : 0.f;
  return m;
}
// Centred dropped probabilities (Pd - c) of one quad, the operand of the head mixing and of every moment.  fp32 map: v holds
// P on entry.  Centred bf16 map: v holds d = P - c, and  keep ? (d + c) * s - c : -c  collapses into ONE FFMA + select per
// element (kd = c * s - c) instead of add / multiply / select / subtract; without dropout d already is the result.
// Returns the keep bits.  The last argument only selects the overload (the map's element type).
__device__ __forceinline__ uint32_t pdc_quad(float4& v, uint32_t ctr, const QuadCtx& q, const float*) {
  const uint32_t m = drop_quad(v, ctr, q);
  v.x -= q.c; v.y -= q.c; v.z -= q.c; v.w -= q.c;
  return m;
}
__device__ __forceinline__ uint32_t pdc_quad(float4& v, uint32_t ctr, const QuadCtx& q, const __nv_bfloat16*) {
  if (!q.thresh) return 0xFu;
  const uint4 rr = Philox::gen_k(q.key, ctr);
  const float kd = fmaf(q.c, q.dscale, -q.c), nc = -q.c;
  const bool k0 = rr.x >= q.thresh, k1 = rr.y >= q.thresh, k2 = rr.z >= q.thresh, k3 = rr.w >= q.thresh;
  v.x = k0 ? fmaf(v.x, q.dscale, kd) : nc; v.y = k1 ? fmaf(v.y, q.dscale, kd) : nc;
  v.z = k2 ? fmaf(v.z, q.dscale, kd) : nc; v.w = k3 ? fmaf(v.w, q.dscale, kd) : nc;
  return (k0 ? 1u : 0u) | (k1 ? 2u : 0u) | (k2 ? 4u : 0u) | (k3 ? 8u : 0u);
}
// 2^x for x <= 0 as one MUFU (ex2() adds a denormal-range rescale around it; a softmax term below 2^-126 may flush)
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void sub4(float4& v, float c) { v.x -= c; v.y -= c; v.z -= c; v.w -= c; }
__device__ __forceinline__ void add4(float4& v, float c) { v.x += c; v.y += c; v.z += c; v.w += c; }
__device__ __forceinline__ void mul4(float4& v, float c) { v.x *= c; v.y *= c; v.z *= c; v.w *= c; }
__device__ __forceinline__ float hsum4(const float4& v) { return (v.x + v.y) + (v.z + v.w); }
__device__ __forceinline__ float hmax4(const float4& v) { return fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)); }
__device__ __forceinline__ float comp(const float4& v, int u) { return u == 0 ? v.x : (u == 1 ? v.y : (u == 2 ? v.z : v.w)); }
__device__ __forceinline__ float4 ldq(const float* p) { return *reinterpret_cast<const float4*>(p); }
constexpr float4 kZero4 = {0.f, 0.f, 0.f, 0.f};
// The probabilities are read either as fp32 P or as CENTRED bf16 (Pc = P - 1/N, see softmax_stats_*): the bf16
// rounding is then relative to the deviation from the uniform row, which is what the head mixing + BatchNorm see.
__device__ __forceinline__ float4 ldp(const float* p, float) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldp(const __nv_bfloat16* p, float c) { float4 v = map_ld(p); add4(v, c); return v; }
// what pdc_quad starts from: P (fp32 map) or d = P - c (centred bf16 map)
__device__ __forceinline__ float4 ldraw(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldraw(const __nv_bfloat16* p) { return map_ld(p); }

__device__ __forceinline__ uint2 pack_bf16x4(const float4& v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 u; u.x = *reinterpret_cast<uint32_t*>(&lo); u.y = *reinterpret_cast<uint32_t*>(&hi);
  return u;
}
__device__ __forceinline__ float4 unpack_bf16x4(const uint2& u) {
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                     __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}
// softmax pass 3 of one quad: p = x * inv is written (fp32 over S, or centred bf16 into Pc) and turned into the centred
// dropped value the moments are taken of -- for the bf16 map from the ROUNDED d, so the moments describe exactly the map the
// later kernels read back: x * inv - c is one FFMA, pack, store, unpack (two shifts / masks), one FFMA + select.
__device__ __forceinline__ float4 emit_pdc(float* __restrict__ S, __nv_bfloat16* __restrict__ Pc, int off, const float4& x,
                                           float inv, uint32_t ctr, const QuadCtx& q) {
  float4 v;
  if (Pc) {
    v = make_float4(fmaf(x.x, inv, -q.c), fmaf(x.y, inv, -q.c), fmaf(x.z, inv, -q.c), fmaf(x.w, inv, -q.c));
    const uint2 u = pack_bf16x4(v);
    *reinterpret_cast<uint2*>(Pc + off) = u;
    v = unpack_bf16x4(u);
    pdc_quad(v, ctr, q, static_cast<const __nv_bfloat16*>(nullptr));
  } else {
    v = make_float4(x.x * inv, x.y * inv, x.z * inv, x.w * inv);
    *reinterpret_cast<float4*>(S + off) = v;
    pdc_quad(v, ctr, q, static_cast<const float*>(nullptr));
  }
  return v;
}
// G'/X' style accumulation in the stats layout: C[e][2k4..2k4+1] += sum_keys A_e[key] * B_n[key] over this lane's 8 keys
__device__ __forceinline__ void mma_keys(float (&c)[4], const float4& a_lo, const float4& a_hi, const float4& b_lo,
                                         const float4& b_hi) {
#pragma unroll
  for (int u = 0; u < 4; ++u)
    mma8(c, rnd(comp(a_lo, u)), 0u, rnd(comp(a_hi, u)), 0u, rnd(comp(b_lo, u)), rnd(comp(b_hi, u)));
}

// sum over the 8 warps of the 8 + 64 statistics accumulators (s[head e] after the 4-lane reduction, M[e][2k4..+1] on
// every lane), one double atomicAdd per entry.  `part` is [8][72] shared floats; warps >= 8 only join the barrier.
__device__ __forceinline__ void reduce_stats(float (*part)[H + H * H], bool consumer, float s, float m0, float m1,
                                             double* __restrict__ out, int nwarps = 8) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, e = lane >> 2, k4 = lane & 3;
  if (consumer) {
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (k4 == 0) part[warp][e] = s;
    part[warp][H + e * H + 2 * k4] = m0;
    part[warp][H + e * H + 2 * k4 + 1] = m1;
  }
  __syncthreads();
  if (threadIdx.x < H + H * H) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < nwarps; ++w) t += (double)part[w][threadIdx.x];
    atomicAdd(out + threadIdx.x, t);
  }
}

// All index arithmetic below is 32-bit: one image of maps has 8*N*N < 2^31 elements (N <= 8192, checked by the
// dispatcher); the 64-bit image base is added to the pointers once per row / CTA.

// ------------------------------------------------------------------ forward: softmax + centred moments (short rows)
// One warp per row (b, i), all 8 heads at once in the stats layout (lane = head e, key quads k4 and k4+4 of each
// 32-key tile).  Sweep A: online (max, sum) of exp2; sweep B (re-read from L1): write P, accumulate
// s'_g = sum (Pd_g - c) and G' = sum (Pd - c)(Pd - c)^T (4 MMAs per tile, K = keys).  Used for N <= 256.
__global__ void __launch_bounds__(256)
softmax_stats_mma_kernel(float* __restrict__ S, __nv_bfloat16* __restrict__ Pc, int B, int N, float scale, QuadCtx q,
                         double* __restrict__ sums) {
  __shared__ float part[8][H + H * H];
  const int lane = threadIdx.x & 31, e = lane >> 2, k4 = lane & 3;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const float sl2 = scale * 1.4426950408889634f;
  const int ld4 = N >> 2, ntiles = (ld4 + 7) >> 3, hs = N * N, rows = B * N;
  float cg[4] = {0.f, 0.f, 0.f, 0.f}, s = 0.f;
  for (int r = wid; r < rows; r += nw) {
    const int b = r / N, i = r - b * N;
    const int64_t base = (int64_t)b * hs * H;
    const uint32_t ctr0 = (uint32_t)((uint64_t)base >> 2);
    const int roff = e * hs + i * N;
    float* Sb = S + base; __nv_bfloat16* Pb = Pc ? Pc + base : nullptr;
    float m = -INFINITY, l = 0.f;
    for (int t = 0; t < ntiles; ++t) {
      const int qa = t * 8 + k4, qb = qa + 4;
      float4 xa = qa < ld4 ? ldq(Sb + roff + 4 * qa) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      float4 xb = qb < ld4 ? ldq(Sb + roff + 4 * qb) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      mul4(xa, sl2); mul4(xb, sl2);
      const float mn = fmaxf(m, fmaxf(hmax4(xa), hmax4(xb)));
      if (mn > -INFINITY) {
        l = l * ex2(m - mn) + ((ex2(xa.x - mn) + ex2(xa.y - mn)) + (ex2(xa.z - mn) + ex2(xa.w - mn))) +
            ((ex2(xb.x - mn) + ex2(xb.y - mn)) + (ex2(xb.z - mn) + ex2(xb.w - mn)));
        m = mn;
      }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m, o), lo = __shfl_xor_sync(0xffffffffu, l, o);
      const float mn = fmaxf(m, mo);
      l = (m > -INFINITY ? l * ex2(m - mn) : 0.f) + (mo > -INFINITY ? lo * ex2(mo - mn) : 0.f);
      m = mn;
    }
    const float inv = 1.0f / l;
    for (int t = 0; t < ntiles; ++t) {
      const int qa = t * 8 + k4, qb = qa + 4;
      float4 pa = kZero4, pb = kZero4;
      if (qa < ld4) {
        const float4 x = ldq(Sb + roff + 4 * qa);
        const float4 ex = make_float4(ex2(fmaf(x.x, sl2, -m)), ex2(fmaf(x.y, sl2, -m)), ex2(fmaf(x.z, sl2, -m)), ex2(fmaf(x.w, sl2, -m)));
        pa = emit_pdc(Sb, Pb, roff + 4 * qa, ex, inv, ctr0 + ((uint32_t)roff >> 2) + qa, q);
      }
      if (qb < ld4) {
        const float4 x = ldq(Sb + roff + 4 * qb);
        const float4 ex = make_float4(ex2(fmaf(x.x, sl2, -m)), ex2(fmaf(x.y, sl2, -m)), ex2(fmaf(x.z, sl2, -m)), ex2(fmaf(x.w, sl2, -m)));
        pb = emit_pdc(Sb, Pb, roff + 4 * qb, ex, inv, ctr0 + ((uint32_t)roff >> 2) + qb, q);
      }
      s += hsum4(pa) + hsum4(pb);
      mma_keys(cg, pa, pb, pa, pb);
    }
  }
  reduce_stats(part, true, s, cg[0], cg[1], sums);
}

// ------------------------------------------------------------------ forward: A = fold . (Pd - c) + shift'
// grid (x, B): flat tiles of 32 consecutive positions of image blockIdx.y (rows are contiguous: ld == N), pair layout.
// Whole tiles (all but possibly the last one of an image) take a predicate-free path whose four pointers simply walk the
// map -- the general path's per-access 64-bit address arithmetic and zero fills were a third of the instructions issued.
template <typename PT, typename MT>
__device__ __forceinline__ void mix_tile(float4& x0, float4& x1, uint32_t ctr, int hs, const QuadCtx& q, uint32_t b0, uint32_t b1,
                                         float sh0, float sh1, MT* a0, MT* a1, const PT* tag) {
  pdc_quad(x0, ctr, q, tag); pdc_quad(x1, ctr + hs, q, tag);
  float4 y0, y1; mix_pair(x0, x1, b0, b1, y0, y1, sh0, sh1);
  map_st(a0, y0); map_st(a1, y1);
}

template <typename PT, typename MT>
__global__ void __launch_bounds__(256)
reattn_mix_mma_kernel(const PT* __restrict__ P, MT* __restrict__ A, const float* __restrict__ fold,
                      int N, QuadCtx q) {
  const int lane = threadIdx.x & 31, e = lane >> 2, k4 = lane & 3;
  uint32_t b0, b1; frag_fwd(fold, e, k4, b0, b1);
  float sh0 = 0.f, sh1 = 0.f;
#pragma unroll
  for (int g = 0; g < H; ++g) { sh0 += fold[k4 * H + g]; sh1 += fold[(k4 + 4) * H + g]; }
  sh0 = fmaf(sh0, q.c, fold[H * H + k4]); sh1 = fmaf(sh1, q.c, fold[H * H + k4 + 4]);
  const int hs = N * N, quads = hs >> 2, tiles = (quads + 7) >> 3, full_tiles = quads >> 3;
  const int64_t base = (int64_t)blockIdx.y * hs * H;
  const uint32_t ctr0 = (uint32_t)((uint64_t)base >> 2);
  const PT* Pi = P + base; MT* Ai = A + base;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  // this lane's walk: heads k4 / k4 + 4, quad t * 8 + e of tile t; consecutive tiles of the warp are `step` elements apart
  const int lane_off = k4 * hs + (wid * 8 + e) * 4, step = 32 * nw;
  const PT* p0 = Pi + lane_off; const PT* p1 = p0 + 4 * hs;
  MT* a0 = Ai + lane_off; MT* a1 = a0 + 4 * hs;
  uint32_t ctr = ctr0 + ((uint32_t)lane_off >> 2);
  int t = wid;
  for (; t + nw < full_tiles; t += 2 * nw) {                  // two whole tiles per iteration, four loads in flight
    float4 x0 = ldraw(p0), x1 = ldraw(p1), z0 = ldraw(p0 + step), z1 = ldraw(p1 + step);
    mix_tile(x0, x1, ctr, hs, q, b0, b1, sh0, sh1, a0, a1, Pi);
    mix_tile(z0, z1, ctr + 8 * nw, hs, q, b0, b1, sh0, sh1, a0 + step, a1 + step, Pi);
    p0 += 2 * step; p1 += 2 * step; a0 += 2 * step; a1 += 2 * step; ctr += 16 * nw;
  }
  for (; t < tiles; t += nw) {                                // what is left: one tile at a time, predicated
    const int quad = t * 8 + e;
    const bool ok = quad < quads;
    const int off = k4 * hs + quad * 4;
    float4 x0 = ok ? ldraw(Pi + off) : kZero4, x1 = ok ? ldraw(Pi + off + 4 * hs) : kZero4;
    pdc_quad(x0, ctr0 + ((uint32_t)off >> 2), q, Pi); pdc_quad(x1, ctr0 + ((uint32_t)off >> 2) + hs, q, Pi);
    float4 y0, y1; mix_pair(x0, x1, b0, b1, y0, y1, sh0, sh1);
    if (ok) { map_st(Ai + off, y0); map_st(Ai + off + 4 * hs, y1); }
  }
}

// ------------------------------------------------------------------ backward pass 1: A = mix(P) recomputed + reductions
// red[h] += sum dA_h ;  red[H + h*H + g] += sum dA_h (Pd_g - c).  The mix (MIX: the forward map was not kept) runs in the
// pair layout, the reductions in the stats layout (second read of the same P tile hits L1).  grid (x, B).  Whole tiles take
// a predicate-free path with walking pointers, as in the forward kernel.
template <typename PT, typename MT, bool MIX>
__global__ void __launch_bounds__(256)
reattn_mix_reduce_mma_kernel(const PT* __restrict__ P, const MT* __restrict__ dA,
                             MT* __restrict__ A, const float* __restrict__ fold, int N, QuadCtx q,
                             double* __restrict__ out) {
  __shared__ float part[8][H + H * H];
  const int lane = threadIdx.x & 31, e = lane >> 2, k4 = lane & 3;
  uint32_t b0 = 0u, b1 = 0u;
  float sh0 = 0.f, sh1 = 0.f;
  if (MIX) {
    frag_fwd(fold, e, k4, b0, b1);
#pragma unroll
    for (int g = 0; g < H; ++g) { sh0 += fold[k4 * H + g]; sh1 += fold[(k4 + 4) * H + g]; }
    sh0 = fmaf(sh0, q.c, fold[H * H + k4]); sh1 = fmaf(sh1, q.c, fold[H * H + k4 + 4]);
  }
  const int hs = N * N, quads = hs >> 2, tiles = (quads + 7) >> 3, full_tiles = quads >> 3;
  const int64_t base = (int64_t)blockIdx.y * hs * H;
  const uint32_t ctr0 = (uint32_t)((uint64_t)base >> 2);
  const PT* Pi = P + base; const MT* Di = dA + base; MT* Ai = MIX ? A + base : nullptr;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  float cx[4] = {0.f, 0.f, 0.f, 0.f}, s1 = 0.f;
  // stats-layout walk: head e, quads t * 8 + k4 and + 4 of tile t
  const int lane_off = e * hs + (wid * 8 + k4) * 4, step = 32 * nw;
  const PT* pp = Pi + lane_off; const MT* dp = Di + lane_off;
  uint32_t ctr = ctr0 + ((uint32_t)lane_off >> 2);
  for (int t = wid; t < tiles; t += nw, pp += step, dp += step, ctr += 8 * nw) {
    if (MIX) {                                   // pair layout: mixed map
      const int quad = t * 8 + e;
      const bool ok = quad < quads;
      const int off = k4 * hs + quad * 4;
      float4 x0 = ok ? ldraw(Pi + off) : kZero4, x1 = ok ? ldraw(Pi + off + 4 * hs) : kZero4;
      pdc_quad(x0, ctr0 + ((uint32_t)off >> 2), q, Pi); pdc_quad(x1, ctr0 + ((uint32_t)off >> 2) + hs, q, Pi);
      float4 y0, y1; mix_pair(x0, x1, b0, b1, y0, y1, sh0, sh1);       // same accumulator start as the forward kernel: bit-equal maps
      if (ok) { map_st(Ai + off, y0); map_st(Ai + off + 4 * hs, y1); }
    }
    float4 pa, pb, da, db;
    if (t < full_tiles) {                        // warp-uniform: all eight quads of the tile exist
      pa = ldraw(pp); pb = ldraw(pp + 16); da = map_ld(dp); db = map_ld(dp + 16);
      pdc_quad(pa, ctr, q, Pi); pdc_quad(pb, ctr + 4, q, Pi);
    } else {
      const int qa = t * 8 + k4, qb = qa + 4;
      const bool va = qa < quads, vb = qb < quads;
      pa = va ? ldraw(pp) : kZero4; pb = vb ? ldraw(pp + 16) : kZero4;
      da = va ? map_ld(dp) : kZero4; db = vb ? map_ld(dp + 16) : kZero4;
      if (va) pdc_quad(pa, ctr, q, Pi);
      if (vb) pdc_quad(pb, ctr + 4, q, Pi);
    }
    s1 += hsum4(da) + hsum4(db);
    mma_keys(cx, da, db, pa, pb);
  }
  reduce_stats(part, true, s1, cx[0], cx[1], out);
}

// ------------------------------------------------------------------ backward pass 2: dA -> dS in place
//   dM_h  = k_h (dA_h - m1_h - Ahat_h m2_h)   (train)   |   k_h dA_h   (eval),   Ahat_h = (M_h - mean_h) invstd_h
//   dPd_g = sum_h W[h][g] dM_h ;  dP_g = keep_g dPd_g / (1-p) ;  r_g = sum_j dP_g P_g ;  dS_g = scale P_g (dP_g - r_g)
// Per-lane constants of the pair layout (heads k4 and k4 + 4).
struct RowsConst {
  uint32_t f0, f1, g0, g1;
  float offp[2], a1[2], a2[2];
};
// The per-head factor k_h = gamma_h invstd_h of dM and the softmax scale of dS are folded into the backward weight
// fragment: the second MMA directly yields  scale * dPd_g;  the keep scale 1/(1-p) of dP rides on the keep factors of rows_tile.
__device__ __forceinline__ RowsConst rows_const(const float* __restrict__ W, const float* __restrict__ bconv,
                                                const float* __restrict__ gamma, const float* __restrict__ saved,
                                                const float* __restrict__ coef, int train, const QuadCtx& q, float scale,
                                                int e, int k4) {
  RowsConst k;
  frag_fwd(W, e, k4, k.f0, k.f1);           // M_h   = sum_g W[h][g] Pd_g
  const float s0 = gamma[k4] * saved[H + k4] * scale, s1 = gamma[k4 + 4] * saved[H + k4 + 4] * scale;
  k.g0 = tf32(W[k4 * H + sigma(e)] * s0);   // scale dPd_g = sum_h (k_h scale W[h][g]) (dA_h - m1_h - Ahat_h m2_h)
  k.g1 = tf32(W[(k4 + 4) * H + sigma(e)] * s1);
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int h = k4 + 4 * v;
    float rs = 0.f;
#pragma unroll
    for (int g = 0; g < H; ++g) rs += W[h * H + g];
    k.offp[v] = bconv[h] - saved[h] + q.c * rs;           // M_h - mean_h = sum_g W_hg (Pd_g - c) + offp
    k.a1[v] = train ? coef[h] : 0.f;
    k.a2[v] = train ? saved[H + h] * coef[H + h] : 0.f;
  }
  return k;
}
// one tile of pass 1: (p0, p1) probabilities, (t0, t1) = dA in, dP out (before bf16 rounding)
__device__ __forceinline__ void rows_tile(const RowsConst& k, const QuadCtx& q, int train, uint32_t ctr, int hs,
                                          const float4& p0, const float4& p1, float4& t0, float4& t1) {
  // keep factors kf = keep ? 1/(1-p) : 0 as FLOATS: the centred dropped probabilities are one FFMA (P kf - c) and the
  // masking of dP one FMUL, with no mask word to pack and to test again (that cost ~6 instructions per element)
  float kf0[4], kf1[4];
  if (q.thresh) {
    const uint4 r0 = Philox::gen_k(q.key, ctr), r1 = Philox::gen_k(q.key, ctr + hs);
    kf0[0] = r0.x >= q.thresh ? q.dscale : 0.f; kf0[1] = r0.y >= q.thresh ? q.dscale : 0.f;
    kf0[2] = r0.z >= q.thresh ? q.dscale : 0.f; kf0[3] = r0.w >= q.thresh ? q.dscale : 0.f;
    kf1[0] = r1.x >= q.thresh ? q.dscale : 0.f; kf1[1] = r1.y >= q.thresh ? q.dscale : 0.f;
    kf1[2] = r1.z >= q.thresh ? q.dscale : 0.f; kf1[3] = r1.w >= q.thresh ? q.dscale : 0.f;
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) { kf0[u] = q.dscale; kf1[u] = q.dscale; }
  }
  if (train) {
    const float nc = -q.c;
    const float4 x0 = make_float4(fmaf(p0.x, kf0[0], nc), fmaf(p0.y, kf0[1], nc), fmaf(p0.z, kf0[2], nc), fmaf(p0.w, kf0[3], nc));
    const float4 x1 = make_float4(fmaf(p1.x, kf1[0], nc), fmaf(p1.y, kf1[1], nc), fmaf(p1.z, kf1[2], nc), fmaf(p1.w, kf1[3], nc));
    float4 M0, M1; mix_pair(x0, x1, k.f0, k.f1, M0, M1);
    const float c0 = -k.a1[0] - k.offp[0] * k.a2[0], c1 = -k.a1[1] - k.offp[1] * k.a2[1];
    t0.x = fmaf(-M0.x, k.a2[0], t0.x + c0); t0.y = fmaf(-M0.y, k.a2[0], t0.y + c0);
    t0.z = fmaf(-M0.z, k.a2[0], t0.z + c0); t0.w = fmaf(-M0.w, k.a2[0], t0.w + c0);
    t1.x = fmaf(-M1.x, k.a2[1], t1.x + c1); t1.y = fmaf(-M1.y, k.a2[1], t1.y + c1);
    t1.z = fmaf(-M1.z, k.a2[1], t1.z + c1); t1.w = fmaf(-M1.w, k.a2[1], t1.w + c1);
  }
  float4 dp0, dp1; mix_pair(t0, t1, k.g0, k.g1, dp0, dp1);       // = scale * dPd (the keep factor carries 1/(1-p)), see rows_const
  t0 = make_float4(dp0.x * kf0[0], dp0.y * kf0[1], dp0.z * kf0[2], dp0.w * kf0[3]);
  t1 = make_float4(dp1.x * kf1[0], dp1.y * kf1[1], dp1.z * kf1[2], dp1.w * kf1[3]);
}
// dS = P (scale dP - scale r): dp and r already carry the softmax scale
__device__ __forceinline__ float4 ds_quad(const float4& p, const float4& dp, float r) {
  return make_float4(p.x * (dp.x - r), p.y * (dp.y - r), p.z * (dp.z - r), p.w * (dp.w - r));
}

// short rows (N <= 256): one warp per row, pair layout, second sweep re-reads P and dP from L1
template <typename PT, typename MT>
__global__ void __launch_bounds__(256)
reattn_bwd_rows_mma_kernel(const PT* __restrict__ P, MT* __restrict__ dA, int B, int N,
                           const float* __restrict__ W, const float* __restrict__ bconv,
                           const float* __restrict__ gamma, const float* __restrict__ saved,
                           const float* __restrict__ coef, int train, float scale, QuadCtx q) {
  const int lane = threadIdx.x & 31, e = lane >> 2, k4 = lane & 3;
  const RowsConst kc = rows_const(W, bconv, gamma, saved, coef, train, q, scale, e, k4);
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int ld4 = N >> 2, ntiles = (ld4 + 7) >> 3, hs = N * N, rows = B * N;
  for (int r = wid; r < rows; r += nw) {
    const int b = r / N, i = r - b * N;
    const int64_t base = (int64_t)b * hs * H;
    const uint32_t ctr0 = (uint32_t)((uint64_t)base >> 2);
    const PT* Pi = P + base; MT* Di = dA + base;
    const int roff = k4 * hs + i * N;
    float rg0 = 0.f, rg1 = 0.f;
    for (int t = 0; t < ntiles; ++t) {
      const int quad = t * 8 + e;
      const bool ok = quad < ld4;
      const int off = roff + 4 * quad;
      const float4 p0 = ok ? ldp(Pi + off, q.c) : kZero4, p1 = ok ? ldp(Pi + off + 4 * hs, q.c) : kZero4;
      float4 t0 = ok ? map_ld(Di + off) : kZero4, t1 = ok ? map_ld(Di + off + 4 * hs) : kZero4;
      rows_tile(kc, q, train, ctr0 + ((uint32_t)off >> 2), hs, p0, p1, t0, t1);
      if (ok) {
        rg0 += dot4(t0, p0); rg1 += dot4(t1, p1);
        map_st(Di + off, t0); map_st(Di + off + 4 * hs, t1);
      }
    }
#pragma unroll
    for (int o = 4; o <= 16; o <<= 1) {
      rg0 += __shfl_xor_sync(0xffffffffu, rg0, o); rg1 += __shfl_xor_sync(0xffffffffu, rg1, o);
    }
    for (int t = 0; t < ntiles; ++t) {
      const int quad = t * 8 + e;
      if (quad < ld4) {
        const int off = roff + 4 * quad;
        const float4 p0 = ldp(Pi + off, q.c), p1 = ldp(Pi + off + 4 * hs, q.c);
        const float4 dp0 = map_ld(Di + off), dp1 = map_ld(Di + off + 4 * hs);
        map_st(Di + off, ds_quad(p0, dp0, rg0)); map_st(Di + off + 4 * hs, ds_quad(p1, dp1, rg1));
      }
    }
  }
}

// ------------------------------------------------------------------ long rows: one CTA (8 warps) per row
// For N > 256 the 8-head row (N * 32 bytes) no longer stays in L1 between the two sweeps of the warp-per-row kernels
// and, with every warp of the GPU on a different row, not even in L2.  This variant spreads the 32-key tiles of a
// row over the NW warps of a CTA (TPW tiles per warp, NW * TPW >= tiles with as little slack as possible) and keeps
// the row in registers between the sweeps: P and dA
// cross HBM exactly once.  The row sums r_g are combined through shared memory.
template <int TPW, int NW, typename PT, int MINB = (NW >= 7 ? 2 : 3)>
__global__ void __launch_bounds__(NW * 32, MINB)
reattn_bwd_rows_mma_cta_kernel(const PT* __restrict__ P, __nv_bfloat16* __restrict__ dA, int B, int N,
                               const float* __restrict__ W, const float* __restrict__ bconv,
                               const float* __restrict__ gamma, const float* __restrict__ saved,
                               const float* __restrict__ coef, int train, float scale, QuadCtx q) {
  __shared__ float srg[2][8][H];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, e = lane >> 2, k4 = lane & 3;
  const RowsConst kc = rows_const(W, bconv, gamma, saved, coef, train, q, scale, e, k4);
  const int ld4 = N >> 2, ntiles = (ld4 + 7) >> 3, hs = N * N, rows = B * N;
  // every quad of this warp's TPW tiles exists (warp-uniform, the same for all rows): predicate-free loads / stores at
  // immediate offsets from four row pointers.  Only the warp that owns the ragged last tile takes the general path.
  const bool allfull = (w + NW * (TPW - 1)) * 8 + 8 <= ld4;
  constexpr int TS = 32 * NW;                                // elements between consecutive tiles of one warp
  int par = 0;
  for (int r = blockIdx.x; r < rows; r += gridDim.x, par ^= 1) {
    const int b = r / N, i = r - b * N;
    const int64_t base = (int64_t)b * hs * H;
    const int loff = k4 * hs + i * N + 4 * (w * 8 + e);      // heads k4 (and k4 + 4 at + 4 hs), quad w * 8 + e of the row
    const uint32_t ctr = (uint32_t)((uint64_t)base >> 2) + ((uint32_t)loff >> 2);
    const PT* pr = P + base + loff; const PT* pr1 = pr + 4 * hs;
    __nv_bfloat16* dr = dA + base + loff; __nv_bfloat16* dr1 = dr + 4 * hs;
    float4 p0[TPW], p1[TPW]; uint2 k0[TPW], k1[TPW];      // k: dA on the way in, dP between the sweeps (bf16 x 4)
    if (allfull) {
#pragma unroll
      for (int tt = 0; tt < TPW; ++tt) {
        p0[tt] = ldp(pr + tt * TS, q.c); p1[tt] = ldp(pr1 + tt * TS, q.c);
        k0[tt] = *reinterpret_cast<const uint2*>(dr + tt * TS);
        k1[tt] = *reinterpret_cast<const uint2*>(dr1 + tt * TS);
      }
    } else {
#pragma unroll
      for (int tt = 0; tt < TPW; ++tt) {
        const bool ok = (w + NW * tt) * 8 + e < ld4;
        p0[tt] = ok ? ldp(pr + tt * TS, q.c) : kZero4; p1[tt] = ok ? ldp(pr1 + tt * TS, q.c) : kZero4;
        k0[tt] = ok ? *reinterpret_cast<const uint2*>(dr + tt * TS) : make_uint2(0u, 0u);
        k1[tt] = ok ? *reinterpret_cast<const uint2*>(dr1 + tt * TS) : make_uint2(0u, 0u);
      }
    }
    float rg0 = 0.f, rg1 = 0.f;
#pragma unroll
    for (int tt = 0; tt < TPW; ++tt) {
      if (!allfull && w + NW * tt >= ntiles) break;                       // warp-uniform
      float4 t0 = unpack_bf16x4(k0[tt]), t1 = unpack_bf16x4(k1[tt]);
      rows_tile(kc, q, train, ctr + 8 * NW * tt, hs, p0[tt], p1[tt], t0, t1);
      rg0 += dot4(t0, p0[tt]); rg1 += dot4(t1, p1[tt]);                   // invalid quads: P == 0
      k0[tt] = pack_bf16x4(t0); k1[tt] = pack_bf16x4(t1);
    }
#pragma unroll
    for (int o = 4; o <= 16; o <<= 1) {
      rg0 += __shfl_xor_sync(0xffffffffu, rg0, o); rg1 += __shfl_xor_sync(0xffffffffu, rg1, o);
    }
    if (e == 0) { srg[par][w][k4] = rg0; srg[par][w][k4 + 4] = rg1; }
    __syncthreads();
    rg0 = 0.f; rg1 = 0.f;
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) { rg0 += srg[par][ww][k4]; rg1 += srg[par][ww][k4 + 4]; }
    if (allfull) {
#pragma unroll
      for (int tt = 0; tt < TPW; ++tt) {
        map_st(dr + tt * TS, ds_quad(p0[tt], unpack_bf16x4(k0[tt]), rg0));
        map_st(dr1 + tt * TS, ds_quad(p1[tt], unpack_bf16x4(k1[tt]), rg1));
      }
    } else {
#pragma unroll
      for (int tt = 0; tt < TPW; ++tt) {
        if ((w + NW * tt) * 8 + e < ld4) {
          map_st(dr + tt * TS, ds_quad(p0[tt], unpack_bf16x4(k0[tt]), rg0));
          map_st(dr1 + tt * TS, ds_quad(p1[tt], unpack_bf16x4(k1[tt]), rg1));
        }
      }
    }
  }
}

// ------------------------------------------------------------------ long rows, asynchronous row pipeline (softmax)
// A ninth warp streams whole 8-head rows into a shared-memory ring with cp.async.bulk (one bulk copy per head row,
// completion on an mbarrier), so the next row is in flight while the eight consumer warps make their three sweeps
// over the current one in shared memory; ~56 registers per thread, 4 CTAs per SM.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bar_consumers(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

static inline size_t bulk_smem_bytes(int stages, int floats_per_stage) {
  return (size_t)stages * floats_per_stage * 4 + 2 * stages * sizeof(uint64_t);
}

// shared-memory accesses of the row ring by 32-bit shared-window address (computed once per row; the generic-pointer
// form re-derived the window base -- S2UR + uniform adds -- in front of every access)
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// NW consumer warps + 1 producer warp (8 by default, see the dispatcher).
template <int STAGES, int NW>
__global__ void __launch_bounds__((NW + 1) * 32)
softmax_stats_mma_bulk_kernel(float* __restrict__ S, __nv_bfloat16* __restrict__ Pc, int B, int N, float scale, QuadCtx q,
                              double* __restrict__ sums) {
  extern __shared__ __align__(128) unsigned char dsm[];
  float* ring = reinterpret_cast<float*>(dsm);
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + (size_t)STAGES * H * N * 4);
  uint64_t* empty = full + STAGES;
  __shared__ float smax[2][8][H], ssum[2][8][H];
  __shared__ float part[8][H + H * H];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, e = lane >> 2, k4 = lane & 3;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int st = 0; st < STAGES; ++st) { mbar_init(full + st, 1); mbar_init(empty + st, NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int hs = N * N, rows = B * N;
  const int nmine = (int)blockIdx.x < rows ? (rows - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  float cg[4] = {0.f, 0.f, 0.f, 0.f}, s = 0.f;
  if (w == NW) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int k = 0; k < nmine; ++k) {
        mbar_wait(empty + st, ph ^ 1u);
        mbar_expect_tx(full + st, (uint32_t)(H * N * 4));
        const int r = blockIdx.x + k * gridDim.x, b = r / N, i = r - b * N;
        const float* src = S + (int64_t)b * hs * H + i * N;
#pragma unroll
        for (int h = 0; h < H; ++h) bulk_g2s(ring + (st * H + h) * N, src + h * hs, (uint32_t)(N * 4), full + st);
        if (++st == STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else {
    const float sl2 = scale * 1.4426950408889634f;
    const int ld4 = N >> 2, ntiles = (ld4 + 7) >> 3;
    const uint32_t ring_s = smem_addr(ring);
    int st = 0; uint32_t ph = 0;
    for (int k = 0; k < nmine; ++k) {
      const int par = k & 1;
      const int r = blockIdx.x + k * gridDim.x, b = r / N, i = r - b * N;
      const int64_t base = (int64_t)b * hs * H;
      const int roff = e * hs + i * N;
      const uint32_t ctr0 = (uint32_t)((uint64_t)base >> 2) + ((uint32_t)roff >> 2);
      float* Srow = S + base + roff; __nv_bfloat16* Prow = Pc ? Pc + base + roff : nullptr;
      const uint32_t rs = ring_s + (uint32_t)((st * H + e) * N) * 4u;  // this lane's head row in shared memory
      mbar_wait(full + st, ph);
      float m = -INFINITY;
      for (int t = w; t < ntiles; t += NW) {
        const int qa = t * 8 + k4, qb = qa + 4;
        if (qa < ld4) m = fmaxf(m, hmax4(lds128(rs + 16 * qa)));
        if (qb < ld4) m = fmaxf(m, hmax4(lds128(rs + 16 * qb)));
      }
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      if (k4 == 0) smax[par][w][e] = m;
      bar_consumers(NW * 32);
#pragma unroll
      for (int ww = 0; ww < NW; ++ww) m = fmaxf(m, smax[par][ww][e]);
      m *= sl2;                                                    // scale > 0: max commutes with the scaling
      float l = 0.f;
      for (int t = w; t < ntiles; t += NW) {
        const int qa = t * 8 + k4, qb = qa + 4;
        if (qa < ld4) {
          float4 x = lds128(rs + 16 * qa);
          x.x = ex2(fmaf(x.x, sl2, -m)); x.y = ex2(fmaf(x.y, sl2, -m)); x.z = ex2(fmaf(x.z, sl2, -m)); x.w = ex2(fmaf(x.w, sl2, -m));
          sts128(rs + 16 * qa, x);
          l += hsum4(x);
        }
        if (qb < ld4) {
          float4 x = lds128(rs + 16 * qb);
          x.x = ex2(fmaf(x.x, sl2, -m)); x.y = ex2(fmaf(x.y, sl2, -m)); x.z = ex2(fmaf(x.z, sl2, -m)); x.w = ex2(fmaf(x.w, sl2, -m));
          sts128(rs + 16 * qb, x);
          l += hsum4(x);
        }
      }
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      if (k4 == 0) ssum[par][w][e] = l;
      bar_consumers(NW * 32);
      l = 0.f;
#pragma unroll
      for (int ww = 0; ww < NW; ++ww) l += ssum[par][ww][e];
      const float inv = 1.0f / l;
      for (int t = w; t < ntiles; t += NW) {
        const int qa = t * 8 + k4, qb = qa + 4;
        float4 pa = kZero4, pb = kZero4;
        if (qa < ld4) pa = emit_pdc(Srow, Prow, 4 * qa, lds128(rs + 16 * qa), inv, ctr0 + qa, q);
        if (qb < ld4) pb = emit_pdc(Srow, Prow, 4 * qb, lds128(rs + 16 * qb), inv, ctr0 + qb, q);
        s += hsum4(pa) + hsum4(pb);
        mma_keys(cg, pa, pb, pa, pb);
      }
      fence_proxy_async();                 // our generic-proxy writes to the stage precede the next bulk copy into it
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + st);
      if (++st == STAGES) { st = 0; ph ^= 1u; }
    }
  }
  reduce_stats(part, w < NW, s, cg[0], cg[1], sums, NW);
}

}  // namespace mma
}  // namespace vu
