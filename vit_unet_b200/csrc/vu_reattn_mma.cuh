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
                                         float4& y0, float4& y1) {
  float c[4] = {0.f, 0.f, 0.f, 0.f}, d[4] = {0.f, 0.f, 0.f, 0.f};
  mma8(c, tf32(x0.x), tf32(x0.y), tf32(x1.x), tf32(x1.y), b0, b1);    // rows e / e+8 = components 0 / 1
  mma8(d, tf32(x0.z), tf32(x0.w), tf32(x1.z), tf32(x1.w), b0, b1);    // rows e / e+8 = components 2 / 3
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

// dropout of one quad in place; returns the keep bits (bit t = component t kept)
__device__ __forceinline__ uint32_t drop_quad(float4& v, uint64_t flat_idx, const QuadCtx& q) {
  if (!q.thresh) return 0xFu;
  const uint4 rr = Philox::gen(q.seed, q.stream, flat_idx >> 2);
  const uint32_t m = (rr.x >= q.thresh ? 1u : 0u) | (rr.y >= q.thresh ? 2u : 0u) | (rr.z >= q.thresh ? 4u : 0u) |
                     (rr.w >= q.thresh ? 8u : 0u);
  v.x = (m & 1u) ? v.x * q.dscale : 0.f; v.y = (m & 2u) ? v.y * q.dscale : 0.f;
  v.z = (m & 4u) ? v.z * q.dscale : 0.f; v.w = (m & 8u) ? v.w * q.dscale : 0.f;
  return m;
}
__device__ __forceinline__ void sub4(float4& v, float c) { v.x -= c; v.y -= c; v.z -= c; v.w -= c; }
__device__ __forceinline__ float comp(const float4& v, int u) { return u == 0 ? v.x : (u == 1 ? v.y : (u == 2 ? v.z : v.w)); }
__device__ __forceinline__ float4 ldq(const float* p) { return *reinterpret_cast<const float4*>(p); }
constexpr float4 kZero4 = {0.f, 0.f, 0.f, 0.f};

// block-wide reduction of the 8 + 64 accumulators of a statistics kernel (s[head e] on lanes k4 == 0, M[e][2k4..+1]
// on every lane) followed by one double atomicAdd per entry.  256 threads.
__device__ __forceinline__ void reduce_stats(float s, float m0, float m1, double* __restrict__ out) {
  __shared__ float part[8][H + H * H];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, e = lane >> 2, k4 = lane & 3;
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if (k4 == 0) part[warp][e] = s;
  part[warp][H + e * H + 2 * k4] = m0;
  part[warp][H + e * H + 2 * k4 + 1] = m1;
  __syncthreads();
  if (threadIdx.x < H + H * H) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += (double)part[w][threadIdx.x];
    atomicAdd(out + threadIdx.x, t);
  }
}

// ------------------------------------------------------------------ forward: softmax + centred moments
// One warp per row (b, i), all 8 heads at once in the stats layout (lane = head e, key quads k4 and k4+4 of each
// 32-key tile).  Sweep A: online (max, sum) of exp2; sweep B (re-read from L1/L2): write P, accumulate
// s'_g = sum (Pd_g - c) and G' = sum (Pd - c)(Pd - c)^T (4 MMAs per tile, K = keys).
__global__ void __launch_bounds__(256)
softmax_stats_mma_kernel(float* __restrict__ S, int B, int N, float scale, QuadCtx q, double* __restrict__ sums) {
  const int lane = threadIdx.x & 31, e = lane >> 2, k4 = lane & 3;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float sl2 = scale * 1.4426950408889634f;
  const int ld4 = N >> 2, ntiles = (ld4 + 7) >> 3;
  const int64_t head_stride = (int64_t)N * N, img_stride = head_stride * H, rows = (int64_t)B * N;
  float cg[4] = {0.f, 0.f, 0.f, 0.f}, s = 0.f;
  for (int64_t r = wid; r < rows; r += nw) {
    const int64_t b = r / N; const int i = (int)(r - b * N);
    const int64_t row_off = b * img_stride + e * head_stride + (int64_t)i * N;
    float* row = S + row_off;
    float m = -INFINITY, l = 0.f;
    for (int t = 0; t < ntiles; ++t) {
      const int qa = t * 8 + k4, qb = qa + 4;
      const bool va = qa < ld4, vb = qb < ld4;
      float4 xa = va ? ldq(row + 4 * qa) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      float4 xb = vb ? ldq(row + 4 * qb) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      xa.x *= sl2; xa.y *= sl2; xa.z *= sl2; xa.w *= sl2; xb.x *= sl2; xb.y *= sl2; xb.z *= sl2; xb.w *= sl2;
      const float mq = fmaxf(fmaxf(fmaxf(xa.x, xa.y), fmaxf(xa.z, xa.w)), fmaxf(fmaxf(xb.x, xb.y), fmaxf(xb.z, xb.w)));
      const float mn = fmaxf(m, mq);
      if (mn > -INFINITY) {
        l = l * exp2f(m - mn) + ((exp2f(xa.x - mn) + exp2f(xa.y - mn)) + (exp2f(xa.z - mn) + exp2f(xa.w - mn))) +
            ((exp2f(xb.x - mn) + exp2f(xb.y - mn)) + (exp2f(xb.z - mn) + exp2f(xb.w - mn)));
        m = mn;
      }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m, o), lo = __shfl_xor_sync(0xffffffffu, l, o);
      const float mn = fmaxf(m, mo);
      l = (m > -INFINITY ? l * exp2f(m - mn) : 0.f) + (mo > -INFINITY ? lo * exp2f(mo - mn) : 0.f);
      m = mn;
    }
    const float inv = 1.0f / l;
    for (int t = 0; t < ntiles; ++t) {
      const int qa = t * 8 + k4, qb = qa + 4;
      const bool va = qa < ld4, vb = qb < ld4;
      float4 pa = kZero4, pb = kZero4;
      if (va) {
        const float4 x = ldq(row + 4 * qa);
        pa = make_float4(exp2f(fmaf(x.x, sl2, -m)) * inv, exp2f(fmaf(x.y, sl2, -m)) * inv,
                         exp2f(fmaf(x.z, sl2, -m)) * inv, exp2f(fmaf(x.w, sl2, -m)) * inv);
        *reinterpret_cast<float4*>(row + 4 * qa) = pa;
        drop_quad(pa, (uint64_t)(row_off + 4 * qa), q); sub4(pa, q.c);
      }
      if (vb) {
        const float4 x = ldq(row + 4 * qb);
        pb = make_float4(exp2f(fmaf(x.x, sl2, -m)) * inv, exp2f(fmaf(x.y, sl2, -m)) * inv,
                         exp2f(fmaf(x.z, sl2, -m)) * inv, exp2f(fmaf(x.w, sl2, -m)) * inv);
        *reinterpret_cast<float4*>(row + 4 * qb) = pb;
        drop_quad(pb, (uint64_t)(row_off + 4 * qb), q); sub4(pb, q.c);
      }
      s += ((pa.x + pa.y) + (pa.z + pa.w)) + ((pb.x + pb.y) + (pb.z + pb.w));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t ua = tf32(comp(pa, u)), ub = tf32(comp(pb, u));
        mma8(cg, ua, 0u, ub, 0u, ua, ub);
      }
    }
  }
  reduce_stats(s, cg[0], cg[1], sums);
}

// ------------------------------------------------------------------ forward: A = fold . (Pd - c) + shift'
// Flat tiles of 32 consecutive positions of one image (rows are contiguous: ld == N), pair layout.
__global__ void __launch_bounds__(256)
reattn_mix_mma_kernel(const float* __restrict__ P, __nv_bfloat16* __restrict__ A, const float* __restrict__ fold,
                      int B, int N, QuadCtx q) {
  const int lane = threadIdx.x & 31, e = lane >> 2, k4 = lane & 3;
  uint32_t b0, b1; frag_fwd(fold, e, k4, b0, b1);
  float sh0 = 0.f, sh1 = 0.f;
#pragma unroll
  for (int g = 0; g < H; ++g) { sh0 += fold[k4 * H + g]; sh1 += fold[(k4 + 4) * H + g]; }
  sh0 = fmaf(sh0, q.c, fold[H * H + k4]); sh1 = fmaf(sh1, q.c, fold[H * H + k4 + 4]);
  const int64_t head_stride = (int64_t)N * N, img_stride = head_stride * H;
  const int64_t quads = head_stride >> 2, tiles = (quads + 7) >> 3, total = tiles * B;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = wid; t < total; t += 2 * nw) {
    int64_t off[2]; bool ok[2]; float4 x0[2], x1[2];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int64_t tt = t + v * nw;
      const int64_t b = tt / tiles, quad = (tt - b * tiles) * 8 + e;
      ok[v] = tt < total && quad < quads;
      off[v] = b * img_stride + k4 * head_stride + quad * 4;
      x0[v] = ok[v] ? ldq(P + off[v]) : kZero4;
      x1[v] = ok[v] ? ldq(P + off[v] + 4 * head_stride) : kZero4;
    }
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      if (t + v * nw >= total) break;                       // warp-uniform
      drop_quad(x0[v], (uint64_t)off[v], q); drop_quad(x1[v], (uint64_t)(off[v] + 4 * head_stride), q);
      sub4(x0[v], q.c); sub4(x1[v], q.c);
      float4 y0, y1; mix_pair(x0[v], x1[v], b0, b1, y0, y1);
      if (ok[v]) {
        y0.x += sh0; y0.y += sh0; y0.z += sh0; y0.w += sh0; y1.x += sh1; y1.y += sh1; y1.z += sh1; y1.w += sh1;
        map_st(A + off[v], y0); map_st(A + off[v] + 4 * head_stride, y1);
      }
    }
  }
}

// ------------------------------------------------------------------ backward pass 1: A = mix(P) recomputed + reductions
// red[h] += sum dA_h ;  red[H + h*H + g] += sum dA_h (Pd_g - c).  The mix runs in the pair layout, the reductions in
// the stats layout (second read of the same 1 KB P tile hits L1).
__global__ void __launch_bounds__(256)
reattn_mix_reduce_mma_kernel(const float* __restrict__ P, const __nv_bfloat16* __restrict__ dA,
                             __nv_bfloat16* __restrict__ A, const float* __restrict__ fold, int B, int N, QuadCtx q,
                             double* __restrict__ out) {
  const int lane = threadIdx.x & 31, e = lane >> 2, k4 = lane & 3;
  uint32_t b0, b1; frag_fwd(fold, e, k4, b0, b1);
  float sh0 = 0.f, sh1 = 0.f;
#pragma unroll
  for (int g = 0; g < H; ++g) { sh0 += fold[k4 * H + g]; sh1 += fold[(k4 + 4) * H + g]; }
  sh0 = fmaf(sh0, q.c, fold[H * H + k4]); sh1 = fmaf(sh1, q.c, fold[H * H + k4 + 4]);
  const int64_t head_stride = (int64_t)N * N, img_stride = head_stride * H;
  const int64_t quads = head_stride >> 2, tiles = (quads + 7) >> 3, total = tiles * B;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float cx[4] = {0.f, 0.f, 0.f, 0.f}, s1 = 0.f;
  for (int64_t t = wid; t < total; t += nw) {
    const int64_t b = t / tiles, tile = t - b * tiles, ibase = b * img_stride;
    // pair layout: mixed map
    const int64_t quad = tile * 8 + e;
    const bool ok = quad < quads;
    const int64_t off = ibase + k4 * head_stride + quad * 4;
    float4 x0 = ok ? ldq(P + off) : kZero4, x1 = ok ? ldq(P + off + 4 * head_stride) : kZero4;
    // stats layout: head e, quads k4 and k4 + 4 of the tile
    const int64_t qa = tile * 8 + k4, qb = qa + 4;
    const bool va = qa < quads, vb = qb < quads;
    const int64_t offa = ibase + e * head_stride + qa * 4, offb = offa + 16;
    float4 pa = va ? ldq(P + offa) : kZero4, pb = vb ? ldq(P + offb) : kZero4;
    const float4 da = va ? map_ld(dA + offa) : kZero4, db = vb ? map_ld(dA + offb) : kZero4;
    drop_quad(x0, (uint64_t)off, q); drop_quad(x1, (uint64_t)(off + 4 * head_stride), q);
    sub4(x0, q.c); sub4(x1, q.c);
    float4 y0, y1; mix_pair(x0, x1, b0, b1, y0, y1);
    if (ok) {
      y0.x += sh0; y0.y += sh0; y0.z += sh0; y0.w += sh0; y1.x += sh1; y1.y += sh1; y1.z += sh1; y1.w += sh1;
      map_st(A + off, y0); map_st(A + off + 4 * head_stride, y1);
    }
    if (va) { drop_quad(pa, (uint64_t)offa, q); sub4(pa, q.c); }
    if (vb) { drop_quad(pb, (uint64_t)offb, q); sub4(pb, q.c); }
    s1 += ((da.x + da.y) + (da.z + da.w)) + ((db.x + db.y) + (db.z + db.w));
#pragma unroll
    for (int u = 0; u < 4; ++u)
      mma8(cx, tf32(comp(da, u)), 0u, tf32(comp(db, u)), 0u, tf32(comp(pa, u)), tf32(comp(pb, u)));
  }
  reduce_stats(s1, cx[0], cx[1], out);
}

// ------------------------------------------------------------------ backward pass 2: dA -> dS in place
// One warp per row (b, i), pair layout:
//   dM_h  = k_h (dA_h - m1_h - Ahat_h m2_h)   (train)   |   k_h dA_h   (eval),   Ahat_h = (M_h - mean_h) invstd_h
//   dPd_g = sum_h W[h][g] dM_h ;  dP_g = keep_g dPd_g / (1-p) ;  r_g = sum_j dP_g P_g ;  dS_g = scale P_g (dP_g - r_g)
__global__ void __launch_bounds__(256)
reattn_bwd_rows_mma_kernel(const float* __restrict__ P, __nv_bfloat16* __restrict__ dA, int B, int N,
                           const float* __restrict__ W, const float* __restrict__ bconv,
                           const float* __restrict__ gamma, const float* __restrict__ saved,
                           const float* __restrict__ coef, int train, float scale, QuadCtx q) {
  const int lane = threadIdx.x & 31, e = lane >> 2, k4 = lane & 3;
  uint32_t f0, f1, g0, g1;
  frag_fwd(W, e, k4, f0, f1);           // M_h   = sum_g W[h][g] Pd_g
  frag_bwd(W, e, k4, g0, g1);           // dPd_g = sum_h W[h][g] dM_h
  float offp[2], a1[2], a2[2], kh[2];
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int h = k4 + 4 * v;
    float rs = 0.f;
#pragma unroll
    for (int g = 0; g < H; ++g) rs += W[h * H + g];
    offp[v] = bconv[h] - saved[h] + q.c * rs;             // M_h - mean_h = sum_g W_hg (Pd_g - c) + offp
    a1[v] = train ? coef[h] : 0.f;
    a2[v] = train ? saved[H + h] * coef[H + h] : 0.f;
    kh[v] = gamma[h] * saved[H + h];
  }
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int ld4 = N >> 2, ntiles = (ld4 + 7) >> 3;
  const int64_t head_stride = (int64_t)N * N, img_stride = head_stride * H, rows = (int64_t)B * N;
  for (int64_t r = wid; r < rows; r += nw) {
    const int64_t b = r / N; const int i = (int)(r - b * N);
    const int64_t row_off = b * img_stride + k4 * head_stride + (int64_t)i * N;
    float rg0 = 0.f, rg1 = 0.f;
    for (int t = 0; t < ntiles; ++t) {
      const int quad = t * 8 + e;
      const bool ok = quad < ld4;
      const int64_t off0 = row_off + 4 * quad, off1 = off0 + 4 * head_stride;
      const float4 p0 = ok ? ldq(P + off0) : kZero4, p1 = ok ? ldq(P + off1) : kZero4;
      const float4 d0 = ok ? map_ld(dA + off0) : kZero4, d1 = ok ? map_ld(dA + off1) : kZero4;
      float4 x0 = p0, x1 = p1;
      const uint32_t m0 = drop_quad(x0, (uint64_t)off0, q), m1 = drop_quad(x1, (uint64_t)off1, q);
      float4 t0 = d0, t1 = d1;
      if (train) {
        sub4(x0, q.c); sub4(x1, q.c);
        float4 M0, M1; mix_pair(x0, x1, f0, f1, M0, M1);
        t0.x = d0.x - a1[0] - (M0.x + offp[0]) * a2[0]; t0.y = d0.y - a1[0] - (M0.y + offp[0]) * a2[0];
        t0.z = d0.z - a1[0] - (M0.z + offp[0]) * a2[0]; t0.w = d0.w - a1[0] - (M0.w + offp[0]) * a2[0];
        t1.x = d1.x - a1[1] - (M1.x + offp[1]) * a2[1]; t1.y = d1.y - a1[1] - (M1.y + offp[1]) * a2[1];
        t1.z = d1.z - a1[1] - (M1.z + offp[1]) * a2[1]; t1.w = d1.w - a1[1] - (M1.w + offp[1]) * a2[1];
      }
      t0.x *= kh[0]; t0.y *= kh[0]; t0.z *= kh[0]; t0.w *= kh[0];
      t1.x *= kh[1]; t1.y *= kh[1]; t1.z *= kh[1]; t1.w *= kh[1];
      float4 dp0, dp1; mix_pair(t0, t1, g0, g1, dp0, dp1);
      dp0.x = (m0 & 1u) ? dp0.x * q.dscale : 0.f; dp0.y = (m0 & 2u) ? dp0.y * q.dscale : 0.f;
      dp0.z = (m0 & 4u) ? dp0.z * q.dscale : 0.f; dp0.w = (m0 & 8u) ? dp0.w * q.dscale : 0.f;
      dp1.x = (m1 & 1u) ? dp1.x * q.dscale : 0.f; dp1.y = (m1 & 2u) ? dp1.y * q.dscale : 0.f;
      dp1.z = (m1 & 4u) ? dp1.z * q.dscale : 0.f; dp1.w = (m1 & 8u) ? dp1.w * q.dscale : 0.f;
      if (ok) {
        rg0 += dot4(dp0, p0); rg1 += dot4(dp1, p1);
        map_st(dA + off0, dp0); map_st(dA + off1, dp1);
      }
    }
#pragma unroll
    for (int o = 4; o <= 16; o <<= 1) {
      rg0 += __shfl_xor_sync(0xffffffffu, rg0, o); rg1 += __shfl_xor_sync(0xffffffffu, rg1, o);
    }
    for (int t = 0; t < ntiles; ++t) {
      const int quad = t * 8 + e;
      if (quad < ld4) {
        const int64_t off0 = row_off + 4 * quad, off1 = off0 + 4 * head_stride;
        const float4 p0 = ldq(P + off0), p1 = ldq(P + off1);
        const float4 dp0 = map_ld(dA + off0), dp1 = map_ld(dA + off1);
        map_st(dA + off0, make_float4(scale * p0.x * (dp0.x - rg0), scale * p0.y * (dp0.y - rg0),
                                      scale * p0.z * (dp0.z - rg0), scale * p0.w * (dp0.w - rg0)));
        map_st(dA + off1, make_float4(scale * p1.x * (dp1.x - rg1), scale * p1.y * (dp1.y - rg1),
                                      scale * p1.z * (dp1.z - rg1), scale * p1.w * (dp1.w - rg1)));
      }
    }
  }
}

// ------------------------------------------------------------------ long rows: one CTA (8 warps) per row
// For N > 256 the 8-head row (N * 32 bytes) no longer stays in L1 between the two sweeps of the warp-per-row kernels
// and, with every warp of the GPU on a different row, not even in L2.  These variants spread the 32-key tiles of a
// row over the 8 warps of a CTA (TPW tiles per warp) and keep the row in registers between the sweeps: S / P / dA
// cross HBM exactly once.  Row-wide quantities (max, sum of exp, r_g) are combined through shared memory.
template <int TPW>
__global__ void __launch_bounds__(256)
softmax_stats_mma_cta_kernel(float* __restrict__ S, int B, int N, float scale, QuadCtx q, double* __restrict__ sums) {
  __shared__ float smax[2][8][H], ssum[2][8][H];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, e = lane >> 2, k4 = lane & 3;
  const float sl2 = scale * 1.4426950408889634f;
  const int ld4 = N >> 2, ntiles = (ld4 + 7) >> 3;
  const int64_t head_stride = (int64_t)N * N, img_stride = head_stride * H, rows = (int64_t)B * N;
  float cg[4] = {0.f, 0.f, 0.f, 0.f}, s = 0.f;
  int par = 0;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x, par ^= 1) {
    const int64_t b = r / N; const int i = (int)(r - b * N);
    const int64_t row_off = b * img_stride + e * head_stride + (int64_t)i * N;
    float* row = S + row_off;
    float4 xa[TPW], xb[TPW];
    float m = -INFINITY;
#pragma unroll
    for (int tt = 0; tt < TPW; ++tt) {
      const int qa = (w + 8 * tt) * 8 + k4, qb = qa + 4;
      xa[tt] = qa < ld4 ? ldq(row + 4 * qa) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      xb[tt] = qb < ld4 ? ldq(row + 4 * qb) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
#pragma unroll
    for (int tt = 0; tt < TPW; ++tt) {
      xa[tt].x *= sl2; xa[tt].y *= sl2; xa[tt].z *= sl2; xa[tt].w *= sl2;
      xb[tt].x *= sl2; xb[tt].y *= sl2; xb[tt].z *= sl2; xb[tt].w *= sl2;
      m = fmaxf(m, fmaxf(fmaxf(fmaxf(xa[tt].x, xa[tt].y), fmaxf(xa[tt].z, xa[tt].w)),
                         fmaxf(fmaxf(xb[tt].x, xb[tt].y), fmaxf(xb[tt].z, xb[tt].w))));
    }
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    if (k4 == 0) smax[par][w][e] = m;
    __syncthreads();
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) m = fmaxf(m, smax[par][ww][e]);
    float l = 0.f;
#pragma unroll
    for (int tt = 0; tt < TPW; ++tt) {
      xa[tt].x = exp2f(xa[tt].x - m); xa[tt].y = exp2f(xa[tt].y - m); xa[tt].z = exp2f(xa[tt].z - m); xa[tt].w = exp2f(xa[tt].w - m);
      xb[tt].x = exp2f(xb[tt].x - m); xb[tt].y = exp2f(xb[tt].y - m); xb[tt].z = exp2f(xb[tt].z - m); xb[tt].w = exp2f(xb[tt].w - m);
      l += ((xa[tt].x + xa[tt].y) + (xa[tt].z + xa[tt].w)) + ((xb[tt].x + xb[tt].y) + (xb[tt].z + xb[tt].w));
    }
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    if (k4 == 0) ssum[par][w][e] = l;
    __syncthreads();
    l = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) l += ssum[par][ww][e];
    const float inv = 1.0f / l;
#pragma unroll
    for (int tt = 0; tt < TPW; ++tt) {
      if (w + 8 * tt >= ntiles) break;                                   // warp-uniform
      const int qa = (w + 8 * tt) * 8 + k4, qb = qa + 4;
      float4 pa = kZero4, pb = kZero4;
      if (qa < ld4) {
        pa = make_float4(xa[tt].x * inv, xa[tt].y * inv, xa[tt].z * inv, xa[tt].w * inv);
        *reinterpret_cast<float4*>(row + 4 * qa) = pa;
        drop_quad(pa, (uint64_t)(row_off + 4 * qa), q); sub4(pa, q.c);
      }
      if (qb < ld4) {
        pb = make_float4(xb[tt].x * inv, xb[tt].y * inv, xb[tt].z * inv, xb[tt].w * inv);
        *reinterpret_cast<float4*>(row + 4 * qb) = pb;
        drop_quad(pb, (uint64_t)(row_off + 4 * qb), q); sub4(pb, q.c);
      }
      s += ((pa.x + pa.y) + (pa.z + pa.w)) + ((pb.x + pb.y) + (pb.z + pb.w));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t ua = tf32(comp(pa, u)), ub = tf32(comp(pb, u));
        mma8(cg, ua, 0u, ub, 0u, ua, ub);
      }
    }
  }
  reduce_stats(s, cg[0], cg[1], sums);
}

__device__ __forceinline__ uint2 pack_bf16x4(const float4& v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 u; u.x = *reinterpret_cast<uint32_t*>(&lo); u.y = *reinterpret_cast<uint32_t*>(&hi);
  return u;
}
__device__ __forceinline__ float4 unpack_bf16x4(const uint2& u) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

template <int TPW>
__global__ void __launch_bounds__(256, 2)
reattn_bwd_rows_mma_cta_kernel(const float* __restrict__ P, __nv_bfloat16* __restrict__ dA, int B, int N,
                               const float* __restrict__ W, const float* __restrict__ bconv,
                               const float* __restrict__ gamma, const float* __restrict__ saved,
                               const float* __restrict__ coef, int train, float scale, QuadCtx q) {
  __shared__ float srg[2][8][H];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, e = lane >> 2, k4 = lane & 3;
  uint32_t f0, f1, g0, g1;
  frag_fwd(W, e, k4, f0, f1);
  frag_bwd(W, e, k4, g0, g1);
  float offp[2], a1[2], a2[2], kh[2];
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int h = k4 + 4 * v;
    float rs = 0.f;
#pragma unroll
    for (int g = 0; g < H; ++g) rs += W[h * H + g];
    offp[v] = bconv[h] - saved[h] + q.c * rs;
    a1[v] = train ? coef[h] : 0.f;
    a2[v] = train ? saved[H + h] * coef[H + h] : 0.f;
    kh[v] = gamma[h] * saved[H + h];
  }
  const int ld4 = N >> 2, ntiles = (ld4 + 7) >> 3;
  const int64_t head_stride = (int64_t)N * N, img_stride = head_stride * H, rows = (int64_t)B * N;
  int par = 0;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x, par ^= 1) {
    const int64_t b = r / N; const int i = (int)(r - b * N);
    const int64_t row_off = b * img_stride + k4 * head_stride + (int64_t)i * N;
    float4 p0[TPW], p1[TPW]; uint2 k0[TPW], k1[TPW];      // k: dA on the way in, dP between the sweeps (bf16 x 4)
#pragma unroll
    for (int tt = 0; tt < TPW; ++tt) {
      const int quad = (w + 8 * tt) * 8 + e;
      const bool ok = quad < ld4;
      const int64_t off0 = row_off + 4 * quad, off1 = off0 + 4 * head_stride;
      p0[tt] = ok ? ldq(P + off0) : kZero4; p1[tt] = ok ? ldq(P + off1) : kZero4;
      k0[tt] = ok ? *reinterpret_cast<const uint2*>(dA + off0) : make_uint2(0u, 0u);
      k1[tt] = ok ? *reinterpret_cast<const uint2*>(dA + off1) : make_uint2(0u, 0u);
    }
    float rg0 = 0.f, rg1 = 0.f;
#pragma unroll
    for (int tt = 0; tt < TPW; ++tt) {
      if (w + 8 * tt >= ntiles) break;                                   // warp-uniform
      const int quad = (w + 8 * tt) * 8 + e;
      const int64_t off0 = row_off + 4 * quad, off1 = off0 + 4 * head_stride;
      float4 x0 = p0[tt], x1 = p1[tt];
      const uint32_t m0 = drop_quad(x0, (uint64_t)off0, q), m1 = drop_quad(x1, (uint64_t)off1, q);
      float4 t0 = unpack_bf16x4(k0[tt]), t1 = unpack_bf16x4(k1[tt]);
      if (train) {
        sub4(x0, q.c); sub4(x1, q.c);
        float4 M0, M1; mix_pair(x0, x1, f0, f1, M0, M1);
        t0.x = t0.x - a1[0] - (M0.x + offp[0]) * a2[0]; t0.y = t0.y - a1[0] - (M0.y + offp[0]) * a2[0];
        t0.z = t0.z - a1[0] - (M0.z + offp[0]) * a2[0]; t0.w = t0.w - a1[0] - (M0.w + offp[0]) * a2[0];
        t1.x = t1.x - a1[1] - (M1.x + offp[1]) * a2[1]; t1.y = t1.y - a1[1] - (M1.y + offp[1]) * a2[1];
        t1.z = t1.z - a1[1] - (M1.z + offp[1]) * a2[1]; t1.w = t1.w - a1[1] - (M1.w + offp[1]) * a2[1];
      }
      t0.x *= kh[0]; t0.y *= kh[0]; t0.z *= kh[0]; t0.w *= kh[0];
      t1.x *= kh[1]; t1.y *= kh[1]; t1.z *= kh[1]; t1.w *= kh[1];
      float4 dp0, dp1; mix_pair(t0, t1, g0, g1, dp0, dp1);
      dp0.x = (m0 & 1u) ? dp0.x * q.dscale : 0.f; dp0.y = (m0 & 2u) ? dp0.y * q.dscale : 0.f;
      dp0.z = (m0 & 4u) ? dp0.z * q.dscale : 0.f; dp0.w = (m0 & 8u) ? dp0.w * q.dscale : 0.f;
      dp1.x = (m1 & 1u) ? dp1.x * q.dscale : 0.f; dp1.y = (m1 & 2u) ? dp1.y * q.dscale : 0.f;
      dp1.z = (m1 & 4u) ? dp1.z * q.dscale : 0.f; dp1.w = (m1 & 8u) ? dp1.w * q.dscale : 0.f;
      rg0 += dot4(dp0, p0[tt]); rg1 += dot4(dp1, p1[tt]);                 // invalid quads: P == 0
      k0[tt] = pack_bf16x4(dp0); k1[tt] = pack_bf16x4(dp1);
    }
#pragma unroll
    for (int o = 4; o <= 16; o <<= 1) {
      rg0 += __shfl_xor_sync(0xffffffffu, rg0, o); rg1 += __shfl_xor_sync(0xffffffffu, rg1, o);
    }
    if (e == 0) { srg[par][w][k4] = rg0; srg[par][w][k4 + 4] = rg1; }
    __syncthreads();
    rg0 = 0.f; rg1 = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) { rg0 += srg[par][ww][k4]; rg1 += srg[par][ww][k4 + 4]; }
#pragma unroll
    for (int tt = 0; tt < TPW; ++tt) {
      const int quad = (w + 8 * tt) * 8 + e;
      if (quad < ld4) {
        const int64_t off0 = row_off + 4 * quad, off1 = off0 + 4 * head_stride;
        const float4 dp0 = unpack_bf16x4(k0[tt]), dp1 = unpack_bf16x4(k1[tt]);
        const float4 a = p0[tt], c = p1[tt];
        map_st(dA + off0, make_float4(scale * a.x * (dp0.x - rg0), scale * a.y * (dp0.y - rg0),
                                      scale * a.z * (dp0.z - rg0), scale * a.w * (dp0.w - rg0)));
        map_st(dA + off1, make_float4(scale * c.x * (dp1.x - rg1), scale * c.y * (dp1.y - rg1),
                                      scale * c.z * (dp1.z - rg1), scale * c.w * (dp1.w - rg1)));
      }
    }
  }
}

}  // namespace mma
}  // namespace vu
