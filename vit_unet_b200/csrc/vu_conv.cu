// 3x3 C->C convolutions of the ViT-UNet path, evaluated directly on the patch layouts:
//   * q/k/v convs: every token is a (C,p,p) image, zero padded at the PATCH border (model.py:137-139,152-154)
//   * PatchEncoder conv (README variant) and the reconstruction conv: whole-image 'same' conv (model.py:428)
// One thread per pixel, all C output channels of all fused convs in registers; the 9*C input taps come from
// L1 (each input float is reused 9*C*nconv times), so HBM sees one read of x and one write per output.
#include <mutex>

#include "vu_common.cuh"
#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace vu {

// Filter weights / biases of the launch in flight live in __constant__ memory: after full unrolling every weight is
// an immediate constant-bank operand of its FFMA (no LDS / register per weight).  They are refreshed with a
// stream-ordered device-to-device cudaMemcpyToSymbolAsync before each launch (972 B at most); FilterGuard (below)
// orders the (upload, launch) pairs of different streams / host threads on the same device.
__constant__ float c_w[3 * 4 * 4 * 9];
__constant__ float c_b[3 * 4];

struct ConvGeom {
  Layout lin, lout;      // storage layouts of the tensor(s) read / written
  int bp;                // border patch (0: image)
  int fast;              // lin.p == bp: taps are at fixed offsets inside the patch
  int64_t per_image, npix_total;   // C*H*W ; B*H*W
};

__device__ __forceinline__ void tap_setup(const ConvGeom& g, int y, int x, int& iy, int& ix, int& limy, int& limx) {
  if (g.bp) {
    if ((g.bp & (g.bp - 1)) == 0) { iy = y & (g.bp - 1); ix = x & (g.bp - 1); }
    else { iy = y % g.bp; ix = x % g.bp; }
    limy = g.bp; limx = g.bp;
  }
  else { iy = y; ix = x; limy = g.lin.H; limx = g.lin.W; }
}

// forward: out_k[co] = bias_k[co] + sum_{ci,ky,kx} x[ci](y+ky-1, x+kx-1) * w_k[co][ci][ky][kx]
template <int C, int NCONV>
__global__ void __launch_bounds__(256)
conv3x3_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ o0, float* __restrict__ o1, float* __restrict__ o2, ConvGeom g) {
  const uint32_t hw = (uint32_t)(g.lin.H * g.lin.W), total = (uint32_t)g.npix_total;     // host checks < 2^31
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const uint32_t b = t / hw, pix = t - b * hw;
    int y, xx; g.lout.pixel(pix, y, xx);
    int iy, ix, limy, limx; tap_setup(g, y, xx, iy, ix, limy, limx);
    const float* xb = x + (int64_t)b * g.per_image;
    float v[C][9];
    const int64_t cs = g.lin.cstride();
    if (g.fast) {
      int64_t base = g.lin.at(0, y, xx);
      int rs = g.bp ? g.bp : g.lin.W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          bool ok = (unsigned)(iy + ky - 1) < (unsigned)limy && (unsigned)(ix + kx - 1) < (unsigned)limx;
#pragma unroll
          for (int ci = 0; ci < C; ++ci)
            v[ci][ky * 3 + kx] = ok ? __ldg(xb + base + ci * cs + (ky - 1) * rs + (kx - 1)) : 0.f;
        }
    } else {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          bool ok = (unsigned)(iy + ky - 1) < (unsigned)limy && (unsigned)(ix + kx - 1) < (unsigned)limx;
          int64_t off = ok ? g.lin.at(0, y + ky - 1, xx + kx - 1) : 0;
#pragma unroll
          for (int ci = 0; ci < C; ++ci) v[ci][ky * 3 + kx] = ok ? __ldg(xb + off + ci * cs) : 0.f;
        }
    }
    int64_t obase = (int64_t)b * g.per_image + g.lout.base_of(pix);
    const int64_t ocs = g.lout.cstride();
#pragma unroll
    for (int k = 0; k < NCONV; ++k) {
      float* op = k == 0 ? o0 : (k == 1 ? o1 : o2);
#pragma unroll
      for (int co = 0; co < C; ++co) {
        float acc = c_b[k * C + co];
#pragma unroll
        for (int ci = 0; ci < C; ++ci)
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) acc = fmaf(v[ci][tp], c_w[((k * C + co) * C + ci) * 9 + tp], acc);
        op[obase + co * ocs] = acc;
      }
    }
  }
}

// backward data: dx[ci](y,x) = sum_k sum_co sum_{ky,kx} dy_k[co](y-ky+1, x-kx+1) * w_k[co][ci][ky][kx]
template <int C, int NCONV>
__global__ void __launch_bounds__(256)
conv3x3_bwd_data_kernel(const float* __restrict__ d0, const float* __restrict__ d1, const float* __restrict__ d2,
                        const float* __restrict__ w, float* __restrict__ dx, ConvGeom g, int accumulate) {
  const uint32_t hw = (uint32_t)(g.lin.H * g.lin.W), total = (uint32_t)g.npix_total;     // host checks < 2^31
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const uint32_t b = t / hw, pix = t - b * hw;
    int y, xx; g.lout.pixel(pix, y, xx);
    int iy, ix, limy, limx; tap_setup(g, y, xx, iy, ix, limy, limx);
    float acc[C];
#pragma unroll
    for (int ci = 0; ci < C; ++ci) acc[ci] = 0.f;
    const int64_t cs = g.lin.cstride();
    int64_t base = g.fast ? g.lin.at(0, y, xx) : 0;
    int rs = g.bp ? g.bp : g.lin.W;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        // source pixel (y - ky + 1, x - kx + 1)
        bool ok = (unsigned)(iy - ky + 1) < (unsigned)limy && (unsigned)(ix - kx + 1) < (unsigned)limx;
        if (!ok) continue;
        int64_t off = g.fast ? base + (1 - ky) * rs + (1 - kx) : g.lin.at(0, y - ky + 1, xx - kx + 1);
#pragma unroll
        for (int k = 0; k < NCONV; ++k) {
          const float* dp = (k == 0 ? d0 : (k == 1 ? d1 : d2)) + (int64_t)b * g.per_image + off;
#pragma unroll
          for (int co = 0; co < C; ++co) {
            float dv = __ldg(dp + co * cs);
#pragma unroll
            for (int ci = 0; ci < C; ++ci) acc[ci] = fmaf(dv, c_w[((k * C + co) * C + ci) * 9 + ky * 3 + kx], acc[ci]);
          }
        }
      }
    int64_t obase = (int64_t)b * g.per_image + g.lout.base_of(pix);
    const int64_t ocs = g.lout.cstride();
#pragma unroll
    for (int ci = 0; ci < C; ++ci) {
      float* p = dx + obase + ci * ocs;
      *p = accumulate ? *p + acc[ci] : acc[ci];
    }
  }
}

// backward weights.  grid.y enumerates the fused convs k (WIDE: every thread keeps the C*C*9 + C partial sums
// of one conv, so the 9*C taps of x are loaded once per conv) or (k, co) pairs (narrow fallback for C == 4).
// Pixels are grid-strided in x's own layout order.  Here g.lin describes x, g.lout describes dy.
template <int C, bool WIDE>
__global__ void __launch_bounds__(128, 3)
conv3x3_bwd_weight_kernel(const float* __restrict__ x, const float* __restrict__ d0, const float* __restrict__ d1,
                          const float* __restrict__ d2, float* __restrict__ dw0, float* __restrict__ dw1,
                          float* __restrict__ dw2, float* __restrict__ dbias, ConvGeom g) {
  constexpr int NCO = WIDE ? C : 1;
  constexpr int NACC = NCO * (C * 9 + 1);
  const int k = WIDE ? blockIdx.y : blockIdx.y / C;
  const int co0 = WIDE ? 0 : blockIdx.y % C;
  const float* dy = k == 0 ? d0 : (k == 1 ? d1 : d2);
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
  const uint32_t hw = (uint32_t)(g.lin.H * g.lin.W), total = (uint32_t)g.npix_total;     // host checks < 2^31
  const int64_t cs = g.lin.cstride(), dcs = g.lout.cstride();
  const int rs = g.bp ? g.bp : g.lin.W;
  // gather one pixel: dy of the NCO output channels and the 9*C taps of x (zeros outside the border patch)
  auto gather = [&](uint32_t t, float (&dv)[NCO], float (&xv)[C][9]) {
    const uint32_t b = t / hw, pix = t - b * hw;
    int y, xx; g.lin.pixel(pix, y, xx);
    int iy, ix, limy, limx; tap_setup(g, y, xx, iy, ix, limy, limx);
    const int64_t dbase = (int64_t)b * g.per_image + g.lout.at(co0, y, xx);
#pragma unroll
    for (int c = 0; c < NCO; ++c) dv[c] = __ldg(dy + dbase + c * dcs);
    const float* xb = x + (int64_t)b * g.per_image;
    const int64_t base = g.lin.base_of(pix);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const bool ok = (unsigned)(iy + ky - 1) < (unsigned)limy && (unsigned)(ix + kx - 1) < (unsigned)limx;
        const int64_t off = !ok ? 0 : (g.fast ? base + (ky - 1) * rs + (kx - 1) : g.lin.at(0, y + ky - 1, xx + kx - 1));
#pragma unroll
        for (int ci = 0; ci < C; ++ci) xv[ci][ky * 3 + kx] = ok ? __ldg(xb + off + ci * cs) : 0.f;
      }
  };
  auto accumulate = [&](const float (&dv)[NCO], const float (&xv)[C][9]) {
#pragma unroll
    for (int c = 0; c < NCO; ++c) {
      acc[c * (C * 9 + 1) + C * 9] += dv[c];
#pragma unroll
      for (int ci = 0; ci < C; ++ci)
#pragma unroll
        for (int tp = 0; tp < 9; ++tp)
          acc[c * (C * 9 + 1) + ci * 9 + tp] = fmaf(dv[c], xv[ci][tp], acc[c * (C * 9 + 1) + ci * 9 + tp]);
    }
  };
  // two pixels per iteration: both gathers are issued before either accumulation (more loads in flight per warp)
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += 2 * stride) {
    float dv0[NCO], xv0[C][9], dv1[NCO], xv1[C][9];
    gather(t, dv0, xv0);
    const bool second = t + stride < total;
    if (second) gather(t + stride, dv1, xv1);
    accumulate(dv0, xv0);
    if (second) accumulate(dv1, xv1);
  }
  __shared__ float red[NACC][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float v = warp_sum(acc[i]);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < NACC) {
    float s = 0.f;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) s += red[threadIdx.x][wv];
    const int c = threadIdx.x / (C * 9 + 1), e = threadIdx.x % (C * 9 + 1);
    const int co = co0 + c;
    float* dw = k == 0 ? dw0 : (k == 1 ? dw1 : dw2);        // gradient block [C][C][3][3] of conv k
    if (e < C * 9) atomicAdd(dw + ((int64_t)co * C) * 9 + e, s);
    else if (dbias) atomicAdd(dbias + k * C + co, s);
  }
}

// ------------------------------------------------------------------ backward weights, patch-tiled fast path
// q/k/v convs with x and dy both in patch layout P == border patch (model.py:152-154 backward).  A CTA walks items of
// 512 pixels (K whole patches, or half a 32x32 patch); the x tile of an item is staged once in shared memory with a
// zero halo, so the 9 taps are fixed-offset LDS with no border predicates, and each thread owns a run of 4
// consecutive pixels of one row: per channel and tap row one LDS.128 + LDS.64 feed 36 FFMAs.  The next item's x and dy
// are prefetched into registers while the current one is accumulated.  grid.y = fused conv k (C*C*9 + C partial
// sums per thread); the three CTAs of an item run side by side on one SM and share its x tile through L1/L2.
template <int C, int P>
struct WgTile {
  static constexpr int PS = P == 4 ? 2 : (P == 8 ? 3 : (P == 16 ? 4 : 5));
  static constexpr int PP = P * P;
  static constexpr int ITEM = 512;                                  // pixels per item = 4 * blockDim
  static constexpr int R = (PP >= ITEM) ? ITEM / P : P;             // rows per unit (power of two)
  static constexpr int RS = R == 4 ? 2 : (R == 8 ? 3 : 4);
  static constexpr int UPP = P / R;                                 // units per patch
  static constexpr int K = ITEM / (R * P);                          // units per item
  static constexpr int ROWS = R + 2;
  static constexpr int PITCH = P == 4 ? 12 : (P < 32 ? P + 32 : 36);  // % 4 == 0; spreads a phase's rows over banks
  static constexpr int F4R = P / 4;
  static constexpr int N4 = K * C * ROWS * F4R;                     // float4 loads per item
  static constexpr int NF = (N4 + 127) / 128;
  static constexpr int SMEM = K * C * ROWS * PITCH;
  static_assert(R == 4 || R == 8 || R == 16, "rows per unit");
};

template <int C, int P>
__global__ void __launch_bounds__(128, 3)
conv3x3_wgrad_patch_kernel(const float* __restrict__ x, const float* __restrict__ d0, const float* __restrict__ d1,
                           const float* __restrict__ d2, float* __restrict__ dw0, float* __restrict__ dw1,
                           float* __restrict__ dw2, float* __restrict__ dbias, int64_t total_units) {
  using T = WgTile<C, P>;
  constexpr int NACC = C * (C * 9 + 1);
  __shared__ __align__(16) float xs[T::SMEM];
  __shared__ float red[NACC][4];
  const int k = blockIdx.y, tid = threadIdx.x;
  const float* __restrict__ dy = k == 0 ? d0 : (k == 1 ? d1 : d2);
  for (int i = tid; i < T::SMEM; i += 128) xs[i] = 0.f;
  float acc[C][C][9], accb[C];
#pragma unroll
  for (int co = 0; co < C; ++co) {
    accb[co] = 0.f;
#pragma unroll
    for (int ci = 0; ci < C; ++ci)
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[co][ci][t] = 0.f;
  }
  // this thread's run of 4 pixels inside an item: unit slot s, row i of the unit, columns j0..j0+3
  const int j0 = (tid * 4) & (P - 1), rowlin = (tid * 4) >> T::PS, s = rowlin >> T::RS, i = rowlin & (T::R - 1);
  const int64_t n_items = (total_units + T::K - 1) / T::K;
  float4 xpre[T::NF], dpre[C];
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto prefetch = [&](int64_t it) {
#pragma unroll
    for (int f = 0; f < T::NF; ++f) {
      const int e = tid + 128 * f;
      xpre[f] = zero4;
      if (e < T::N4) {
        const int c4 = e % T::F4R, rr = (e / T::F4R) % T::ROWS, ci = (e / (T::F4R * T::ROWS)) % C,
                  ss = e / (T::F4R * T::ROWS * C);
        const int64_t u = it * T::K + ss;
        const int gr = (int)(u % T::UPP) * T::R + rr - 1;
        if (u < total_units && (unsigned)gr < (unsigned)P)
          xpre[f] = __ldg(reinterpret_cast<const float4*>(x + (u / T::UPP) * (C * T::PP) + ci * T::PP + gr * P + 4 * c4));
      }
    }
    const int64_t u = it * T::K + s;
    const bool ok = u < total_units;
    const float* dp = dy + (u / T::UPP) * (C * T::PP) + ((int)(u % T::UPP) * T::R + i) * P + j0;
#pragma unroll
    for (int co = 0; co < C; ++co) dpre[co] = ok ? __ldg(reinterpret_cast<const float4*>(dp + co * T::PP)) : zero4;
  };
  int64_t it = blockIdx.x;
  if (it < n_items) prefetch(it);
  const float* xb = xs + (s * C * T::ROWS + i) * T::PITCH + j0;
  for (; it < n_items; it += gridDim.x) {
    __syncthreads();                                   // everyone is done reading the previous tile
#pragma unroll
    for (int f = 0; f < T::NF; ++f) {
      const int e = tid + 128 * f;
      if (e < T::N4) {
        const int c4 = e % T::F4R, row = e / T::F4R;    // row = (ss*C + ci)*ROWS + rr
        float* dst = xs + row * T::PITCH + 1 + 4 * c4;
        dst[0] = xpre[f].x; dst[1] = xpre[f].y; dst[2] = xpre[f].z; dst[3] = xpre[f].w;
      }
    }
    float4 d[C];
#pragma unroll
    for (int co = 0; co < C; ++co) d[co] = dpre[co];
    __syncthreads();
    if (it + gridDim.x < n_items) prefetch(it + gridDim.x);
#pragma unroll
    for (int co = 0; co < C; ++co) accb[co] += (d[co].x + d[co].y) + (d[co].z + d[co].w);
#pragma unroll
    for (int ci = 0; ci < C; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float* r = xb + (ci * T::ROWS + ky) * T::PITCH;
        const float4 a = *reinterpret_cast<const float4*>(r);
        const float2 b = *reinterpret_cast<const float2*>(r + 4);
        const float w[6] = {a.x, a.y, a.z, a.w, b.x, b.y};
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int co = 0; co < C; ++co) {
            float v = acc[co][ci][ky * 3 + kx];
            v = fmaf(d[co].x, w[kx], v); v = fmaf(d[co].y, w[kx + 1], v);
            v = fmaf(d[co].z, w[kx + 2], v); v = fmaf(d[co].w, w[kx + 3], v);
            acc[co][ci][ky * 3 + kx] = v;
          }
      }
  }
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int co = 0; co < C; ++co) {
#pragma unroll
    for (int ci = 0; ci < C; ++ci)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float v = warp_sum(acc[co][ci][t]);
        if (lane == 0) red[co * (C * 9 + 1) + ci * 9 + t][warp] = v;
      }
    const float v = warp_sum(accb[co]);
    if (lane == 0) red[co * (C * 9 + 1) + C * 9][warp] = v;
  }
  __syncthreads();
  if (tid < NACC) {
    const float sum = (red[tid][0] + red[tid][1]) + (red[tid][2] + red[tid][3]);
    const int co = tid / (C * 9 + 1), e = tid % (C * 9 + 1);
    float* dw = k == 0 ? dw0 : (k == 1 ? dw1 : dw2);
    if (e < C * 9) atomicAdd(dw + ((int64_t)co * C) * 9 + e, sum);
    else if (dbias) atomicAdd(dbias + k * C + co, sum);
  }
}

template <int C, int P>
static void launch_wgrad_patch(const float* x, const float* d0, const float* d1, const float* d2, int nconv,
                               float* dw0, float* dw1, float* dw2, float* dbias, int64_t patches, cudaStream_t s) {
  using T = WgTile<C, P>;
  const int64_t units = patches * T::UPP, items = (units + T::K - 1) / T::K;
  const int per_sm = nconv == 3 ? 1 : (nconv == 2 ? 2 : 3);        // 3 resident CTAs per SM over (x, y)
  const int bx = (int)std::min<int64_t>(items, (int64_t)sm_count() * per_sm);
  conv3x3_wgrad_patch_kernel<C, P><<<dim3(bx, nconv), 128, 0, s>>>(x, d0, d1, d2, dw0, dw1, dw2, dbias, units);
}

// ------------------------------------------------------------------ forward / backward-data, patch-tiled fast paths
// Same tiling as the weight-gradient kernel above: items of 512 pixels, source planes staged in shared memory with a
// zero halo (x: C planes; dy: NCONV*C planes), one run of 4 consecutive pixels per thread, filters as constant-bank
// FFMA operands, float4 stores.  Used when source, destination and border patch agree (the q/k/v convs).
template <int P, int NPL>
struct PatchTile {
  static constexpr int PS = P == 4 ? 2 : (P == 8 ? 3 : (P == 16 ? 4 : 5));
  static constexpr int PP = P * P;
  static constexpr int ITEM = 512;
  static constexpr int R = (PP >= ITEM) ? ITEM / P : P;
  static constexpr int RS = R == 4 ? 2 : (R == 8 ? 3 : 4);
  static constexpr int UPP = P / R;
  static constexpr int K = ITEM / (R * P);
  static constexpr int ROWS = R + 2;
  static constexpr int PITCH = P + 4;
  static constexpr int F4R = P / 4;
  static constexpr int N4 = K * NPL * ROWS * F4R;
  static constexpr int NF = (N4 + 127) / 128;
  static constexpr int SMEM = K * NPL * ROWS * PITCH;
};

// stage helpers shared by both kernels: plane pl of unit u lives at src(pl) + patch*CPP + (pl % C)*PP
template <int C, int NCONV, int P, bool DGRAD>
__global__ void __launch_bounds__(128, DGRAD ? 3 : 4)
conv3x3_patch_kernel(const float* __restrict__ s0, const float* __restrict__ s1, const float* __restrict__ s2,
                     float* __restrict__ o0, float* __restrict__ o1, float* __restrict__ o2,
                     int64_t total_units, int accumulate) {
  constexpr int NPL = DGRAD ? NCONV * C : C;
  using T = PatchTile<P, NPL>;
  __shared__ __align__(16) float xs[T::SMEM];
  const int tid = threadIdx.x;
  for (int i = tid; i < T::SMEM; i += 128) xs[i] = 0.f;
  const int j0 = (tid * 4) & (P - 1), rowlin = (tid * 4) >> T::PS, s = rowlin >> T::RS, i = rowlin & (T::R - 1);
  const int64_t n_items = (total_units + T::K - 1) / T::K;
  float4 pre[T::NF];
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto prefetch = [&](int64_t it) {
#pragma unroll
    for (int f = 0; f < T::NF; ++f) {
      const int e = tid + 128 * f;
      pre[f] = zero4;
      if (e < T::N4) {
        const int c4 = e % T::F4R, rr = (e / T::F4R) % T::ROWS, pl = (e / (T::F4R * T::ROWS)) % NPL,
                  ss = e / (T::F4R * T::ROWS * NPL);
        const int64_t u = it * T::K + ss;
        const int gr = (int)(u % T::UPP) * T::R + rr - 1;
        const float* src = DGRAD ? (pl / C == 0 ? s0 : (pl / C == 1 ? s1 : s2)) : s0;
        if (u < total_units && (unsigned)gr < (unsigned)P)
          pre[f] = __ldg(reinterpret_cast<const float4*>(src + (u / T::UPP) * (C * T::PP) + (pl % C) * T::PP + gr * P + 4 * c4));
      }
    }
  };
  int64_t it = blockIdx.x;
  if (it < n_items) prefetch(it);
  const float* xb = xs + (s * NPL * T::ROWS + i) * T::PITCH + j0;
  for (; it < n_items; it += gridDim.x) {
    __syncthreads();
#pragma unroll
    for (int f = 0; f < T::NF; ++f) {
      const int e = tid + 128 * f;
      if (e < T::N4) {
        float* dst = xs + (e / T::F4R) * T::PITCH + 1 + 4 * (e % T::F4R);
        dst[0] = pre[f].x; dst[1] = pre[f].y; dst[2] = pre[f].z; dst[3] = pre[f].w;
      }
    }
    __syncthreads();
    const int64_t u = it * T::K + s;
    if (it + gridDim.x < n_items) prefetch(it + gridDim.x);
    constexpr int NOUT = DGRAD ? C : NCONV * C;
    float4 acc[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) { const float b = DGRAD ? 0.f : c_b[o]; acc[o] = make_float4(b, b, b, b); }
#pragma unroll
    for (int pl = 0; pl < NPL; ++pl)
#pragma unroll
      for (int wy = 0; wy < 3; ++wy) {
        const float* r = xb + (pl * T::ROWS + wy) * T::PITCH;
        const float4 a = *reinterpret_cast<const float4*>(r);
        const float2 b = *reinterpret_cast<const float2*>(r + 4);
        const float w[6] = {a.x, a.y, a.z, a.w, b.x, b.y};
#pragma unroll
        for (int wx = 0; wx < 3; ++wx)
#pragma unroll
          for (int o = 0; o < NOUT; ++o) {
            // forward: plane = ci, o = (k, co), tap (wy, wx).  dgrad: plane = (k, co), o = ci, tap (2-wy, 2-wx).
            const float wt = DGRAD ? c_w[(pl * C + o) * 9 + (2 - wy) * 3 + (2 - wx)]
                                   : c_w[(o * C + pl) * 9 + wy * 3 + wx];
            acc[o].x = fmaf(w[wx], wt, acc[o].x); acc[o].y = fmaf(w[wx + 1], wt, acc[o].y);
            acc[o].z = fmaf(w[wx + 2], wt, acc[o].z); acc[o].w = fmaf(w[wx + 3], wt, acc[o].w);
          }
      }
    if (u < total_units) {
      const int64_t off = (u / T::UPP) * (C * T::PP) + ((int)(u % T::UPP) * T::R + i) * P + j0;
#pragma unroll
      for (int o = 0; o < NOUT; ++o) {
        float* op = DGRAD ? o0 : (o / C == 0 ? o0 : (o / C == 1 ? o1 : o2));
        float4* dst = reinterpret_cast<float4*>(op + off + (o % C) * T::PP);
        float4 v = acc[o];
        if (DGRAD && accumulate) { const float4 t = *dst; v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
        *dst = v;
      }
    }
  }
}

template <int C, int NCONV, bool DGRAD>
static bool launch_patch_conv(int p, const float* s0, const float* s1, const float* s2, float* o0, float* o1, float* o2,
                              int64_t patches, int accumulate, cudaStream_t st) {
  auto go = [&](auto tag) {
    constexpr int P = decltype(tag)::value;
    using T = PatchTile<P, DGRAD ? NCONV * C : C>;
    const int64_t units = patches * T::UPP, items = (units + T::K - 1) / T::K;
    const int bx = (int)std::min<int64_t>(items, (int64_t)sm_count() * (DGRAD ? 3 : 4));
    conv3x3_patch_kernel<C, NCONV, P, DGRAD><<<bx, 128, 0, st>>>(s0, s1, s2, o0, o1, o2, units, accumulate);
  };
  constexpr int NPL = DGRAD ? NCONV * C : C;
  if constexpr (PatchTile<4, NPL>::SMEM * 4 <= 48 * 1024) {        // static shared memory limit
    if (p == 4) { go(std::integral_constant<int, 4>()); return true; }
  }
  if (p == 8) { go(std::integral_constant<int, 8>()); return true; }
  if (p == 16) { go(std::integral_constant<int, 16>()); return true; }
  if (p == 32) { go(std::integral_constant<int, 32>()); return true; }
  return false;
}

// ------------------------------------------------------------------ q/k/v convs as implicit GEMMs on warp MMAs (TF32 class)
// Tensor-core formulation of the patch-tiled forward conv for the tf32 / bf16 precision modes (the FP32 mode keeps the
// FFMA kernel above).  Per 8 pixels of a patch row:  out^T[n_out, pixel] = W[n_out, k] . X[k, pixel],  n_out = (conv, c_out)
// (NCONV * C <= 9 rows of the 16-row MMA tile), k = (c_in, tap) (C * 9 <= 27 of 32: four m16n8k8 TF32 MMAs), X gathered
// from the same zero-haloed shared-memory tile as the FFMA kernel: B fragment = two LDS.32 per k-step at per-lane constant
// tap offsets, A fragments = the filters, 16 registers loaded once per CTA.  A lane ends up with two consecutive pixels of
// output plane gid (and of plane gid + 8): 8-byte stores, eight full 32-byte sectors per store instruction.
// 64 MMAs + 128 LDS per 128 pixels instead of 972 FFMAs.
__device__ __forceinline__ uint32_t conv_tf32(float x) {
  uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r;
}
__device__ __forceinline__ void conv_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int C, int NCONV, int P>
__global__ void __launch_bounds__(128, 4)
conv3x3_patch_mma_fwd_kernel(const float* __restrict__ x, float* __restrict__ o0, float* __restrict__ o1,
                             float* __restrict__ o2, int64_t total_units) {
  using T = PatchTile<P, C>;
  static_assert(P >= 8, "an 8-pixel MMA column tile must lie inside one patch row");
  constexpr int NOUT = NCONV * C, KTOT = C * 9, KSTEPS = (KTOT + 7) / 8;
  __shared__ __align__(16) float xs[T::SMEM];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
  for (int i = tid; i < T::SMEM; i += 128) xs[i] = 0.f;
  // filters as A fragments: rows n_out = gid / gid + 8, columns k = 8 s + tig / + 4
  uint32_t af[KSTEPS][4];
#pragma unroll
  for (int s = 0; s < KSTEPS; ++s) {
    const int k0 = 8 * s + tig, k1 = k0 + 4;
    af[s][0] = (gid < NOUT && k0 < KTOT) ? conv_tf32(c_w[gid * KTOT + k0]) : 0u;
    af[s][1] = (gid + 8 < NOUT && k0 < KTOT) ? conv_tf32(c_w[(gid + 8) * KTOT + k0]) : 0u;
    af[s][2] = (gid < NOUT && k1 < KTOT) ? conv_tf32(c_w[gid * KTOT + k1]) : 0u;
    af[s][3] = (gid + 8 < NOUT && k1 < KTOT) ? conv_tf32(c_w[(gid + 8) * KTOT + k1]) : 0u;
  }
  // tap offsets of this lane's k indices inside the haloed tile (pads read tap 0: their filter entries are zero)
  int koff[KSTEPS][2];
#pragma unroll
  for (int s = 0; s < KSTEPS; ++s)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int k = 8 * s + tig + 4 * h;
      if (k >= KTOT) k = 0;
      const int ci = k / 9, tap = k - ci * 9, dy = tap / 3, dx = tap - dy * 3;
      koff[s][h] = (ci * T::ROWS + dy) * T::PITCH + dx;
    }
  const float bias0 = gid < NOUT ? c_b[gid] : 0.f, bias1 = gid + 8 < NOUT ? c_b[gid + 8] : 0.f;
  float* const ob0 = gid < NOUT ? ((gid / C == 0 ? o0 : (gid / C == 1 ? o1 : o2)) + (gid % C) * T::PP) : nullptr;
  float* const ob1 = gid + 8 < NOUT ? (((gid + 8) / C == 0 ? o0 : ((gid + 8) / C == 1 ? o1 : o2)) + ((gid + 8) % C) * T::PP) : nullptr;

  const int64_t n_items = (total_units + T::K - 1) / T::K;
  float4 pre[T::NF];
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto prefetch = [&](int64_t it) {
#pragma unroll
    for (int f = 0; f < T::NF; ++f) {
      const int e = tid + 128 * f;
      pre[f] = zero4;
      if (e < T::N4) {
        const int c4 = e % T::F4R, rr = (e / T::F4R) % T::ROWS, pl = (e / (T::F4R * T::ROWS)) % C,
                  ss = e / (T::F4R * T::ROWS * C);
        const int64_t u = it * T::K + ss;
        const int gr = (int)(u % T::UPP) * T::R + rr - 1;
        if (u < total_units && (unsigned)gr < (unsigned)P)
          pre[f] = __ldg(reinterpret_cast<const float4*>(x + (u / T::UPP) * (C * T::PP) + pl * T::PP + gr * P + 4 * c4));
      }
    }
  };
  int64_t it = blockIdx.x;
  if (it < n_items) prefetch(it);
  for (; it < n_items; it += gridDim.x) {
    __syncthreads();
#pragma unroll
    for (int f = 0; f < T::NF; ++f) {
      const int e = tid + 128 * f;
      if (e < T::N4) {
        float* dst = xs + (e / T::F4R) * T::PITCH + 1 + 4 * (e % T::F4R);
        dst[0] = pre[f].x; dst[1] = pre[f].y; dst[2] = pre[f].z; dst[3] = pre[f].w;
      }
    }
    __syncthreads();
    if (it + gridDim.x < n_items) prefetch(it + gridDim.x);
    // this warp: pixels [warp * 128, warp * 128 + 128) of the item = 16 column tiles of 8 pixels
#pragma unroll 2
    for (int t = 0; t < 16; ++t) {
      const int pix = warp * 128 + t * 8;                         // first pixel of the tile inside the item
      const int col = pix & (P - 1), rowlin = pix >> T::PS, slot = rowlin >> T::RS, i = rowlin & (T::R - 1);
      const int64_t u = it * T::K + slot;
      const float* xb = xs + (slot * C * T::ROWS + i) * T::PITCH + col + gid;
      float c[4] = {bias0, bias0, bias1, bias1};
#pragma unroll
      for (int s = 0; s < KSTEPS; ++s) {
        const uint32_t b0 = __float_as_uint(xb[koff[s][0]]) + 0x1000u, b1 = __float_as_uint(xb[koff[s][1]]) + 0x1000u;
        conv_mma(c, af[s], b0, b1);
      }
      if (u < total_units) {
        const int64_t off = (u / T::UPP) * (C * T::PP) + ((int)(u % T::UPP) * T::R + i) * P + col + 2 * tig;
        if (ob0) *reinterpret_cast<float2*>(ob0 + off) = make_float2(c[0], c[1]);
        if (ob1) *reinterpret_cast<float2*>(ob1 + off) = make_float2(c[2], c[3]);
      }
    }
  }
}

template <int C, int NCONV>
static bool launch_patch_mma_fwd(int p, const float* x, float* o0, float* o1, float* o2, int64_t patches, cudaStream_t st) {
  auto go = [&](auto tag) {
    constexpr int P = decltype(tag)::value;
    using T = PatchTile<P, C>;
    const int64_t units = patches * T::UPP, items = (units + T::K - 1) / T::K;
    const int bx = (int)std::min<int64_t>(items, (int64_t)sm_count() * 4);
    conv3x3_patch_mma_fwd_kernel<C, NCONV, P><<<bx, 128, 0, st>>>(x, o0, o1, o2, units);
  };
  if (p == 8) { go(std::integral_constant<int, 8>()); return true; }
  if (p == 16) { go(std::integral_constant<int, 16>()); return true; }
  if (p == 32) { go(std::integral_constant<int, 32>()); return true; }
  return false;
}

// Data gradient on warp MMAs.  The roles are swapped with respect to the forward kernel because there are only C <= 3
// output planes:  dx[pixel, ci] = DY[pixel, k] . W'[k, ci]  with M = 16 pixels (two runs of 8 of one patch row, or of two
// consecutive rows when P == 8), k = (source plane (conv, c_out), tap) (NCONV * C * 9 <= 81 of 88: eleven k-steps),
// N = ci (C of the 8 columns).  A fragment = four LDS.32 per k-step at per-lane constant tap offsets of the zero-haloed
// dy tile, B fragments = the flipped filters (22 registers, loaded once per CTA): 11 MMAs + 44 LDS per 16 pixels instead
// of 3888 FFMAs.  A lane with tig == 0 ends up with planes 0 / 1 of pixels gid and gid + 8, tig == 1 with plane 2.
template <int C, int NCONV, int P>
__global__ void __launch_bounds__(128, 3)
conv3x3_patch_mma_dgrad_kernel(const float* __restrict__ s0, const float* __restrict__ s1, const float* __restrict__ s2,
                               float* __restrict__ dx, int64_t total_units, int accumulate) {
  constexpr int NPL = NCONV * C, KTOT = NPL * 9, KSTEPS = (KTOT + 7) / 8;
  using T = PatchTile<P, NPL>;
  static_assert(P >= 8 && C <= 3, "16-pixel MMA row tiles; planes 0..2");
  __shared__ __align__(16) float xs[T::SMEM];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
  for (int i = tid; i < T::SMEM; i += 128) xs[i] = 0.f;
  // flipped filters as B fragments: rows k = 8 s + tig / + 4, column n = ci = gid
  uint32_t bw[KSTEPS][2];
  int koff[KSTEPS][2];
#pragma unroll
  for (int s = 0; s < KSTEPS; ++s)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = 8 * s + tig + 4 * h;
      const bool kv = k < KTOT;
      const int kk = kv ? k : 0;
      const int pl = kk / 9, tap = kk - pl * 9, wy = tap / 3, wx = tap - wy * 3;
      koff[s][h] = (pl * T::ROWS + wy) * T::PITCH + wx;
      bw[s][h] = (kv && gid < C) ? conv_tf32(c_w[(pl * C + gid) * 9 + (2 - wy) * 3 + (2 - wx)]) : 0u;
    }
  const int64_t n_items = (total_units + T::K - 1) / T::K;
  float4 pre[T::NF];
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto prefetch = [&](int64_t it) {
#pragma unroll
    for (int f = 0; f < T::NF; ++f) {
      const int e = tid + 128 * f;
      pre[f] = zero4;
      if (e < T::N4) {
        const int c4 = e % T::F4R, rr = (e / T::F4R) % T::ROWS, pl = (e / (T::F4R * T::ROWS)) % NPL,
                  ss = e / (T::F4R * T::ROWS * NPL);
        const int64_t u = it * T::K + ss;
        const int gr = (int)(u % T::UPP) * T::R + rr - 1;
        const float* src = pl / C == 0 ? s0 : (pl / C == 1 ? s1 : s2);
        if (u < total_units && (unsigned)gr < (unsigned)P)
          pre[f] = __ldg(reinterpret_cast<const float4*>(src + (u / T::UPP) * (C * T::PP) + (pl % C) * T::PP + gr * P + 4 * c4));
      }
    }
  };
  // the two pixel runs of a 16-pixel tile: 8 columns apart in one row (P >= 16) or the same columns of the next row (P == 8)
  constexpr int SECOND_PIX = 8;                                  // pixel index distance inside the item
  int64_t it = blockIdx.x;
  if (it < n_items) prefetch(it);
  for (; it < n_items; it += gridDim.x) {
    __syncthreads();
#pragma unroll
    for (int f = 0; f < T::NF; ++f) {
      const int e = tid + 128 * f;
      if (e < T::N4) {
        float* dst = xs + (e / T::F4R) * T::PITCH + 1 + 4 * (e % T::F4R);
        dst[0] = pre[f].x; dst[1] = pre[f].y; dst[2] = pre[f].z; dst[3] = pre[f].w;
      }
    }
    __syncthreads();
    if (it + gridDim.x < n_items) prefetch(it + gridDim.x);
#pragma unroll 2
    for (int t = 0; t < 8; ++t) {
      const int pixA = warp * 128 + t * 16 + gid, pixB = pixA + SECOND_PIX;
      const int colA = pixA & (P - 1), rlA = pixA >> T::PS, slotA = rlA >> T::RS, iA = rlA & (T::R - 1);
      const int colB = pixB & (P - 1), rlB = pixB >> T::PS, slotB = rlB >> T::RS, iB = rlB & (T::R - 1);
      const float* xa = xs + (slotA * NPL * T::ROWS + iA) * T::PITCH + colA;
      const float* xb = xs + (slotB * NPL * T::ROWS + iB) * T::PITCH + colB;
      float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int s = 0; s < KSTEPS; ++s) {
        uint32_t a[4];
        a[0] = __float_as_uint(xa[koff[s][0]]) + 0x1000u; a[1] = __float_as_uint(xb[koff[s][0]]) + 0x1000u;
        a[2] = __float_as_uint(xa[koff[s][1]]) + 0x1000u; a[3] = __float_as_uint(xb[koff[s][1]]) + 0x1000u;
        conv_mma(c, a, bw[s][0], bw[s][1]);
      }
      // c[0], c[1]: planes 2 tig, 2 tig + 1 of pixel A; c[2], c[3]: the same planes of pixel B
      const int64_t uA = it * T::K + slotA, uB = it * T::K + slotB;
      if (2 * tig < C) {
        float* pa = dx + (uA / T::UPP) * (C * T::PP) + ((int)(uA % T::UPP) * T::R + iA) * P + colA + (2 * tig) * T::PP;
        float* pb = dx + (uB / T::UPP) * (C * T::PP) + ((int)(uB % T::UPP) * T::R + iB) * P + colB + (2 * tig) * T::PP;
        if (uA < total_units) {
          pa[0] = accumulate ? pa[0] + c[0] : c[0];
          if (2 * tig + 1 < C) pa[T::PP] = accumulate ? pa[T::PP] + c[1] : c[1];
        }
        if (uB < total_units) {
          pb[0] = accumulate ? pb[0] + c[2] : c[2];
          if (2 * tig + 1 < C) pb[T::PP] = accumulate ? pb[T::PP] + c[3] : c[3];
        }
      }
    }
  }
}

template <int C, int NCONV>
static bool launch_patch_mma_dgrad(int p, const float* s0, const float* s1, const float* s2, float* dx, int64_t patches,
                                   int accumulate, cudaStream_t st) {
  auto go = [&](auto tag) {
    constexpr int P = decltype(tag)::value;
    using T = PatchTile<P, NCONV * C>;
    const int64_t units = patches * T::UPP, items = (units + T::K - 1) / T::K;
    const int bx = (int)std::min<int64_t>(items, (int64_t)sm_count() * 3);
    conv3x3_patch_mma_dgrad_kernel<C, NCONV, P><<<bx, 128, 0, st>>>(s0, s1, s2, dx, units, accumulate);
  };
  if (p == 8) { go(std::integral_constant<int, 8>()); return true; }
  if (p == 16) { go(std::integral_constant<int, 16>()); return true; }
  if (p == 32) { go(std::integral_constant<int, 32>()); return true; }
  return false;
}

// Weight gradient on warp MMAs:  dW[n_out, k] = sum over pixels DY[n_out, pixel] . X[k, pixel]:  M = n_out = (conv, c_out)
// (NCONV * C <= 9 of 16 rows -- one CTA serves all fused convs, x is staged once), N = k = (c_in, tap) (C * 9 <= 27 of 32:
// four n-tiles), contraction over 8 pixels of a patch row per MMA.  A fragment = dy from a shared-memory tile laid out
// [plane][pixel of the item] (pitch 516: the eight planes of a quarter-warp land on disjoint banks), B fragment = x gathered
// from the zero-haloed tile at per-lane constant tap offsets.  4 MMAs + 12 LDS per 8 pixels and conv triple instead of
// 1944 FFMAs; the accumulators (16 registers) live for the whole CTA and are reduced once (shared-memory atomics, then one
// global atomicAdd per filter entry).  Bias gradients: plain sums of the staged dy quads.
template <int C, int NCONV, int P>
__global__ void __launch_bounds__(128, 3)
conv3x3_wgrad_mma_kernel(const float* __restrict__ x, const float* __restrict__ d0, const float* __restrict__ d1,
                         const float* __restrict__ d2, float* __restrict__ dw0, float* __restrict__ dw1,
                         float* __restrict__ dw2, float* __restrict__ dbias, int64_t total_units) {
  using T = PatchTile<P, C>;
  constexpr int NOUT = NCONV * C, KTOT = C * 9, NT = (KTOT + 7) / 8, DP = T::ITEM + 4;
  static_assert(P >= 8 && T::ITEM == 512, "8-pixel MMA k-steps inside one patch row; 4 pixels per thread");
  __shared__ __align__(16) float xs[T::SMEM];
  __shared__ __align__(16) float ds[(NOUT + 1) * DP];             // plane NOUT stays zero (rows of the tile that do not exist)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
  for (int i = tid; i < T::SMEM; i += 128) xs[i] = 0.f;
  for (int i = tid; i < (NOUT + 1) * DP; i += 128) ds[i] = 0.f;
  int koff[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    int k = 8 * nt + gid;
    if (k >= KTOT) k = 0;                                         // pad columns: finite values nobody reads
    const int ci = k / 9, tap = k - ci * 9, dy = tap / 3, dxx = tap - dy * 3;
    koff[nt] = (ci * T::ROWS + dy) * T::PITCH + dxx + tig;
  }
  const float* const da0 = ds + (gid < NOUT ? gid : NOUT) * DP + tig;
  const float* const da1 = ds + (gid + 8 < NOUT ? gid + 8 : NOUT) * DP + tig;
  float acc[NT][4], accb[NOUT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = 0.f; acc[nt][1] = 0.f; acc[nt][2] = 0.f; acc[nt][3] = 0.f; }
#pragma unroll
  for (int o = 0; o < NOUT; ++o) accb[o] = 0.f;
  // this thread's dy quad inside an item: unit slot s, row i of the unit, columns j0..j0+3 (item pixel index 4 * tid)
  const int j0 = (tid * 4) & (P - 1), rowlin = (tid * 4) >> T::PS, sq = rowlin >> T::RS, iq = rowlin & (T::R - 1);
  const int64_t n_items = (total_units + T::K - 1) / T::K;
  float4 xpre[T::NF], dpre[NOUT];
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto prefetch = [&](int64_t it) {
#pragma unroll
    for (int f = 0; f < T::NF; ++f) {
      const int e = tid + 128 * f;
      xpre[f] = zero4;
      if (e < T::N4) {
        const int c4 = e % T::F4R, rr = (e / T::F4R) % T::ROWS, ci = (e / (T::F4R * T::ROWS)) % C,
                  ss = e / (T::F4R * T::ROWS * C);
        const int64_t u = it * T::K + ss;
        const int gr = (int)(u % T::UPP) * T::R + rr - 1;
        if (u < total_units && (unsigned)gr < (unsigned)P)
          xpre[f] = __ldg(reinterpret_cast<const float4*>(x + (u / T::UPP) * (C * T::PP) + ci * T::PP + gr * P + 4 * c4));
      }
    }
    const int64_t u = it * T::K + sq;
    const bool ok = u < total_units;
    const int64_t off = (u / T::UPP) * (C * T::PP) + ((int)(u % T::UPP) * T::R + iq) * P + j0;
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
      const float* dy = o / C == 0 ? d0 : (o / C == 1 ? d1 : d2);
      dpre[o] = ok ? __ldg(reinterpret_cast<const float4*>(dy + off + (o % C) * T::PP)) : zero4;
    }
  };
  int64_t it = blockIdx.x;
  if (it < n_items) prefetch(it);
  for (; it < n_items; it += gridDim.x) {
    __syncthreads();                                   // everyone is done reading the previous tiles
#pragma unroll
    for (int f = 0; f < T::NF; ++f) {
      const int e = tid + 128 * f;
      if (e < T::N4) {
        float* dst = xs + (e / T::F4R) * T::PITCH + 1 + 4 * (e % T::F4R);
        dst[0] = xpre[f].x; dst[1] = xpre[f].y; dst[2] = xpre[f].z; dst[3] = xpre[f].w;
      }
    }
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
      *reinterpret_cast<float4*>(ds + o * DP + 4 * tid) = dpre[o];
      accb[o] += (dpre[o].x + dpre[o].y) + (dpre[o].z + dpre[o].w);
    }
    __syncthreads();
    if (it + gridDim.x < n_items) prefetch(it + gridDim.x);
#pragma unroll 4
    for (int t = 0; t < 16; ++t) {
      const int pix = warp * 128 + t * 8;                         // first pixel of the k-step inside the item
      const int col = pix & (P - 1), rl = pix >> T::PS, slot = rl >> T::RS, i = rl & (T::R - 1);
      const float* xb = xs + (slot * C * T::ROWS + i) * T::PITCH + col;
      uint32_t a[4];
      a[0] = __float_as_uint(da0[pix]) + 0x1000u; a[1] = __float_as_uint(da1[pix]) + 0x1000u;
      a[2] = __float_as_uint(da0[pix + 4]) + 0x1000u; a[3] = __float_as_uint(da1[pix + 4]) + 0x1000u;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const uint32_t b0 = __float_as_uint(xb[koff[nt]]) + 0x1000u, b1 = __float_as_uint(xb[koff[nt] + 4]) + 0x1000u;
        conv_mma(acc[nt], a, b0, b1);
      }
    }
  }
  // ---- reduce: the four warps' fragments through shared-memory atomics, then one global atomic per filter entry
  __syncthreads();
  float* red = ds;                                                // [16][NT * 8] + [NOUT]
  for (int i = tid; i < 16 * NT * 8 + NOUT; i += 128) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int kc = 8 * nt + 2 * tig;
    atomicAdd(red + gid * (NT * 8) + kc, acc[nt][0]);
    atomicAdd(red + gid * (NT * 8) + kc + 1, acc[nt][1]);
    if (gid + 8 < NOUT) {
      atomicAdd(red + (gid + 8) * (NT * 8) + kc, acc[nt][2]);
      atomicAdd(red + (gid + 8) * (NT * 8) + kc + 1, acc[nt][3]);
    }
  }
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
    const float v = warp_sum(accb[o]);
    if (lane == 0) atomicAdd(red + 16 * NT * 8 + o, v);
  }
  __syncthreads();
  for (int e = tid; e < NOUT * KTOT + NOUT; e += 128) {
    if (e < NOUT * KTOT) {
      const int o = e / KTOT, kidx = e - o * KTOT, k = o / C, co = o - k * C;
      float* dw = k == 0 ? dw0 : (k == 1 ? dw1 : dw2);
      atomicAdd(dw + (int64_t)co * KTOT + kidx, red[o * (NT * 8) + kidx]);
    } else if (dbias) {
      atomicAdd(dbias + (e - NOUT * KTOT), red[16 * NT * 8 + (e - NOUT * KTOT)]);
    }
  }
}

template <int C, int NCONV>
static bool launch_wgrad_mma(int p, const float* x, const float* d0, const float* d1, const float* d2,
                             float* dw0, float* dw1, float* dw2, float* dbias, int64_t patches, cudaStream_t st) {
  auto go = [&](auto tag) {
    constexpr int P = decltype(tag)::value;
    using T = PatchTile<P, C>;
    const int64_t units = patches * T::UPP, items = (units + T::K - 1) / T::K;
    const int bx = (int)std::min<int64_t>(items, (int64_t)sm_count() * 3);
    conv3x3_wgrad_mma_kernel<C, NCONV, P><<<bx, 128, 0, st>>>(x, d0, d1, d2, dw0, dw1, dw2, dbias, units);
  };
  if (p == 8) { go(std::integral_constant<int, 8>()); return true; }
  if (p == 16) { go(std::integral_constant<int, 16>()); return true; }
  if (p == 32) { go(std::integral_constant<int, 32>()); return true; }
  return false;
}

// c_w / c_b are per-device globals, and a stream-ordered upload only orders work inside ONE stream.  FilterGuard
// makes every (upload, launch) pair on a device wait for the previous conv kernel that read the filters, whatever
// stream or host thread launched it: a per-device mutex around the enqueue + an event recorded after each launch that
// the next user's stream waits on (a no-op when it is the same stream).  Two streams or threads can therefore run the
// model on one GPU without seeing each other's filters.
struct FilterSlot { std::mutex mu; cudaEvent_t ev = nullptr; bool recorded = false; };
static FilterSlot g_filter_slot[64];
struct FilterGuard {
  FilterSlot* slot; cudaStream_t s; std::unique_lock<std::mutex> lk;
  explicit FilterGuard(cudaStream_t st) : s(st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) dev = 0;
    slot = &g_filter_slot[dev];
    lk = std::unique_lock<std::mutex>(slot->mu);
    if (!slot->ev) cudaEventCreateWithFlags(&slot->ev, cudaEventDisableTiming);
    if (slot->recorded) cudaStreamWaitEvent(s, slot->ev, 0);
  }
  ~FilterGuard() { if (slot->ev && cudaEventRecord(slot->ev, s) == cudaSuccess) slot->recorded = true; }
};

// w1 == NULL (and nconv > 1): the filters of all convs are contiguous at w0; else conv k's filters are at w_k -- the three
// nn.Conv2d weights are uploaded straight from their own parameter tensors (no concatenation pass)
static int upload_filters(const char* fn, const float* w0, const float* w1, const float* w2, const float* bias, int nconv,
                          int C, cudaStream_t s) {
  const size_t one = sizeof(float) * C * C * 9;
  if (nconv == 1 || !w1) {
    if (cudaMemcpyToSymbolAsync(c_w, w0, one * nconv, 0, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return check_launch(fn);
  } else {
    const float* ws[3] = {w0, w1, w2};
    for (int k = 0; k < nconv; ++k)
      if (cudaMemcpyToSymbolAsync(c_w, ws[k], one, one * k, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return check_launch(fn);
  }
  if (bias) {
    if (cudaMemcpyToSymbolAsync(c_b, bias, sizeof(float) * nconv * C, 0, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
      return check_launch(fn);
  } else {
    static const float zeros[12] = {0};
    if (cudaMemcpyToSymbolAsync(c_b, zeros, sizeof(float) * nconv * C, 0, cudaMemcpyHostToDevice, s) != cudaSuccess)
      return check_launch(fn);
  }
  return VU_OK;
}

static int make_geom(const char* fn, ConvGeom& g, int p_in, int p_out, int border_p, int B, int C, int H, int W) {
  VU_REQUIRE(B > 0 && H > 0 && W > 0, fn, "empty shape");
  VU_REQUIRE((int64_t)B * H * W < (int64_t)1 << 31, fn, "B*H*W must be below 2^31 pixels per call");
  VU_REQUIRE(C >= 1 && C <= 4, fn, "num_channels must be 1..4");
  auto ok = [&](int p) { return p == 0 || (p > 0 && H % p == 0 && W % p == 0); };
  VU_REQUIRE(ok(p_in) && ok(p_out) && ok(border_p), fn, "patch size must divide the image height and width");
  g.lin = Layout(C, H, W, p_in); g.lout = Layout(C, H, W, p_out);
  g.bp = border_p; g.fast = (p_in == border_p);
  g.per_image = (int64_t)C * H * W; g.npix_total = (int64_t)B * H * W;
  return VU_OK;
}

}  // namespace vu

#define VU_DISPATCH_C(C, ...)                  \
  switch (C) {                                 \
    case 1: { constexpr int CC = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int CC = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int CC = 3; __VA_ARGS__; } break; \
    default: { constexpr int CC = 4; __VA_ARGS__; } break; \
  }

extern "C" int vu_conv3x3_fwd(const float* x, int p_x, const float* w, const float* w1, const float* w2, const float* bias,
                              int nconv, float* out0, float* out1, float* out2, int p_out, int border_p,
                              int B, int C, int H, int W, int tf32, void* stream) {
  using namespace vu;
  const char* fn = "vu_conv3x3_fwd";
  VU_REQUIRE(x && w && out0 && nconv >= 1 && nconv <= 3, fn, "null pointer or nconv outside 1..3");
  VU_REQUIRE((nconv < 2 || out1) && (nconv < 3 || out2), fn, "missing output pointer");
  VU_REQUIRE(!w1 || nconv < 3 || w2, fn, "separate filter blocks: missing w2");
  ConvGeom g; int rc = make_geom(fn, g, p_x, p_out, border_p, B, C, H, W); if (rc) return rc;
  int threads = 256;
  int blocks = (int)std::min<int64_t>(cdiv(g.npix_total, threads), (int64_t)sm_count() * 32);
  cudaStream_t s = as_stream(stream);
  FilterGuard guard(s);           // released (event recorded) after the launch below, on every return path
  rc = upload_filters(fn, w, w1, w2, bias, nconv, C, s); if (rc) return rc;
  if (tf32 && g.fast && p_x == p_out && C <= 3 && p_x >= 8 && !getenv("VU_CONV_GENERIC") && !getenv("VU_CONV_NO_MMA")) {
    // tensor-core class (tf32 / bf16 modes): implicit GEMM on warp MMAs
    const int64_t patches = (int64_t)B * (H / p_x) * (W / p_x);
    bool done = false;
#define VU_PM(CC, NC) done = launch_patch_mma_fwd<CC, NC>(p_x, x, out0, out1, out2, patches, s)
#define VU_PM_N(CC) do { if (nconv == 1) VU_PM(CC, 1); else if (nconv == 2) VU_PM(CC, 2); else VU_PM(CC, 3); } while (0)
    if (C == 1) VU_PM_N(1); else if (C == 2) VU_PM_N(2); else VU_PM_N(3);
#undef VU_PM_N
#undef VU_PM
    if (done) return check_launch(fn);
  }
  if (g.fast && p_x == p_out && C <= 3 && p_x >= 4 && !getenv("VU_CONV_GENERIC")) {
    const int64_t patches = (int64_t)B * (H / p_x) * (W / p_x);
    bool done = false;
#define VU_PC(CC, NC) done = launch_patch_conv<CC, NC, false>(p_x, x, nullptr, nullptr, out0, out1, out2, patches, 0, s)
#define VU_PC_N(CC) do { if (nconv == 1) VU_PC(CC, 1); else if (nconv == 2) VU_PC(CC, 2); else VU_PC(CC, 3); } while (0)
    if (C == 1) VU_PC_N(1); else if (C == 2) VU_PC_N(2); else VU_PC_N(3);
#undef VU_PC_N
#undef VU_PC
    if (done) return check_launch(fn);
  }
  VU_DISPATCH_C(C,
    if (nconv == 1) conv3x3_fwd_kernel<CC, 1><<<blocks, threads, 0, s>>>(x, w, bias, out0, out1, out2, g);
    else if (nconv == 2) conv3x3_fwd_kernel<CC, 2><<<blocks, threads, 0, s>>>(x, w, bias, out0, out1, out2, g);
    else conv3x3_fwd_kernel<CC, 3><<<blocks, threads, 0, s>>>(x, w, bias, out0, out1, out2, g));
  return check_launch(fn);
}

extern "C" int vu_conv3x3_bwd_data(const float* dy0, const float* dy1, const float* dy2, int p_dy,
                                   const float* w, const float* w1, const float* w2, int nconv, float* dx, int p_dx,
                                   int border_p, int B, int C, int H, int W, int accumulate, int tf32, void* stream) {
  using namespace vu;
  const char* fn = "vu_conv3x3_bwd_data";
  VU_REQUIRE(dy0 && w && dx && nconv >= 1 && nconv <= 3, fn, "null pointer or nconv outside 1..3");
  VU_REQUIRE((nconv < 2 || dy1) && (nconv < 3 || dy2), fn, "missing gradient pointer");
  ConvGeom g; int rc = make_geom(fn, g, p_dy, p_dx, border_p, B, C, H, W); if (rc) return rc;
  int threads = 256;
  int blocks = (int)std::min<int64_t>(cdiv(g.npix_total, threads), (int64_t)sm_count() * 32);
  cudaStream_t s = as_stream(stream);
  FilterGuard guard(s);
  VU_REQUIRE(!w1 || nconv < 3 || w2, fn, "separate filter blocks: missing w2");
  rc = upload_filters(fn, w, w1, w2, nullptr, nconv, C, s); if (rc) return rc;
  if (tf32 && g.fast && p_dy == p_dx && C <= 3 && p_dy >= 8 && !getenv("VU_CONV_GENERIC") && !getenv("VU_CONV_NO_MMA")) {
    const int64_t patches = (int64_t)B * (H / p_dy) * (W / p_dy);
    bool done = false;
#define VU_PM(CC, NC) done = launch_patch_mma_dgrad<CC, NC>(p_dy, dy0, dy1, dy2, dx, patches, accumulate, s)
#define VU_PM_N(CC) do { if (nconv == 1) VU_PM(CC, 1); else if (nconv == 2) VU_PM(CC, 2); else VU_PM(CC, 3); } while (0)
    if (C == 1) VU_PM_N(1); else if (C == 2) VU_PM_N(2); else VU_PM_N(3);
#undef VU_PM_N
#undef VU_PM
    if (done) return check_launch(fn);
  }
  if (g.fast && p_dy == p_dx && C <= 3 && p_dy >= 4 && !getenv("VU_CONV_GENERIC")) {
    const int64_t patches = (int64_t)B * (H / p_dy) * (W / p_dy);
    bool done = false;
#define VU_PC(CC, NC) done = launch_patch_conv<CC, NC, true>(p_dy, dy0, dy1, dy2, dx, nullptr, nullptr, patches, accumulate, s)
#define VU_PC_N(CC) do { if (nconv == 1) VU_PC(CC, 1); else if (nconv == 2) VU_PC(CC, 2); else VU_PC(CC, 3); } while (0)
    if (C == 1) VU_PC_N(1); else if (C == 2) VU_PC_N(2); else VU_PC_N(3);
#undef VU_PC_N
#undef VU_PC
    if (done) return check_launch(fn);
  }
  VU_DISPATCH_C(C,
    if (nconv == 1) conv3x3_bwd_data_kernel<CC, 1><<<blocks, threads, 0, s>>>(dy0, dy1, dy2, w, dx, g, accumulate);
    else if (nconv == 2) conv3x3_bwd_data_kernel<CC, 2><<<blocks, threads, 0, s>>>(dy0, dy1, dy2, w, dx, g, accumulate);
    else conv3x3_bwd_data_kernel<CC, 3><<<blocks, threads, 0, s>>>(dy0, dy1, dy2, w, dx, g, accumulate));
  return check_launch(fn);
}

extern "C" int vu_conv3x3_bwd_weight(const float* x, int p_x, const float* dy0, const float* dy1,
                                     const float* dy2, int p_dy, int nconv, float* dw, float* dw1, float* dw2, float* dbias,
                                     int border_p, int B, int C, int H, int W, int tf32, void* stream) {
  using namespace vu;
  const char* fn = "vu_conv3x3_bwd_weight";
  VU_REQUIRE(x && dy0 && dw && nconv >= 1 && nconv <= 3, fn, "null pointer or nconv outside 1..3");
  VU_REQUIRE((nconv < 2 || dy1) && (nconv < 3 || dy2), fn, "missing gradient pointer");
  ConvGeom g; int rc = make_geom(fn, g, p_x, p_dy, border_p, B, C, H, W); if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  // dw1 == NULL: the gradient blocks of all convs are contiguous at dw; else conv k accumulates into its own block dw_k
  // (the three nn.Conv2d gradients live at their own offsets of the flat gradient buffer)
  VU_REQUIRE(!dw1 || nconv < 3 || dw2, fn, "separate gradient blocks: missing dw2");
  if (!dw1) { dw1 = dw + (int64_t)C * C * 9; dw2 = dw + (int64_t)2 * C * C * 9; }
  if (tf32 && g.fast && p_x == p_dy && C <= 3 && p_x >= 8 && !getenv("VU_CONV_WGRAD_GENERIC") && !getenv("VU_CONV_NO_MMA")) {
    const int64_t patches = (int64_t)B * (H / p_x) * (W / p_x);
    bool done = false;
#define VU_PM(CC, NC) done = launch_wgrad_mma<CC, NC>(p_x, x, dy0, dy1, dy2, dw, dw1, dw2, dbias, patches, s)
#define VU_PM_N(CC) do { if (nconv == 1) VU_PM(CC, 1); else if (nconv == 2) VU_PM(CC, 2); else VU_PM(CC, 3); } while (0)
    if (C == 1) VU_PM_N(1); else if (C == 2) VU_PM_N(2); else VU_PM_N(3);
#undef VU_PM_N
#undef VU_PM
    if (done) return check_launch(fn);
  }
  if (g.fast && p_x == p_dy && C <= 3 && (p_x == 4 || p_x == 8 || p_x == 16 || p_x == 32) && !getenv("VU_CONV_WGRAD_GENERIC")) {
    const int64_t patches = (int64_t)B * (H / p_x) * (W / p_x);
#define VU_WG(CC, PPX) launch_wgrad_patch<CC, PPX>(x, dy0, dy1, dy2, nconv, dw, dw1, dw2, dbias, patches, s)
#define VU_WG_P(CC) (p_x == 4 ? VU_WG(CC, 4) : p_x == 8 ? VU_WG(CC, 8) : p_x == 16 ? VU_WG(CC, 16) : VU_WG(CC, 32))
    if (C == 1) VU_WG_P(1); else if (C == 2) VU_WG_P(2); else VU_WG_P(3);
#undef VU_WG_P
#undef VU_WG
    return check_launch(fn);
  }
  int threads = 128;
  int bx = (int)std::min<int64_t>(cdiv(g.npix_total, threads * 8), (int64_t)sm_count() * 6);
  if (bx < 1) bx = 1;
  VU_DISPATCH_C(C,
    if (CC <= 3) conv3x3_bwd_weight_kernel<CC, (CC <= 3)><<<dim3(bx, nconv), threads, 0, s>>>(x, dy0, dy1, dy2, dw, dw1, dw2, dbias, g);
    else conv3x3_bwd_weight_kernel<CC, false><<<dim3(bx, nconv * C), threads, 0, s>>>(x, dy0, dy1, dy2, dw, dw1, dw2, dbias, g));
  return check_launch(fn);
}
