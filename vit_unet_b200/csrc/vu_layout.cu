// Patch-layout permutations: patch / unpatch / downsampling / upsampling (model.py:8-53) and the
// PatchEncoder add (model.py:84-91; ViT_UNet.ipynb c16).  Pure HBM-bound gathers: one read, one write.
// A run of 4 pixels along x that starts at x % 4 == 0 is contiguous in every layout whose patch size is a
// multiple of 4, so the common case moves float4s; otherwise a scalar path is used.
#include <cuda_bf16.h>

#include "vu_common.cuh"

namespace vu {

template <int V, bool ADD>
__global__ void __launch_bounds__(256)
repatch_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ table,
               Layout li, Layout lo, Layout lt, int64_t per_image, int64_t total) {
  // enumerate output elements in the output layout's own order -> fully coalesced stores
  for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V; e < total;
       e += (int64_t)gridDim.x * blockDim.x * V) {
    int64_t b = e / per_image;
    int64_t r = e - b * per_image;
    int c, y, x;
    if (lo.p == 0) {
      int hw = lo.H * lo.W;
      c = (int)(r / hw); int rem = (int)(r - (int64_t)c * hw);
      y = rem / lo.W; x = rem - y * lo.W;
    } else {
      int d = lo.C * lo.pp;
      int n = (int)(r / d); int f = (int)(r - (int64_t)n * d);
      c = f / lo.pp; int rem = f - c * lo.pp;
      int i = rem / lo.p, j = rem - i * lo.p;
      int rr = n / lo.gw, q = n - rr * lo.gw;
      y = rr * lo.p + i; x = q * lo.p + j;
    }
    const float* src = in + b * per_image + li.at(c, y, x);
    if (V == 4) {
      float4 v = *reinterpret_cast<const float4*>(src);
      if (ADD) {
        float4 t = *reinterpret_cast<const float4*>(table + lt.at(c, y, x));
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
      }
      *reinterpret_cast<float4*>(out + e) = v;
    } else {
      float v = *src;
      if (ADD) v += table[lt.at(c, y, x)];
      out[e] = v;
    }
  }
}

// dtable[layout lt] (+)= sum_b dout[b][layout lo]; one thread per table element (quad), loop over batch.
template <int V>
__global__ void __launch_bounds__(256)
pe_bwd_table_kernel(const float* __restrict__ dout, float* __restrict__ dtable, Layout lo, Layout lt,
                    int64_t per_image, int B, int accumulate) {
  int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (e >= per_image) return;
  int c, y, x;
  if (lt.p == 0) {
    int hw = lt.H * lt.W;
    c = (int)(e / hw); int rem = (int)(e - (int64_t)c * hw);
    y = rem / lt.W; x = rem - y * lt.W;
  } else {
    int d = lt.C * lt.pp;
    int n = (int)(e / d); int f = (int)(e - (int64_t)n * d);
    c = f / lt.pp; int rem = f - c * lt.pp;
    int i = rem / lt.p, j = rem - i * lt.p;
    int rr = n / lt.gw, q = n - rr * lt.gw;
    y = rr * lt.p + i; x = q * lt.p + j;
  }
  int64_t off = lo.at(c, y, x);
  if (V == 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
      float4 v = *reinterpret_cast<const float4*>(dout + (int64_t)b * per_image + off);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float4* dst = reinterpret_cast<float4*>(dtable + e);
    if (accumulate) { float4 o = *dst; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
    *dst = acc;
  } else {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += dout[(int64_t)b * per_image + off];
    dtable[e] = accumulate ? dtable[e] + acc : acc;
  }
}

// dst[b][h][e][n] (bf16, row pitch ldn) = src[b][n][h*hd + e] (fp32): per-head TRANSPOSED bf16 copy of a token tensor,
// so that it can be the K-major B operand (k = token index contiguous) of the bf16 tensor-core GEMMs
// A.V, dS.K, dS^T.Q and A^T.dO.  32x32 smem tile transpose: coalesced reads along e, coalesced writes along n.
__global__ void __launch_bounds__(256)
heads_transpose_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int N, int D, int hd, int h,
                            int ldn) {
  __shared__ float tile[32][33];
  const int z = blockIdx.z, b = z / h, hh = z - b * h;
  const int n0 = blockIdx.x * 32, e0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    const int n = n0 + ty + r, e = e0 + tx;
    tile[ty + r][tx] = (n < N && e < hd) ? src[((int64_t)b * N + n) * D + hh * hd + e] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    const int e = e0 + ty + r, n = n0 + tx;
    if (e < hd && n < N) dst[(((int64_t)b * h + hh) * hd + e) * ldn + n] = __float2bfloat16_rn(tile[tx][ty + r]);
  }
}

// Wide variant (the default): one CTA moves 64 tokens x 64 columns of the (N, D) token matrix of one image, whatever the
// head width -- the loads walk whole 256-byte row segments across head boundaries (the per-head kernel above reads 96-byte
// head rows at hd = 24 and leaves a quarter of its lanes idle), the stores are 16-byte chunks of 8 tokens, eight lanes
// completing a 128-byte line of one (head, e) row.  bf16 tile in shared memory, [column][token] with a pitch of 33 words:
// both the transposing stores (lane = (token % 8, float4 of the row)) and the read-back (lane = (8-token group, column))
// touch 32 distinct banks.  Tokens >= N load as zeros and land in the row padding (ldn = N rounded up to 8).
__global__ void __launch_bounds__(256)
heads_transpose_bf16_wide_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int N, int D, int hd, int ldn) {
  __shared__ __align__(16) uint32_t tile[64 * 33];               // tile[c][t / 2] packs tokens t, t + 1 of column c
  const int b = blockIdx.z, n0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __nv_bfloat16* th = reinterpret_cast<__nv_bfloat16*>(tile);
  // loads: warp w, pass r -> tokens 8 (w + 8 r) .. + 7 (lane % 8), float4 column group lane / 8 + 4 q (q = 0..3)
#pragma unroll
  for (int r = 0; r < 1; ++r) {
    const int t = 8 * warp + (lane & 7), n = n0 + t;
    float4 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = c0 + 4 * ((lane >> 3) + 4 * q);
      v[q] = (n < N && c < D) ? __ldg(reinterpret_cast<const float4*>(src + ((int64_t)b * N + n) * D + c))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int cl = 4 * ((lane >> 3) + 4 * q);
      th[(cl + 0) * 66 + t] = __float2bfloat16_rn(v[q].x);
      th[(cl + 1) * 66 + t] = __float2bfloat16_rn(v[q].y);
      th[(cl + 2) * 66 + t] = __float2bfloat16_rn(v[q].z);
      th[(cl + 3) * 66 + t] = __float2bfloat16_rn(v[q].w);
    }
  }
  __syncthreads();
  // stores: lane -> 8-token group j = lane % 8 of column 4 (warp + 8 r) + lane / 8
  const int j = lane & 7;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int cl = 4 * (warp + 8 * r) + (lane >> 3), c = c0 + cl, n = n0 + 8 * j;
    if (c < D && n < N) {
      const uint32_t* row = tile + cl * 33 + 4 * j;
      uint4 o; o.x = row[0]; o.y = row[1]; o.z = row[2]; o.w = row[3];
      const int hh = c / hd, e = c - hh * hd;
      *reinterpret_cast<uint4*>(dst + (((int64_t)b * (D / hd) + hh) * hd + e) * ldn + n) = o;
    }
  }
}

static bool layout_ok(int H, int W, int p) { return p == 0 || (p > 0 && H % p == 0 && W % p == 0); }
static bool vec_ok(int W, int p) { return p == 0 ? (W % 4 == 0) : (p % 4 == 0); }

static int launch_repatch(const char* fn, const float* in, int p_in, const float* table, int p_table,
                          float* out, int p_out, int B, int C, int H, int W, void* stream) {
  VU_REQUIRE(in && out && B > 0 && C > 0 && H > 0 && W > 0, fn, "null pointer or empty shape");
  VU_REQUIRE(layout_ok(H, W, p_in) && layout_ok(H, W, p_out) && layout_ok(H, W, p_table), fn,
             "patch size must divide the image height and width");
  Layout li(C, H, W, p_in), lo(C, H, W, p_out), lt(C, H, W, p_table);
  int64_t per = (int64_t)C * H * W, total = per * B;
  bool vec = vec_ok(W, p_in) && vec_ok(W, p_out) && (!table || vec_ok(W, p_table)) &&
             ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0) && (!table || (uintptr_t)table % 16 == 0);
  int threads = 256;
  int64_t work = vec ? total / 4 : total;
  int blocks = (int)std::min<int64_t>(cdiv(work, threads), (int64_t)sm_count() * 16);
  cudaStream_t s = as_stream(stream);
  if (vec) {
    if (table) repatch_kernel<4, true><<<blocks, threads, 0, s>>>(in, out, table, li, lo, lt, per, total);
    else repatch_kernel<4, false><<<blocks, threads, 0, s>>>(in, out, nullptr, li, lo, lt, per, total);
  } else {
    if (table) repatch_kernel<1, true><<<blocks, threads, 0, s>>>(in, out, table, li, lo, lt, per, total);
    else repatch_kernel<1, false><<<blocks, threads, 0, s>>>(in, out, nullptr, li, lo, lt, per, total);
  }
  return check_launch(fn);
}

}  // namespace vu

extern "C" int vu_repatch(const float* in, float* out, int B, int C, int H, int W, int p_in, int p_out,
                          void* stream) {
  return vu::launch_repatch("vu_repatch", in, p_in, nullptr, 0, out, p_out, B, C, H, W, stream);
}

extern "C" int vu_pe_fwd(const float* in, int p_in, const float* table, int p_table, float* out, int p_out,
                         int B, int C, int H, int W, void* stream) {
  VU_REQUIRE(table != nullptr, "vu_pe_fwd", "null table");
  return vu::launch_repatch("vu_pe_fwd", in, p_in, table, p_table, out, p_out, B, C, H, W, stream);
}

extern "C" int vu_pe_bwd_table(const float* dout, int p_out, float* dtable, int p_table,
                               int B, int C, int H, int W, int accumulate, void* stream) {
  using namespace vu;
  const char* fn = "vu_pe_bwd_table";
  VU_REQUIRE(dout && dtable && B > 0 && C > 0 && H > 0 && W > 0, fn, "null pointer or empty shape");
  VU_REQUIRE(layout_ok(H, W, p_out) && layout_ok(H, W, p_table), fn, "patch size must divide the image");
  Layout lo(C, H, W, p_out), lt(C, H, W, p_table);
  int64_t per = (int64_t)C * H * W;
  bool vec = vec_ok(W, p_out) && vec_ok(W, p_table) && ((uintptr_t)dout % 16 == 0) && ((uintptr_t)dtable % 16 == 0);
  int threads = 256;
  cudaStream_t s = as_stream(stream);
  if (vec) pe_bwd_table_kernel<4><<<(int)cdiv(per / 4, threads), threads, 0, s>>>(dout, dtable, lo, lt, per, B, accumulate);
  else pe_bwd_table_kernel<1><<<(int)cdiv(per, threads), threads, 0, s>>>(dout, dtable, lo, lt, per, B, accumulate);
  return check_launch(fn);
}

extern "C" int vu_heads_transpose_bf16(const float* src, void* dst, int B, int N, int D, int h, int ldn, void* stream) {
  using namespace vu;
  const char* fn = "vu_heads_transpose_bf16";
  VU_REQUIRE(src && dst && B > 0 && N > 0 && D > 0 && h > 0 && D % h == 0 && ldn >= N, fn, "bad arguments");
  const int hd = D / h;
  static const bool wide_on = []() { const char* e = getenv("VU_TRANSPOSE_WIDE"); return !(e && e[0] == '0'); }();
  if (wide_on && D % 4 == 0 && ldn % 8 == 0 && ldn >= ((N + 7) & ~7) && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0 &&
      cdiv(D, 64) <= 65535 && B <= 65535) {
    dim3 grid((unsigned)cdiv(N, 64), (unsigned)cdiv(D, 64), (unsigned)B);
    heads_transpose_bf16_wide_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, (__nv_bfloat16*)dst, N, D, hd, ldn);
    return check_launch(fn);
  }
  dim3 grid((unsigned)cdiv(N, 32), (unsigned)cdiv(hd, 32), (unsigned)(B * h));
  heads_transpose_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, (__nv_bfloat16*)dst, N, D, hd, h, ldn);
  return check_launch(fn);
}
