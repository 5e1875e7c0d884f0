// Shared device/host helpers for the vit_unet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/vit_unet_b200.h"

namespace vu {

// ------------------------------------------------------------------ host-side error plumbing
void set_error(const std::string& s);
int fail_arg(const char* fn, const char* what);
int check_launch(const char* fn);     // cudaPeekAtLastError -> VU_ERR_CUDA
int sm_count();

#define VU_REQUIRE(cond, fn, msg) do { if (!(cond)) return vu::fail_arg(fn, msg); } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
// Per-kernel function attributes (opt-in dynamic shared memory) are per DEVICE: `seen` is a bit mask of the devices a
// call site has already configured.  True the first time the current device shows up (benign if two threads race).
static inline bool first_use_on_device(uint64_t& seen) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;
  const uint64_t bit = 1ull << dev;
  if (seen & bit) return false;
  seen |= bit;
  return true;
}
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------ patch-layout addressing
// Offset (inside one image of C*H*W floats) of pixel (c, y, x) in patch layout p (p == 0: NCHW).
// reference: patch()/unflatten()/unpatch(), model.py:8-35.
struct Layout {
  int C, H, W, p, gw, pp;   // gw = W/p, pp = p*p
  int ps;                   // log2(p) when p is a power of two (shift/mask index math), else -1
  __host__ __device__ Layout() {}
  __host__ __device__ Layout(int C_, int H_, int W_, int p_) : C(C_), H(H_), W(W_), p(p_) {
    gw = p ? W / p : 1; pp = p * p;
    ps = -1;
    if (p > 0 && (p & (p - 1)) == 0) { ps = 0; while ((1 << ps) < p) ++ps; }
  }
  __device__ __forceinline__ int64_t at(int c, int y, int x) const {
    if (p == 0) return ((int64_t)c * H + y) * W + x;
    if (ps >= 0) {
      const int r = y >> ps, q = x >> ps, i = y & (p - 1), j = x & (p - 1);
      return (((int64_t)(r * gw + q) * C + c) << (2 * ps)) + (i << ps) + j;
    }
    int r = y / p, q = x / p;
    int i = y - r * p, j = x - q * p;
    return ((int64_t)(r * gw + q) * C + c) * pp + i * p + j;
  }
  // inverse: linear pixel index (channel 0 plane ordering of this layout) -> (y, x).
  // pix enumerates (token, i, j) for p>0 and (y, x) for p==0.  (pix < H*W always fits 32 bits.)
  __device__ __forceinline__ void pixel(uint32_t pix, int& y, int& x) const {
    if (p == 0) { y = (int)(pix / (uint32_t)W); x = (int)(pix - (uint32_t)y * W); return; }
    int n, i, j;
    if (ps >= 0) {
      n = (int)(pix >> (2 * ps)); const int rem = (int)(pix & (uint32_t)(pp - 1));
      i = rem >> ps; j = rem & (p - 1);
    } else {
      n = (int)(pix / (uint32_t)pp); const int rem = (int)(pix - (uint32_t)n * pp);
      i = rem / p; j = rem - i * p;
    }
    int r = n / gw, q = n - r * gw;
    y = r * p + i; x = q * p + j;
  }
  // offset of channel-0 value for linear pixel index pix, and stride between channels
  __device__ __forceinline__ int64_t base_of(uint32_t pix) const {
    if (p == 0) return pix;
    if (ps >= 0) { const uint32_t n = pix >> (2 * ps); return ((int64_t)n * C << (2 * ps)) + (pix & (uint32_t)(pp - 1)); }
    uint32_t n = pix / (uint32_t)pp; uint32_t rem = pix - n * pp;
    return (int64_t)n * C * pp + rem;
  }
  __device__ __forceinline__ int64_t cstride() const { return p == 0 ? (int64_t)H * W : pp; }
};

// ------------------------------------------------------------------ counter-based dropout RNG
// One 32-bit multiply-xorshift chain (murmur3 finaliser) per QUAD of consecutive elements, keyed by (seed, stream id)
// and counted by the low 32 bits of (element index / 4), yields four 16-bit uniforms: element e of the quad is kept
// iff its 16 bits >= thresh16 (= round(p * 65536)).  ~15 integer instructions per quad, all 32-bit (the attention-map
// kernels regenerate the mask in five passes instead of storing it, and were issue-bound on a 64-bit hash).  The same
// function is used by every kernel that needs the mask of a given element, so masks never have to be stored; the
// mask pattern repeats every 2^34 elements of one stream.  (Struct keeps its historical name.)
struct Philox {
  __host__ __device__ __forceinline__ static uint32_t key(uint64_t seed, uint32_t stream) {
    uint32_t k = (uint32_t)seed * 0x9E3779B1u ^ (uint32_t)(seed >> 32) * 0x85EBCA77u;
    k ^= stream * 0xC2B2AE3Du + 0x27D4EB2Fu;
    k ^= k >> 15; k *= 0x2C1B3C6Du; k ^= k >> 12; k *= 0x297A2D39u; k ^= k >> 15;
    return k;
  }
  __device__ __forceinline__ static uint4 gen_k(uint32_t key, uint32_t ctr) {
    uint32_t h = ctr * 0x9E3779B1u + key;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    uint32_t g = (h ^ key) * 0x27D4EB2Fu;
    g ^= g >> 15; g *= 0x165667B1u; g ^= g >> 13;
    return make_uint4(h & 0xFFFFu, h >> 16, g & 0xFFFFu, g >> 16);
  }
  __device__ __forceinline__ static uint4 gen(uint64_t seed, uint32_t stream, uint64_t ctr) {
    return gen_k(key(seed, stream), (uint32_t)ctr);
  }
  // keep test for one element index; thresh = p * 2^16 (drop if r < thresh)
  __device__ __forceinline__ static bool keep(uint64_t seed, uint32_t stream, uint64_t idx, uint32_t thresh) {
    uint4 r = gen(seed, stream, idx >> 2);
    uint32_t v = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
    return v >= thresh;
  }
};
// 16-bit threshold and the matching unbiased keep-scale 1 / (1 - thresh/65536)
static inline uint32_t drop_threshold(float p) {
  double t = (double)p * 65536.0 + 0.5;
  if (t < 0) t = 0; if (t > 65535.0) t = 65535.0;
  return (uint32_t)t;
}
static inline float drop_keep_scale(float p) {
  if (!(p > 0.f)) return 1.0f;
  return (float)(1.0 / (1.0 - (double)drop_threshold(p) / 65536.0));
}

// ------------------------------------------------------------------ reductions
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum of NV values per thread (double), result valid in thread 0. blockDim.x <= 1024.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /* >= NV*32 doubles */) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) smem[i * 32 + warp] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double t = lane < nw ? smem[i * 32 + lane] : 0.0;
      v[i] = warp_sum(t);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float gelu_exact(float x) {          // torch.nn.GELU() default (erf form)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_exact_grad(float x) {
  float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

}  // namespace vu
