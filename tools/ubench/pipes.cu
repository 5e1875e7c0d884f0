// Micro-benchmarks that decide how the streamed Re-Attention kernels spend their issue slots on B200:
// FFMA vs packed FFMA2 (fma.rn.f32x2), MUFU.EX2, mma.sync m16n8k8 tf32 / m16n8k16 bf16, the dropout hash.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/pipes tools/ubench/pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int n) {
  float a[16], b = threadIdx.x * 1e-3f + 1.0f, c = 0.5f;
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x + i;
  if (MODE == 0) {           // 16 independent FFMA chains
    for (int it = 0; it < n; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    }
  } else if (MODE == 1) {    // 8 independent FFMA2 chains (16 FMAs per 8 instructions)
    unsigned long long v[8], bb, cc;
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
    for (int it = 0; it < n; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(bb), "l"(cc));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(v[i]));
  } else if (MODE == 2) {    // MUFU.EX2
    for (int it = 0; it < n; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = exp2f(a[i] * 0.001f);
    }
  } else if (MODE == 3) {    // mixed: 1 MUFU + 1 FFMA per element (exp2 of a scaled score)
    for (int it = 0; it < n; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = exp2f(fmaf(a[i], b, -c));
    }
  } else if (MODE == 4) {    // the dropout hash (vu_common.cuh Philox::gen_k) + 4 selects per quad
    uint32_t key = 0x1234567u + threadIdx.x, ctr = blockIdx.x * 977u + threadIdx.x;
    for (int it = 0; it < n; ++it) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t h = (ctr + i) * 0x9E3779B1u + key;
        h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
        uint32_t g = (h ^ key) * 0x27D4EB2Fu;
        g ^= g >> 15; g *= 0x165667B1u; g ^= g >> 13;
        a[4 * i + 0] = (h << 16) >= (13107u << 16) ? a[4 * i + 0] + 1.f : 0.f;
        a[4 * i + 1] = h >= (13107u << 16) ? a[4 * i + 1] + 1.f : 0.f;
        a[4 * i + 2] = (g << 16) >= (13107u << 16) ? a[4 * i + 2] + 1.f : 0.f;
        a[4 * i + 3] = g >= (13107u << 16) ? a[4 * i + 3] + 1.f : 0.f;
      }
      ctr += 4;
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
__global__ void __launch_bounds__(256) kmma(float* out, int n) {
  float c[4][4];
#pragma unroll
  for (int t = 0; t < 4; ++t) for (int i = 0; i < 4; ++i) c[t][i] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b0 = 0x3f800000u, b1 = 0x3f000000u;
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (MODE == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
  }
  float s = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t) for (int i = 0; i < 4; ++i) s += c[t][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  int dev_clk; cudaDeviceGetAttribute(&dev_clk, cudaDevAttrClockRate, 0);
  const int grid = 148 * 8, n = ITERS;          // 8 CTAs x 8 warps per SM = 16 warps per SMSP
  const double warps = (double)grid * 8, smsp = 148 * 4;
  auto rep = [&](const char* name, float ms, double instr_per_iter_per_warp, double ops_per_instr) {
    double winstr = warps * n * instr_per_iter_per_warp;
    printf("%-28s %8.3f ms  %7.2f warp-instr/ns  = %.3f per SMSP per clk @1.9GHz ; %.1f Tops/s\n", name, ms,
           winstr / (ms * 1e6), winstr / (ms * 1e-3) / smsp / 1.9e9, winstr * 32 * ops_per_instr / (ms * 1e-3) / 1e12);
  };
  rep("FFMA x16", timeit([&] { k<0><<<grid, 256>>>(out, n); }), 16, 1);
  rep("FFMA2 x8 (f32x2)", timeit([&] { k<1><<<grid, 256>>>(out, n); }), 8, 2);
  rep("MUFU.EX2 (+FMUL)", timeit([&] { k<2><<<grid, 256>>>(out, n); }), 16, 1);
  rep("FFMA+MUFU.EX2 pairs", timeit([&] { k<3><<<grid, 256>>>(out, n); }), 16, 1);
  rep("dropout hash per quad x4", timeit([&] { k<4><<<grid, 256>>>(out, n); }), 4, 4);
  rep("mma m16n8k8 tf32 x4", timeit([&] { kmma<0><<<grid, 256>>>(out, n); }), 4, 1024.0 / 32 * 2);
  rep("mma m16n8k16 bf16 x4", timeit([&] { kmma<1><<<grid, 256>>>(out, n); }), 4, 2048.0 / 32 * 2);
  printf("(device clock rate attr %d kHz)\n", dev_clk);
  return 0;
}
