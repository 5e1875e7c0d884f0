// Tensor-core GEMM for sm_100a: tcgen05.mma (kind::tf32, FP32 accumulate in TMEM) fed by TMA through an
// mbarrier pipeline.  Same vu_gemm_desc contract (batch strides, transposes, fused epilogues, split-K) as the
// CUDA-core kernel in vu_gemm_simt.cu, so every contraction of the ViT-UNet block -- QK^T / PV, proj,
// FeedForward and all their data / weight gradients -- runs here when precision == VU_PREC_TF32.
//
// Operands stay FP32 in HBM; TMA (data type TFLOAT32) lands them in shared memory in the canonical
// SWIZZLE_128B layouts the UMMA smem descriptors expect:
//   K-major  operand (k contiguous in memory):  one box of {32 k, ROWS} per stage; row r at r*128 B, 8-row swizzle
//            atoms 1024 B apart (SBO).  One MMA consumes 8 k = 32 B: descriptor start advances 32 B per k-step.
//   MN-major operand (m or n contiguous):       ROWS/32 boxes of {32 mn, 32 k} per stage; each box is 32 k-rows of
//            128 B.  For 32-bit MN-major operands the only legal layout is SWIZZLE_128B_BASE32B (32-byte swizzle
//            atoms, TMA mode SWIZZLE_128B_ATOM_32B): slabs 4096 B apart (LBO), 4-k groups 512 B apart (SBO);
//            start advances 1024 B (8 k-rows) per k-step.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..5 = epilogue (TMEM -> registers -> fused epilogue -> global), each owning its TMEM lane quarter.
// One 128 x BLOCK_N output tile per CTA; two CTAs per SM so one CTA's epilogue overlaps the other's main loop.
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <cstdlib>

#include "vu_common.cuh"

namespace vu {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 32;          // floats per stage along K (= 128 bytes = one swizzle span)
constexpr int TC_UMMA_K = 8;            // tf32: 32 bytes per MMA

struct TcArgs {
  float* C;
  const float* bias; const float* residual; const float* aux_in; float* aux_out;
  int M, N, K;
  int64_t ldc, ldr, ldaux;
  int batch_inner;
  int64_t sCo, sCi;
  float alpha; int act; int accumulate; int split_k; int k_per_split;
  float drop_scale; uint32_t drop_thresh; uint64_t drop_seed; uint32_t drop_stream;
  uint32_t mn_lbo, mn_sbo;      // MN-major descriptor strides (bytes)
  int c_bf16;                   // C is __nv_bfloat16 (no accumulate / split-K; the fused epilogues apply)
  int aux_bf16;                 // aux_in / aux_out are __nv_bfloat16
  int tiles_m, tiles_n, total_tiles;   // persistent kernel: tile id -> (n tile fastest, m tile, batch * split)
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 %%rx;\n"
      ".reg .pred %%px;\n"
      "elect.sync %%rx|%%px, %1;\n"
      "@%%px mov.s32 %0, 1;\n"
      "}\n" : "+r"(pred) : "r"(0xffffffffu));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;   // 2 = SWIZZLE_128B (16 B atoms), 1 = SWIZZLE_128B_BASE32B (32 B atoms)
  return d;
}
// cute::UMMA::InstrDescriptor: F32 accumulate, TF32 x TF32, M=128
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn, bool bf16) {
  const uint32_t fmt = bf16 ? 1u : 2u;        // F16F32Format: 1 = BF16, 2 = TF32
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BLOCK_M >> 4) << 24);
}

// Second half of the epilogue, shared by the one-tile-per-CTA kernel and the persistent kernel: the warp's 32 x BLOCK_N
// sub-tile sits in shared memory (`stage`, row pitch BLOCK_N + 4 floats, alpha already applied); rows leave as coalesced
// 16-byte-per-lane accesses through the fused epilogue (bias, GELU / GELU', dropout, residual, fp32 / bf16 / atomic store).
template <int BLOCK_N, int UN>
__device__ __forceinline__ void epilogue_rows(const TcArgs& g, const float* stage, const int lane, const int q, const int m0,
                                              const int n0, const int z, const int ks, const int64_t coff) {
  constexpr int LDS = BLOCK_N + 4;
  float* C = g.C + coff;
  constexpr int LPR = BLOCK_N / 4;                      // lanes covering one output row
  constexpr int RPI = 32 / LPR;                         // rows handled per iteration
  const int cl = (lane % LPR) * 4;                      // this lane's 4 columns inside the tile
  const int n = n0 + cl;
  const int nv = min(4, g.N - n);                       // <= 0: nothing to do for this lane
  const bool vec_c = ((uintptr_t)C % 16 == 0) && (g.ldc % 4 == 0) && nv == 4;
  float bias4[4] = {0.f, 0.f, 0.f, 0.f};
  if (g.bias && (g.split_k <= 1 || ks == 0)) {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (j < nv) bias4[j] = g.bias[n + j];
  }
  const uint32_t drop_key = Philox::key(g.drop_seed, g.drop_stream);
  const bool vec_r = g.residual && ((uintptr_t)(g.residual + coff) % 16 == 0) && (g.ldr % 4 == 0) && nv == 4;
  const float* axp = g.aux_in ? g.aux_in : g.aux_out;
  // aux rows as one vector access: float4 (fp32) or 4 x bf16 = 8 bytes
  const bool vec_x = axp && nv == 4 && (g.ldaux % 4 == 0) &&
                     (g.aux_bf16 ? ((uintptr_t)(reinterpret_cast<const __nv_bfloat16*>(axp) + coff) % 8 == 0)
                                 : ((uintptr_t)(axp + coff) % 16 == 0));
  // ---- fast path: no activation, no split-K, whole aligned quads (every proj / FeedForward-2 / data-gradient / dense
  // weight-gradient launch of the path).  The general row routine below costs ~100 executed instructions per row with one
  // epilogue warp per scheduler and nothing to overlap them with -- measured 16 us per 128 x 128 tile, the whole lifetime
  // of a thin GEMM's CTA.  Here a row is ~12 instructions, the loads of a batch of rows (residual, and C itself when
  // accumulating) are issued together, and the addresses are strength-reduced.
  {
    const bool c_ok = g.c_bf16 ? ((uintptr_t)(reinterpret_cast<const __nv_bfloat16*>(g.C) + coff) % 8 == 0 && g.ldc % 4 == 0) : vec_c;
    if (g.split_k <= 1 && g.act == VU_ACT_NONE && nv == 4 && c_ok && (!g.residual || vec_r) && !(g.accumulate && g.c_bf16) &&
        (!g.drop_thresh || g.N % 4 == 0)) {
      constexpr int ITER = 32 / RPI;                       // row iterations of this warp
      constexpr int UB = ITER < UN ? ITER : UN;            // rows per batch of loads
      const int r_in = lane / LPR;
      const int mrow0 = m0 + q * 32 + r_in;
      const float* sp = stage + r_in * LDS + cl;
      const float4 b4 = make_float4(bias4[0], bias4[1], bias4[2], bias4[3]);
      const float* rp = g.residual ? g.residual + coff + (int64_t)mrow0 * g.ldr + n : nullptr;
      float* cp = C + (int64_t)mrow0 * g.ldc + n;
      __nv_bfloat16* cb = reinterpret_cast<__nv_bfloat16*>(g.C) + coff + (int64_t)mrow0 * g.ldc + n;
      const int64_t rstep = (int64_t)RPI * g.ldr, cstep = (int64_t)RPI * g.ldc;
      const bool acc = g.accumulate != 0;
      const uint64_t didx0 = (uint64_t)z * g.M * g.N + (uint64_t)mrow0 * g.N + n;      // dropout counter of the first row
      const uint32_t dstep = (uint32_t)((RPI * g.N) >> 2);
#pragma unroll 1
      for (int ib = 0; ib < ITER; ib += UB) {
        float4 rr[UB], cc[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const bool ok = mrow0 + (ib + u) * RPI < g.M;
          rr[u] = (rp && ok) ? *reinterpret_cast<const float4*>(rp + (ib + u) * rstep) : make_float4(0.f, 0.f, 0.f, 0.f);
          cc[u] = (acc && ok) ? *reinterpret_cast<const float4*>(cp + (ib + u) * cstep) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          if (mrow0 + (ib + u) * RPI >= g.M) continue;
          float4 t = *reinterpret_cast<const float4*>(sp + (ib + u) * RPI * LDS);
          t.x += b4.x; t.y += b4.y; t.z += b4.z; t.w += b4.w;
          if (g.drop_thresh) {
            const uint4 r = Philox::gen_k(drop_key, (uint32_t)(didx0 >> 2) + (uint32_t)(ib + u) * dstep);
            t.x = r.x >= g.drop_thresh ? t.x * g.drop_scale : 0.f; t.y = r.y >= g.drop_thresh ? t.y * g.drop_scale : 0.f;
            t.z = r.z >= g.drop_thresh ? t.z * g.drop_scale : 0.f; t.w = r.w >= g.drop_thresh ? t.w * g.drop_scale : 0.f;
          }
          t.x += rr[u].x + cc[u].x; t.y += rr[u].y + cc[u].y; t.z += rr[u].z + cc[u].z; t.w += rr[u].w + cc[u].w;
          if (g.c_bf16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(t.x, t.y), hi = __floats2bfloat162_rn(t.z, t.w);
            uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(cb + (ib + u) * cstep) = pk;
          } else {
            *reinterpret_cast<float4*>(cp + (ib + u) * cstep) = t;
          }
        }
      }
      return;
    }
  }
  // One output row (this lane's 4 columns) through the fused epilogue.  pr / px: the residual and GELU' pre-activation
  // quads of the row, loaded by the caller a whole batch of rows ahead (have_r / have_x), else fetched here.
  auto process = [&](const int row, const bool have_r, const float4 pr, const bool have_x, const float4 px) {
    const int m = m0 + q * 32 + row;
    if (m >= g.M || nv <= 0) return;
    const float4 t4 = *reinterpret_cast<const float4*>(stage + row * LDS + cl);
    float v[4] = {t4.x + bias4[0], t4.y + bias4[1], t4.z + bias4[2], t4.w + bias4[3]};
    float* dst = C + (int64_t)m * g.ldc + n;
    if (g.split_k > 1) {
      if (g.residual && ks == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j < nv) v[j] += g.residual[coff + (int64_t)m * g.ldr + n + j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) if (j < nv) atomicAdd(dst + j, v[j]);
      return;
    }
    const int64_t xoff = coff + (int64_t)m * g.ldaux + n;
    if (g.act == VU_ACT_GELU) {
      if (g.aux_out) {
        if (g.aux_bf16) {
          __nv_bfloat16* ax = reinterpret_cast<__nv_bfloat16*>(g.aux_out) + xoff;
          if (vec_x) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
            uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(ax) = pk;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < nv) ax[j] = __float2bfloat16_rn(v[j]);
          }
        } else {
          float* ax = g.aux_out + xoff;
          if (vec_x) *reinterpret_cast<float4*>(ax) = make_float4(v[0], v[1], v[2], v[3]);
          else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < nv) ax[j] = v[j];
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = gelu_exact(v[j]);
    } else if (g.act == VU_ACT_GELU_BWD) {
      float a[4] = {px.x, px.y, px.z, px.w};
      if (have_x) {
      } else if (g.aux_bf16) {
        const __nv_bfloat16* ax = reinterpret_cast<const __nv_bfloat16*>(g.aux_in) + xoff;
        if (vec_x) {
          const uint2 pk = *reinterpret_cast<const uint2*>(ax);
          const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
          const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
          a[0] = lo.x; a[1] = lo.y; a[2] = hi.x; a[3] = hi.y;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (j < nv) a[j] = __bfloat162float(ax[j]);
        }
      } else {
        const float* ax = g.aux_in + xoff;
        if (vec_x) { float4 t = *reinterpret_cast<const float4*>(ax); a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w; }
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (j < nv) a[j] = ax[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= gelu_exact_grad(a[j]);
    }
    if (g.drop_thresh) {
      const uint64_t idx0 = (uint64_t)z * g.M * g.N + (uint64_t)m * g.N + n;
      if ((idx0 & 3) == 0) {       // the lane's 4 columns are one RNG quad: one hash instead of four
        const uint4 r = Philox::gen_k(drop_key, (uint32_t)(idx0 >> 2));
        v[0] = r.x >= g.drop_thresh ? v[0] * g.drop_scale : 0.f;
        v[1] = r.y >= g.drop_thresh ? v[1] * g.drop_scale : 0.f;
        v[2] = r.z >= g.drop_thresh ? v[2] * g.drop_scale : 0.f;
        v[3] = r.w >= g.drop_thresh ? v[3] * g.drop_scale : 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          v[j] = Philox::keep(g.drop_seed, g.drop_stream, idx0 + j, g.drop_thresh) ? v[j] * g.drop_scale : 0.f;
      }
    }
    if (g.residual) {
      const float* rp = g.residual + coff + (int64_t)m * g.ldr + n;
      if (have_r) { v[0] += pr.x; v[1] += pr.y; v[2] += pr.z; v[3] += pr.w; }
      else if (vec_r) { float4 t = *reinterpret_cast<const float4*>(rp); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j < nv) v[j] += rp[j];
      }
    }
    if (g.c_bf16) {          // bf16 output (host side rejects accumulate / split-K)
      __nv_bfloat16* cb = reinterpret_cast<__nv_bfloat16*>(g.C) + coff + (int64_t)m * g.ldc + n;
      if (nv == 4 && ((uintptr_t)cb % 8 == 0)) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
        uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(cb) = pk;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j < nv) cb[j] = __float2bfloat16_rn(v[j]);
      }
    } else if (vec_c) {
      float4 o = make_float4(v[0], v[1], v[2], v[3]);
      if (g.accumulate) { float4 c = *reinterpret_cast<float4*>(dst); o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
      *reinterpret_cast<float4*>(dst) = o;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nv) dst[j] = g.accumulate ? dst[j] + v[j] : v[j];
    }
  };
  // Rows go through in batches of UN: the batch's residual / pre-activation quads are requested from global memory
  // first (UN independent 16-byte loads per lane in flight), then the rows are finished.  With one load per row issued
  // right where it is consumed the epilogue of an output-bound product (proj forward: bias + dropout + residual) ran at
  // one memory latency per row: 251 us instead of 106 us for the same product without a residual (M = 200704, N = K = 192).
  const bool pre_r = vec_r && g.split_k <= 1;
  const bool pre_x = vec_x && g.act == VU_ACT_GELU_BWD;
#pragma unroll 1
  for (int rb = 0; rb < 32; rb += RPI * UN) {
    float4 pr[UN], px[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int row = rb + u * RPI + lane / LPR;
      const int m = m0 + q * 32 + row;
      pr[u] = make_float4(0.f, 0.f, 0.f, 0.f); px[u] = pr[u];
      if (m < g.M && nv > 0) {
        if (pre_r) pr[u] = *reinterpret_cast<const float4*>(g.residual + coff + (int64_t)m * g.ldr + n);
        if (pre_x) {
          if (g.aux_bf16) {
            const uint2 pk = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(g.aux_in) + coff + (int64_t)m * g.ldaux + n);
            const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
            const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
            px[u] = make_float4(lo.x, lo.y, hi.x, hi.y);
          } else {
            px[u] = *reinterpret_cast<const float4*>(g.aux_in + coff + (int64_t)m * g.ldaux + n);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) process(rb + u * RPI + lane / LPR, pre_r, pr[u], pre_x, px[u]);
  }
}

// BF16 = true: both operands are __nv_bfloat16 (kind::f16, 64 elements per 128-byte k-block, UMMA_K = 16).  Either
// operand may be MN-major with the ordinary SWIZZLE_128B atoms (64 elements x 8 k-rows; ROWS / 64 slabs of 8 KB per
// stage, LBO = 8192 between slabs, SBO = 1024 between 8-k-row groups): the weight-gradient products dW = dY^T X read
// both token tensors as they lie in memory.  An MN-major bf16 B needs BLOCK_N >= 64 (one slab is 64 columns wide).
// EPI_UN: rows per batch of the epilogue's global loads (residual / GELU' pre-activation): 4 for launches that read such a
// tensor, 1 (no batching, 71 instead of 96 registers -> one more resident CTA for the map-reading products) otherwise.
template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN, bool BF16, int EPI_UN>
__global__ void __launch_bounds__(192, 3)
gemm_tf32_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcArgs g) {
  static_assert(!(BF16 && B_MN) || BLOCK_N >= 64, "an MN-major bf16 B tile is made of 64-column slabs");
  constexpr int KB = BF16 ? 64 : 32;                    // elements per 128-byte k-block
  constexpr uint32_t A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 4;
  constexpr uint32_t B_BYTES = BLOCK_N * TC_BLOCK_K * 4;
  constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned carve-up (SWIZZLE_128B atoms are 1024 B)
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  // barriers live past BOTH the pipeline ring and the epilogue staging area that later reuses the ring's space
  constexpr uint32_t RING_BYTES = STAGES * (A_BYTES + B_BYTES);
  constexpr uint32_t STAGING_BYTES = 4u * 32u * (BLOCK_N + 4) * 4u;
  constexpr uint32_t BAR_OFFSET = ((RING_BYTES > STAGING_BYTES ? RING_BYTES : STAGING_BYTES) + 15u) & ~15u;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BLOCK_M, n0 = blockIdx.x * BLOCK_N;
  int z = blockIdx.z, ks = 0;
  if (g.split_k > 1) { ks = z % g.split_k; z /= g.split_k; }
  const int zo = z / g.batch_inner, zi = z % g.batch_inner;
  const int kbeg = ks * g.k_per_split;
  const int kend = min(g.K, kbeg + g.k_per_split);
  const int nkb = (kend - kbeg + KB - 1) / KB;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty_bar + s, ph ^ 1);
        mbar_expect_tx(full_bar + s, A_BYTES + B_BYTES);
        const int k0 = kbeg + kb * KB;
        uint8_t* a_dst = sA + s * A_BYTES;
        uint8_t* b_dst = sB + s * B_BYTES;
        if (!A_MN) tma_load_4d(a_dst, &tmA, full_bar + s, k0, m0, zi, zo);
        else if (BF16) {
#pragma unroll
          for (int sl = 0; sl < TC_BLOCK_M / 64; ++sl) tma_load_4d(a_dst + sl * 8192, &tmA, full_bar + s, m0 + sl * 64, k0, zi, zo);
        } else {
#pragma unroll
          for (int sl = 0; sl < TC_BLOCK_M / 32; ++sl) tma_load_4d(a_dst + sl * 4096, &tmA, full_bar + s, m0 + sl * 32, k0, zi, zo);
        }
        if (!B_MN) tma_load_4d(b_dst, &tmB, full_bar + s, k0, n0, zi, zo);
        else if (BF16) {
#pragma unroll
          for (int sl = 0; sl < BLOCK_N / 64; ++sl) tma_load_4d(b_dst + sl * 8192, &tmB, full_bar + s, n0 + sl * 64, k0, zi, zo);
        } else {
#pragma unroll
          for (int sl = 0; sl < BLOCK_N / 32; ++sl) tma_load_4d(b_dst + sl * 4096, &tmB, full_bar + s, n0 + sl * 32, k0, zi, zo);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    constexpr uint32_t idesc = make_idesc(BLOCK_N, A_MN, B_MN, BF16);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(full_bar + s, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_base = smem_u32(sA + s * A_BYTES), b_base = smem_u32(sB + s * B_BYTES);
#pragma unroll
        for (int kk = 0; kk < TC_BLOCK_K / TC_UMMA_K; ++kk) {
          const uint64_t ad = !A_MN ? make_smem_desc(a_base + kk * 32, 16, 1024, 2)
                              : (BF16 ? make_smem_desc(a_base + kk * 2048, 8192, 1024, 2)
                                      : make_smem_desc(a_base + kk * 1024, g.mn_lbo, g.mn_sbo, 1));
          const uint64_t bd = !B_MN ? make_smem_desc(b_base + kk * 32, 16, 1024, 2)
                              : (BF16 ? make_smem_desc(b_base + kk * 2048, 8192, 1024, 2)
                                      : make_smem_desc(b_base + kk * 1024, g.mn_lbo, g.mn_sbo, 1));
          if (BF16) umma_bf16(tmem_base, ad, bd, idesc, (kb | kk) != 0 ? 1u : 0u);
          else umma_tf32(tmem_base, ad, bd, idesc, (kb | kk) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar + s);                       // frees the smem slot once these MMAs retire
        if (kb == nkb - 1) umma_commit(tmem_full_bar);    // accumulator complete -> epilogue
      }
      __syncwarp();
    }
  } else {
    // ===================================================== epilogue (warps 2..5)
    // TMEM -> registers (one accumulator row per lane) -> shared-memory staging (the pipeline buffers are idle
    // once the accumulator is complete) -> row-contiguous float4 global accesses: every load of residual / aux
    // and every store of C is a fully coalesced 16 B-per-lane transaction.
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    const int64_t coff = zo * g.sCo + zi * g.sCi;
    constexpr int LDS = BLOCK_N + 4;                      // padded staging row (floats)
    float* stage = reinterpret_cast<float*>(smem) + (size_t)q * 32 * LDS;
    if (EPI_UN > 1 && g.residual && g.split_k <= 1) {
      // the residual rows of this warp's 32 x BLOCK_N sub-tile are pulled into L2 while the main loop runs
      const float* rbase = g.residual + coff + (int64_t)(m0 + q * 32) * g.ldr + n0;
      constexpr int LINES = BLOCK_N / 32;                   // 128-byte lines per row
      for (int i = lane; i < 32 * LINES; i += 32) {
        const int rr = i / LINES, cc = (i - rr * LINES) * 32;
        if (m0 + q * 32 + rr < g.M && n0 + cc < g.N)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(rbase + (int64_t)rr * g.ldr + cc));
      }
    }
    if (nkb > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 16) {
      uint32_t r[16];
      if (nkb > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = 0u;
      }
#pragma unroll
      for (int j4 = 0; j4 < 16; j4 += 4)
        *reinterpret_cast<float4*>(stage + lane * LDS + c0 + j4) =
            make_float4(__uint_as_float(r[j4]) * g.alpha, __uint_as_float(r[j4 + 1]) * g.alpha,
                        __uint_as_float(r[j4 + 2]) * g.alpha, __uint_as_float(r[j4 + 3]) * g.alpha);
    }
    __syncwarp();
    epilogue_rows<BLOCK_N, EPI_UN>(g, stage, lane, q, m0, n0, z, ks, coff);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------ persistent variant (token GEMMs)
// One CTA per SM walks the output tiles (tile = blockIdx.x + i * gridDim.x, n tile fastest so that concurrently running
// CTAs share A rows through L2).  The accumulator is DOUBLE-BUFFERED in TMEM (2 x BLOCK_N columns): while the four epilogue
// warps drain tile i (TMEM -> registers -> their own staging rows in shared memory -> fused epilogue -> global), the MMA
// warp already accumulates tile i+1 into the other buffer and the TMA warp keeps the STAGES-deep operand ring full across
// tile boundaries.  Barrier init, tensor-map prefetch and the TMEM allocation are paid once per SM instead of once per
// tile -- what the one-tile-per-CTA kernel loses on the thin (K <= 768) and output-bound shapes -- and the epilogue of a
// large tile no longer idles the tensor pipe.  With one CTA per SM the epilogue can afford 8-row load batches (UN = 8).
template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN, bool BF16>
__global__ void __launch_bounds__(192, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcArgs g) {
  static_assert(!(BF16 && B_MN) || BLOCK_N >= 64, "an MN-major bf16 B tile is made of 64-column slabs");
  constexpr int KB = BF16 ? 64 : 32;
  constexpr uint32_t A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 4;
  constexpr uint32_t B_BYTES = BLOCK_N * TC_BLOCK_K * 4;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;              // two accumulators (64 -> 128, 128 -> 256 columns)
  constexpr int LDS = BLOCK_N + 4;
  constexpr uint32_t RING_BYTES = STAGES * (A_BYTES + B_BYTES);
  constexpr uint32_t STAGING_BYTES = 4u * 32u * LDS * 4u;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  float* staging = reinterpret_cast<float*>(smem + RING_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + RING_BYTES + STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;            // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full_bar + b, 1); mbar_init(tmem_empty_bar + b, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile id -> coordinates (identical in the three roles)
  auto decode = [&](int tile, int& m0, int& n0, int& z, int& ks, int& zo, int& zi, int& kbeg, int& nkb) {
    const int tn = tile % g.tiles_n; tile /= g.tiles_n;
    const int tm = tile % g.tiles_m; z = tile / g.tiles_m;
    m0 = tm * TC_BLOCK_M; n0 = tn * BLOCK_N;
    ks = 0;
    if (g.split_k > 1) { ks = z % g.split_k; z /= g.split_k; }
    zo = z / g.batch_inner; zi = z % g.batch_inner;
    kbeg = ks * g.k_per_split;
    const int kend = min(g.K, kbeg + g.k_per_split);
    nkb = (kend - kbeg + KB - 1) / KB;
  };

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      uint32_t it = 0;                                     // k-blocks issued so far (ring position across tiles)
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        int m0, n0, z, ks, zo, zi, kbeg, nkb;
        decode(tile, m0, n0, z, ks, zo, zi, kbeg, nkb);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty_bar + s, ph ^ 1);
          mbar_expect_tx(full_bar + s, A_BYTES + B_BYTES);
          const int k0 = kbeg + kb * KB;
          uint8_t* a_dst = sA + s * A_BYTES;
          uint8_t* b_dst = sB + s * B_BYTES;
          if (!A_MN) tma_load_4d(a_dst, &tmA, full_bar + s, k0, m0, zi, zo);
          else if (BF16) {
#pragma unroll
            for (int sl = 0; sl < TC_BLOCK_M / 64; ++sl) tma_load_4d(a_dst + sl * 8192, &tmA, full_bar + s, m0 + sl * 64, k0, zi, zo);
          } else {
#pragma unroll
            for (int sl = 0; sl < TC_BLOCK_M / 32; ++sl) tma_load_4d(a_dst + sl * 4096, &tmA, full_bar + s, m0 + sl * 32, k0, zi, zo);
          }
          if (!B_MN) tma_load_4d(b_dst, &tmB, full_bar + s, k0, n0, zi, zo);
          else if (BF16) {
#pragma unroll
            for (int sl = 0; sl < BLOCK_N / 64; ++sl) tma_load_4d(b_dst + sl * 8192, &tmB, full_bar + s, n0 + sl * 64, k0, zi, zo);
          } else {
#pragma unroll
            for (int sl = 0; sl < BLOCK_N / 32; ++sl) tma_load_4d(b_dst + sl * 4096, &tmB, full_bar + s, n0 + sl * 32, k0, zi, zo);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    constexpr uint32_t idesc = make_idesc(BLOCK_N, A_MN, B_MN, BF16);
    uint32_t it = 0, t = 0;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++t) {
      int m0, n0, z, ks, zo, zi, kbeg, nkb;
      decode(tile, m0, n0, z, ks, zo, zi, kbeg, nkb);
      const uint32_t buf = t & 1;
      mbar_wait(tmem_empty_bar + buf, ((t >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator (2 tiles ago)
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * BLOCK_N;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(full_bar + s, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_base = smem_u32(sA + s * A_BYTES), b_base = smem_u32(sB + s * B_BYTES);
#pragma unroll
          for (int kk = 0; kk < TC_BLOCK_K / TC_UMMA_K; ++kk) {
            const uint64_t ad = !A_MN ? make_smem_desc(a_base + kk * 32, 16, 1024, 2)
                                : (BF16 ? make_smem_desc(a_base + kk * 2048, 8192, 1024, 2)
                                        : make_smem_desc(a_base + kk * 1024, g.mn_lbo, g.mn_sbo, 1));
            const uint64_t bd = !B_MN ? make_smem_desc(b_base + kk * 32, 16, 1024, 2)
                                : (BF16 ? make_smem_desc(b_base + kk * 2048, 8192, 1024, 2)
                                        : make_smem_desc(b_base + kk * 1024, g.mn_lbo, g.mn_sbo, 1));
            if (BF16) umma_bf16(tmem_d, ad, bd, idesc, (kb | kk) != 0 ? 1u : 0u);
            else umma_tf32(tmem_d, ad, bd, idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar + s);
          if (kb == nkb - 1) umma_commit(tmem_full_bar + buf);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..5)
    const int q = warp & 3;
    float* stage = staging + (size_t)q * 32 * LDS;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++t) {
      int m0, n0, z, ks, zo, zi, kbeg, nkb;
      decode(tile, m0, n0, z, ks, zo, zi, kbeg, nkb);
      const int64_t coff = zo * g.sCo + zi * g.sCi;
      const uint32_t buf = t & 1;
      if (g.residual && g.split_k <= 1) {                  // pull this warp's residual rows into L2 ahead of their use
        const float* rbase = g.residual + coff + (int64_t)(m0 + q * 32) * g.ldr + n0;
        constexpr int LINES = BLOCK_N / 32;
        for (int i = lane; i < 32 * LINES; i += 32) {
          const int rr = i / LINES, cc = (i - rr * LINES) * 32;
          if (m0 + q * 32 + rr < g.M && n0 + cc < g.N)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rbase + (int64_t)rr * g.ldr + cc));
        }
      }
      mbar_wait(tmem_full_bar + buf, (t >> 1) & 1);
      tc_fence_after();
      __syncwarp();                                        // the previous tile's rows have left this warp's staging area
#pragma unroll
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {           // two 16-column TMEM loads in flight per wait
        uint32_t r0[16], r1[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BLOCK_N + (uint32_t)c0, r0);
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BLOCK_N + (uint32_t)c0 + 16u, r1);
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
          *reinterpret_cast<float4*>(stage + lane * LDS + c0 + j4) =
              make_float4(__uint_as_float(r0[j4]) * g.alpha, __uint_as_float(r0[j4 + 1]) * g.alpha,
                          __uint_as_float(r0[j4 + 2]) * g.alpha, __uint_as_float(r0[j4 + 3]) * g.alpha);
          *reinterpret_cast<float4*>(stage + lane * LDS + c0 + 16 + j4) =
              make_float4(__uint_as_float(r1[j4]) * g.alpha, __uint_as_float(r1[j4 + 1]) * g.alpha,
                          __uint_as_float(r1[j4 + 2]) * g.alpha, __uint_as_float(r1[j4 + 3]) * g.alpha);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar + buf);    // accumulator free again: the MMA warp may start tile t + 2 in it
      epilogue_rows<BLOCK_N, 8>(g, stage, lane, q, m0, n0, z, ks, coff);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

// 4-D view of one operand: (contiguous extent, rows extent [stride ld], inner batch [stride sI], outer batch [stride sO])
static bool encode_operand(CUtensorMap* tm, const void* base, int64_t contig, int64_t rows, int64_t ld, int bi,
                           int64_t sI, int bo, int64_t sO, int box_contig, int box_rows, bool mn_major, bool bf16) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  const int esz = bf16 ? 2 : 4, al = 16 / esz;          // strides must be multiples of 16 bytes
  if ((uintptr_t)base % 16 != 0 || ld % al != 0) return false;
  if (bi > 1 && sI % al != 0) return false;
  if (bo > 1 && sO % al != 0) return false;
  cuuint64_t dims[4] = {(cuuint64_t)contig, (cuuint64_t)rows, (cuuint64_t)(bi > 0 ? bi : 1), (cuuint64_t)(bo > 0 ? bo : 1)};
  // unused batch dims still need a legal (multiple of 16 B, non-zero) stride
  cuuint64_t st_i = (bi > 1 ? (cuuint64_t)sI : (cuuint64_t)ld * (cuuint64_t)rows) * esz;
  cuuint64_t st_o = (bo > 1 ? (cuuint64_t)sO : (cuuint64_t)ld * (cuuint64_t)rows) * esz;
  if (st_i == 0) st_i = 16; if (st_o == 0) st_o = 16;
  cuuint64_t strides[3] = {(cuuint64_t)ld * esz, st_i, st_o};
  cuuint32_t box[4] = {(cuuint32_t)box_contig, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   (mn_major && !bf16) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BLOCK_N, int STAGES, bool BF16 = false>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& g, bool a_mn, bool b_mn, int nbatch,
                     cudaStream_t s) {
  constexpr size_t ring = (size_t)STAGES * (TC_BLOCK_M * TC_BLOCK_K * 4 + BLOCK_N * TC_BLOCK_K * 4);
  constexpr size_t staging = (size_t)4 * 32 * (BLOCK_N + 4) * 4;     // epilogue staging reuses the ring
  constexpr size_t smem = (ring > staging ? ring : staging) + 1024 + 256;
  dim3 grid((unsigned)cdiv(g.N, BLOCK_N), (unsigned)cdiv(g.M, TC_BLOCK_M), (unsigned)(nbatch * g.split_k));
  dim3 block(192);
  // launches whose epilogue reads a residual or a pre-activation tensor take the variant that batches those loads
  const bool batched_epi = g.split_k <= 1 && (g.residual != nullptr || g.act == VU_ACT_GELU_BWD);
#define VU_TC_LAUNCH(AMN, BMN)                                                                                   \
  do {                                                                                                            \
    auto kfn = gemm_tf32_tc_kernel<BLOCK_N, STAGES, AMN, BMN, BF16, 1>;                                                 \
    auto kfn4 = gemm_tf32_tc_kernel<BLOCK_N, STAGES, AMN, BMN, BF16, 4>;                                                \
    static uint64_t seen = 0;                                                                                     \
    if (first_use_on_device(seen)) {                                                                              \
      if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||     \
          cudaFuncSetAttribute(kfn4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)      \
        return check_launch("vu_gemm(tc attr)");                                                                  \
    }                                                                                                             \
    if (batched_epi) kfn4<<<grid, block, smem, s>>>(tmA, tmB, g);                                                 \
    else kfn<<<grid, block, smem, s>>>(tmA, tmB, g);                                                              \
  } while (0)
  if constexpr (BF16 && BLOCK_N < 64) {
    if (b_mn) return fail_arg("vu_gemm", "MN-major bf16 B operand needs a 64-column tile");
    if (!a_mn) VU_TC_LAUNCH(false, false);
    else VU_TC_LAUNCH(true, false);
  } else {
    if (!a_mn && !b_mn) VU_TC_LAUNCH(false, false);
    else if (!a_mn && b_mn) VU_TC_LAUNCH(false, true);
    else if (a_mn && b_mn) VU_TC_LAUNCH(true, true);
    else VU_TC_LAUNCH(true, false);
  }
#undef VU_TC_LAUNCH
  return check_launch("vu_gemm(tc)");
}

// persistent launch: BLOCK_N in {64, 128}; stages fill what the 227 KB of shared memory leave after the staging rows
template <int BLOCK_N, bool BF16>
static int launch_tc_persistent(const CUtensorMap& tmA, const CUtensorMap& tmB, TcArgs g, bool a_mn, bool b_mn, int nbatch,
                                cudaStream_t s) {
  constexpr int STAGES = BLOCK_N == 128 ? 4 : 6;
  constexpr size_t ring = (size_t)STAGES * (TC_BLOCK_M * TC_BLOCK_K * 4 + BLOCK_N * TC_BLOCK_K * 4);
  constexpr size_t staging = (size_t)4 * 32 * (BLOCK_N + 4) * 4;
  constexpr size_t smem = ring + staging + 1024 + 256;
  static_assert(smem <= 227 * 1024, "persistent GEMM: shared memory budget");
  g.tiles_m = (int)cdiv(g.M, TC_BLOCK_M); g.tiles_n = (int)cdiv(g.N, BLOCK_N);
  const int64_t total = (int64_t)g.tiles_m * g.tiles_n * nbatch * g.split_k;
  if (total > 0x7fffffff) return fail_arg("vu_gemm", "too many tiles");
  g.total_tiles = (int)total;
  const unsigned grid = (unsigned)std::min<int64_t>(total, sm_count());
#define VU_TCP_LAUNCH(AMN, BMN)                                                                                  \
  do {                                                                                                            \
    auto kfn = gemm_tc_persistent_kernel<BLOCK_N, STAGES, AMN, BMN, BF16>;                                        \
    static uint64_t seen = 0;                                                                                     \
    if (first_use_on_device(seen)) {                                                                              \
      if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)       \
        return check_launch("vu_gemm(tc persistent attr)");                                                       \
    }                                                                                                             \
    kfn<<<grid, 192, smem, s>>>(tmA, tmB, g);                                                                     \
  } while (0)
  if (!a_mn && !b_mn) VU_TCP_LAUNCH(false, false);
  else if (!a_mn && b_mn) VU_TCP_LAUNCH(false, true);
  else if (a_mn && b_mn) VU_TCP_LAUNCH(true, true);
  else VU_TCP_LAUNCH(true, false);
#undef VU_TCP_LAUNCH
  return check_launch("vu_gemm(tc persistent)");
}

int gemm_tc(const vu_gemm_desc& d, cudaStream_t s, bool* handled) {
  *handled = false;
  const bool bf16 = d.a_bf16 != 0;
  if ((d.a_bf16 != 0) != (d.b_bf16 != 0)) return fail_arg("vu_gemm", "A and B must have the same element type");
  const bool a_mn = d.trans_a != 0;        // A(m,k) = A[k*lda + m]  -> m contiguous
  const bool b_mn = d.trans_b == 0;        // B(k,n) = B[k*ldb + n]  -> n contiguous
  if (d.c_bf16 && (d.split_k > 1 || d.accumulate))
    return fail_arg("vu_gemm", "bf16 output cannot be accumulated into (no accumulate / split_k)");
  if (d.aux_bf16 && !bf16) return fail_arg("vu_gemm", "bf16 aux tensors need bf16 operands");
  const int bi = d.batch_inner > 0 ? d.batch_inner : 1, bo = d.batch_outer > 0 ? d.batch_outer : 1;
  const int kb_elems = bf16 ? 64 : TC_BLOCK_K;
  int block_n = d.N <= 32 ? 32 : (d.N <= 64 ? 64 : 128);
  if (bf16 && b_mn && block_n < 64) block_n = 64;      // one MN-major bf16 slab is 64 columns (TMA zero-fills past N)
  const int kps0 = (int)cdiv(cdiv(d.K, d.split_k > 1 ? d.split_k : 1), kb_elems) * kb_elems;
  // one k-block contractions (K <= 32 floats / 64 bf16: the second FeedForward product and the first one's data gradient)
  // are output-bound: 64-column tiles with a single stage let 6 CTAs share an SM
  const bool one_kb_narrow = block_n == 128 && kps0 <= kb_elems && getenv("VU_TC_QK128") == nullptr;
  if (one_kb_narrow) block_n = 64;      // box width of the B operand must match the kernel's BLOCK_N
  CUtensorMap tmA, tmB;
  bool ok;
  if (!a_mn) ok = encode_operand(&tmA, d.A, d.K, d.M, d.lda, bi, d.sAi, bo, d.sAo, kb_elems, TC_BLOCK_M, false, bf16);
  else ok = encode_operand(&tmA, d.A, d.M, d.K, d.lda, bi, d.sAi, bo, d.sAo, bf16 ? 64 : 32, kb_elems, true, bf16);
  if (!ok) return bf16 ? fail_arg("vu_gemm", "bf16 operand A is not TMA-addressable (16-byte alignment of base/strides)") : VU_OK;
  if (!b_mn) ok = encode_operand(&tmB, d.B, d.K, d.N, d.ldb, bi, d.sBi, bo, d.sBo, kb_elems, block_n, false, bf16);
  else ok = encode_operand(&tmB, d.B, d.N, d.K, d.ldb, bi, d.sBi, bo, d.sBo, bf16 ? 64 : 32, kb_elems, true, bf16);
  if (!ok) return bf16 ? fail_arg("vu_gemm", "bf16 operand B is not TMA-addressable (16-byte alignment of base/strides)") : VU_OK;

  TcArgs g;
  g.C = d.C; g.bias = d.bias; g.residual = d.residual; g.aux_in = d.aux_in; g.aux_out = d.aux_out;
  g.M = d.M; g.N = d.N; g.K = d.K; g.ldc = d.ldc; g.ldr = d.ldr; g.ldaux = d.ldaux;
  g.batch_inner = bi; g.sCo = d.sCo; g.sCi = d.sCi;
  g.alpha = d.alpha; g.act = d.act; g.accumulate = d.accumulate;
  g.split_k = d.split_k > 1 ? d.split_k : 1;
  int kps = (int)cdiv(cdiv(g.K, g.split_k), kb_elems) * kb_elems;
  g.k_per_split = kps;
  g.split_k = (int)cdiv(g.K, kps);
  g.drop_thresh = d.drop_p > 0.f ? drop_threshold(d.drop_p) : 0u;
  g.drop_scale = drop_keep_scale(d.drop_p);
  g.drop_seed = d.drop_seed; g.drop_stream = d.drop_stream;
  g.mn_lbo = 4096; g.mn_sbo = 512;
  g.c_bf16 = d.c_bf16 != 0;
  g.aux_bf16 = d.aux_bf16 != 0;
  const int nbatch = bi * bo;
  int rc;
  // OPT-IN (VU_TC_PERSISTENT=1, read per call): token GEMMs (no (image, head) batch) on the persistent kernel, 64-column
  // tiles for N <= 192 (192 = 3 x 64 exactly).  Measured on B200 inside the Base step at 256 images (tools/ab_persistent.sh,
  // bf16 mode): 8.75 ms per step for the token GEMMs against 7.76 ms with the one-tile-per-CTA kernel below -- once the
  // epilogue's fast path made a tile's epilogue cheap, three resident CTAs per SM (12 epilogue warps, three independent
  // MMA streams) beat one persistent CTA (4 epilogue warps, a 4-stage ring that covers only ~0.55 us of TMA latency);
  // the persistent kernel wins only on the K = 768 projections (86 vs 94 us).  It would need 128 x 256 tiles / CTA pairs.
  const char* pe = getenv("VU_TC_PERSISTENT");
  const bool persistent_on = pe && pe[0] == '1';
  if (persistent_on && nbatch == 1 && d.N > 32) {
    const bool n64 = d.N <= 64 || d.N == 192 || (bf16 && b_mn && d.N < 128);
    const int pbn = n64 ? 64 : 128;
    bool okp = true;
    if (pbn != block_n) {        // the B tensor map was encoded for another tile width: re-encode
      if (!b_mn) okp = encode_operand(&tmB, d.B, d.K, d.N, d.ldb, bi, d.sBi, bo, d.sBo, kb_elems, pbn, false, bf16);
      else okp = encode_operand(&tmB, d.B, d.N, d.K, d.ldb, bi, d.sBi, bo, d.sBo, bf16 ? 64 : 32, kb_elems, true, bf16);
    }
    if (okp) {
      if (bf16) rc = pbn == 64 ? launch_tc_persistent<64, true>(tmA, tmB, g, a_mn, b_mn, nbatch, s)
                               : launch_tc_persistent<128, true>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
      else rc = pbn == 64 ? launch_tc_persistent<64, false>(tmA, tmB, g, a_mn, b_mn, nbatch, s)
                          : launch_tc_persistent<128, false>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
      *handled = true;
      return rc;
    }
  }
  if (bf16) {
    // map-reading products (N = head_dim <= 32, K = tokens): pure HBM streams of the bf16 map.  Two pipeline stages
    // (41 KB of shared memory) let 5 CTAs share an SM and hide each other's prologue / epilogue: 4.27 vs 3.89 TB/s
    // with four stages and 2 CTAs per SM (VU_TC_BF16_STAGES overrides)
    static const int st32 = []() { const char* e = getenv("VU_TC_BF16_STAGES"); return e ? atoi(e) : 2; }();
    if (block_n == 32 && st32 == 2) rc = launch_tc<32, 2, true>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
    else if (block_n == 32 && st32 == 3) rc = launch_tc<32, 3, true>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
    else if (block_n == 32) rc = launch_tc<32, 4, true>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
    // token GEMMs in the bf16 mode: short contractions (<= 32 k-blocks of 64) are output-bound -> two stages, more CTAs per SM
    else if (one_kb_narrow) rc = launch_tc<64, 1, true>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
    else if (block_n == 64) {
      if (g.k_per_split <= 2048) rc = launch_tc<64, 2, true>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
      else rc = launch_tc<64, 4, true>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
    } else {
      if (g.k_per_split <= 2048) rc = launch_tc<128, 2, true>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
      else rc = launch_tc<128, 3, true>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
    }
  } else if (block_n == 32) {
    if (g.k_per_split <= 1024) rc = launch_tc<32, 2>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
    else rc = launch_tc<32, 4>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
  }
  else if (one_kb_narrow) rc = launch_tc<64, 1>(tmA, tmB, g, a_mn, b_mn, nbatch, s);   // one k-block, output-bound: 6 CTAs/SM
  else if (block_n == 64) {
    static const int st64 = []() { const char* e = getenv("VU_TC_STAGES64"); return e ? atoi(e) : 0; }();
    const bool two = st64 ? st64 == 2 : g.k_per_split <= 1024;
    if (two) rc = launch_tc<64, 2>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
    else rc = launch_tc<64, 4>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
  }
  else if (g.k_per_split <= TC_BLOCK_K) rc = launch_tc<128, 1>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
  else {
    // short contractions are output / HBM bound: two stages (3 CTAs per SM overlap each other's epilogue) beat three
    // (measured per shape with tools/gemm_bench.py: K <= 768 gains 6-21 %, K >= 3072 loses 3-25 %)
    static const int st128 = []() { const char* e = getenv("VU_TC_STAGES128"); return e ? atoi(e) : 0; }();
    const bool two = st128 ? st128 == 2 : g.k_per_split <= 1024;
    if (two) rc = launch_tc<128, 2>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
    else rc = launch_tc<128, 3>(tmA, tmB, g, a_mn, b_mn, nbatch, s);
  }
  *handled = true;
  return rc;
}

}  // namespace vu
