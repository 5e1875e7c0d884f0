// Device half of the reference's input pipeline: DenoisingDataset.__getitem__ (dataset.py:56-70) and the
// albumentations transforms of run_denoising.py:52-58, per batch instead of per sample in DataLoader workers:
//   cv2.resize(img, (im_size, im_size))                      -> vu_resize_u8hwc   (INTER_LINEAR, half-pixel centres, uint8 out)
//   ShiftScaleRotate(border_mode=BORDER_CONSTANT)            -> vu_warp_u8hwc_to_chw with a per-image inverse affine map
//   Normalize(mean, std, max_pixel_value=255) ; /255 ; HWC -> CHW float   (fused into the same kernel's store)
// Both kernels are HBM-bound gathers over uint8 images: one thread per output pixel (all channels), reads hit L1/L2
// (each source texel is used by ~4 outputs), writes are coalesced along x.
#include "vu_common.cuh"

namespace vu {

__global__ void __launch_bounds__(256)
resize_u8hwc_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int B, int C, int Hs, int Ws, int Hd, int Wd,
                    float sy, float sx) {
  const int64_t total = (int64_t)B * Hd * Wd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wd), y = (int)((i / Wd) % Hd), b = (int)(i / ((int64_t)Wd * Hd));
    // cv2.resize INTER_LINEAR: source coordinate of the pixel CENTRE, clamped to the image (replicated border)
    float fy = (y + 0.5f) * sy - 0.5f, fx = (x + 0.5f) * sx - 0.5f;
    int y0 = (int)floorf(fy), x0 = (int)floorf(fx);
    float wy = fy - y0, wx = fx - x0;
    if (y0 < 0) { y0 = 0; wy = 0.f; }
    if (x0 < 0) { x0 = 0; wx = 0.f; }
    int y1 = y0 + 1, x1 = x0 + 1;
    if (y1 >= Hs) { y1 = Hs - 1; if (y0 >= Hs - 1) { y0 = Hs - 1; wy = 0.f; } }
    if (x1 >= Ws) { x1 = Ws - 1; if (x0 >= Ws - 1) { x0 = Ws - 1; wx = 0.f; } }
    const uint8_t* s = src + (int64_t)b * Hs * Ws * C;
    uint8_t* d = dst + i * C;
    for (int c = 0; c < C; ++c) {
      const float v00 = s[((int64_t)y0 * Ws + x0) * C + c], v01 = s[((int64_t)y0 * Ws + x1) * C + c];
      const float v10 = s[((int64_t)y1 * Ws + x0) * C + c], v11 = s[((int64_t)y1 * Ws + x1) * C + c];
      const float v = (v00 * (1.f - wx) + v01 * wx) * (1.f - wy) + (v10 * (1.f - wx) + v11 * wx) * wy;
      d[c] = (uint8_t)min(255, max(0, __float2int_rn(v)));
    }
  }
}

// dst[b, c, y, x] = ((sample(src_b, M_b (x, y)) * scale) - mean) / std * post ; M_b: 2x3 map from OUTPUT pixel to SOURCE pixel
// coordinates (the inverse of the matrix cv2.warpAffine is given).  interp 1: bilinear over the 4 neighbours, each
// outside-the-image neighbour contributing `border` (BORDER_CONSTANT); interp 0: nearest.  round_u8: round the sampled
// value to an integer grey level first (what cv2 returns for uint8 images).
__global__ void __launch_bounds__(256)
warp_u8hwc_to_chw_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, const float* __restrict__ mats, int B, int C,
                         int Hs, int Ws, int Hd, int Wd, int interp, float border, int round_u8, float scale, float mean,
                         float inv_std, float post) {
  const int64_t total = (int64_t)B * Hd * Wd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wd), y = (int)((i / Wd) % Hd), b = (int)(i / ((int64_t)Wd * Hd));
    float fx = (float)x, fy = (float)y;
    if (mats) {
      const float* m = mats + b * 6;
      fx = m[0] * x + m[1] * y + m[2]; fy = m[3] * x + m[4] * y + m[5];
    }
    const uint8_t* s = src + (int64_t)b * Hs * Ws * C;
    auto at = [&](int yy, int xx, int c) -> float {
      return ((unsigned)yy < (unsigned)Hs && (unsigned)xx < (unsigned)Ws) ? (float)s[((int64_t)yy * Ws + xx) * C + c] : border;
    };
    for (int c = 0; c < C; ++c) {
      float v;
      if (interp) {
        const int y0 = (int)floorf(fy), x0 = (int)floorf(fx);
        const float wy = fy - y0, wx = fx - x0;
        v = (at(y0, x0, c) * (1.f - wx) + at(y0, x0 + 1, c) * wx) * (1.f - wy) +
            (at(y0 + 1, x0, c) * (1.f - wx) + at(y0 + 1, x0 + 1, c) * wx) * wy;
      } else {
        v = at(__float2int_rn(fy), __float2int_rn(fx), c);
      }
      if (round_u8) v = fminf(255.f, fmaxf(0.f, rintf(v)));
      dst[(((int64_t)b * C + c) * Hd + y) * Wd + x] = (v * scale - mean) * inv_std * post;
    }
  }
}

static int ew_blocks(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 256), (int64_t)sm_count() * 16)); }

}  // namespace vu

extern "C" int vu_resize_u8hwc(const uint8_t* src, uint8_t* dst, int B, int C, int Hs, int Ws, int Hd, int Wd, void* stream) {
  using namespace vu;
  const char* fn = "vu_resize_u8hwc";
  VU_REQUIRE(src && dst && B > 0 && C > 0 && C <= 4 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0, fn, "bad arguments");
  resize_u8hwc_kernel<<<ew_blocks((int64_t)B * Hd * Wd), 256, 0, as_stream(stream)>>>(src, dst, B, C, Hs, Ws, Hd, Wd,
                                                                                      (float)Hs / (float)Hd, (float)Ws / (float)Wd);
  return check_launch(fn);
}

extern "C" int vu_warp_u8hwc_to_chw(const uint8_t* src, float* dst, const float* mats, int B, int C, int Hs, int Ws, int Hd,
                                    int Wd, int interp, float border, int round_u8, float scale, float mean, float std,
                                    float post, void* stream) {
  using namespace vu;
  const char* fn = "vu_warp_u8hwc_to_chw";
  VU_REQUIRE(src && dst && B > 0 && C > 0 && C <= 4 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0, fn, "bad arguments");
  VU_REQUIRE(std != 0.f && (interp == 0 || interp == 1), fn, "std must be non-zero, interp 0 (nearest) or 1 (bilinear)");
  warp_u8hwc_to_chw_kernel<<<ew_blocks((int64_t)B * Hd * Wd), 256, 0, as_stream(stream)>>>(
      src, dst, mats, B, C, Hs, Ws, Hd, Wd, interp, border, round_u8, scale, mean, 1.0f / std, post);
  return check_launch(fn);
}
