// Device-side roi_align arithmetic shared by crop.cu (stand-alone crop kernel) and raster.cu (crop fused
// into the render resolve pass).  torchvision.ops.roi_align semantics as used by the reference
// (src/megapose/lib3d/cropping.py:113-144: sampling_ratio 4, aligned=False, spatial_scale 1).
#pragma once
#include "mpx_common.cuh"

namespace mpx {

struct RoiParams {
  float x1, y1, bin_w, bin_h;
};

__device__ __forceinline__ RoiParams make_roi(const float* __restrict__ box, int oh, int ow) {
  RoiParams r;
  const float x1 = box[0], y1 = box[1], x2 = box[2], y2 = box[3];
  const float roi_w = fmaxf(x2 - x1, 1.f), roi_h = fmaxf(y2 - y1, 1.f);
  r.x1 = x1;
  r.y1 = y1;
  r.bin_w = roi_w / static_cast<float>(ow);
  r.bin_h = roi_h / static_cast<float>(oh);
  return r;
}

// Average of the 4x4 bilinear samples of output pixel (ph, pw); acc.w carries depth, vacc the averaged
// depth-validity mask.  The 2x2 texel block of the previous sample is kept in registers: with bins smaller
// than a pixel most of the 16 samples share it, which removes ~3/4 of the loads without changing a single
// floating-point operation.
__device__ __forceinline__ void roi_align_pixel(const float4* __restrict__ img, int h, int w, const RoiParams& r,
                                                int ph, int pw, float4& acc, float& vacc) {
  acc = make_float4(0.f, 0.f, 0.f, 0.f);
  vacc = 0.f;
  int cy0 = -1, cy1 = -1, cx0 = -1, cx1 = -1;
  float4 v1 = acc, v2 = acc, v3 = acc, v4 = acc;
#pragma unroll
  for (int iy = 0; iy < 4; ++iy) {
    float y = r.y1 + ph * r.bin_h + (iy + 0.5f) * r.bin_h / 4.f;
#pragma unroll
    for (int ix = 0; ix < 4; ++ix) {
      float x = r.x1 + pw * r.bin_w + (ix + 0.5f) * r.bin_w / 4.f;
      float yy = y;
      if (yy < -1.0f || yy > static_cast<float>(h) || x < -1.0f || x > static_cast<float>(w)) continue;
      if (yy <= 0.f) yy = 0.f;
      if (x <= 0.f) x = 0.f;
      int y_low = static_cast<int>(yy), x_low = static_cast<int>(x);
      int y_high, x_high;
      if (y_low >= h - 1) {
        y_high = y_low = h - 1;
        yy = static_cast<float>(y_low);
      } else {
        y_high = y_low + 1;
      }
      if (x_low >= w - 1) {
        x_high = x_low = w - 1;
        x = static_cast<float>(x_low);
      } else {
        x_high = x_low + 1;
      }
      if (y_low != cy0 || y_high != cy1 || x_low != cx0 || x_high != cx1) {
        v1 = __ldg(img + y_low * w + x_low);
        v2 = __ldg(img + y_low * w + x_high);
        v3 = __ldg(img + y_high * w + x_low);
        v4 = __ldg(img + y_high * w + x_high);
        cy0 = y_low; cy1 = y_high; cx0 = x_low; cx1 = x_high;
      }
      const float ly = yy - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
      const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
      acc.x += w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
      acc.y += w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
      acc.z += w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
      acc.w += w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
      vacc += w1 * (v1.w > 0.f ? 1.f : 0.f) + w2 * (v2.w > 0.f ? 1.f : 0.f) + w3 * (v3.w > 0.f ? 1.f : 0.f) +
              w4 * (v4.w > 0.f ? 1.f : 0.f);
    }
  }
  acc.x /= 16.f;
  acc.y /= 16.f;
  acc.z /= 16.f;
  acc.w /= 16.f;
  vacc /= 16.f;
}

}  // namespace mpx
