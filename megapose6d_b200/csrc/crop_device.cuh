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

// One axis of a bilinear sample, torchvision rules: coordinate -> (low, high, l, h) or invalid.
struct AxisTap {
  int lo, hi;
  float l, h;
  bool ok;
};
__device__ __forceinline__ AxisTap axis_tap(float v, int size) {
  AxisTap t;
  t.ok = !(v < -1.0f || v > static_cast<float>(size));
  if (v <= 0.f) v = 0.f;
  int lo = static_cast<int>(v);
  if (lo >= size - 1) {
    t.hi = lo = size - 1;
    v = static_cast<float>(lo);
  } else {
    t.hi = lo + 1;
  }
  t.lo = lo;
  t.l = v - lo;
  t.h = 1.f - t.l;
  return t;
}

// Average of the 4x4 bilinear samples of output pixel (ph, pw); acc.w carries depth, vacc the averaged
// depth-validity mask (only accumulated when WITH_DEPTH).  The per-axis work (bounds, floor, weights) is done
// once per sample row / column instead of once per sample, and the 2x2 texel block of the previous sample is kept
// in registers: with bins smaller than a pixel most of the 16 samples share it.  Every floating-point operation
// and its order are those of torchvision's kernel.
template <bool WITH_DEPTH>
__device__ __forceinline__ void roi_align_pixel(const float4* __restrict__ img, int h, int w, const RoiParams& r,
                                                int ph, int pw, float4& acc, float& vacc) {
  acc = make_float4(0.f, 0.f, 0.f, 0.f);
  vacc = 0.f;
  AxisTap tx[4];
#pragma unroll
  for (int ix = 0; ix < 4; ++ix) tx[ix] = axis_tap(r.x1 + pw * r.bin_w + (ix + 0.5f) * r.bin_w / 4.f, w);
  int cy0 = -1, cy1 = -1, cx0 = -1, cx1 = -1;
  float4 v1 = acc, v2 = acc, v3 = acc, v4 = acc;
#pragma unroll
  for (int iy = 0; iy < 4; ++iy) {
    const AxisTap ty = axis_tap(r.y1 + ph * r.bin_h + (iy + 0.5f) * r.bin_h / 4.f, h);
    if (!ty.ok) continue;
#pragma unroll
    for (int ix = 0; ix < 4; ++ix) {
      if (!tx[ix].ok) continue;
      if (ty.lo != cy0 || ty.hi != cy1 || tx[ix].lo != cx0 || tx[ix].hi != cx1) {
        v1 = __ldg(img + ty.lo * w + tx[ix].lo);
        v2 = __ldg(img + ty.lo * w + tx[ix].hi);
        v3 = __ldg(img + ty.hi * w + tx[ix].lo);
        v4 = __ldg(img + ty.hi * w + tx[ix].hi);
        cy0 = ty.lo; cy1 = ty.hi; cx0 = tx[ix].lo; cx1 = tx[ix].hi;
      }
      const float w1 = ty.h * tx[ix].h, w2 = ty.h * tx[ix].l, w3 = ty.l * tx[ix].h, w4 = ty.l * tx[ix].l;
      acc.x += w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
      acc.y += w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
      acc.z += w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
      if (WITH_DEPTH) {
        acc.w += w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
        vacc += w1 * (v1.w > 0.f ? 1.f : 0.f) + w2 * (v2.w > 0.f ? 1.f : 0.f) + w3 * (v3.w > 0.f ? 1.f : 0.f) +
                w4 * (v4.w > 0.f ? 1.f : 0.f);
      }
    }
  }
  acc.x /= 16.f;
  acc.y /= 16.f;
  acc.z /= 16.f;
  if (WITH_DEPTH) {
    acc.w /= 16.f;
    vacc /= 16.f;
  }
}

}  // namespace mpx
