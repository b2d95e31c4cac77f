// Observation crops: torchvision.ops.roi_align exactly as the reference calls it
// (reference: src/megapose/lib3d/cropping.py:113-144 crop_images -- output_size (240,320),
//  spatial_scale 1, sampling_ratio 4, aligned=False; RGB-D branch masks crop depth where the
//  averaged validity mask < 0.99).
//
// The reference first materialises observation.images[batch_im_ids] (one full-resolution copy per
// hypothesis, inference/pose_estimator.py:389); here the crop kernel indexes the frame through
// d_im_idx instead, and the frame is packed once to NHWC4 so that each bilinear tap is one 16-byte
// load that hits L2.  For single-view samples (coarse / scoring) the crop is computed inside the
// rasteriser's resolve pass instead (raster.cu) so that whole pixel vectors are written once.
#include "crop_device.cuh"

namespace mpx {

__global__ void image_to_nhwc4_kernel(const float* __restrict__ in, int b, int c, int h, int w,
                                      float4* __restrict__ out) {
  const long long total = static_cast<long long>(b) * h * w;
  const long long plane = static_cast<long long>(h) * w;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long img = i / plane, pix = i - img * plane;
    const float* p = in + img * c * plane + pix;
    float4 v;
    v.x = p[0];
    v.y = p[plane];
    v.z = p[2 * plane];
    v.w = c > 3 ? p[3 * plane] : 0.f;
    out[i] = v;
  }
}

int image_to_nhwc4(const float* in, int b, int c, int h, int w, float* out, cudaStream_t stream) {
  MPX_REQUIRE(c == 3 || c == 4, "image_to_nhwc4: C=%d must be 3 or 4", c);
  const long long total = static_cast<long long>(b) * h * w;
  if (total == 0) return MPX_OK;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  image_to_nhwc4_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(in, b, c, h, w,
                                                                       reinterpret_cast<float4*>(out));
  MPX_CHECK_CUDA(cudaGetLastError());
  ++g_launches;
  return MPX_OK;
}

__global__ void __launch_bounds__(256)
roi_align_kernel(const float4* __restrict__ images, int b, int h, int w, const int* __restrict__ im_idx,
                 const float* __restrict__ boxes, int n, int c, int oh, int ow, CropOut out) {
  const int roi = blockIdx.x;
  const int npix = oh * ow;
  const RoiParams rp = make_roi(boxes + 4 * roi, oh, ow);
  int im = im_idx ? im_idx[roi] : roi;
  const bool im_ok = im >= 0 && im < b;
  const float4* img = images + static_cast<size_t>(im_ok ? im : 0) * h * w;
  __shared__ AxisW s_axis[kAxisTableMax];
  bool collapsed = false;  // block-uniform
  if (im_ok && oh + ow <= kAxisTableMax) collapsed = build_axis_tables(rp, oh, ow, h, w, s_axis, s_axis + oh);
  for (int pix = blockIdx.y * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.y * blockDim.x) {
    const int ph = pix / ow, pw = pix - ph * ow;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float vacc = 0.f;
    if (collapsed) {
      if (c == 4) roi_align_pixel_collapsed<true>(img, w, s_axis[ph], s_axis[oh + pw], acc, vacc);
      else roi_align_pixel_collapsed<false>(img, w, s_axis[ph], s_axis[oh + pw], acc, vacc);
    } else if (im_ok) {
      if (c == 4) roi_align_pixel<true>(img, h, w, rp, ph, pw, acc, vacc);
      else roi_align_pixel<false>(img, h, w, rp, ph, pw, acc, vacc);
    }
    if (c == 4 && vacc < 0.99f) acc.w = 0.f;
    if (out.nchw) {
      float* o = out.nchw + static_cast<size_t>(roi) * c * npix + pix;
      o[0] = acc.x; o[npix] = acc.y; o[2 * npix] = acc.z;
      if (c == 4) o[3 * npix] = acc.w;
    }
    if (out.x) {
      const int hs = oh >> 1, ws = ow >> 1;
      act_t* o = out.x + ((static_cast<size_t>(roi) * hs + (ph >> 1)) * ws + (pw >> 1)) * (4 * out.c_pad) +
                         ((ph & 1) * 2 + (pw & 1)) * out.c_pad;
      o[0] = to_act(acc.x);
      o[1] = to_act(acc.y);
      o[2] = to_act(acc.z);
      if (c == 4) {
        float d = acc.w;
        if (out.depth_norm_z) d = depth_norm(d, __ldg(out.depth_norm_z + roi), out.depth_norm_kind);
        o[3] = to_act(d);
      }
    }
  }
}

int roi_align_launch(const float* images, int b, int h, int w, const int* im_idx, const float* boxes, int n,
                     int c, int oh, int ow, const CropOut& out, cudaStream_t stream) {
  MPX_REQUIRE(c == 3 || c == 4, "roi_align: C=%d must be 3 or 4", c);
  MPX_REQUIRE(oh > 0 && ow > 0, "roi_align: empty output");
  if (out.x) MPX_REQUIRE(oh % 2 == 0 && ow % 2 == 0, "roi_align: fused output needs even size");
  if (n == 0) return MPX_OK;
  const int npix = oh * ow;
  int bx = (npix + 255) / 256;
  if (bx > 64) bx = 64;
  dim3 grid(n, bx);
  roi_align_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(images), b, h, w, im_idx, boxes, n,
                                             c, oh, ow, out);
  MPX_CHECK_CUDA(cudaGetLastError());
  ++g_launches;
  return MPX_OK;
}

}  // namespace mpx
