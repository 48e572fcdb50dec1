// Stand-alone homography warp: tf_transform_homography (homography_warping.py:211-253, zero-fill
// bilinear of tf.contrib.image.transform) and the legacy homography_warping / interpolate
// (:131-210, clamp-gather).  The fused cost-volume kernel (cost_volume.cu) is the hot path; this
// one backs the reference's public warp functions and the "warped features <= 1e-5" parity gate.
#include "geometry.cuh"

namespace mvsb200 {

// legacy `interpolate` (homography_warping.py:131-174): weights from the CLAMPED corners, a,b,c,d
// summed in add_n order, every op rounded separately so the cancellations of Appendix A.3 are exact.
template <int VEC>
__device__ __forceinline__ void legacy_sample(const float* __restrict__ img, int width, int height, int channels,
                                              float xw, float yw, int c, float* out) {
  float x = sub_(xw, 0.5f), y = sub_(yw, 0.5f);
  int x0 = floor_to_int(x), y0 = floor_to_int(y);
  int x1 = x0 + 1, y1 = y0 + 1;
  x0 = min(max(x0, 0), width - 1);  x1 = min(max(x1, 0), width - 1);
  y0 = min(max(y0, 0), height - 1); y1 = min(max(y1, 0), height - 1);
  float x0f = (float)x0, x1f = (float)x1, y0f = (float)y0, y1f = (float)y1;
  float wa = mul_(sub_(y1f, y), sub_(x1f, x));
  float wb = mul_(sub_(y1f, y), sub_(x, x0f));
  float wc = mul_(sub_(y, y0f), sub_(x1f, x));
  float wd = mul_(sub_(y, y0f), sub_(x, x0f));
  const float* pa = img + ((size_t)y0 * width + x0) * channels + c;
  const float* pb = img + ((size_t)y0 * width + x1) * channels + c;
  const float* pc = img + ((size_t)y1 * width + x0) * channels + c;
  const float* pd = img + ((size_t)y1 * width + x1) * channels + c;
#pragma unroll
  for (int k = 0; k < VEC; ++k)
    out[k] = add_(add_(add_(mul_(wa, __ldg(pa + k)), mul_(wb, __ldg(pb + k))), mul_(wc, __ldg(pc + k))),
                  mul_(wd, __ldg(pd + k)));
}

template <int VEC>
__device__ __forceinline__ void transform_sample(const float* __restrict__ img, int width, int height,
                                                 int channels, float ix, float iy, int c, float* out) {
  Footprint f = make_footprint(ix, iy, width, height);
  float p00[VEC], p01[VEC], p10[VEC], p11[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) p00[k] = p01[k] = p10[k] = p11[k] = 0.0f;
  const float* base = img + ((int64_t)f.y0 * width + f.x0) * channels + c;
  const int64_t row = (int64_t)width * channels;
  if (f.vy0 && f.vx0) for (int k = 0; k < VEC; ++k) p00[k] = __ldg(base + k);
  if (f.vy0 && f.vx1) for (int k = 0; k < VEC; ++k) p01[k] = __ldg(base + channels + k);
  if (f.vy1 && f.vx0) for (int k = 0; k < VEC; ++k) p10[k] = __ldg(base + row + k);
  if (f.vy1 && f.vx1) for (int k = 0; k < VEC; ++k) p11[k] = __ldg(base + row + channels + k);
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    float v0 = f.wxl * p00[k] + f.wxr * p01[k];
    float v1 = f.wxl * p10[k] + f.wxr * p11[k];
    out[k] = f.wyl * v0 + f.wyr * v1;
  }
}

template <int VEC, int SAMPLER>
__global__ void warp_kernel(const float* __restrict__ image, size_t image_stride,
                            const float* __restrict__ homographies, int height, int width, int channels,
                            float* __restrict__ out) {
  const int groups = channels / VEC;
  const size_t per_image = (size_t)height * width * groups;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= per_image) return;
  int g = (int)(idx % groups);
  size_t pix = idx / groups;
  int x = (int)(pix % width), y = (int)(pix / width);
  float h[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) h[i] = __ldg(homographies + (size_t)b * 9 + i);
  const float* img = image + (size_t)b * image_stride;
  float res[VEC];
  if (SAMPLER == MVSB200_SAMPLER_TRANSFORM) {
    float t[8], ix, iy;
    transform_coefs(h, t);
    transform_coords(t, (float)x, (float)y, ix, iy);
    transform_sample<VEC>(img, width, height, channels, ix, iy, g * VEC, res);
  } else {
    float xw, yw;
    legacy_coords(h, x, y, width, height, xw, yw);
    legacy_sample<VEC>(img, width, height, channels, xw, yw, g * VEC, res);
  }
  float* o = out + ((size_t)b * height * width + pix) * channels + g * VEC;
#pragma unroll
  for (int k = 0; k < VEC; ++k) o[k] = res[k];
}

template <int SAMPLER>
__global__ void sample_coords_kernel(const float* __restrict__ homographies, int height, int width,
                                     float* __restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= (size_t)height * width) return;
  int x = (int)(idx % width), y = (int)(idx / width);
  float h[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) h[i] = __ldg(homographies + (size_t)b * 9 + i);
  float cx, cy;
  if (SAMPLER == MVSB200_SAMPLER_TRANSFORM) {
    float t[8];
    transform_coefs(h, t);
    transform_coords(t, (float)x, (float)y, cx, cy);
  } else {
    legacy_coords(h, x, y, width, height, cx, cy);
  }
  float* o = out + ((size_t)b * height * width + idx) * 2;
  o[0] = cx;
  o[1] = cy;
}

// legacy interpolate() on caller-supplied image coordinates (homography_warping.py:131-174)
template <int VEC>
__global__ void interpolate_kernel(const float* __restrict__ image, const float* __restrict__ xs,
                                   const float* __restrict__ ys, int batch, int height, int width, int channels,
                                   float* __restrict__ out) {
  const int groups = channels / VEC;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)batch * height * width * groups;
  if (idx >= total) return;
  int g = (int)(idx % groups);
  size_t pix = idx / groups;                         // flat b*H*W index, as in the reference
  int b = (int)(pix / ((size_t)height * width));
  float res[VEC];
  legacy_sample<VEC>(image + (size_t)b * height * width * channels, width, height, channels, xs[pix], ys[pix],
                     g * VEC, res);
#pragma unroll
  for (int k = 0; k < VEC; ++k) out[pix * channels + g * VEC + k] = res[k];
}

// get_pixel_grids (homography_warping.py:108-117): [x(H*W) | y(H*W) | 1(H*W)]
__global__ void pixel_grids_kernel(int height, int width, float* __restrict__ out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int n = height * width;
  if (idx >= n) return;
  int x = idx % width, y = idx / width;
  out[idx] = tf_linspace_at(0.5f, sub_((float)width, 0.5f), width, x);
  out[n + idx] = tf_linspace_at(0.5f, sub_((float)height, 0.5f), height, y);
  out[2 * n + idx] = 1.0f;
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_interpolate(const float* image, const float* xs, const float* ys, int batch, int height,
                                   int width, int channels, float* out, void* stream) {
  MVS_CHECK_ARG(image && xs && ys && out, "interpolate: NULL pointer");
  MVS_CHECK_ARG(batch >= 0 && height > 0 && width > 0 && channels > 0, "interpolate: bad shape");
  if (batch == 0) return MVSB200_OK;
  const int vec = channels % 4 == 0 ? 4 : 1;
  size_t total = (size_t)batch * height * width * (channels / vec);
  unsigned blocks = (unsigned)((total + 255) / 256);
  cudaStream_t s = (cudaStream_t)stream;
  if (vec == 4) interpolate_kernel<4><<<blocks, 256, 0, s>>>(image, xs, ys, batch, height, width, channels, out);
  else interpolate_kernel<1><<<blocks, 256, 0, s>>>(image, xs, ys, batch, height, width, channels, out);
  MVS_LAUNCH_CHECK("interpolate_kernel");
  return MVSB200_OK;
}

extern "C" int mvsb200_pixel_grids(int height, int width, float* out, void* stream) {
  MVS_CHECK_ARG(out && height > 0 && width > 0, "pixel_grids: bad arguments");
  pixel_grids_kernel<<<ceil_div(height * width, 256), 256, 0, (cudaStream_t)stream>>>(height, width, out);
  MVS_LAUNCH_CHECK("pixel_grids_kernel");
  return MVSB200_OK;
}

extern "C" int mvsb200_warp(const float* image, int image_count, const float* homographies, int hom_count,
                            int height, int width, int channels, int sampler, float* out, void* stream) {
  MVS_CHECK_ARG(image && homographies && out, "warp: NULL pointer");
  MVS_CHECK_ARG(height > 0 && width > 0 && channels > 0 && hom_count >= 0, "warp: bad shape %dx%dx%d", height,
                width, channels);
  MVS_CHECK_ARG(image_count == hom_count || image_count == 1,
                "warp: image_count (%d) must be 1 or equal hom_count (%d)", image_count, hom_count);
  MVS_CHECK_ARG(sampler == MVSB200_SAMPLER_TRANSFORM || sampler == MVSB200_SAMPLER_LEGACY, "warp: bad sampler %d",
                sampler);
  if (hom_count == 0) return MVSB200_OK;
  MVS_CHECK_ARG(hom_count <= 65535, "warp: hom_count %d exceeds 65535 per call", hom_count);
  cudaStream_t s = (cudaStream_t)stream;
  size_t image_stride = image_count == 1 ? 0 : (size_t)height * width * channels;
  const int vec = (channels % 4 == 0) ? 4 : 1;
  size_t per_image = (size_t)height * width * (channels / vec);
  dim3 grid((unsigned)((per_image + 255) / 256), (unsigned)hom_count);
#define LAUNCH_WARP(V, S) \
  warp_kernel<V, S><<<grid, 256, 0, s>>>(image, image_stride, homographies, height, width, channels, out)
  if (sampler == MVSB200_SAMPLER_TRANSFORM) {
    if (vec == 4) LAUNCH_WARP(4, MVSB200_SAMPLER_TRANSFORM); else LAUNCH_WARP(1, MVSB200_SAMPLER_TRANSFORM);
  } else {
    if (vec == 4) LAUNCH_WARP(4, MVSB200_SAMPLER_LEGACY); else LAUNCH_WARP(1, MVSB200_SAMPLER_LEGACY);
  }
#undef LAUNCH_WARP
  MVS_LAUNCH_CHECK("warp_kernel");
  return MVSB200_OK;
}

extern "C" int mvsb200_sample_coords(const float* homographies, int hom_count, int height, int width,
                                     int sampler, float* out, void* stream) {
  MVS_CHECK_ARG(homographies && out && height > 0 && width > 0 && hom_count >= 0, "sample_coords: bad arguments");
  MVS_CHECK_ARG(sampler == MVSB200_SAMPLER_TRANSFORM || sampler == MVSB200_SAMPLER_LEGACY,
                "sample_coords: bad sampler %d", sampler);
  if (hom_count == 0) return MVSB200_OK;
  MVS_CHECK_ARG(hom_count <= 65535, "sample_coords: hom_count %d exceeds 65535 per call", hom_count);
  dim3 grid((unsigned)(((size_t)height * width + 255) / 256), (unsigned)hom_count);
  cudaStream_t s = (cudaStream_t)stream;
  if (sampler == MVSB200_SAMPLER_TRANSFORM)
    sample_coords_kernel<MVSB200_SAMPLER_TRANSFORM><<<grid, 256, 0, s>>>(homographies, height, width, out);
  else
    sample_coords_kernel<MVSB200_SAMPLER_LEGACY><<<grid, 256, 0, s>>>(homographies, height, width, out);
  MVS_LAUNCH_CHECK("sample_coords_kernel");
  return MVSB200_OK;
}
