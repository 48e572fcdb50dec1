// Kernel 1: plane-sweep homographies and pixel-coordinate transform coefficients.
// Replaces get_homographies / get_homographies_inv_depth (homography_warping.py:10-106) and the
// coefficient half of tf_transform_homography (:216-250).  One thread per (source view, plane).
#include "geometry.cuh"

namespace mvsb200 {

__global__ void homographies_kernel(const float* __restrict__ cams, int n_views, int depth_num,
                                    float depth_start, float depth_step, int inverse_depth,
                                    float* __restrict__ homographies, float* __restrict__ transforms) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int total = (n_views - 1) * depth_num;
  if (idx >= total) return;
  int v = idx / depth_num, d = idx - v * depth_num;
  float left[32], right[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    left[i] = cams[i];
    right[i] = cams[(v + 1) * 32 + i];
  }
  float depth = plane_depth(d, depth_num, depth_start, depth_step, inverse_depth);
  float H[9];
  plane_homography(left, right, depth, H);
  if (homographies) {
#pragma unroll
    for (int i = 0; i < 9; ++i) homographies[idx * 9 + i] = H[i];
  }
  if (transforms) {
    float t[8];
    transform_coefs(H, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) transforms[idx * 8 + i] = t[i];
  }
}

__global__ void transform_coefs_kernel(const float* __restrict__ homographies, int count,
                                       float* __restrict__ transforms) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  float h[9], t[8];
#pragma unroll
  for (int i = 0; i < 9; ++i) h[i] = homographies[idx * 9 + i];
  transform_coefs(h, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) transforms[idx * 8 + i] = t[i];
}

int launch_homographies(const float* cams, int n_views, int depth_num, float depth_start, float depth_step,
                        int inverse_depth, float* homographies, float* transforms, cudaStream_t s) {
  MVS_CHECK_ARG(cams != nullptr, "homographies: cams is NULL");
  MVS_CHECK_ARG(n_views >= 2 && depth_num >= 1, "homographies: need n_views>=2 and depth_num>=1 (got %d, %d)",
                n_views, depth_num);
  MVS_CHECK_ARG(homographies || transforms, "homographies: both outputs are NULL");
  int total = (n_views - 1) * depth_num;
  homographies_kernel<<<ceil_div(total, 64), 64, 0, s>>>(cams, n_views, depth_num, depth_start, depth_step,
                                                         inverse_depth, homographies, transforms);
  MVS_LAUNCH_CHECK("homographies_kernel");
  return MVSB200_OK;
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_homographies(const float* cams, int n_views, int depth_num, float depth_start,
                                    float depth_step, int inverse_depth, float* homographies,
                                    float* transforms, void* stream) {
  return launch_homographies(cams, n_views, depth_num, depth_start, depth_step, inverse_depth, homographies,
                             transforms, (cudaStream_t)stream);
}

extern "C" int mvsb200_transform_coefs(const float* homographies, int count, float* transforms, void* stream) {
  MVS_CHECK_ARG(homographies && transforms && count >= 0, "transform_coefs: bad arguments");
  if (count == 0) return MVSB200_OK;
  transform_coefs_kernel<<<ceil_div(count, 128), 128, 0, (cudaStream_t)stream>>>(homographies, count, transforms);
  MVS_LAUNCH_CHECK("transform_coefs_kernel");
  return MVSB200_OK;
}
