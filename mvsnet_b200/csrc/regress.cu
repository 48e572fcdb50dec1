// Kernel 4: softmax(-F) over depth, soft-argmin depth and the 4-neighbour probability map.
// Replaces tf.nn.softmax + linspace/tile/reduce_sum (model.py:472-495) and the four gather_nd of
// get_probability_map_slice (model.py:45-144).  The [D, PIX] column tile of F is staged once in
// shared memory (coalesced 128-bit loads), so F is read from HBM exactly once; each thread then
// owns one pixel and walks its column three times (max, sum, weighted sum) from shared memory in
// the same op order as the reference: P_i = e_i / sum, depth = sum_i samples_i * P_i, i ascending.
#include "geometry.cuh"

namespace mvsb200 {

constexpr int kRegPix = 64;   // pixels per block

__device__ __forceinline__ float depth_sample(int i, int depth_num, float depth_start, float depth_end,
                                              int inverse_depth) {
  // model.py:480-490: tf.linspace(start, end, D) or 1/linspace(1/start, 1/end, D)
  if (inverse_depth) return plane_depth(i, depth_num, depth_start, depth_end, 1);
  return tf_linspace_at(depth_start, depth_end, depth_num, i);
}

// Bucket indices of get_probability_map_slice (model.py:83-120).
__device__ __forceinline__ void prob_buckets(float depth, int D, float depth_start, float depth_interval,
                                             float depth_end, int inverse_depth, int& l0, int& l1, int& r0,
                                             int& r1) {
  if (inverse_depth) {
    float inv_start = div_(1.0f, depth_start), inv_end = div_(1.0f, depth_end);
    float inv_interval = div_(sub_(inv_start, inv_end), sub_((float)D, 1.0f));
    float v = div_(sub_(div_(1.0f, depth), inv_end), inv_interval);
    l0 = D - floor_to_int(ceilf(v)) - 1;
    r0 = D - floor_to_int(v) - 1;
  } else {
    float v = div_(sub_(depth, depth_start), depth_interval);
    l0 = floor_to_int(v);
    r0 = floor_to_int(ceilf(v));
  }
  l0 = min(max(l0, 0), D - 1);
  r0 = min(max(r0, 0), D - 1);
  l1 = min(max(l0 - 1, 0), D - 1);
  r1 = min(max(r0 + 1, 0), D - 1);
}

template <bool SMEM>
__global__ void __launch_bounds__(kRegPix)
depth_regress_kernel(const float* __restrict__ filtered, int D, int npix, float depth_start, float depth_interval,
                     float depth_end, int inverse_depth, int num_buckets, float* __restrict__ depth_map,
                     float* __restrict__ prob_map, float* __restrict__ prob_volume) {
  extern __shared__ float s_col[];   // [D][kRegPix]
  const int pix0 = blockIdx.x * kRegPix;
  const int tid = threadIdx.x;
  const int pix = pix0 + tid;
  if (SMEM) {
    const int valid = min(kRegPix, npix - pix0);
    if (valid == kRegPix && (npix % 4) == 0) {
      // 16 threads cover one 64-pixel row with float4; 4 rows per pass
      const int q = tid & 15, r0 = tid >> 4;
      for (int d = r0; d < D; d += 4) {
        float4 v = __ldg(reinterpret_cast<const float4*>(filtered + (size_t)d * npix + pix0) + q);
        *reinterpret_cast<float4*>(&s_col[d * kRegPix + q * 4]) = v;
      }
    } else {
      for (int d = 0; d < D; ++d)
        if (tid < valid) s_col[d * kRegPix + tid] = __ldg(filtered + (size_t)d * npix + pix);
    }
    __syncthreads();
  }
  if (pix >= npix) return;
  auto F = [&](int d) -> float {
    return SMEM ? s_col[d * kRegPix + tid] : __ldg(filtered + (size_t)d * npix + pix);
  };
  // x = -F ; m = max x
  float m = -F(0);
  for (int d = 1; d < D; ++d) m = fmaxf(m, -F(d));
  float sum = 0.0f;
  for (int d = 0; d < D; ++d) sum = add_(sum, expf(sub_(-F(d), m)));
  float depth = 0.0f;
  for (int d = 0; d < D; ++d) {
    float p = div_(expf(sub_(-F(d), m)), sum);
    if (prob_volume) prob_volume[(size_t)d * npix + pix] = p;
    depth = add_(depth, mul_(depth_sample(d, D, depth_start, depth_end, inverse_depth), p));
  }
  depth_map[pix] = depth;
  int l0, l1, r0, r1;
  prob_buckets(depth, D, depth_start, depth_interval, depth_end, inverse_depth, l0, l1, r0, r1);
  auto P = [&](int d) -> float { return div_(expf(sub_(-F(d), m)), sum); };
  float prob = add_(P(l0), P(r0));                               // model.py:128-130
  if (num_buckets == 4) prob = add_(prob, add_(P(l1), P(r1)));   // model.py:132-140
  prob_map[pix] = prob;
}

__global__ void probability_map_kernel(const float* __restrict__ prob_volume, const float* __restrict__ depth_map,
                                       int D, int npix, float depth_start, float depth_interval, float depth_end,
                                       int inverse_depth, int num_buckets, float* __restrict__ prob_map) {
  int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  int l0, l1, r0, r1;
  prob_buckets(depth_map[pix], D, depth_start, depth_interval, depth_end, inverse_depth, l0, l1, r0, r1);
  auto P = [&](int d) -> float { return __ldg(prob_volume + (size_t)d * npix + pix); };
  float prob = add_(P(l0), P(r0));
  if (num_buckets == 4) prob = add_(prob, add_(P(l1), P(r1)));
  prob_map[pix] = prob;
}

static float host_depth_end(int depth_num, float depth_start, float depth_interval) {
  // model.py:378-379 in fp32: start + (float(D) - 1) * interval
  volatile float dm1 = (float)depth_num - 1.0f;
  volatile float prod = dm1 * depth_interval;
  volatile float end = depth_start + prod;
  return end;
}

int launch_depth_regress(const float* filtered, int depth_num, int hf, int wf, float depth_start,
                         float depth_interval, int inverse_depth, int num_buckets, float* depth_map,
                         float* prob_map, float* prob_volume, cudaStream_t s) {
  MVS_CHECK_ARG(filtered && depth_map && prob_map, "depth_regress: NULL pointer");
  MVS_CHECK_ARG(depth_num >= 1 && hf >= 1 && wf >= 1, "depth_regress: bad shape D=%d %dx%d", depth_num, hf, wf);
  MVS_CHECK_ARG(num_buckets == 2 || num_buckets == 4, "depth_regress: num_buckets must be 2 or 4 (got %d)",
                num_buckets);
  const int npix = hf * wf;
  const float depth_end = host_depth_end(depth_num, depth_start, depth_interval);
  const size_t smem = (size_t)depth_num * kRegPix * sizeof(float);
  const int blocks = ceil_div(npix, kRegPix);
  if (smem <= 200 * 1024) {
    static bool attr_set = false;
    if (!attr_set) {
      MVS_CUDA(cudaFuncSetAttribute(depth_regress_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    200 * 1024));
      attr_set = true;
    }
    depth_regress_kernel<true><<<blocks, kRegPix, smem, s>>>(filtered, depth_num, npix, depth_start, depth_interval,
                                                            depth_end, inverse_depth, num_buckets, depth_map,
                                                            prob_map, prob_volume);
  } else {
    depth_regress_kernel<false><<<blocks, kRegPix, 0, s>>>(filtered, depth_num, npix, depth_start, depth_interval,
                                                          depth_end, inverse_depth, num_buckets, depth_map,
                                                          prob_map, prob_volume);
  }
  MVS_LAUNCH_CHECK("depth_regress_kernel");
  return MVSB200_OK;
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_depth_regress(const float* filtered, int depth_num, int hf, int wf, float depth_start,
                                     float depth_interval, int inverse_depth, int num_buckets, float* depth_map,
                                     float* prob_map, float* prob_volume, void* stream) {
  return launch_depth_regress(filtered, depth_num, hf, wf, depth_start, depth_interval, inverse_depth, num_buckets,
                              depth_map, prob_map, prob_volume, (cudaStream_t)stream);
}

extern "C" int mvsb200_probability_map(const float* prob_volume, const float* depth_map, int depth_num, int height,
                                       int width, float depth_start, float depth_interval, int inverse_depth,
                                       int num_buckets, float* prob_map, void* stream) {
  MVS_CHECK_ARG(prob_volume && depth_map && prob_map, "probability_map: NULL pointer");
  MVS_CHECK_ARG(depth_num >= 1 && height >= 1 && width >= 1, "probability_map: bad shape");
  MVS_CHECK_ARG(num_buckets == 2 || num_buckets == 4, "probability_map: num_buckets must be 2 or 4 (got %d)",
                num_buckets);
  const int npix = height * width;
  const float depth_end = host_depth_end(depth_num, depth_start, depth_interval);
  probability_map_kernel<<<ceil_div(npix, 128), 128, 0, (cudaStream_t)stream>>>(
      prob_volume, depth_map, depth_num, npix, depth_start, depth_interval, depth_end, inverse_depth, num_buckets,
      prob_map);
  MVS_LAUNCH_CHECK("probability_map_kernel");
  return MVSB200_OK;
}
