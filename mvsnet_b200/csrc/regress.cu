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

// ---- D-slab mode: the softmax over depth split over ranks -------------------------------------------------------
// Each rank reduces its own planes to (m, s, w) = (max of -F, sum of exp(-F - m), sum of d_i exp(-F - m)) per pixel;
// the partials of all ranks are combined as M = max m_r, S = sum s_r e^(m_r - M), W = sum w_r e^(m_r - M),
// depth = W / S (model.py:472-495 up to the association of the sums); the probability map adds, on each rank, the
// buckets of get_probability_map_slice (model.py:113-140) that fall into its slab, and the host sums the ranks.
__global__ void regress_partial_kernel(const float* __restrict__ filtered, int dl, int d0, int D, int npix,
                                       float depth_start, float depth_end, int inverse_depth,
                                       float* __restrict__ partial) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  float m = -__ldg(filtered + pix);
  for (int d = 1; d < dl; ++d) m = fmaxf(m, -__ldg(filtered + (size_t)d * npix + pix));
  float s = 0.0f, w = 0.0f;
  for (int d = 0; d < dl; ++d) {
    const float e = expf(sub_(-__ldg(filtered + (size_t)d * npix + pix), m));
    s = add_(s, e);
    w = add_(w, mul_(depth_sample(d0 + d, D, depth_start, depth_end, inverse_depth), e));
  }
  partial[pix] = m;
  partial[npix + pix] = s;
  partial[2 * (size_t)npix + pix] = w;
}

__global__ void regress_combine_kernel(const float* __restrict__ partials, int slabs, const float* __restrict__ filtered,
                                       int dl, int d0, int D, int npix, float depth_start, float depth_interval,
                                       float depth_end, int inverse_depth, int num_buckets,
                                       float* __restrict__ depth_map, float* __restrict__ prob_partial) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  float M = partials[pix];
  for (int r = 1; r < slabs; ++r) M = fmaxf(M, partials[(size_t)r * 3 * npix + pix]);
  float S = 0.0f, W = 0.0f;
  for (int r = 0; r < slabs; ++r) {
    const float* pr = partials + (size_t)r * 3 * npix;
    const float sc = expf(sub_(pr[pix], M));
    S = add_(S, mul_(pr[npix + pix], sc));
    W = add_(W, mul_(pr[2 * (size_t)npix + pix], sc));
  }
  const float depth = div_(W, S);
  depth_map[pix] = depth;
  int b[4];
  prob_buckets(depth, D, depth_start, depth_interval, depth_end, inverse_depth, b[0], b[2], b[1], b[3]);   // l0 r0 | l1 r1
  float prob = 0.0f;
  const int nb = num_buckets == 4 ? 4 : 2;
  for (int k = 0; k < nb; ++k) {
    const int i = b[k] - d0;
    if (i >= 0 && i < dl) prob = add_(prob, div_(expf(sub_(-__ldg(filtered + (size_t)i * npix + pix), M)), S));
  }
  prob_partial[pix] = prob;
}

int launch_regress_partial(const float* filtered, int dl, int d0, int D, int npix, float depth_start,
                           float depth_interval, int inverse_depth, float* partial, cudaStream_t s);
int launch_regress_combine(const float* partials, int slabs, const float* filtered, int dl, int d0, int D, int npix,
                           float depth_start, float depth_interval, int inverse_depth, int num_buckets, float* depth_map,
                           float* prob_partial, cudaStream_t s);

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
    static std::atomic<uint64_t> attr_set{0};      // per device of this process
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr_set.load(std::memory_order_acquire) >> (dev & 63) & 1u)) {
      MVS_CUDA(cudaFuncSetAttribute(depth_regress_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    200 * 1024));
      attr_set.fetch_or(1ull << (dev & 63), std::memory_order_release);
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

int launch_regress_partial(const float* filtered, int dl, int d0, int D, int npix, float depth_start,
                           float depth_interval, int inverse_depth, float* partial, cudaStream_t s) {
  MVS_CHECK_ARG(filtered && partial && dl >= 1 && d0 >= 0 && d0 + dl <= D && npix >= 1, "regress_partial: bad arguments");
  regress_partial_kernel<<<ceil_div(npix, 128), 128, 0, s>>>(filtered, dl, d0, D, npix, depth_start,
                                                             host_depth_end(D, depth_start, depth_interval),
                                                             inverse_depth, partial);
  MVS_LAUNCH_CHECK("regress_partial_kernel");
  return MVSB200_OK;
}

int launch_regress_combine(const float* partials, int slabs, const float* filtered, int dl, int d0, int D, int npix,
                           float depth_start, float depth_interval, int inverse_depth, int num_buckets, float* depth_map,
                           float* prob_partial, cudaStream_t s) {
  MVS_CHECK_ARG(partials && filtered && depth_map && prob_partial && slabs >= 1, "regress_combine: bad arguments");
  MVS_CHECK_ARG(num_buckets == 2 || num_buckets == 4, "regress_combine: num_buckets must be 2 or 4 (got %d)", num_buckets);
  regress_combine_kernel<<<ceil_div(npix, 128), 128, 0, s>>>(partials, slabs, filtered, dl, d0, D, npix, depth_start,
                                                             depth_interval, host_depth_end(D, depth_start, depth_interval),
                                                             inverse_depth, num_buckets, depth_map, prob_partial);
  MVS_LAUNCH_CHECK("regress_combine_kernel");
  return MVSB200_OK;
}

}  // namespace mvsb200

using namespace mvsb200;

// D-slab mode: soft-argmin partials of planes [d0, d0 + dl) of a depth_num-plane sweep: filtered [dl, npix] ->
// partial [3, npix] = (max of -F, sum of exp, depth-weighted sum)
extern "C" int mvsb200_regress_partial(const float* filtered, int dl, int d0, int depth_num, int npix, float depth_start,
                                       float depth_interval, int inverse_depth, float* partial, void* stream) {
  return launch_regress_partial(filtered, dl, d0, depth_num, npix, depth_start, depth_interval, inverse_depth, partial,
                                (cudaStream_t)stream);
}

// partials [slabs, 3, npix] of all ranks -> depth_map [npix] (identical on every rank) and this rank's share of the
// probability map (sum the shares of all ranks)
extern "C" int mvsb200_regress_combine(const float* partials, int slabs, const float* filtered, int dl, int d0,
                                       int depth_num, int npix, float depth_start, float depth_interval,
                                       int inverse_depth, int num_buckets, float* depth_map, float* prob_partial,
                                       void* stream) {
  return launch_regress_combine(partials, slabs, filtered, dl, d0, depth_num, npix, depth_start, depth_interval,
                                inverse_depth, num_buckets, depth_map, prob_partial, (cudaStream_t)stream);
}

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
