// Kernel 2: fused homography warp + N-view variance cost volume.
// Replaces the D x (N-1) tf_transform_homography launches and the running-sum ops of
// model.py:423-463 (inference_mem) / model.py:315-334 (inference).  The (N-1) x D x H x W x C
// warped volume is never written: each thread keeps the running sum and squared sum of its
// voxels in registers across views and stores only the variance.
//
// Transform coefficients live in __constant__ memory (one 8-float row per (view, plane),
// written once per call by prepare_table_kernel + a device-to-device symbol copy).
#include "geometry.cuh"

namespace mvsb200 {

constexpr int kTableFloats = 14336;  // 56 KB: (N-1)*D*8 for N=8, D=256 (inference.py:29-31 defaults)
__constant__ float c_table[kTableFloats];
__device__ float g_table[kTableFloats];  // staging copy in global memory (generic kernel reads this one)

__global__ void prepare_table_kernel(const float* __restrict__ homographies, int count) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  float h[9], t[8];
#pragma unroll
  for (int i = 0; i < 9; ++i) h[i] = homographies[idx * 9 + i];
  transform_coefs(h, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) g_table[idx * 8 + i] = t[i];
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float variance(float S, float Q, float n_f, float nn_f, int order) {
  if (order == MVSB200_ORDER_MEM) {
    float A = (S * S) / nn_f;       // model.py:458
    return Q / n_f - A;             // model.py:460
  }
  float mean = S / n_f, mean2 = Q / n_f;   // model.py:330-331
  return mean2 - mean * mean;              // model.py:332
}

__device__ __forceinline__ void store_cost4(void* out, size_t elem, float4 c, bool bf16) {
  if (bf16) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(c.x, c.y), hi = __floats2bfloat162_rn(c.z, c.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + elem) = pk;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + elem) = c;
  }
}

// ------------------------------------------------------------------------------------------------
// Fast path, C = 32: block = 32 reference pixels x 8 channel groups (float4), DC = 8 planes per
// thread.  The 8 lanes of a pixel each build the complete bilinear footprint of a different plane
// (sample position with 2 IEEE divisions, the four weights, the clamped tap coordinates) and
// broadcast it with shuffles, so that work is done once per voxel and view instead of once per
// channel group.  A tap outside the image gets weight 0 and a clamped (always mapped) address,
// which is arithmetically identical to the reference's zero fill (0 * finite feature) and keeps
// the loads unpredicated.  Taps are 128-bit loads through L1 (the features are L2-resident:
// 5 x 216 x 288 x 32 fp32 = 40 MB at config 2); the running sum and squared sum of the 8 planes
// stay in registers across views.
// ------------------------------------------------------------------------------------------------
constexpr int kDC = 8;
constexpr int kMaxSrcViews = 7;

// OUT: 0 = fp32 [D,Hf,Wf,32], 1 = bf16 [D,Hf,Wf,32], 2 = bf16 in the regularizer's planar layouts (conv3d_tc.cu):
// out = CP8 [D][4][Hf][Wf][8] and out2 = PS8 [D][4][4][Hs][Ws][8] (either may be NULL)
template <int OUT, int TX, int TY>
__global__ void __launch_bounds__(256, 2)
cost_volume_c32_kernel(const float* __restrict__ feats, int n_views, int D, int Hf, int Wf, int order,
                       void* __restrict__ out, void* __restrict__ out2) {
  static_assert(TX * TY == 32, "tile must hold 32 pixels");
  __shared__ float s_coef[kMaxSrcViews * kDC * 8];
  const int tid = threadIdx.x;
  const int g = tid & 7;            // channel group (4 channels) and plane slot for the footprint
  const int p = tid >> 3;           // pixel within the tile
  const int x = blockIdx.x * TX + (p % TX);
  const int y = blockIdx.y * TY + (p / TX);
  const int d0 = blockIdx.z * kDC;
  const int n_src = n_views - 1;
  for (int i = tid; i < n_src * kDC * 8; i += 256) {
    int v = i / (kDC * 8), r = i - v * (kDC * 8);
    int d = min(d0 + (r >> 3), D - 1);
    s_coef[i] = c_table[(v * D + d) * 8 + (r & 7)];
  }
  __syncthreads();
  const bool active = (x < Wf) && (y < Hf);
  const int xc = min(x, Wf - 1), yc = min(y, Hf - 1);
  const size_t plane = (size_t)Hf * Wf * 32;
  const float4 r = ldg4(feats + ((size_t)yc * Wf + xc) * 32 + g * 4);
  float4 S[kDC], Q[kDC];
#pragma unroll
  for (int dd = 0; dd < kDC; ++dd) {
    S[dd] = r;
    Q[dd] = make_float4(r.x * r.x, r.y * r.y, r.z * r.z, r.w * r.w);
  }
  const unsigned lane_base = (threadIdx.x & 31) & ~7u;
  for (int v = 0; v < n_src; ++v) {
    const float* img = feats + (size_t)(v + 1) * plane + g * 4;
    // this lane's plane: position, weights (0 where the tap is outside), clamped tap rows / columns
    float ix, iy;
    transform_coords(&s_coef[(v * kDC + g) * 8], (float)xc, (float)yc, ix, iy);
    const Footprint f = make_footprint(ix, iy, Wf, Hf);
    const float l_wxl = f.vx0 ? f.wxl : 0.0f, l_wxr = f.vx1 ? f.wxr : 0.0f;
    const float l_wyl = f.vy0 ? f.wyl : 0.0f, l_wyr = f.vy1 ? f.wyr : 0.0f;
    const int cx0 = min(max(f.x0, 0), Wf - 1), cx1 = min(max(f.x0 + 1, 0), Wf - 1);
    const int cy0 = min(max(f.y0, 0), Hf - 1), cy1 = min(max(f.y0 + 1, 0), Hf - 1);
    const int l_col = cx0 | (cx1 << 16), l_row = cy0 | (cy1 << 16);
#pragma unroll
    for (int dd = 0; dd < kDC; ++dd) {
      const unsigned src = lane_base | dd;
      const float wxl = __shfl_sync(0xffffffffu, l_wxl, src), wxr = __shfl_sync(0xffffffffu, l_wxr, src);
      const float wyl = __shfl_sync(0xffffffffu, l_wyl, src), wyr = __shfl_sync(0xffffffffu, l_wyr, src);
      const int col = __shfl_sync(0xffffffffu, l_col, src), row = __shfl_sync(0xffffffffu, l_row, src);
      const int x0 = col & 0xffff, x1 = col >> 16;
      const int r0 = (row & 0xffff) * Wf, r1 = (row >> 16) * Wf;
      const float4 p00 = ldg4(img + (size_t)(r0 + x0) * 32), p01 = ldg4(img + (size_t)(r0 + x1) * 32);
      const float4 p10 = ldg4(img + (size_t)(r1 + x0) * 32), p11 = ldg4(img + (size_t)(r1 + x1) * 32);
      float4 w;
      w.x = wyl * (wxl * p00.x + wxr * p01.x) + wyr * (wxl * p10.x + wxr * p11.x);
      w.y = wyl * (wxl * p00.y + wxr * p01.y) + wyr * (wxl * p10.y + wxr * p11.y);
      w.z = wyl * (wxl * p00.z + wxr * p01.z) + wyr * (wxl * p10.z + wxr * p11.z);
      w.w = wyl * (wxl * p00.w + wxr * p01.w) + wyr * (wxl * p10.w + wxr * p11.w);
      S[dd].x += w.x; S[dd].y += w.y; S[dd].z += w.z; S[dd].w += w.w;
      Q[dd].x += w.x * w.x; Q[dd].y += w.y * w.y; Q[dd].z += w.z * w.z; Q[dd].w += w.w * w.w;
    }
  }
  if (!active) return;
  // variance with reciprocal multiplies (<= 1 ulp from the reference's divisions, model.py:458-461 / :330-332)
  const float inv_n = 1.0f / (float)n_views, inv_nn = 1.0f / (float)(n_views * n_views);
#pragma unroll
  for (int dd = 0; dd < kDC; ++dd) {
    const int d = d0 + dd;
    if (d < D) {
      float4 c;
      if (order == MVSB200_ORDER_MEM) {
        c.x = Q[dd].x * inv_n - (S[dd].x * S[dd].x) * inv_nn;
        c.y = Q[dd].y * inv_n - (S[dd].y * S[dd].y) * inv_nn;
        c.z = Q[dd].z * inv_n - (S[dd].z * S[dd].z) * inv_nn;
        c.w = Q[dd].w * inv_n - (S[dd].w * S[dd].w) * inv_nn;
      } else {
        const float mx = S[dd].x * inv_n, my = S[dd].y * inv_n, mz = S[dd].z * inv_n, mw = S[dd].w * inv_n;
        c.x = Q[dd].x * inv_n - mx * mx;
        c.y = Q[dd].y * inv_n - my * my;
        c.z = Q[dd].z * inv_n - mz * mz;
        c.w = Q[dd].w * inv_n - mw * mw;
      }
      if (OUT < 2) {
        store_cost4(out, (((size_t)d * Hf + y) * Wf + x) * 32 + g * 4, c, OUT == 1);
      } else {
        // 8-byte half of the 16-byte cell (chunk g>>1) of this voxel
        const size_t zc = (size_t)d * 4 + (g >> 1);
        if (out) store_cost4(out, ((zc * Hf + y) * Wf + x) * 8 + (g & 1) * 4, c, true);
        if (out2) {
          const int Hs = (Hf + 1) >> 1, Ws = (Wf + 1) >> 1;
          store_cost4(out2, (((zc * 4 + (y & 1) * 2 + (x & 1)) * Hs + (y >> 1)) * Ws + (x >> 1)) * 8 + (g & 1) * 4, c, true);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Generic path: any channel count, both samplers.  One thread per (plane, pixel, channel group).
// ------------------------------------------------------------------------------------------------
template <int VEC, int SAMPLER, bool BF16OUT>
__global__ void __launch_bounds__(256)
cost_volume_generic_kernel(const float* __restrict__ feats, const float* __restrict__ homographies, int n_views,
                           int D, int Hf, int Wf, int C, int order, void* __restrict__ out) {
  const int groups = C / VEC;
  const size_t total = (size_t)D * Hf * Wf * groups;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  size_t vox = idx / groups;
  const int x = (int)(vox % Wf);
  const int y = (int)((vox / Wf) % Hf);
  const int d = (int)(vox / ((size_t)Wf * Hf));
  const size_t plane = (size_t)Hf * Wf * C;
  const int c0 = g * VEC;
  float S[VEC], Q[VEC];
  const float* ref = feats + ((size_t)y * Wf + x) * C + c0;
#pragma unroll
  for (int k = 0; k < VEC; ++k) { float r = __ldg(ref + k); S[k] = r; Q[k] = r * r; }
  for (int v = 0; v < n_views - 1; ++v) {
    const float* img = feats + (size_t)(v + 1) * plane;
    float w[VEC];
    if (SAMPLER == MVSB200_SAMPLER_TRANSFORM) {
      float ix, iy;
      transform_coords(&g_table[(v * D + d) * 8], (float)x, (float)y, ix, iy);
      Footprint f = make_footprint(ix, iy, Wf, Hf);
      const float* base = img + ((int64_t)f.y0 * Wf + f.x0) * C + c0;
      const int64_t row = (int64_t)Wf * C;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        float p00 = (f.vy0 && f.vx0) ? __ldg(base + k) : 0.f;
        float p01 = (f.vy0 && f.vx1) ? __ldg(base + C + k) : 0.f;
        float p10 = (f.vy1 && f.vx0) ? __ldg(base + row + k) : 0.f;
        float p11 = (f.vy1 && f.vx1) ? __ldg(base + row + C + k) : 0.f;
        w[k] = f.wyl * (f.wxl * p00 + f.wxr * p01) + f.wyr * (f.wxl * p10 + f.wxr * p11);
      }
    } else {
      // legacy sampler: every op rounded separately (Appendix A.3 cancellations)
      float h[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) h[i] = __ldg(homographies + ((size_t)v * D + d) * 9 + i);
      float xw, yw;
      legacy_coords(h, x, y, Wf, Hf, xw, yw);
      float xs = sub_(xw, 0.5f), ys = sub_(yw, 0.5f);
      int x0 = floor_to_int(xs), y0 = floor_to_int(ys);
      int x1 = x0 + 1, y1 = y0 + 1;
      x0 = min(max(x0, 0), Wf - 1); x1 = min(max(x1, 0), Wf - 1);
      y0 = min(max(y0, 0), Hf - 1); y1 = min(max(y1, 0), Hf - 1);
      float wa = mul_(sub_((float)y1, ys), sub_((float)x1, xs));
      float wb = mul_(sub_((float)y1, ys), sub_(xs, (float)x0));
      float wc = mul_(sub_(ys, (float)y0), sub_((float)x1, xs));
      float wd = mul_(sub_(ys, (float)y0), sub_(xs, (float)x0));
      const float* pa = img + ((size_t)y0 * Wf + x0) * C + c0;
      const float* pb = img + ((size_t)y0 * Wf + x1) * C + c0;
      const float* pc = img + ((size_t)y1 * Wf + x0) * C + c0;
      const float* pd = img + ((size_t)y1 * Wf + x1) * C + c0;
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        w[k] = add_(add_(add_(mul_(wa, __ldg(pa + k)), mul_(wb, __ldg(pb + k))), mul_(wc, __ldg(pc + k))),
                    mul_(wd, __ldg(pd + k)));
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) { S[k] += w[k]; Q[k] += w[k] * w[k]; }
  }
  const float n_f = (float)n_views, nn_f = (float)(n_views * n_views);
  const size_t o = vox * C + c0;
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    float c = variance(S[k], Q[k], n_f, nn_f, order);
    if (BF16OUT) reinterpret_cast<__nv_bfloat16*>(out)[o + k] = __float2bfloat16_rn(c);
    else reinterpret_cast<float*>(out)[o + k] = c;
  }
}

// planar_ps8 != NULL or planar != 0: write the bf16 planar layouts (out = CP8, planar_ps8 = PS8); fast path only
static int launch_cost_volume_any(const float* feats, const float* homographies, int n_views, int depth_num, int hf,
                                  int wf, int channels, int order, int sampler, int out_dtype, void* out, int variant,
                                  int planar, void* planar_ps8, cudaStream_t s) {
  MVS_CHECK_ARG(feats && homographies && (out || planar_ps8), "cost_volume: NULL pointer");
  MVS_CHECK_ARG(n_views >= 2 && depth_num >= 1 && hf >= 1 && wf >= 1 && channels >= 1,
                "cost_volume: bad shape N=%d D=%d %dx%dx%d", n_views, depth_num, hf, wf, channels);
  MVS_CHECK_ARG(order == MVSB200_ORDER_MEM || order == MVSB200_ORDER_TRAIN, "cost_volume: bad order %d", order);
  MVS_CHECK_ARG(sampler == MVSB200_SAMPLER_TRANSFORM || sampler == MVSB200_SAMPLER_LEGACY,
                "cost_volume: bad sampler %d", sampler);
  MVS_CHECK_ARG(out_dtype == MVSB200_F32 || out_dtype == MVSB200_BF16, "cost_volume: bad out_dtype %d", out_dtype);
  const int rows = (n_views - 1) * depth_num;
  const bool bf16 = out_dtype == MVSB200_BF16;
  if (sampler == MVSB200_SAMPLER_TRANSFORM) {
    if (rows * 8 > kTableFloats) {
      set_error("cost_volume: (n_views-1)*depth_num = %d exceeds the %d-row coefficient table", rows,
                kTableFloats / 8);
      return MVSB200_ERR_UNSUPPORTED;
    }
    prepare_table_kernel<<<ceil_div(rows, 128), 128, 0, s>>>(homographies, rows);
    MVS_LAUNCH_CHECK("prepare_table_kernel");
  }
  // variant: 0 auto, 1 generic, 2 fast path with a 32x1 pixel tile, 3 fast path with a 16x2 tile
  bool fast_ok = sampler == MVSB200_SAMPLER_TRANSFORM && channels == 32 && n_views - 1 <= kMaxSrcViews &&
                 hf < 65536 && wf < 32768;
  if (variant == 0) variant = fast_ok ? 3 : 1;
  if (planar && (!fast_ok || variant < 2)) {
    set_error("cost_volume: the planar output needs sampler=transform, C=32, n_views<=8");
    return MVSB200_ERR_UNSUPPORTED;
  }
  MVS_CHECK_ARG(variant >= 1 && variant <= 3, "cost_volume: bad variant %d", variant);
  if (variant >= 2 && !fast_ok) {
    set_error("cost_volume: variant %d needs sampler=transform, C=32, n_views<=8", variant);
    return MVSB200_ERR_UNSUPPORTED;
  }
  if (variant >= 2) {
    void* g_ptr = nullptr;
    MVS_CUDA(cudaGetSymbolAddress(&g_ptr, g_table));
    MVS_CUDA(cudaMemcpyToSymbolAsync(c_table, g_ptr, (size_t)rows * 8 * sizeof(float), 0,
                                     cudaMemcpyDeviceToDevice, s));
    const bool wide = variant == 2;
    const int tx = wide ? 32 : 16, ty = wide ? 1 : 2;
    dim3 grid(ceil_div(wf, tx), ceil_div(hf, ty), ceil_div(depth_num, kDC));
    MVS_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "cost_volume: grid too large");
#define CV_FAST(O, TX_, TY_) \
  cost_volume_c32_kernel<O, TX_, TY_><<<grid, 256, 0, s>>>(feats, n_views, depth_num, hf, wf, order, out, planar_ps8)
    if (planar && ((hf | wf) & 1) && planar_ps8)
      MVS_CUDA(cudaMemsetAsync(planar_ps8, 0, (size_t)depth_num * 16 * ((hf + 1) / 2) * ((wf + 1) / 2) * 16, s));
    if (planar) { if (wide) CV_FAST(2, 32, 1); else CV_FAST(2, 16, 2); }
    else if (wide) { if (bf16) CV_FAST(1, 32, 1); else CV_FAST(0, 32, 1); }
    else           { if (bf16) CV_FAST(1, 16, 2); else CV_FAST(0, 16, 2); }
#undef CV_FAST
    MVS_LAUNCH_CHECK("cost_volume_c32_kernel");
    return MVSB200_OK;
  }
  const int vec = channels % 4 == 0 ? 4 : 1;
  size_t total = (size_t)depth_num * hf * wf * (channels / vec);
  MVS_CHECK_ARG((total + 255) / 256 <= 0x7fffffffu, "cost_volume: problem too large for one launch");
  unsigned blocks = (unsigned)((total + 255) / 256);
#define CV_GEN(V, SM, BF)                                                                                  \
  cost_volume_generic_kernel<V, SM, BF><<<blocks, 256, 0, s>>>(feats, homographies, n_views, depth_num, hf, \
                                                               wf, channels, order, out)
  if (sampler == MVSB200_SAMPLER_TRANSFORM) {
    if (vec == 4) { if (bf16) CV_GEN(4, MVSB200_SAMPLER_TRANSFORM, true); else CV_GEN(4, MVSB200_SAMPLER_TRANSFORM, false); }
    else          { if (bf16) CV_GEN(1, MVSB200_SAMPLER_TRANSFORM, true); else CV_GEN(1, MVSB200_SAMPLER_TRANSFORM, false); }
  } else {
    if (vec == 4) { if (bf16) CV_GEN(4, MVSB200_SAMPLER_LEGACY, true); else CV_GEN(4, MVSB200_SAMPLER_LEGACY, false); }
    else          { if (bf16) CV_GEN(1, MVSB200_SAMPLER_LEGACY, true); else CV_GEN(1, MVSB200_SAMPLER_LEGACY, false); }
  }
#undef CV_GEN
  MVS_LAUNCH_CHECK("cost_volume_generic_kernel");
  return MVSB200_OK;
}

int launch_cost_volume(const float* feats, const float* homographies, int n_views, int depth_num, int hf,
                       int wf, int channels, int order, int sampler, int out_dtype, void* out, int variant,
                       cudaStream_t s) {
  return launch_cost_volume_any(feats, homographies, n_views, depth_num, hf, wf, channels, order, sampler, out_dtype,
                                out, variant, 0, nullptr, s);
}

bool cost_volume_planar_ok(int n_views, int hf, int wf, int channels, int sampler) {
  return sampler == MVSB200_SAMPLER_TRANSFORM && channels == 32 && n_views - 1 <= kMaxSrcViews && hf < 65536 &&
         wf < 32768;
}

int launch_cost_volume_planar(const float* feats, const float* homographies, int n_views, int depth_num, int hf,
                              int wf, int channels, int order, int sampler, void* cp8, void* ps8, cudaStream_t s) {
  return launch_cost_volume_any(feats, homographies, n_views, depth_num, hf, wf, channels, order, sampler,
                                MVSB200_BF16, cp8, 0, 1, ps8, s);
}

}  // namespace mvsb200

extern "C" int mvsb200_cost_volume(const float* feats, const float* homographies, int n_views, int depth_num,
                                   int hf, int wf, int channels, int order, int sampler, int out_dtype, void* out,
                                   int variant, void* stream) {
  return mvsb200::launch_cost_volume(feats, homographies, n_views, depth_num, hf, wf, channels, order, sampler,
                                     out_dtype, out, variant, (cudaStream_t)stream);
}
