// Image feature tower UNetDS2GN (cnn_wrapper/mvsnetworks.py:53-115, called per view at model.py:392-406), fp32 parity
// mode: direct 2-D convolution / stride-2 transposed convolution on CUDA cores with TF SAME padding and no bias
// (network.py:172-215, :295-331), channel concatenation folded into the read (two sources, network.py:452-454), the
// group statistics of the result reduced in the epilogue, and group normalisation (+ ReLU) applied in place in the
// reference's op order (network.py:237-276, :349-409: groups of 8 channels, biased variance, eps 1e-5; conv_gn has a
// ReLU, deconv_gn has none).  This is the step before the hot path; the tcgen05 version
// is round-2 work, this file is the parity-first CUDA implementation (8.8 ms for 5 views at 1152x864).
#include <stdlib.h>

#include "common.cuh"
#include "feature2d_plan.h"

namespace mvsb200 {

namespace {

constexpr int kCoT = 8;            // output channels per block = one normalisation group (group_channel = 8)
constexpr int kPxT = 4;            // output rows per thread (8 rows: 198 registers and no faster, measured)
constexpr int kThreads2d = 128;

// One thread = kPxT output pixels (the same column of kPxT consecutive rows) x kCoT output channels; a warp covers
// 32 consecutive columns, so its loads and stores run along x.  The weights of the block's channel chunk sit in
// shared memory as [tap][ci][kCoT].  Sources a and b are the two halves of a channel concatenation (cb = 0: single
// source).  blockIdx = (warp tiles of 32 columns x kPxT rows, output-channel chunk, view).
template <bool TRANSPOSED>
__global__ void __launch_bounds__(kThreads2d)
conv2d_direct_kernel(const float* __restrict__ xa, int ca, const float* __restrict__ xb, int cb,
                     const float* __restrict__ kernel_tf, int H, int W, int Cout, int K, int stride, int Ho, int Wo,
                     int pad_h, int pad_w, float* __restrict__ y, double* __restrict__ stats) {
  extern __shared__ float s_w[];                       // [K*K][Cin][kCoT]
  __shared__ float s_red[2][kThreads2d / 32];
  const int Cin = ca + cb;
  const int co0 = blockIdx.y * kCoT, n = blockIdx.z;
  for (int i = threadIdx.x; i < K * K * Cin * kCoT; i += blockDim.x) {
    const int co = i % kCoT, ci = (i / kCoT) % Cin, tap = i / (kCoT * Cin);
    s_w[i] = TRANSPOSED ? kernel_tf[((size_t)tap * Cout + (co0 + co)) * Cin + ci]       // [kh,kw,Cout,Cin]
                        : kernel_tf[((size_t)tap * Cin + ci) * Cout + (co0 + co)];      // [kh,kw,Cin,Cout]
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int xt = (Wo + 31) / 32;                                   // warp tiles per row band
  const int q = blockIdx.x * (kThreads2d / 32) + warp;             // warp tile index
  const int oy0 = (q / xt) * kPxT, ox = (q % xt) * 32 + lane;
  const bool live = oy0 < Ho && ox < Wo;
  float acc[kPxT][kCoT];
#pragma unroll
  for (int j = 0; j < kPxT; ++j)
#pragma unroll
    for (int k = 0; k < kCoT; ++k) acc[j][k] = 0.0f;
  if (live) {
    const float* xa_n = xa + (size_t)n * H * W * ca;
    const float* xb_n = cb ? xb + (size_t)n * H * W * cb : nullptr;
    for (int kw = 0; kw < K; ++kw) {
      int ix;
      if (TRANSPOSED) { const int t = ox - kw; ix = (t < 0 || (t & 1)) ? -1 : (t >> 1); }
      else ix = ox * stride + kw - pad_w;
      if (ix < 0 || ix >= W) continue;
      for (int kh = 0; kh < K; ++kh) {
        int iy[kPxT];
        bool any = false;
#pragma unroll
        for (int j = 0; j < kPxT; ++j) {
          int v;
          if (TRANSPOSED) { const int t = oy0 + j - kh; v = (t < 0 || (t & 1)) ? -1 : (t >> 1); }
          else v = (oy0 + j) * stride + kh - pad_h;
          if (v >= H || oy0 + j >= Ho) v = -1;
          iy[j] = v;
          any |= v >= 0;
        }
        if (!any) continue;
        const float* wt = s_w + (size_t)(kh * K + kw) * Cin * kCoT;
        // source a, then source b (concat order), 4 input channels at a time when the counts allow it
        for (int src = 0; src < 2; ++src) {
          const float* xs = src ? xb_n : xa_n;
          const int cs = src ? cb : ca;
          if (cs == 0) continue;
          const float* ws = wt + (src ? ca * kCoT : 0);
          const float* col = xs + (size_t)ix * cs;
          const size_t rowp = (size_t)W * cs;
          if ((cs & 3) == 0) {
            for (int ci = 0; ci < cs; ci += 4) {
              float4 v[kPxT];
#pragma unroll
              for (int j = 0; j < kPxT; ++j)
                v[j] = iy[j] >= 0 ? __ldg(reinterpret_cast<const float4*>(col + (size_t)iy[j] * rowp + ci))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const float4 w0 = *reinterpret_cast<const float4*>(ws + (ci + c) * kCoT);
                const float4 w1 = *reinterpret_cast<const float4*>(ws + (ci + c) * kCoT + 4);
#pragma unroll
                for (int j = 0; j < kPxT; ++j) {
                  const float a = c == 0 ? v[j].x : (c == 1 ? v[j].y : (c == 2 ? v[j].z : v[j].w));
                  acc[j][0] = fmaf(a, w0.x, acc[j][0]); acc[j][1] = fmaf(a, w0.y, acc[j][1]);
                  acc[j][2] = fmaf(a, w0.z, acc[j][2]); acc[j][3] = fmaf(a, w0.w, acc[j][3]);
                  acc[j][4] = fmaf(a, w1.x, acc[j][4]); acc[j][5] = fmaf(a, w1.y, acc[j][5]);
                  acc[j][6] = fmaf(a, w1.z, acc[j][6]); acc[j][7] = fmaf(a, w1.w, acc[j][7]);
                }
              }
            }
          } else {
            for (int ci = 0; ci < cs; ++ci) {
              const float4 w0 = *reinterpret_cast<const float4*>(ws + ci * kCoT);
              const float4 w1 = *reinterpret_cast<const float4*>(ws + ci * kCoT + 4);
#pragma unroll
              for (int j = 0; j < kPxT; ++j) {
                const float a = iy[j] >= 0 ? __ldg(col + (size_t)iy[j] * rowp + ci) : 0.0f;
                acc[j][0] = fmaf(a, w0.x, acc[j][0]); acc[j][1] = fmaf(a, w0.y, acc[j][1]);
                acc[j][2] = fmaf(a, w0.z, acc[j][2]); acc[j][3] = fmaf(a, w0.w, acc[j][3]);
                acc[j][4] = fmaf(a, w1.x, acc[j][4]); acc[j][5] = fmaf(a, w1.y, acc[j][5]);
                acc[j][6] = fmaf(a, w1.z, acc[j][6]); acc[j][7] = fmaf(a, w1.w, acc[j][7]);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kPxT; ++j)
      if (oy0 + j < Ho) {
        float4* o = reinterpret_cast<float4*>(y + (((size_t)n * Ho + oy0 + j) * Wo + ox) * Cout + co0);
        o[0] = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
        o[1] = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
      }
  }
  if (stats == nullptr) return;
  // statistics of the (view, group) this block belongs to: warp shuffle -> shared -> one fp64 atomic pair per block
  float s = 0.0f, sq = 0.0f;
  if (live) {
#pragma unroll
    for (int j = 0; j < kPxT; ++j)
      if (oy0 + j < Ho) {
#pragma unroll
        for (int k = 0; k < kCoT; ++k) { s += acc[j][k]; sq = fmaf(acc[j][k], acc[j][k], sq); }
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  if (lane == 0) { s_red[0][warp] = s; s_red[1][warp] = sq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ds = 0.0, dq = 0.0;
    for (int w = 0; w < kThreads2d / 32; ++w) { ds += (double)s_red[0][w]; dq += (double)s_red[1][w]; }
    double* st = stats + ((size_t)n * (Cout / kCoT) + blockIdx.y) * 2;
    atomicAdd(st, ds);
    atomicAdd(st + 1, dq);
  }
}

// y [N][HW][C] in place: ((y - mean) / sqrt(var + eps)) * gamma + beta, then ReLU (network.py:252-275), one op per
// rounding like the reference's separate TF ops.  stats [N][C/8][2] = (sum, sum of squares) per (view, group).
__global__ void group_norm_kernel(float* __restrict__ y, const double* __restrict__ stats, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, int hw, int C, float eps, int relu, size_t total4) {
  const double cnt = (double)hw * kCoT;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i * 4;
    const int c = (int)(e % C);
    const int n = (int)(e / ((size_t)hw * C));
    const double* st = stats + ((size_t)n * (C / kCoT) + c / kCoT) * 2;
    const double mean64 = st[0] / cnt;
    double var64 = st[1] / cnt - mean64 * mean64;          // biased variance (tf.nn.moments), fp64 accumulation
    if (var64 < 0.0) var64 = 0.0;
    const float mean = (float)mean64, sd = __fsqrt_rn(__fadd_rn((float)var64, eps));
    float4 v = *reinterpret_cast<float4*>(y + e);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c)), b = __ldg(reinterpret_cast<const float4*>(beta + c));
    v.x = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(v.x, mean), sd), g.x), b.x);
    v.y = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(v.y, mean), sd), g.y), b.y);
    v.z = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(v.z, mean), sd), g.z), b.z);
    v.w = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(v.w, mean), sd), g.w), b.w);
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    *reinterpret_cast<float4*>(y + e) = v;
  }
}

// Shared-memory tiled variant for the ordinary convolutions whose sources have multiples of 8 channels (28 of the 32
// layers): a block computes 32 columns x 4 R rows x kCoT output channels (R = 4 rows per thread; R = 8 at stride 1
// needs 126 registers and was slower: 10.3 vs 8.8 ms for the tower).  Eight input channels at a time, the haloed
// input tile is staged in shared memory (coalesced loads; split in two 4-channel planes so that a warp's 128-bit
// reads run along x) next to that chunk's weights; a thread keeps the NR = (R - 1) S + K input rows of its column in
// registers and reuses them for its R output rows and the K vertical taps.  SAME padding = zeros staged for the
// pixels outside the image.
// R = output rows per thread (the block covers 4 R rows): each weight read from shared memory feeds R pixels
template <int K, int S, int R>
__global__ void __launch_bounds__(kThreads2d)
conv2d_tile_kernel(const float* __restrict__ xa, int ca, const float* __restrict__ xb, int cb,
                   const float* __restrict__ kernel_tf, int H, int W, int Cout, int Ho, int Wo, int pad_h, int pad_w,
                   float* __restrict__ y, double* __restrict__ stats) {
  constexpr int kTileRows = 4 * R;
  constexpr int IR = (kTileRows - 1) * S + K, IC = 31 * S + K, NR = (R - 1) * S + K;
  extern __shared__ float4 s_tile[];                  // [2 halves][IR][IC] float4, then weights [K*K][8 ci][8 co]
  float* s_wt = reinterpret_cast<float*>(s_tile + 2 * IR * IC);
  __shared__ float s_red[2][kThreads2d / 32];
  const int Cin = ca + cb;
  const int co0 = blockIdx.y * kCoT, n = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int xt = (Wo + 31) / 32;
  const int oy_t = (blockIdx.x / xt) * kTileRows, ox_t = (blockIdx.x % xt) * 32;
  const int oy0 = oy_t + warp * R, ox = ox_t + lane;
  const int gy0 = oy_t * S - pad_h, gx0 = ox_t * S - pad_w;        // image position of tile cell (0, 0)
  float acc[R][kCoT];
#pragma unroll
  for (int j = 0; j < R; ++j)
#pragma unroll
    for (int k = 0; k < kCoT; ++k) acc[j][k] = 0.0f;
  for (int c8 = 0; c8 < Cin; c8 += 8) {
    const bool from_b = c8 >= ca;
    const float* xs = from_b ? xb + (size_t)n * H * W * cb : xa + (size_t)n * H * W * ca;
    const int cs = from_b ? cb : ca, cl = from_b ? c8 - ca : c8;   // channel offset inside the source
    __syncthreads();                                               // the previous chunk has been consumed
    for (int i = threadIdx.x; i < 2 * IR * IC; i += kThreads2d) {
      const int half = i & 1, cell = i >> 1;
      const int r = cell / IC, c = cell - r * IC;
      const int gy = gy0 + r, gx = gx0 + c;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = __ldg(reinterpret_cast<const float4*>(xs + ((size_t)gy * W + gx) * cs + cl + half * 4));
      s_tile[(half * IR + r) * IC + c] = v;
    }
    for (int i = threadIdx.x; i < K * K * 64; i += kThreads2d) {
      const int co = i & 7, ci = (i >> 3) & 7, tap = i >> 6;
      s_wt[i] = kernel_tf[((size_t)tap * Cin + c8 + ci) * Cout + co0 + co];      // [kh,kw,Cin,Cout]
    }
    __syncthreads();
#pragma unroll
    for (int kw = 0; kw < K; ++kw) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float4 rowv[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) rowv[r] = s_tile[(half * IR + warp * R * S + r) * IC + lane * S + kw];
#pragma unroll
        for (int kh = 0; kh < K; ++kh) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float* wp = s_wt + ((kh * K + kw) * 8 + half * 4 + c) * 8;
            const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
            for (int j = 0; j < R; ++j) {
              const float4 rv = rowv[j * S + kh];
              const float a = c == 0 ? rv.x : (c == 1 ? rv.y : (c == 2 ? rv.z : rv.w));
              acc[j][0] = fmaf(a, w0.x, acc[j][0]); acc[j][1] = fmaf(a, w0.y, acc[j][1]);
              acc[j][2] = fmaf(a, w0.z, acc[j][2]); acc[j][3] = fmaf(a, w0.w, acc[j][3]);
              acc[j][4] = fmaf(a, w1.x, acc[j][4]); acc[j][5] = fmaf(a, w1.y, acc[j][5]);
              acc[j][6] = fmaf(a, w1.z, acc[j][6]); acc[j][7] = fmaf(a, w1.w, acc[j][7]);
            }
          }
        }
      }
    }
  }
  const bool live = ox < Wo;
  float sm = 0.0f, sq = 0.0f;
  if (live) {
#pragma unroll
    for (int j = 0; j < R; ++j)
      if (oy0 + j < Ho) {
        float4* o = reinterpret_cast<float4*>(y + (((size_t)n * Ho + oy0 + j) * Wo + ox) * Cout + co0);
        o[0] = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
        o[1] = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
#pragma unroll
        for (int k = 0; k < kCoT; ++k) { sm += acc[j][k]; sq = fmaf(acc[j][k], acc[j][k], sq); }
      }
  }
  if (stats == nullptr) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sm += __shfl_xor_sync(0xffffffffu, sm, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  if (lane == 0) { s_red[0][warp] = sm; s_red[1][warp] = sq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ds = 0.0, dq = 0.0;
    for (int w = 0; w < kThreads2d / 32; ++w) { ds += (double)s_red[0][w]; dq += (double)s_red[1][w]; }
    double* st = stats + ((size_t)n * (Cout / kCoT) + blockIdx.y) * 2;
    atomicAdd(st, ds);
    atomicAdd(st + 1, dq);
  }
}

template <int K, int S, int R>
int launch_conv2d_tile(const float* xa, int ca, const float* xb, int cb, const float* kernel_tf, int n, int h, int w,
                       int cout, int ho, int wo, float* y, double* stats, cudaStream_t s) {
  constexpr int kTileRows = 4 * R;
  constexpr int IR = (kTileRows - 1) * S + K, IC = 31 * S + K;
  const size_t smem = (size_t)2 * IR * IC * sizeof(float4) + (size_t)K * K * 64 * sizeof(float);
  MVS_CUDA(cudaFuncSetAttribute(conv2d_tile_kernel<K, S, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(wo, 32) * ceil_div(ho, kTileRows), cout / kCoT, n);
  conv2d_tile_kernel<K, S, R><<<grid, kThreads2d, smem, s>>>(xa, ca, xb, cb, kernel_tf, h, w, cout, ho, wo,
                                                          tf_same_pad_before(h, K, S), tf_same_pad_before(w, K, S), y, stats);
  MVS_LAUNCH_CHECK("conv2d_tile_kernel");
  return MVSB200_OK;
}

int out_extent(int in, int stride, int transposed) { return transposed ? in * 2 : (in + stride - 1) / stride; }

int launch_conv2d(const float* xa, int ca, const float* xb, int cb, const float* kernel_tf, int n, int h, int w, int cout,
                  int k, int stride, int transposed, float* y, double* stats, cudaStream_t s) {
  MVS_CHECK_ARG(xa && kernel_tf && y && ca >= 1 && cb >= 0 && (cb == 0 || xb), "conv2d: NULL pointer");
  MVS_CHECK_ARG(n >= 1 && h >= 1 && w >= 1, "conv2d: bad shape N=%d %dx%d", n, h, w);
  MVS_CHECK_ARG(cout >= kCoT && cout % kCoT == 0, "conv2d: Cout=%d must be a multiple of %d", cout, kCoT);
  MVS_CHECK_ARG((k == 3 || k == 5) && (stride == 1 || stride == 2), "conv2d: kernel %d stride %d unsupported", k, stride);
  MVS_CHECK_ARG(!transposed || (k == 3 && stride == 2), "conv2d: the transposed convolution is 3x3 stride 2");
  MVS_CHECK_ARG(n <= 65535, "conv2d: too many views");
  const int ho = out_extent(h, stride, transposed), wo = out_extent(w, stride, transposed);
  const bool no_tile = tuning().unet_no_tile != 0;       // development switch
  if (!transposed && ca % 8 == 0 && cb % 8 == 0 && !no_tile) {
    if (k == 3 && stride == 1) return launch_conv2d_tile<3, 1, 4>(xa, ca, xb, cb, kernel_tf, n, h, w, cout, ho, wo, y, stats, s);
    if (k == 3 && stride == 2) return launch_conv2d_tile<3, 2, 4>(xa, ca, xb, cb, kernel_tf, n, h, w, cout, ho, wo, y, stats, s);
    if (k == 5 && stride == 2) return launch_conv2d_tile<5, 2, 4>(xa, ca, xb, cb, kernel_tf, n, h, w, cout, ho, wo, y, stats, s);
  }
  const size_t smem = (size_t)k * k * (ca + cb) * kCoT * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("conv2d: %d input channels exceed the shared-memory weight tile", ca + cb);
    return MVSB200_ERR_UNSUPPORTED;
  }
  const int tiles = ceil_div(wo, 32) * ceil_div(ho, kPxT);          // one warp per 32 columns x kPxT rows
  dim3 grid(ceil_div(tiles, kThreads2d / 32), cout / kCoT, n);
  if (transposed)
    conv2d_direct_kernel<true><<<grid, kThreads2d, smem, s>>>(xa, ca, xb, cb, kernel_tf, h, w, cout, k, stride, ho, wo, 0, 0,
                                                              y, stats);
  else
    conv2d_direct_kernel<false><<<grid, kThreads2d, smem, s>>>(xa, ca, xb, cb, kernel_tf, h, w, cout, k, stride, ho, wo,
                                                               tf_same_pad_before(h, k, stride),
                                                               tf_same_pad_before(w, k, stride), y, stats);
  MVS_LAUNCH_CHECK("conv2d_direct_kernel");
  return MVSB200_OK;
}

int launch_group_norm(float* y, const double* stats, const float* gamma, const float* beta, int n, int hw, int c, float eps,
                      int relu, cudaStream_t s) {
  MVS_CHECK_ARG(y && stats && gamma && beta && n >= 1 && hw >= 1 && c >= kCoT && c % kCoT == 0,
                "group_norm: bad arguments (C=%d must be a multiple of %d)", c, kCoT);
  const size_t total4 = (size_t)n * hw * c / 4;
  const int blocks = (int)((total4 + 255) / 256 < (size_t)sm_count_current() * 16 ? (total4 + 255) / 256 : (size_t)sm_count_current() * 16);
  group_norm_kernel<<<blocks, 256, 0, s>>>(y, stats, gamma, beta, hw, c, eps, relu, total4);
  MVS_LAUNCH_CHECK("group_norm_kernel");
  return MVSB200_OK;
}

// ---- whole tower ----------------------------------------------------------------------------------------------
struct UnetPlan {
  int h[MVSB200_UNET_LAYERS], w[MVSB200_UNET_LAYERS], c[MVSB200_UNET_LAYERS];     // output extents per layer
  size_t off[MVSB200_UNET_LAYERS];
  size_t stats_off, stats_bytes, total;
};

int make_unet_plan(int n, int H, int W, int base_filter, UnetPlan* p) {
  MVS_CHECK_ARG(n >= 1 && H >= 16 && W >= 16 && base_filter >= 8 && base_filter % 8 == 0,
                "unet: bad shape N=%d %dx%d base_filter=%d (groups of 8 channels need base_filter %% 8 == 0)", n, H, W,
                base_filter);
  MVS_CHECK_ARG(H % 16 == 0 && W % 16 == 0, "unet: H=%d and W=%d must be multiples of 16 (the concatenations of the "
                "reference graph do not close otherwise)", H, W);
  size_t off = 0;
  for (int l = 0; l < MVSB200_UNET_LAYERS; ++l) {
    const F2Layer& L = kUnet[l];
    const int ih = L.src_a < 0 ? H : p->h[L.src_a], iw = L.src_a < 0 ? W : p->w[L.src_a];
    p->h[l] = out_extent(ih, L.stride, L.transposed);
    p->w[l] = out_extent(iw, L.stride, L.transposed);
    p->c[l] = base_filter * L.mult;
    p->off[l] = off;
    off += align_up((size_t)n * p->h[l] * p->w[l] * p->c[l] * sizeof(float), 256);
  }
  p->stats_off = off;
  p->stats_bytes = (size_t)MVSB200_UNET_LAYERS * n * (16 * base_filter / kCoT) * 2 * sizeof(double);
  p->total = off + align_up(p->stats_bytes, 256);
  return MVSB200_OK;
}

}  // namespace

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_conv2d_layer(const float* xa, int ca, const float* xb, int cb, const float* kernel_tf, int n_views,
                                    int height, int width, int cout, int ksize, int stride, int transposed, float* y,
                                    double* stats, void* stream) {
  return launch_conv2d(xa, ca, xb, cb, kernel_tf, n_views, height, width, cout, ksize, stride, transposed, y, stats,
                       (cudaStream_t)stream);
}

extern "C" int mvsb200_group_norm(float* y, const double* stats, const float* gamma, const float* beta, int n_views,
                                  int pixels, int channels, float eps, int relu, void* stream) {
  return launch_group_norm(y, stats, gamma, beta, n_views, pixels, channels, eps, relu, (cudaStream_t)stream);
}

extern "C" size_t mvsb200_unet_workspace_bytes(int n_views, int height, int width, int base_filter) {
  UnetPlan p;
  if (make_unet_plan(n_views, height, width, base_filter, &p)) return 0;
  return p.total;
}

extern "C" int mvsb200_unet_layer_output(int n_views, int height, int width, int base_filter, int layer, size_t* offset,
                                         int* dims) {
  UnetPlan p;
  int rc = make_unet_plan(n_views, height, width, base_filter, &p);
  if (rc) return rc;
  MVS_CHECK_ARG(layer >= 0 && layer < MVSB200_UNET_LAYERS && offset && dims, "unet_layer_output: bad layer %d", layer);
  *offset = p.off[layer];
  dims[0] = p.h[layer]; dims[1] = p.w[layer]; dims[2] = p.c[layer];
  return MVSB200_OK;
}

extern "C" int mvsb200_unet_forward(const float* images, const mvsb200_unet_params* params, int n_views, int height,
                                    int width, int base_filter, float gn_eps, float* feats, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  MVS_CHECK_ARG(images && params && feats && workspace, "unet_forward: NULL pointer");
  UnetPlan p;
  int rc = make_unet_plan(n_views, height, width, base_filter, &p);
  if (rc) return rc;
  if (workspace_bytes < p.total) {
    set_error("unet_forward: workspace %zu < required %zu bytes", workspace_bytes, p.total);
    return MVSB200_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  double* stats = (double*)(ws + p.stats_off);
  MVS_CUDA(cudaMemsetAsync(stats, 0, p.stats_bytes, s));
  const int gmax = 16 * base_filter / kCoT;
  // development aid: MVSB200_UNET_PROFILE=1 prints per-layer device times (synchronises; not for timed runs)
  const bool profile = tuning().unet_profile != 0;
  cudaEvent_t pev[MVSB200_UNET_LAYERS + 1];
  if (profile) {
    for (int i = 0; i <= MVSB200_UNET_LAYERS; ++i) cudaEventCreate(&pev[i]);
    cudaEventRecord(pev[0], s);
  }
  for (int l = 0; l < MVSB200_UNET_LAYERS; ++l) {
    const F2Layer& L = kUnet[l];
    MVS_CHECK_ARG(params->kernel[l] != nullptr, "unet_forward: kernel[%d] (%s) is NULL", l, L.name);
    if (L.gn) MVS_CHECK_ARG(params->gamma[l] && params->beta[l], "unet_forward: gamma/beta[%d] (%s) is NULL", l, L.name);
    const float* xa = L.src_a < 0 ? images : (const float*)(ws + p.off[L.src_a]);
    const int ca = L.src_a < 0 ? 3 : p.c[L.src_a];
    const int ih = L.src_a < 0 ? height : p.h[L.src_a], iw = L.src_a < 0 ? width : p.w[L.src_a];
    const float* xb = L.src_b >= 0 ? (const float*)(ws + p.off[L.src_b]) : nullptr;
    const int cb = L.src_b >= 0 ? p.c[L.src_b] : 0;
    if (L.src_b >= 0)
      MVS_CHECK_ARG(p.h[L.src_b] == ih && p.w[L.src_b] == iw, "unet_forward: concat extents differ at %s", L.name);
    const bool last = l == MVSB200_UNET_LAYERS - 1;
    float* y = last ? feats : (float*)(ws + p.off[l]);
    double* st = L.gn ? stats + (size_t)l * n_views * gmax * 2 : nullptr;
    rc = launch_conv2d(xa, ca, xb, cb, params->kernel[l], n_views, ih, iw, p.c[l], L.k, L.stride, L.transposed, y, st, s);
    if (rc) return rc;
    if (L.gn) {
      rc = launch_group_norm(y, st, params->gamma[l], params->beta[l], n_views, p.h[l] * p.w[l], p.c[l], gn_eps, L.relu, s);
      if (rc) return rc;
    }
    if (profile) cudaEventRecord(pev[l + 1], s);
  }
  if (profile) {
    cudaStreamSynchronize(s);
    float total = 0.f;
    for (int l = 0; l < MVSB200_UNET_LAYERS; ++l) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, pev[l], pev[l + 1]);
      total += ms;
      fprintf(stderr, "[unet] %-10s %4dx%-4d C=%-3d %.3f ms\n", kUnet[l].name, p.h[l], p.w[l], p.c[l], ms);
    }
    fprintf(stderr, "[unet] total %.3f ms\n", total);
    for (int i = 0; i <= MVSB200_UNET_LAYERS; ++i) cudaEventDestroy(pev[i]);
  }
  return MVSB200_OK;
}
