// Regularizer layer, fp32 parity mode: direct 3x3x3 convolution / stride-2 transposed convolution
// on CUDA cores with TF SAME padding (network.py:210, :327), the producer's batch-norm + ReLU and
// the skip-connection add folded into the input read (network.py:459, :496-508), and the batch
// statistics of the result reduced in the epilogue.  This is the bit-faithful fp32 path used for
// parity against the oracle; the bf16 tcgen05 path lives in conv3d_tc.cu.
#include "common.cuh"

namespace mvsb200 {

template <typename T> __device__ __forceinline__ float load_as_float(const T* p);
template <> __device__ __forceinline__ float load_as_float<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}

// act(t) = relu(t*scale + shift) (BN + ReLU of the producing layer) or identity when scale == NULL
__device__ __forceinline__ float act(float v, const float* scale, const float* shift, int c) {
  if (scale == nullptr) return v;
  return fmaxf(fmaf(v, scale[c], shift[c]), 0.0f);
}

// One thread = one output voxel x CO_T output channels.  Weights of the channel chunk are staged
// in shared memory as [27][Cin][CO_T] fp32.
template <typename TIN, typename TOUT, int CO_T, bool TRANSPOSED>
__global__ void __launch_bounds__(128)
conv3d_direct_kernel(const TIN* __restrict__ x, const float* __restrict__ x_scale, const float* __restrict__ x_shift,
                     const TIN* __restrict__ skip, const float* __restrict__ skip_scale,
                     const float* __restrict__ skip_shift, const float* __restrict__ kernel_tf, int D, int H, int W,
                     int Cin, int Cout, int stride, int Do, int Ho, int Wo, int pad_d, int pad_h, int pad_w,
                     TOUT* __restrict__ y, double* __restrict__ stats, int accumulate) {
  extern __shared__ float s_w[];                      // [27][Cin][CO_T]
  __shared__ float s_red[2][4][CO_T];
  const int co0 = blockIdx.y * CO_T;
  for (int i = threadIdx.x; i < 27 * Cin * CO_T; i += blockDim.x) {
    int co = i % CO_T, ci = (i / CO_T) % Cin, tap = i / (CO_T * Cin);
    float w = 0.0f;
    if (co0 + co < Cout)
      w = TRANSPOSED ? kernel_tf[((size_t)tap * Cout + (co0 + co)) * Cin + ci]     // [kd,kh,kw,Cout,Cin]
                     : kernel_tf[((size_t)tap * Cin + ci) * Cout + (co0 + co)];    // [kd,kh,kw,Cin,Cout]
    s_w[i] = w;
  }
  __syncthreads();
  const size_t nvox = (size_t)Do * Ho * Wo;
  const size_t vox = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = vox < nvox;
  float acc[CO_T];
#pragma unroll
  for (int k = 0; k < CO_T; ++k) acc[k] = 0.0f;
  if (valid) {
    const int ox = (int)(vox % Wo), oy = (int)((vox / Wo) % Ho), oz = (int)(vox / ((size_t)Wo * Ho));
    for (int kd = 0; kd < 3; ++kd) {
      int iz;
      if (TRANSPOSED) { int t = oz - kd; if (t < 0 || (t & 1)) continue; iz = t >> 1; }
      else iz = oz * stride + kd - pad_d;
      if (iz < 0 || iz >= D) continue;
      for (int kh = 0; kh < 3; ++kh) {
        int iy;
        if (TRANSPOSED) { int t = oy - kh; if (t < 0 || (t & 1)) continue; iy = t >> 1; }
        else iy = oy * stride + kh - pad_h;
        if (iy < 0 || iy >= H) continue;
        for (int kw = 0; kw < 3; ++kw) {
          int ix;
          if (TRANSPOSED) { int t = ox - kw; if (t < 0 || (t & 1)) continue; ix = t >> 1; }
          else ix = ox * stride + kw - pad_w;
          if (ix < 0 || ix >= W) continue;
          const size_t in_off = (((size_t)iz * H + iy) * W + ix) * Cin;
          const float* wt = s_w + (size_t)((kd * 3 + kh) * 3 + kw) * Cin * CO_T;
          for (int ci = 0; ci < Cin; ++ci) {
            float v = act(load_as_float<TIN>(x + in_off + ci), x_scale, x_shift, ci);
            if (skip) v += act(load_as_float<TIN>(skip + in_off + ci), skip_scale, skip_shift, ci);
#pragma unroll
            for (int k = 0; k < CO_T; ++k) acc[k] = fmaf(v, wt[ci * CO_T + k], acc[k]);
          }
        }
      }
    }
    TOUT* yo = y + vox * Cout + co0;
#pragma unroll
    for (int k = 0; k < CO_T; ++k)
      if (co0 + k < Cout) {
        if (sizeof(TOUT) == 2) reinterpret_cast<__nv_bfloat16*>(yo)[k] = __float2bfloat16_rn(acc[k]);
        else reinterpret_cast<float*>(yo)[k] = accumulate ? reinterpret_cast<float*>(yo)[k] + acc[k] : acc[k];
      }
  }
  if (stats == nullptr) return;
  // batch statistics of the fp32 result: warp shuffle -> shared -> one double atomic per channel
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < CO_T; ++k) {
    float s = valid ? acc[k] : 0.0f, q = valid ? acc[k] * acc[k] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) { s_red[0][warp][k] = s; s_red[1][warp][k] = q; }
  }
  __syncthreads();
  if (threadIdx.x < CO_T && co0 + threadIdx.x < Cout) {
    double s = 0.0, q = 0.0;
    for (int w = 0; w < 4; ++w) { s += (double)s_red[0][w][threadIdx.x]; q += (double)s_red[1][w][threadIdx.x]; }
    atomicAdd(stats + co0 + threadIdx.x, s);
    atomicAdd(stats + Cout + co0 + threadIdx.x, q);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int C, double count, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mean = stats[c] / count;
  double var = stats[C + c] / count - mean * mean;     // biased variance (tf.nn.moments), fp64
  if (var < 0.0) var = 0.0;
  float meanf = (float)mean, varf = (float)var;
  // tf.nn.batch_normalization: inv = rsqrt(var+eps)*gamma; y = x*inv + (beta - mean*inv)  (Appendix A.6)
  float inv = __fmul_rn(__fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(varf, eps))), gamma[c]);
  scale[c] = inv;
  shift[c] = __fsub_rn(beta[c], __fmul_rn(meanf, inv));
}

int launch_bn_finalize(const double* stats, const float* gamma, const float* beta, int channels, double count,
                       float eps, float* scale, float* shift, cudaStream_t s) {
  MVS_CHECK_ARG(stats && gamma && beta && scale && shift && channels > 0 && count > 0, "bn_finalize: bad arguments");
  bn_finalize_kernel<<<ceil_div(channels, 64), 64, 0, s>>>(stats, gamma, beta, channels, count, eps, scale, shift);
  MVS_LAUNCH_CHECK("bn_finalize_kernel");
  return MVSB200_OK;
}

// all layers of a network in one launch (block = layer); scale / shift rows are `cpad` apart, stats rows 2*cpad
struct BnAll {
  const float* gamma[16]; const float* beta[16]; int channels[16]; int channels_true[16]; double count[16];
};
__global__ void bn_finalize_all_kernel(const double* __restrict__ stats, const __grid_constant__ BnAll a, int cpad, int reps,
                                       int layers, float eps, float* __restrict__ scale, float* __restrict__ shift) {
  const int l = blockIdx.x, C = a.channels[l];      // C: row pitch of the statistics; gamma / beta hold channels_true
  for (int c = threadIdx.x; c < a.channels_true[l]; c += blockDim.x) {
    double sm = 0.0, sq = 0.0;
    for (int r = 0; r < reps; ++r) {          // partial copies [rep][layer][2*cpad]
      const double* st = stats + ((size_t)r * layers + l) * 2 * cpad;
      sm += st[c]; sq += st[C + c];
    }
    double mean = sm / a.count[l];
    double var = sq / a.count[l] - mean * mean;
    if (var < 0.0) var = 0.0;
    float meanf = (float)mean, varf = (float)var;
    float inv = __fmul_rn(__fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(varf, eps))), a.gamma[l][c]);
    scale[(size_t)l * cpad + c] = inv;
    shift[(size_t)l * cpad + c] = __fsub_rn(a.beta[l][c], __fmul_rn(meanf, inv));
  }
}

int launch_bn_finalize_all(const double* stats, const float* const* gamma, const float* const* beta, const int* channels,
                           const int* channels_true, const double* counts, int layers, int cpad, int reps, float eps,
                           float* scale, float* shift, cudaStream_t s) {
  MVS_CHECK_ARG(layers > 0 && layers <= 16, "bn_finalize_all: bad layer count");
  BnAll a;
  for (int i = 0; i < 16; ++i) {
    a.gamma[i] = i < layers ? gamma[i] : nullptr; a.beta[i] = i < layers ? beta[i] : nullptr;
    a.channels[i] = i < layers ? channels[i] : 0; a.count[i] = i < layers ? counts[i] : 1.0;
    a.channels_true[i] = i < layers ? channels_true[i] : 0;
  }
  bn_finalize_all_kernel<<<layers, 64, 0, s>>>(stats, a, cpad, reps, layers, eps, scale, shift);
  MVS_LAUNCH_CHECK("bn_finalize_all_kernel");
  return MVSB200_OK;
}

template <typename TIN, typename TOUT, bool TRANSPOSED>
static int launch_direct_t(const void* x, const float* xs, const float* xb, const void* skip, const float* ss,
                           const float* sb, const float* kernel_tf, int D, int H, int W, int cin, int cout,
                           int stride, void* y, double* stats, cudaStream_t s, int accumulate) {
  int Do, Ho, Wo, pd = 0, ph = 0, pw = 0;
  if (TRANSPOSED) { Do = 2 * D; Ho = 2 * H; Wo = 2 * W; }
  else {
    Do = ceil_div(D, stride); Ho = ceil_div(H, stride); Wo = ceil_div(W, stride);
    pd = tf_same_pad_before(D, 3, stride); ph = tf_same_pad_before(H, 3, stride); pw = tf_same_pad_before(W, 3, stride);
  }
  const size_t nvox = (size_t)Do * Ho * Wo;
  const int co_t = cout >= 16 ? 16 : (cout >= 8 ? 8 : 1);
  const int chunks = ceil_div(cout, co_t);
  const size_t smem = (size_t)27 * cin * co_t * sizeof(float);
  MVS_CHECK_ARG(smem <= 200 * 1024, "conv3d(fp32): Cin=%d too large for the weight stage", cin);
  dim3 grid((unsigned)((nvox + 127) / 128), (unsigned)chunks);
#define DIRECT(CO)                                                                                               \
  do {                                                                                                           \
    auto kfn = conv3d_direct_kernel<TIN, TOUT, CO, TRANSPOSED>;                                                  \
    MVS_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));                \
    kfn<<<grid, 128, smem, s>>>((const TIN*)x, xs, xb, (const TIN*)skip, ss, sb, kernel_tf, D, H, W, cin, cout,  \
                                stride, Do, Ho, Wo, pd, ph, pw, (TOUT*)y, stats, accumulate);                    \
  } while (0)
  if (co_t == 16) DIRECT(16); else if (co_t == 8) DIRECT(8); else DIRECT(1);
#undef DIRECT
  MVS_LAUNCH_CHECK("conv3d_direct_kernel");
  return MVSB200_OK;
}

int launch_conv3d_direct(const void* x, int x_dtype, const float* xs, const float* xb, const void* skip,
                         const float* ss, const float* sb, const float* kernel_tf, int D, int H, int W, int cin,
                         int cout, int stride, int transposed, void* y, int y_dtype, double* stats, cudaStream_t s,
                         int accumulate) {
  // accumulate != 0 (fp32 output only): y += result instead of y = result (the backward pass sums the input gradients
  // of every consumer of an activation)
  MVS_CHECK_ARG(!accumulate || y_dtype == MVSB200_F32, "conv3d(fp32): accumulate needs an fp32 output");
#define GO(TI, TO)                                                                                                   \
  (transposed ? launch_direct_t<TI, TO, true>(x, xs, xb, skip, ss, sb, kernel_tf, D, H, W, cin, cout, stride, y,     \
                                              stats, s, accumulate)                                                  \
              : launch_direct_t<TI, TO, false>(x, xs, xb, skip, ss, sb, kernel_tf, D, H, W, cin, cout, stride, y,    \
                                               stats, s, accumulate))
  if (x_dtype == MVSB200_F32 && y_dtype == MVSB200_F32) return GO(float, float);
  if (x_dtype == MVSB200_BF16 && y_dtype == MVSB200_F32) return GO(__nv_bfloat16, float);
  if (x_dtype == MVSB200_F32 && y_dtype == MVSB200_BF16) return GO(float, __nv_bfloat16);
  return GO(__nv_bfloat16, __nv_bfloat16);
#undef GO
}

}  // namespace mvsb200
