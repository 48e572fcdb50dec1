// Training step of the hot path (BASELINE config 4): forward of model.py:257-372 (`inference`) in the
// fp32 parity mode, mvsnet_regression_loss (loss.py:15-29,190-220, loss_type 'original') and the gradients
// `opt.compute_gradients(loss)` asks for (train.py:429): every RegNetUS0 variable and the feature maps the path receives.
//
//   loss            masked mean absolute error in units of (depth_end - depth_start) / 191, + less-one / less-three
//   soft-argmin     dF[d] = -P[d] * (sample_d - depth) * g            (softmax of -F, model.py:345-366)
//   BN + ReLU       batch statistics are differentiated through (tf.layers.batch_normalization(training=True)):
//                   dx = gamma * rstd * (dy - mean(dy) - xhat * mean(dy * xhat)), dgamma = sum(dy * xhat), dbeta = sum(dy)
//   conv / deconv   input gradient = the SAME kernels run the other way round -- a stride-2 conv's dgrad is the
//                   transposed conv with the same filter (that is how TF defines conv3d_transpose), a transposed conv's
//                   dgrad is the stride-2 conv, a stride-1 conv's dgrad is a stride-1 conv with the filter flipped --
//                   through conv3d_direct_kernel; weight gradient = wgrad_kernel below
//   variance        d w_v = dS + 2 w_v dQ with dQ = dc / N, dS = -2 S dc / N^2   (both op orders, model.py:330-332 / :458-461)
//   warp            the exact adjoint of the zero-fill bilinear gather: the gradient is scattered with the same four
//                   weights ("bilinear-warp backward scatter", BASELINE config 4).  TF 1.12 itself registers a different
//                   gradient for ImageProjectiveTransform (the gradient image resampled with the inverse transform,
//                   SURVEY A.3); oracle/backward_oracle.py has both flavours, this file implements the adjoint.
// fp32 on CUDA cores throughout (the role conv3d_direct.cu plays for the forward): correctness against the oracle
// first; tensor-core dgrad / wgrad would reuse conv3d_tc.cu's fold planner.
#include "geometry.cuh"
#include "regnet_plan.h"

namespace mvsb200 {

int launch_conv3d_direct(const void* x, int x_dtype, const float* xs, const float* xb, const void* skip,
                         const float* ss, const float* sb, const float* kernel_tf, int D, int H, int W, int cin,
                         int cout, int stride, int transposed, void* y, int y_dtype, double* stats, cudaStream_t s,
                         int accumulate);
int launch_cost_volume_coef(const float* feats, const float* homographies, const float* coef_table, int n_views,
                            int depth_num, int hf, int wf, int channels, int order, int sampler, int out_dtype,
                            void* out, cudaStream_t s);
int regnet_forward_impl(const void* cost, int cost_dtype, int cost_planar, const mvsb200_regnet_params* params, int D,
                        int H, int W, int cin, int b, float eps, int precision, float* filtered, void* workspace,
                        size_t workspace_bytes, cudaStream_t s, TcRegress* regress, bool inspect, bool prepared);

constexpr int kMaxViewsBw = 7;       // source views (as the forward kernels)

namespace bw {

static unsigned blocks_for(size_t items, int per_block = 256) {
  const size_t cap = (size_t)sm_count_current() * 32, want = (items + per_block - 1) / per_block;
  return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

// ---------------------------------------------------------------------------------------------------------- loss
// acc[0] = valid pixels, acc[1] = sum |gt - est| over them, acc[2] / acc[3] = pixels within one / three intervals
__global__ void loss_reduce_kernel(const float* __restrict__ gt, const float* __restrict__ est, int n, float interval,
                                   double* __restrict__ acc) {
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float t = gt[i];
    if (t != 0.0f) {
      const float e = fabsf(t - est[i]);
      a[0] += 1.0; a[1] += (double)e;
      const float rel = e / interval;
      if (rel <= 1.0f) a[2] += 1.0;
      if (rel <= 3.0f) a[3] += 1.0;
    }
  }
  __shared__ double s_a[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double v = a[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_a[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += s_a[threadIdx.x][w];
    atomicAdd(acc + threadIdx.x, v);
  }
}

// g = d loss / d est (sign(0) = 0, as tf.abs differentiates) and the three metrics
__global__ void loss_grad_kernel(const float* __restrict__ gt, const float* __restrict__ est, int n, float interval,
                                 const double* __restrict__ acc, float* __restrict__ g, float* __restrict__ metrics) {
  const double denom = fabs(acc[0]) + 1e-6;
  const float k = (float)(1.0 / ((double)interval * denom));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float t = gt[i], d = t - est[i];
    g[i] = t != 0.0f ? (d > 0.0f ? -k : (d < 0.0f ? k : 0.0f)) : 0.0f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && metrics) {
    metrics[0] = (float)((acc[1] / (double)interval) / denom);
    metrics[1] = (float)(acc[2] / denom);
    metrics[2] = (float)(acc[3] / denom);
  }
}

// dF[d, p] = -P[d, p] * (sample_d - depth_p) * g_p, in place over the probability volume
__global__ void soft_argmin_backward_kernel(float* __restrict__ P, const float* __restrict__ depth,
                                            const float* __restrict__ g, int D, int npix, float start, float step) {
  const size_t total = (size_t)D * npix;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(i / npix), p = (int)(i - (size_t)d * npix);
    const float sample = start + step * (float)d;                      // tf.linspace (model.py:358)
    P[i] = -P[i] * (sample - depth[p]) * g[p];
  }
}

// ---------------------------------------------------------------------------------------------------------- BN + ReLU
// mr[c] = mean, mr[C + c] = 1 / sqrt(var + eps) from the forward's statistics (sum | sum of squares)
__global__ void bn_moments_kernel(const double* __restrict__ stats, int C, double count, float eps, float* __restrict__ mr) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = stats[c] / count;
  double var = stats[C + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  mr[c] = (float)mean;
  mr[C + c] = 1.0f / sqrtf((float)var + eps);
}

// sums[c] = sum dyh, sums[C + c] = sum dyh * xhat with dyh = G where the activation is positive (ReLU) else 0
__global__ void bn_backward_reduce_kernel(const float* __restrict__ G, const float* __restrict__ raw,
                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                          const float* __restrict__ mr, int C, size_t nvox, double* __restrict__ sums) {
  // thread = (voxel lane, channel): channels fastest so that loads coalesce; C is a power of two <= 64
  const int c = threadIdx.x % C, lanes = blockDim.x / C, vl = threadIdx.x / C;
  const float sc = scale[c], sh = shift[c], mean = mr[c], rstd = mr[C + c];
  double s0 = 0.0, s1 = 0.0;
  for (size_t v = (size_t)blockIdx.x * lanes + vl; v < nvox; v += (size_t)gridDim.x * lanes) {
    const float x = raw[v * C + c];
    const float dy = fmaf(x, sc, sh) > 0.0f ? G[v * C + c] : 0.0f;
    s0 += (double)dy;
    s1 += (double)(dy * ((x - mean) * rstd));
  }
  __shared__ double s_s[2][256];
  s_s[0][threadIdx.x] = s0; s_s[1][threadIdx.x] = s1;
  __syncthreads();
  if (threadIdx.x < C) {
    double a = 0.0, b = 0.0;
    for (int l = 0; l < lanes; ++l) { a += s_s[0][l * C + threadIdx.x]; b += s_s[1][l * C + threadIdx.x]; }
    atomicAdd(sums + threadIdx.x, a);
    atomicAdd(sums + C + threadIdx.x, b);
  }
}

// G <- gradient with respect to the raw (pre-BN) output; block 0 also writes dgamma / dbeta
__global__ void bn_backward_apply_kernel(float* __restrict__ G, const float* __restrict__ raw, const float* __restrict__ scale,
                                         const float* __restrict__ shift, const float* __restrict__ mr,
                                         const float* __restrict__ gamma, const double* __restrict__ sums, int C, size_t nvox,
                                         float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const size_t total = nvox * C;
  const double inv_n = 1.0 / (double)nvox;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float x = raw[i];
    const float dy = fmaf(x, scale[c], shift[c]) > 0.0f ? G[i] : 0.0f;
    const float xhat = (x - mr[c]) * mr[C + c];
    G[i] = gamma[c] * mr[C + c] * (dy - (float)(sums[c] * inv_n) - xhat * (float)(sums[C + c] * inv_n));
  }
  if (blockIdx.x == 0 && threadIdx.x < C) {
    if (dbeta) dbeta[threadIdx.x] = (float)sums[threadIdx.x];
    if (dgamma) dgamma[threadIdx.x] = (float)sums[C + threadIdx.x];
  }
}

// ---------------------------------------------------------------------------------------------------------- conv helpers
// stride-1 conv, dgrad: the same correlation with the filter flipped in space and its channel roles swapped:
// out[tap][co][ci] = w[26 - tap][ci][co]   (w: [27][Cin][Cout] -> a conv filter [27][Cout][Cin] applied to dY)
__global__ void flip_filter_kernel(const float* __restrict__ w, int cin, int cout, float* __restrict__ out) {
  const int total = 27 * cin * cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % cin, co = (i / cin) % cout, tap = i / (cin * cout);
    out[i] = w[((size_t)(26 - tap) * cin + ci) * cout + co];
  }
}

__device__ __forceinline__ float act_in(const float* x, const float* scale, const float* shift, size_t off, int c) {
  const float v = __ldg(x + off);
  return scale ? fmaxf(fmaf(v, scale[c], shift[c]), 0.0f) : v;
}

// Weight gradient of one layer.  "Base" voxels b run over the conv's OUTPUT volume (conv: x index = b * stride + tap - pad,
// r index = b) or over the transposed conv's INPUT volume (x index = b, r index = 2 b + tap); dW[tap][ci][co] (conv) or
// dW[tap][co][ci] (transposed) += sum_b X[x index][ci] * R[r index][co], X = relu(bn(raw)) (+ the skip activation).
// grid = (voxel chunks, 27 taps); a block stages tiles of 64 base voxels (X and R rows) in shared memory; a thread owns a
// 4 x TCO patch of the Cin x Cout matrix, the groups of threads that cover the matrix split the tile's voxels.
constexpr int kWgTile = 256;      // (64: three block barriers per 64 voxels paced the kernel)
template <bool TRANSPOSED>
__global__ void __launch_bounds__(256)
wgrad_kernel(const float* __restrict__ x, const float* __restrict__ xs, const float* __restrict__ xb,
             const float* __restrict__ skip, const float* __restrict__ ss, const float* __restrict__ sb,
             const float* __restrict__ r, int Dx, int Hx, int Wx, int Dr, int Hr, int Wr, int Db, int Hb, int Wb, int cin,
             int cout, int stride, int pad_d, int pad_h, int pad_w, int chunk, float* __restrict__ dw) {
  extern __shared__ __align__(16) float s_mem[];
  float* s_x = s_mem;                       // [kWgTile][cin]
  float* s_r = s_mem + kWgTile * cin;       // [kWgTile][cout]
  const int tap = blockIdx.y, kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
  const int tco = cout >= 4 ? 4 : cout;
  const int tiles_co = cout / tco, tiles = (cin / 4) * tiles_co;      // threads that cover the matrix once
  const int groups = 256 / tiles, grp = threadIdx.x / tiles, tin = threadIdx.x % tiles;
  const bool worker = grp < groups;
  const int ci0 = (tin / tiles_co) * 4, co0 = (tin % tiles_co) * tco;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
  const size_t nb = (size_t)Db * Hb * Wb;
  const size_t b_begin = (size_t)blockIdx.x * chunk, b_end = b_begin + chunk < nb ? b_begin + chunk : nb;
  // voxel offsets of the tile (X element offset / R element offset, -1 = outside), worked out once per voxel and tap
  __shared__ long long s_off[2][kWgTile];
  const bool vec4 = (cin & 3) == 0 && (cout & 3) == 0;
  for (size_t t0 = b_begin; t0 < b_end; t0 += kWgTile) {
    __syncthreads();
    if (threadIdx.x < kWgTile) {
      const size_t b = t0 + threadIdx.x;
      long long ox = -1, orr = -1;
      if (b < b_end) {
        const int bx = (int)(b % Wb), by = (int)((b / Wb) % Hb), bz = (int)(b / ((size_t)Wb * Hb));
        int xz, xy, xx, rz, ry, rx;
        if (TRANSPOSED) { xz = bz; xy = by; xx = bx; rz = 2 * bz + kd; ry = 2 * by + kh; rx = 2 * bx + kw; }
        else { xz = bz * stride + kd - pad_d; xy = by * stride + kh - pad_h; xx = bx * stride + kw - pad_w; rz = bz; ry = by; rx = bx; }
        if (xz >= 0 && xz < Dx && xy >= 0 && xy < Hx && xx >= 0 && xx < Wx && rz < Dr && ry < Hr && rx < Wr) {
          ox = (long long)((((size_t)xz * Hx + xy) * Wx + xx) * cin);
          orr = (long long)((((size_t)rz * Hr + ry) * Wr + rx) * cout);
        }
      }
      s_off[0][threadIdx.x] = ox; s_off[1][threadIdx.x] = orr;
    }
    __syncthreads();
    // stage the tile: thread -> (voxel, group of 4 channels), channels fastest (scalar channels when Cin / Cout % 4 != 0)
    if (vec4) {
      const int gx4 = cin >> 2, gr4 = cout >> 2;
      for (int i = threadIdx.x; i < kWgTile * (gx4 + gr4); i += 256) {
        const bool is_x = i < kWgTile * gx4;
        const int j = is_x ? i : i - kWgTile * gx4, G4 = is_x ? gx4 : gr4;
        const int v = j / G4, c = (j - v * G4) * 4;
        const long long off = s_off[is_x ? 0 : 1][v];
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (off >= 0) {
          if (is_x) {
            val = __ldg(reinterpret_cast<const float4*>(x + off + c));
            if (xs) {
              val.x = fmaxf(fmaf(val.x, xs[c], xb[c]), 0.0f); val.y = fmaxf(fmaf(val.y, xs[c + 1], xb[c + 1]), 0.0f);
              val.z = fmaxf(fmaf(val.z, xs[c + 2], xb[c + 2]), 0.0f); val.w = fmaxf(fmaf(val.w, xs[c + 3], xb[c + 3]), 0.0f);
            }
            if (skip) {
              float4 k = __ldg(reinterpret_cast<const float4*>(skip + off + c));
              if (ss) {
                k.x = fmaxf(fmaf(k.x, ss[c], sb[c]), 0.0f); k.y = fmaxf(fmaf(k.y, ss[c + 1], sb[c + 1]), 0.0f);
                k.z = fmaxf(fmaf(k.z, ss[c + 2], sb[c + 2]), 0.0f); k.w = fmaxf(fmaf(k.w, ss[c + 3], sb[c + 3]), 0.0f);
              }
              val.x += k.x; val.y += k.y; val.z += k.z; val.w += k.w;
            }
          } else {
            val = __ldg(reinterpret_cast<const float4*>(r + off + c));
          }
        }
        *reinterpret_cast<float4*>((is_x ? s_x + v * cin : s_r + v * cout) + c) = val;
      }
    } else {
      for (int i = threadIdx.x; i < kWgTile * (cin + cout); i += 256) {
        const bool is_x = i < kWgTile * cin;
        const int j = is_x ? i : i - kWgTile * cin, C = is_x ? cin : cout;
        const int v = j / C, c = j - v * C;
        const long long off = s_off[is_x ? 0 : 1][v];
        float val = 0.0f;
        if (off >= 0) {
          if (is_x) {
            val = act_in(x, xs, xb, (size_t)off + c, c);
            if (skip) val += act_in(skip, ss, sb, (size_t)off + c, c);
          } else {
            val = __ldg(r + off + c);
          }
        }
        (is_x ? s_x : s_r)[j] = val;
      }
    }
    __syncthreads();
    if (worker) {
      for (int v = grp; v < kWgTile; v += groups) {
        const float4 xv = *reinterpret_cast<const float4*>(s_x + v * cin + ci0);
        float rv[4] = {0.f, 0.f, 0.f, 0.f};
        if (tco == 4) { const float4 t = *reinterpret_cast<const float4*>(s_r + v * cout + co0); rv[0] = t.x; rv[1] = t.y; rv[2] = t.z; rv[3] = t.w; }
        else for (int b = 0; b < tco; ++b) rv[b] = s_r[v * cout + co0 + b];
        const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(xa[a], rv[b], acc[a][b]);
      }
    }
  }
  if (worker) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < tco; ++b) {
        const int ci = ci0 + a, co = co0 + b;
        float* dst = TRANSPOSED ? dw + ((size_t)tap * cout + co) * cin + ci : dw + ((size_t)tap * cin + ci) * cout + co;
        atomicAdd(dst, acc[a][b]);
      }
  }
}

// ---------------------------------------------------------------------------------------------------------- cost volume
// Backward of the variance + the warp (exact adjoint): thread = (reference pixel, 4-channel group), loop over the planes.
// Per plane the taps are gathered again (the forward keeps no warped volume), d w_v = dS + 2 w_v dQ is scattered with the
// four bilinear weights (float4 atomics), the reference view's share accumulates in registers.
__global__ void __launch_bounds__(256)
cost_backward_kernel(const float* __restrict__ feats, const float* __restrict__ coef, const float* __restrict__ dcost,
                     int n_views, int D, int Hf, int Wf, float* __restrict__ dfeats) {
  const int groups = 8;                                   // C = 32
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)Hf * Wf * groups) return;
  const int g = (int)(idx % groups);
  const size_t pix = idx / groups;
  const int x = (int)(pix % Wf), y = (int)(pix / Wf);
  const size_t plane = (size_t)Hf * Wf * 32;
  const float4 r = __ldg(reinterpret_cast<const float4*>(feats + pix * 32 + g * 4));
  const float inv_n = 1.0f / (float)n_views, inv_nn = 1.0f / (float)(n_views * n_views);
  float4 dref = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int d = 0; d < D; ++d) {
    const float4 dc = __ldg(reinterpret_cast<const float4*>(dcost + ((size_t)d * Hf * Wf + pix) * 32 + g * 4));
    float4 S = r;
    float4 w[kMaxViewsBw];
    Footprint fp[kMaxViewsBw];
    for (int v = 0; v < n_views - 1; ++v) {
      float ix, iy;
      transform_coords(coef + ((size_t)v * D + d) * 8, (float)x, (float)y, ix, iy);
      const Footprint f = make_footprint(ix, iy, Wf, Hf);
      fp[v] = f;
      const float* img = feats + (size_t)(v + 1) * plane + g * 4;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 p00 = (f.vy0 && f.vx0) ? __ldg(reinterpret_cast<const float4*>(img + ((size_t)f.y0 * Wf + f.x0) * 32)) : z;
      const float4 p01 = (f.vy0 && f.vx1) ? __ldg(reinterpret_cast<const float4*>(img + ((size_t)f.y0 * Wf + f.x0 + 1) * 32)) : z;
      const float4 p10 = (f.vy1 && f.vx0) ? __ldg(reinterpret_cast<const float4*>(img + ((size_t)(f.y0 + 1) * Wf + f.x0) * 32)) : z;
      const float4 p11 = (f.vy1 && f.vx1) ? __ldg(reinterpret_cast<const float4*>(img + ((size_t)(f.y0 + 1) * Wf + f.x0 + 1) * 32)) : z;
      float4 wv;
      wv.x = f.wyl * (f.wxl * p00.x + f.wxr * p01.x) + f.wyr * (f.wxl * p10.x + f.wxr * p11.x);
      wv.y = f.wyl * (f.wxl * p00.y + f.wxr * p01.y) + f.wyr * (f.wxl * p10.y + f.wxr * p11.y);
      wv.z = f.wyl * (f.wxl * p00.z + f.wxr * p01.z) + f.wyr * (f.wxl * p10.z + f.wxr * p11.z);
      wv.w = f.wyl * (f.wxl * p00.w + f.wxr * p01.w) + f.wyr * (f.wxl * p10.w + f.wxr * p11.w);
      w[v] = wv;
      S.x += wv.x; S.y += wv.y; S.z += wv.z; S.w += wv.w;
    }
    // cost = Q/N - S^2/N^2 (either op order): dQ = dc/N, dS = -2 S dc / N^2
    const float4 dQ = make_float4(dc.x * inv_n, dc.y * inv_n, dc.z * inv_n, dc.w * inv_n);
    const float4 dS = make_float4(-2.0f * S.x * dc.x * inv_nn, -2.0f * S.y * dc.y * inv_nn, -2.0f * S.z * dc.z * inv_nn,
                                  -2.0f * S.w * dc.w * inv_nn);
    dref.x += dS.x + 2.0f * r.x * dQ.x; dref.y += dS.y + 2.0f * r.y * dQ.y;
    dref.z += dS.z + 2.0f * r.z * dQ.z; dref.w += dS.w + 2.0f * r.w * dQ.w;
    for (int v = 0; v < n_views - 1; ++v) {
      const Footprint f = fp[v];
      const float4 dw = make_float4(dS.x + 2.0f * w[v].x * dQ.x, dS.y + 2.0f * w[v].y * dQ.y, dS.z + 2.0f * w[v].z * dQ.z,
                                    dS.w + 2.0f * w[v].w * dQ.w);
      float* out = dfeats + (size_t)(v + 1) * plane + g * 4;
      auto scatter = [&](bool ok, int yy, int xx, float wgt) {
        if (!ok || wgt == 0.0f) return;
        atomicAdd(reinterpret_cast<float4*>(out + ((size_t)yy * Wf + xx) * 32),
                  make_float4(wgt * dw.x, wgt * dw.y, wgt * dw.z, wgt * dw.w));
      };
      scatter(f.vy0 && f.vx0, f.y0, f.x0, f.wyl * f.wxl);
      scatter(f.vy0 && f.vx1, f.y0, f.x0 + 1, f.wyl * f.wxr);
      scatter(f.vy1 && f.vx0, f.y0 + 1, f.x0, f.wyr * f.wxl);
      scatter(f.vy1 && f.vx1, f.y0 + 1, f.x0 + 1, f.wyr * f.wxr);
    }
  }
  // the reference view receives no scatter: plain store
  *reinterpret_cast<float4*>(dfeats + pix * 32 + g * 4) = dref;
}

}  // namespace bw

// ---------------------------------------------------------------------------------------------------------- host side
namespace {
struct TrainPlan {
  size_t hom_off, coef_off, cost_off, filtered_off, prob_off, pmap_off, g_off, regnet_off, regnet_bytes;
  size_t grad_off[MVSB200_REGNET_LAYERS];     // G_i: gradient with respect to layer i's activation / raw output
  size_t dcost_off, flip_off, sums_off, mr_off, acc_off, total;
};
void make_train_plan(int n_views, int D, int hf, int wf, int C, int b, TrainPlan* tp, RegnetPlan* rp) {
  make_plan(D, hf, wf, C, b, MVSB200_PRECISION_FP32, rp);
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += align_up(bytes, 256); return o; };
  const size_t vol = (size_t)D * hf * wf;
  tp->hom_off = take((size_t)(n_views - 1) * D * 9 * sizeof(float));
  tp->coef_off = take((size_t)(n_views - 1) * D * 8 * sizeof(float));
  tp->cost_off = take(vol * C * sizeof(float));
  tp->filtered_off = take(vol * sizeof(float));
  tp->prob_off = take(vol * sizeof(float));             // softmax(-F), then dF in place
  tp->pmap_off = take((size_t)hf * wf * sizeof(float));
  tp->g_off = take((size_t)hf * wf * sizeof(float));
  tp->regnet_bytes = rp->total;
  tp->regnet_off = take(rp->total);
  for (int i = 0; i < MVSB200_REGNET_LAYERS; ++i)
    tp->grad_off[i] = i == MVSB200_L_3DCONV6_2 ? 0 : take(rp->vox[rp->layer[i].out_level] * rp->layer[i].cout * sizeof(float));
  tp->dcost_off = take(vol * C * sizeof(float));
  tp->flip_off = take((size_t)27 * 64 * 64 * sizeof(float));
  tp->sums_off = take((size_t)MVSB200_REGNET_LAYERS * 2 * 64 * sizeof(double));
  tp->mr_off = take((size_t)MVSB200_REGNET_LAYERS * 2 * 64 * sizeof(float));
  tp->acc_off = take(4 * sizeof(double));
  tp->total = off;
}
}  // namespace
}  // namespace mvsb200

using namespace mvsb200;

extern "C" size_t mvsb200_train_workspace_bytes(int n_views, int depth_num, int hf, int wf, int channels, int base_filter) {
  if (n_views < 2 || depth_num <= 0 || hf <= 0 || wf <= 0 || channels <= 0 || base_filter <= 0) return 0;
  TrainPlan tp;
  RegnetPlan rp;
  make_train_plan(n_views, depth_num, hf, wf, channels, base_filter, &tp, &rp);
  return tp.total;
}

extern "C" int mvsb200_train_step(const float* feats, const float* cams, const float* gt_depth, int n_views, int depth_num,
                                  int hf, int wf, int channels, float depth_start, float depth_interval, int order,
                                  const mvsb200_regnet_params* params, int base_filter, float bn_eps,
                                  const mvsb200_regnet_grads* grads, float* dfeats, float* depth_map, float* metrics,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  using namespace bw;
  MVS_CHECK_ARG(feats && cams && gt_depth && params && grads && dfeats && depth_map && metrics && workspace,
                "train_step: NULL pointer");
  MVS_CHECK_ARG(n_views >= 2 && n_views - 1 <= kMaxViewsBw, "train_step: 2..%d views (got %d)", kMaxViewsBw + 1, n_views);
  MVS_CHECK_ARG(channels == 32, "train_step: 32 feature channels (got %d)", channels);
  MVS_CHECK_ARG(base_filter % 4 == 0 && base_filter <= 8, "train_step: base_filter must be 4 or 8 (got %d)", base_filter);
  int rc = check_regnet_shape(depth_num, hf, wf, channels, base_filter);
  if (rc) return rc;
  TrainPlan tp;
  RegnetPlan p;
  make_train_plan(n_views, depth_num, hf, wf, channels, base_filter, &tp, &p);
  if (workspace_bytes < tp.total) {
    set_error("train_step: workspace %zu < required %zu bytes", workspace_bytes, tp.total);
    return MVSB200_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float* homs = (float*)(ws + tp.hom_off);
  float* coefs = (float*)(ws + tp.coef_off);
  float* cost = (float*)(ws + tp.cost_off);
  float* filtered = (float*)(ws + tp.filtered_off);
  float* prob = (float*)(ws + tp.prob_off);
  float* pmap = (float*)(ws + tp.pmap_off);
  float* g = (float*)(ws + tp.g_off);
  char* rws = ws + tp.regnet_off;
  const int npix = hf * wf;
  const size_t vol = (size_t)depth_num * npix;

  // ---- forward (fp32 parity mode; the training graph's variance order is the caller's choice, model.py:330-332) ----
  rc = launch_homographies(cams, n_views, depth_num, depth_start, depth_interval, 0, homs, coefs, s);
  if (rc) return rc;
  rc = launch_cost_volume_coef(feats, homs, coefs, n_views, depth_num, hf, wf, channels, order, MVSB200_SAMPLER_TRANSFORM,
                               MVSB200_F32, cost, s);
  if (rc) return rc;
  rc = regnet_forward_impl(cost, MVSB200_F32, 0, params, depth_num, hf, wf, channels, base_filter, bn_eps,
                           MVSB200_PRECISION_FP32, filtered, rws, tp.regnet_bytes, s, nullptr, true, false);
  if (rc) return rc;
  rc = launch_depth_regress(filtered, depth_num, hf, wf, depth_start, depth_interval, 0, 4, depth_map, pmap, prob, s);
  if (rc) return rc;

  // ---- loss (loss.py:15-29, 190-220) ----
  volatile float dm1 = (float)depth_num - 1.0f;
  volatile float span = dm1 * depth_interval;                       // depth_end - depth_start (model.py:378)
  const float loss_interval = (float)span / 191.0f;                 // loss.py:194
  const float lin_step = depth_num > 1 ? (float)span / (float)dm1 : 0.0f;
  double* acc = (double*)(ws + tp.acc_off);
  MVS_CUDA(cudaMemsetAsync(acc, 0, 4 * sizeof(double), s));
  loss_reduce_kernel<<<blocks_for(npix), 256, 0, s>>>(gt_depth, depth_map, npix, loss_interval, acc);
  MVS_LAUNCH_CHECK("loss_reduce_kernel");
  loss_grad_kernel<<<blocks_for(npix), 256, 0, s>>>(gt_depth, depth_map, npix, loss_interval, acc, g, metrics);
  MVS_LAUNCH_CHECK("loss_grad_kernel");
  soft_argmin_backward_kernel<<<blocks_for(vol), 256, 0, s>>>(prob, depth_map, g, depth_num, npix, depth_start, lin_step);
  MVS_LAUNCH_CHECK("soft_argmin_backward_kernel");

  // ---- RegNetUS0 backward, layers in reverse order (every activation's consumers come first) ----
  const int cpad = plan_cpad(p);
  const double* stats = (const double*)(rws + p.stats_off);
  const float* scale = (const float*)(rws + p.scale_off);
  const float* shift = (const float*)(rws + p.shift_off);
  double* sums = (double*)(ws + tp.sums_off);
  float* mr = (float*)(ws + tp.mr_off);
  float* flip = (float*)(ws + tp.flip_off);
  float* dcost = (float*)(ws + tp.dcost_off);
  MVS_CUDA(cudaMemsetAsync(sums, 0, (size_t)MVSB200_REGNET_LAYERS * 2 * 64 * sizeof(double), s));
  bool seeded[MVSB200_REGNET_LAYERS + 1] = {};       // has G_i (index 11: dcost) received its first contribution?
  for (int i = MVSB200_REGNET_LAYERS - 1; i >= 0; --i) {
    const LayerDesc& L = p.layer[i];
    const bool last = i == MVSB200_L_3DCONV6_2;
    const int* di = p.dims[L.in_level];
    const int* dout = p.dims[L.out_level];
    const size_t nvox_out = p.vox[L.out_level];
    float* R = last ? prob : (float*)(ws + tp.grad_off[i]);      // gradient w.r.t. the raw output, after the block below
    if (!last) {
      MVS_CHECK_ARG(seeded[i], "train_step: internal: layer %d has no gradient", i);
      const float* raw = (const float*)(rws + p.raw_off[i]);
      float* mri = mr + (size_t)i * 2 * 64;
      double* smi = sums + (size_t)i * 2 * 64;
      bn_moments_kernel<<<1, 64, 0, s>>>(stats + (size_t)i * 2 * cpad, L.cout, (double)nvox_out, bn_eps, mri);
      MVS_LAUNCH_CHECK("bn_moments_kernel");
      bn_backward_reduce_kernel<<<blocks_for(nvox_out * L.cout), 256, 0, s>>>(R, raw, scale + (size_t)i * cpad,
                                                                             shift + (size_t)i * cpad, mri, L.cout, nvox_out, smi);
      MVS_LAUNCH_CHECK("bn_backward_reduce_kernel");
      bn_backward_apply_kernel<<<blocks_for(nvox_out * L.cout), 256, 0, s>>>(R, raw, scale + (size_t)i * cpad,
                                                                            shift + (size_t)i * cpad, mri, params->gamma[i], smi,
                                                                            L.cout, nvox_out, grads->gamma[i], grads->beta[i]);
      MVS_LAUNCH_CHECK("bn_backward_apply_kernel");
    }
    // the layer's input: relu(bn(raw_src)) (+ relu(bn(raw_skip))), or the cost volume
    const float* x = L.src < 0 ? cost : (const float*)(rws + p.raw_off[L.src]);
    const float* xs = L.src < 0 ? nullptr : scale + (size_t)L.src * cpad;
    const float* xb = L.src < 0 ? nullptr : shift + (size_t)L.src * cpad;
    const float* sk = L.skip < 0 ? nullptr : (const float*)(rws + p.raw_off[L.skip]);
    const float* ss = L.skip < 0 ? nullptr : scale + (size_t)L.skip * cpad;
    const float* sb = L.skip < 0 ? nullptr : shift + (size_t)L.skip * cpad;
    // weight gradient
    MVS_CHECK_ARG(grads->kernel[i] != nullptr, "train_step: grads->kernel[%d] is NULL", i);
    MVS_CUDA(cudaMemsetAsync(grads->kernel[i], 0, (size_t)27 * L.cin * L.cout * sizeof(float), s));
    {
      const int* db = L.transposed ? di : dout;          // base voxels: transposed conv walks its input volume
      const size_t nb = (size_t)db[0] * db[1] * db[2];
      const int chunk = 4096;
      dim3 grid((unsigned)((nb + chunk - 1) / chunk), 27);
      const size_t smem = (size_t)kWgTile * (L.cin + L.cout) * sizeof(float);
      const int pd = L.transposed ? 0 : tf_same_pad_before(di[0], 3, L.stride), ph = L.transposed ? 0 : tf_same_pad_before(di[1], 3, L.stride),
                pw = L.transposed ? 0 : tf_same_pad_before(di[2], 3, L.stride);
      // (a tile of 256 voxels x up to 128 channels exceeds the 48 KB default; the attribute is per function and device)
      MVS_CUDA(cudaFuncSetAttribute((const void*)(L.transposed ? wgrad_kernel<true> : wgrad_kernel<false>),
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      if (L.transposed)
        wgrad_kernel<true><<<grid, 256, smem, s>>>(x, xs, xb, sk, ss, sb, R, di[0], di[1], di[2], dout[0], dout[1], dout[2], db[0],
                                                   db[1], db[2], L.cin, L.cout, 2, 0, 0, 0, chunk, grads->kernel[i]);
      else
        wgrad_kernel<false><<<grid, 256, smem, s>>>(x, xs, xb, sk, ss, sb, R, di[0], di[1], di[2], dout[0], dout[1], dout[2], db[0],
                                                    db[1], db[2], L.cin, L.cout, L.stride, pd, ph, pw, chunk, grads->kernel[i]);
      MVS_LAUNCH_CHECK("wgrad_kernel");
    }
    // input gradient, added to the gradient of every activation the input is made of
    const int targets[2] = {L.src < 0 ? MVSB200_REGNET_LAYERS : L.src, L.skip};
    if (!L.transposed && L.stride == 1) {
      flip_filter_kernel<<<blocks_for((size_t)27 * L.cin * L.cout), 256, 0, s>>>(params->kernel[i], L.cin, L.cout, flip);
      MVS_LAUNCH_CHECK("flip_filter_kernel");
    }
    for (int t = 0; t < 2; ++t) {
      const int j = targets[t];
      if (j < 0) continue;
      float* Gj = j == MVSB200_REGNET_LAYERS ? dcost : (float*)(ws + tp.grad_off[j]);
      const int accumulate = seeded[j] ? 1 : 0;
      if (L.transposed) {
        // forward: transposed conv (kernel [27][Cout][Cin]); dgrad = the stride-2 conv of R with the same array read
        // as a conv filter [27][Cin' = Cout][Cout' = Cin]
        rc = launch_conv3d_direct(R, MVSB200_F32, nullptr, nullptr, nullptr, nullptr, nullptr, params->kernel[i], dout[0], dout[1],
                                  dout[2], L.cout, L.cin, 2, 0, Gj, MVSB200_F32, nullptr, s, accumulate);
      } else if (L.stride == 2) {
        // forward: stride-2 conv (kernel [27][Cin][Cout]); dgrad = the transposed conv of R with the same array read as
        // a transposed-conv filter [27][Cout' = Cin][Cin' = Cout] (even extents: SAME pads nothing in front)
        rc = launch_conv3d_direct(R, MVSB200_F32, nullptr, nullptr, nullptr, nullptr, nullptr, params->kernel[i], dout[0], dout[1],
                                  dout[2], L.cout, L.cin, 2, 1, Gj, MVSB200_F32, nullptr, s, accumulate);
      } else {
        rc = launch_conv3d_direct(R, MVSB200_F32, nullptr, nullptr, nullptr, nullptr, nullptr, flip, dout[0], dout[1], dout[2],
                                  L.cout, L.cin, 1, 0, Gj, MVSB200_F32, nullptr, s, accumulate);
      }
      if (rc) return rc;
      seeded[j] = true;
    }
  }
  // ---- variance + warp backward ----
  MVS_CUDA(cudaMemsetAsync(dfeats, 0, (size_t)n_views * npix * channels * sizeof(float), s));
  cost_backward_kernel<<<(unsigned)(((size_t)npix * 8 + 255) / 256), 256, 0, s>>>(feats, coefs, dcost, n_views, depth_num, hf, wf,
                                                                                 dfeats);
  MVS_LAUNCH_CHECK("cost_backward_kernel");
  return MVSB200_OK;
}
