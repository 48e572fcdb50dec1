// Refinement glue after the hot path: model.py:753-811 `depth_refine` -- normalise the depth map to
// [0,1] over the sweep, resize depth / image / probability with tf.image.resize_bilinear, run the refinement tower on
// concat(image, depth[, prob]), scale the residual back and add it -- and the tower of network_type 'original',
// RefineNetConv (mvsnetworks.py:178-193: four 3x3 SAME convolutions WITH bias, ReLU after the first three).
// fp32, NHWC; these are small maps (one per reference view): straightforward kernels, no tensor cores.
#include "common.cuh"

namespace mvsb200 {

// tf.image.resize_bilinear of TF 1.x (align_corners = False, no half-pixel centres): in = out * (in_size / out_size),
// lower = floor(in), upper = min(lower + 1, in_size - 1), lerp = in - lower; top / bottom rows interpolated along x
// first, then along y, each as a + (b - a) * lerp.  Optional affine on the way: y = (resized - sub) * mul.
__global__ void resize_bilinear_kernel(const float* __restrict__ x, int n, int h, int w, int c, float* __restrict__ y,
                                       int oh, int ow, float sy, float sx, float sub, float mul) {
  const size_t total = (size_t)n * oh * ow * c;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    size_t t = i / c;
    const int ox = (int)(t % ow); t /= ow;
    const int oy = (int)(t % oh);
    const int b = (int)(t / oh);
    const float fy = __fmul_rn((float)oy, sy), fx = __fmul_rn((float)ox, sx);
    const int y0 = (int)floorf(fy), x0 = (int)floorf(fx);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = __fsub_rn(fy, (float)y0), lx = __fsub_rn(fx, (float)x0);
    const float* img = x + (size_t)b * h * w * c + ch;
    const float tl = img[((size_t)y0 * w + x0) * c], tr = img[((size_t)y0 * w + x1) * c];
    const float bl = img[((size_t)y1 * w + x0) * c], br = img[((size_t)y1 * w + x1) * c];
    const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
    const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
    const float v = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
    y[i] = __fmul_rn(__fsub_rn(v, sub), mul);
  }
}

// y = x * mul + add_scale * add (element-wise; add may be NULL): the residual scaled back to millimetres and added to
// the initial depth map (model.py:803-809)
__global__ void scale_add_kernel(const float* __restrict__ x, float mul, const float* __restrict__ add, size_t total,
                                 float* __restrict__ scaled, float* __restrict__ sum) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float r = __fmul_rn(x[i], mul);
    if (scaled) scaled[i] = r;
    if (sum) sum[i] = add ? __fadd_rn(r, add[i]) : r;
  }
}

// 3x3 SAME stride-1 convolution of concat(xa, xb) with bias and optional ReLU (tf.layers.conv2d(use_bias=True),
// network.py:171-206): one thread per output pixel and group of 4 output channels, taps in (kh, kw, ci) order
template <int CO>
__global__ void conv3x3_bias_kernel(const float* __restrict__ xa, int ca, const float* __restrict__ xb, int cb,
                                    const float* __restrict__ kernel, const float* __restrict__ bias, int n, int h, int w,
                                    int cout, int relu, float* __restrict__ y) {
  const int groups = (cout + CO - 1) / CO;
  const size_t total = (size_t)n * h * w * groups;
  const int cin = ca + cb;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    size_t t = i / groups;
    const int ox = (int)(t % w); t /= w;
    const int oy = (int)(t % h);
    const int b = (int)(t / h);
    float acc[CO];
#pragma unroll
    for (int k = 0; k < CO; ++k) acc[k] = 0.0f;
    for (int kh = 0; kh < 3; ++kh) {
      const int iy = oy + kh - 1;
      if (iy < 0 || iy >= h) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int ix = ox + kw - 1;
        if (ix < 0 || ix >= w) continue;
        const size_t pix = ((size_t)b * h + iy) * w + ix;
        const float* wrow = kernel + (size_t)(kh * 3 + kw) * cin * cout + g * CO;
        for (int ci = 0; ci < cin; ++ci) {
          const float v = ci < ca ? __ldg(xa + pix * ca + ci) : __ldg(xb + pix * cb + (ci - ca));
#pragma unroll
          for (int k = 0; k < CO; ++k)
            if (g * CO + k < cout) acc[k] = fmaf(v, __ldg(wrow + (size_t)ci * cout + k), acc[k]);
        }
      }
    }
    float* out = y + (((size_t)b * h + oy) * w + ox) * cout + g * CO;
#pragma unroll
    for (int k = 0; k < CO; ++k)
      if (g * CO + k < cout) {
        float v = acc[k] + (bias ? __ldg(bias + g * CO + k) : 0.0f);
        out[k] = relu ? fmaxf(v, 0.0f) : v;
      }
  }
}

static unsigned blocks_for(size_t items) {
  const size_t cap = (size_t)sm_count_current() * 16, want = (items + 255) / 256;
  return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_resize_bilinear(const float* x, int n, int height, int width, int channels, float* y,
                                       int out_height, int out_width, float subtract, float multiply, void* stream) {
  MVS_CHECK_ARG(x && y, "resize_bilinear: NULL pointer");
  MVS_CHECK_ARG(n > 0 && height > 0 && width > 0 && channels > 0 && out_height > 0 && out_width > 0,
                "resize_bilinear: bad shape %dx%dx%dx%d -> %dx%d", n, height, width, channels, out_height, out_width);
  // the scales as TF computes them: float(in) / float(out)
  const float sy = (float)height / (float)out_height, sx = (float)width / (float)out_width;
  const size_t total = (size_t)n * out_height * out_width * channels;
  resize_bilinear_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(x, n, height, width, channels, y, out_height,
                                                                              out_width, sy, sx, subtract, multiply);
  MVS_LAUNCH_CHECK("resize_bilinear_kernel");
  return MVSB200_OK;
}

extern "C" int mvsb200_scale_add(const float* x, float multiply, const float* add, size_t count, float* scaled,
                                 float* sum, void* stream) {
  MVS_CHECK_ARG(x && (scaled || sum) && count > 0, "scale_add: bad arguments");
  scale_add_kernel<<<blocks_for(count), 256, 0, (cudaStream_t)stream>>>(x, multiply, add, count, scaled, sum);
  MVS_LAUNCH_CHECK("scale_add_kernel");
  return MVSB200_OK;
}

extern "C" int mvsb200_conv2d_bias(const float* xa, int ca, const float* xb, int cb, const float* kernel_tf,
                                   const float* bias, int n, int height, int width, int cout, int relu, float* y,
                                   void* stream) {
  MVS_CHECK_ARG(xa && kernel_tf && y && ca > 0 && cb >= 0 && (cb == 0 || xb), "conv2d_bias: bad arguments");
  MVS_CHECK_ARG(n > 0 && height > 0 && width > 0 && cout > 0, "conv2d_bias: bad shape");
  const size_t total = (size_t)n * height * width * ((cout + 3) / 4);
  conv3x3_bias_kernel<4><<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(xa, ca, xb, cb, kernel_tf, bias, n, height,
                                                                              width, cout, relu, y);
  MVS_LAUNCH_CHECK("conv3x3_bias_kernel");
  return MVSB200_OK;
}
