// Kernel 2: fused homography warp + N-view variance cost volume.
// Replaces the D x (N-1) tf_transform_homography launches and the running-sum ops of
// model.py:423-463 (inference_mem) / model.py:315-334 (inference).  The (N-1) x D x H x W x C
// warped volume is never written: each thread keeps the running sum and squared sum of its
// voxels in registers across views and stores only the variance.
//
// Transform coefficients: one 8-float row per (view, plane) in global memory, staged per block in
// shared memory.  The whole-path entry points pass the table the homography kernel wrote into the
// caller's workspace (so calls on different streams do not share state); the stand-alone entry point
// derives it from the homographies into a stream-ordered temporary of its own (cudaMallocAsync / cudaFreeAsync on the
// caller's stream: calls in flight on different streams share nothing).
#include "geometry.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace mvsb200 {

__global__ void prepare_table_kernel(const float* __restrict__ homographies, int count, float* __restrict__ table) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  float h[9], t[8];
#pragma unroll
  for (int i = 0; i < 9; ++i) h[i] = homographies[idx * 9 + i];
  transform_coefs(h, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) table[idx * 8 + i] = t[i];
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: one issue slot for two lanes' worth of fp32 math)
__device__ __forceinline__ unsigned long long f2_bits(float2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 bits_f2(unsigned long long v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
  return bits_f2(d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(d);
}

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

// x-paired fp16 copy of the source-view features (see TAPS16 below): out [N][Hf][Wf+1][8][2][4]
// pad_rows = 0: out [N][Hf][Wf+1][8][2][4]; pad_rows = 1: [N][Hf+3][...], image row y at row y + 2, rows 0, 1 and
// Hf + 2 zero (the 8-byte footprint records of HINT = 2 rely on them instead of validity masks)
__global__ void pair_features_kernel(const float* __restrict__ feats, int n_views, int Hf, int Wf, int pad_rows,
                                     __half* __restrict__ out) {
  const int rows = pad_rows ? Hf + 3 : Hf;
  const size_t total = (size_t)n_views * rows * (Wf + 1) * 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i & 7);
    size_t t = i >> 3;
    const int j = (int)(t % (Wf + 1)); t /= (Wf + 1);     // pair j = pixels (j-1, j)
    const int y = (int)(t % rows) - (pad_rows ? 2 : 0);
    const size_t row = (t / rows) * Hf + (size_t)max(y, 0);    // view * Hf + y
    const bool in_y = y >= 0 && y < Hf;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (in_y && j >= 1) a = ldg4(feats + (row * Wf + (j - 1)) * 32 + g * 4);
    if (in_y && j < Wf) b = ldg4(feats + (row * Wf + j) * 32 + g * 4);
    __half2 h[4] = {__floats2half2_rn(a.x, a.y), __floats2half2_rn(a.z, a.w), __floats2half2_rn(b.x, b.y),
                    __floats2half2_rn(b.z, b.w)};
    *reinterpret_cast<uint4*>(out + i * 8) = *reinterpret_cast<const uint4*>(h);
  }
}

// ------------------------------------------------------------------------------------------------
// Fast path, C = 32: block = 32 reference pixels x 8 channel groups (float4), KDC planes per thread and
// group, a block walks kPlaneChunk consecutive planes (consecutive planes of a pixel sample neighbouring
// source pixels, so their taps hit in L1).  The KDC x (N-1) bilinear footprints of a pixel and plane group
// (sample position with 2 IEEE divisions, the four weights, the clamped tap coordinates) are built once,
// 8 at a time by the 8 lanes of the pixel, and handed to the other lanes through shared memory (one
// 16-byte and one 8-byte broadcast load per footprint; shuffles would cost as many L1 data-pipe wavefronts
// as the taps).  KDC is the smallest of 2 / 4 / 8 for which KDC*(N-1) fills whole rounds of 8: few
// accumulators per thread (KDC = 2 at N = 5) leave registers for 4 resident blocks per SM.  A tap outside
// the image gets weight 0 and a clamped (always mapped) address, which is arithmetically identical to the
// reference's zero fill (0 * finite feature) and keeps the loads unpredicated.  Taps are 128-bit loads
// through L1 (the features are L2-resident: 5 x 216 x 288 x 32 fp32 = 40 MB at config 2); the running sum
// and squared sum stay in registers across views, accumulated in view order.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxSrcViews = 7;
constexpr int kPlaneChunk = 48;

// OUT: 0 = fp32 [D,Hf,Wf,32], 1 = bf16 [D,Hf,Wf,32], 2 = bf16 in the regularizer's planar layouts (conv3d_tc.cu):
// out = CP8 [D][4][Hf][Wf][8] and out2 = PS8 [D][4][4][Hs][Ws][8] (either may be NULL)
// TAPS16: the source views are read from the x-paired fp16 copy built by pair_features_kernel (product mode):
// feats16 [N][Hf][Wf+1][8 groups][2 pixels][4 ch], pair j = pixels (j-1, j) with zeros outside the image, so the
// two horizontal taps of a footprint row arrive in ONE 16-byte load per lane and one 128-byte line per pixel
// (half the L1 wavefronts and load instructions of the fp32 path).  The reference view stays fp32.
// HINT (with TAPS16): the 4-tap blend itself runs in packed fp16 (HFMA2, two channels per instruction, no
// per-tap conversions); the blended value is widened once and the running sums stay fp32.  The blend error
// (~1e-3 relative) is below the bf16 rounding of the volume this variant writes.  HINT = 2: the footprint travels as
// an 8-byte record (offset of the first row pair + the two fractions as halves; the copy has zero guard rows, so no
// validity masks are needed and the second row is always one row further): half the broadcast wavefronts.
template <int OUT, int TX, int TY, int KDC, bool TAPS16, int HINT, int MINB = (KDC == 2 ? 4 : (KDC == 4 ? 3 : 2))>
__global__ void __launch_bounds__(256, MINB)
cost_volume_c32_kernel(const float* __restrict__ feats, const __half* __restrict__ feats16,
                       const float* __restrict__ coef, int n_views, int D, int d0g, int Dloc, int Hf, int Wf,
                       int order, void* __restrict__ out, void* __restrict__ out2) {
  // plane window (D-slab mode): local plane l = global plane d0g + l of the D-plane sweep, Dloc local planes are
  // written (those whose global index falls outside [0, D) are left alone); whole volume: d0g = 0, Dloc = D
  static_assert(TX * TY == 32, "tile must hold 32 pixels");
  static_assert(KDC == 2 || KDC == 4 || KDC == 8, "planes per thread");
  constexpr int VPR = 8 / KDC;            // source views per footprint round
  __shared__ float s_coef[kMaxSrcViews * kPlaneChunk * 8];
  __shared__ float4 s_fw[32][9];          // per pixel: bilinear weights of one round (+1: bank spread)
  __shared__ int2 s_fi[32][9];            // per pixel: packed clamped tap columns / rows
  __shared__ uint2 s_f8[HINT == 2 ? 32 : 1][9];   // HINT = 2: (byte offset of the first row pair, half2(fx, fy))
  __shared__ uint4 s_cells[OUT == 2 ? KDC * 4 * 32 + 16 : 1];   // one group of output cells (planar layouts)
  const int tid = threadIdx.x;
  const int g = tid & 7;            // channel group (4 channels) and footprint slot of the round
  const int p = tid >> 3;           // pixel within the tile
  const int x = blockIdx.x * TX + (p % TX);
  const int y = blockIdx.y * TY + (p / TX);
  const int dbeg = blockIdx.z * kPlaneChunk;
  const int n_src = n_views - 1;
  for (int i = tid; i < n_src * kPlaneChunk * 8; i += 256) {
    int v = i / (kPlaneChunk * 8), r = i - v * (kPlaneChunk * 8);
    int d = min(max(d0g + dbeg + (r >> 3), 0), D - 1);
    s_coef[i] = __ldg(coef + (v * D + d) * 8 + (r & 7));
  }
  __syncthreads();
  const bool active = (x < Wf) && (y < Hf);
  const int xc = min(x, Wf - 1), yc = min(y, Hf - 1);
  const size_t plane = (size_t)Hf * Wf * 32;
  const float4 r = ldg4(feats + ((size_t)yc * Wf + xc) * 32 + g * 4);
  const float4 rq = make_float4(r.x * r.x, r.y * r.y, r.z * r.z, r.w * r.w);
  const float inv_n = 1.0f / (float)n_views, inv_nn = 1.0f / (float)(n_views * n_views);
  const int dend = min(Dloc, dbeg + kPlaneChunk);
  const int rounds = (KDC * n_src) / 8;       // the launcher picks KDC so that this is exact
  const char* base16 = reinterpret_cast<const char*>(feats16) + g * 16;
  const unsigned row16 = (unsigned)(Wf + 1) * 128u;      // bytes per row of the paired copy
  // footprint slot g of a round: view (round*VPR + g / KDC), plane g % KDC of the group
  const int my_dd = g % KDC, my_dv = g / KDC;
  for (int d0 = dbeg; d0 < dend; d0 += KDC) {
    float4 S[KDC], Q[KDC];
#pragma unroll
    for (int dd = 0; dd < KDC; ++dd) { S[dd] = r; Q[dd] = rq; }
    for (int rd = 0; rd < rounds; ++rd) {
      {
        const int v = rd * VPR + my_dv;
        float ix, iy;
        transform_coords(&s_coef[(v * kPlaneChunk + (d0 - dbeg) + my_dd) * 8], (float)xc, (float)yc, ix, iy);
        const Footprint f = make_footprint(ix, iy, Wf, Hf);
        const float l_wxl = f.vx0 ? f.wxl : 0.0f, l_wxr = f.vx1 ? f.wxr : 0.0f;
        const float l_wyl = f.vy0 ? f.wyl : 0.0f, l_wyr = f.vy1 ? f.wyr : 0.0f;
        const int cx0 = min(max(f.x0, 0), Wf - 1), cx1 = min(max(f.x0 + 1, 0), Wf - 1);
        const int cy0 = min(max(f.y0, 0), Hf - 1), cy1 = min(max(f.y0 + 1, 0), Hf - 1);
        __syncwarp();
        if (TAPS16) {
          // the four tap weights as products, and the byte offsets of the two row pairs inside feats16: pair index
          // of (x0, x0+1) in [0, Wf] (either pixel may be the zero guard; its weight is 0 there)
          const int xp = min(max(f.x0 + 1, 0), Wf);
          const unsigned vbase = (unsigned)(v + 1) * (unsigned)(Hf * (Wf + 1));
          const unsigned off0 = (vbase + (unsigned)(cy0 * (Wf + 1) + xp)) * 128u;
          const unsigned off1 = (vbase + (unsigned)(cy1 * (Wf + 1) + xp)) * 128u;
          if (HINT == 2) {
            // 8-byte record: footprints with no tap inside the image point at the two zero rows on top
            const bool any = (f.vx0 || f.vx1) && (f.vy0 || f.vy1);
            const unsigned cell = (unsigned)(v + 1) * (unsigned)((Hf + 3) * (Wf + 1)) +
                                  (unsigned)((any ? f.y0 + 2 : 0) * (Wf + 1) + (any ? f.x0 + 1 : 0));
            const __half2 fr = __floats2half2_rn(any ? f.wxr : 0.0f, any ? f.wyr : 0.0f);
            s_f8[p][g] = make_uint2(cell * 128u, *reinterpret_cast<const uint32_t*>(&fr));
          } else if (HINT) {
            // one 16-byte record per footprint (a 128-bit shared load costs 4 data-pipe wavefronts per warp whatever
            // it broadcasts, so the weights travel as four halves next to the two offsets)
            const __half2 hx = __floats2half2_rn(l_wyl * l_wxl, l_wyl * l_wxr);
            const __half2 hy = __floats2half2_rn(l_wyr * l_wxl, l_wyr * l_wxr);
            float4 packed;
            packed.x = __uint_as_float(*reinterpret_cast<const uint32_t*>(&hx));
            packed.y = __uint_as_float(*reinterpret_cast<const uint32_t*>(&hy));
            packed.z = __uint_as_float(off0);
            packed.w = __uint_as_float(off1);
            s_fw[p][g] = packed;
          } else {
            s_fw[p][g] = make_float4(l_wyl * l_wxl, l_wyl * l_wxr, l_wyr * l_wxl, l_wyr * l_wxr);
            s_fi[p][g] = make_int2((int)off0, (int)off1);
          }
        } else {
          s_fw[p][g] = make_float4(l_wxl, l_wxr, l_wyl, l_wyr);
          s_fi[p][g] = make_int2(cx0 | (cx1 << 16), cy0 | (cy1 << 16));
        }
        __syncwarp();
      }
      uint4 ta = make_uint4(0u, 0u, 0u, 0u), tb = ta;
      int2 prev_fi = make_int2(-1, -1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        constexpr int kdc = KDC;
        const int dd = j % kdc;                       // compile-time after unrolling
        float4 fw;
        int2 fi;
        if (HINT == 2) {
          const uint2 rec = s_f8[p][j];
          const __half2 fr = *reinterpret_cast<const __half2*>(&rec.y);          // (fx, fy)
          const __half2 gr = __hsub2(__float2half2_rn(1.0f), fr);                 // (1 - fx, 1 - fy)
          const __half2 xw = __halves2half2(__low2half(gr), __low2half(fr));      // (wxl, wxr)
          const __half2 r0 = __hmul2(xw, __high2half2(gr)), r1 = __hmul2(xw, __high2half2(fr));
          fw.x = __uint_as_float(*reinterpret_cast<const uint32_t*>(&r0));        // (w00, w01)
          fw.y = __uint_as_float(*reinterpret_cast<const uint32_t*>(&r1));        // (w10, w11)
          fw.z = fw.w = 0.0f;
          fi = make_int2((int)rec.x, (int)(rec.x + row16));
        } else {
          fw = s_fw[p][j];
          if (TAPS16 && HINT) fi = make_int2(__float_as_int(fw.z), __float_as_int(fw.w));
          else fi = s_fi[p][j];
        }
        float4 w;
        if (TAPS16) {
          // consecutive planes of a pixel often fall into the same source cell (the sample moves a fraction of a
          // pixel per plane): a later plane of the group then reuses the taps already in registers (a predicated
          // load makes no L1 traffic for the lanes that skip it)
          if (!(HINT && (j % kdc) != 0 && fi.x == prev_fi.x)) ta = __ldg(reinterpret_cast<const uint4*>(base16 + (unsigned)fi.x));
          if (!(HINT && (j % kdc) != 0 && fi.y == prev_fi.y)) tb = __ldg(reinterpret_cast<const uint4*>(base16 + (unsigned)fi.y));
          prev_fi = fi;
          const uint4 a = ta, b = tb;
          if (HINT) {
            const __half2 hx = *reinterpret_cast<const __half2*>(&fw.x), hy = *reinterpret_cast<const __half2*>(&fw.y);
            const __half2 w00 = __low2half2(hx), w01 = __high2half2(hx), w10 = __low2half2(hy), w11 = __high2half2(hy);
            const __half2 hlo = __hfma2(w11, *reinterpret_cast<const __half2*>(&b.z),
                                __hfma2(w10, *reinterpret_cast<const __half2*>(&b.x),
                                __hfma2(w01, *reinterpret_cast<const __half2*>(&a.z),
                                __hmul2(w00, *reinterpret_cast<const __half2*>(&a.x)))));
            const __half2 hhi = __hfma2(w11, *reinterpret_cast<const __half2*>(&b.w),
                                __hfma2(w10, *reinterpret_cast<const __half2*>(&b.y),
                                __hfma2(w01, *reinterpret_cast<const __half2*>(&a.w),
                                __hmul2(w00, *reinterpret_cast<const __half2*>(&a.y)))));
            const float2 lo = __half22float2(hlo), hi = __half22float2(hhi);
            float2* S2 = reinterpret_cast<float2*>(&S[dd]);
            float2* Q2 = reinterpret_cast<float2*>(&Q[dd]);
            S2[0] = fadd2(S2[0], lo); S2[1] = fadd2(S2[1], hi);
            Q2[0] = ffma2(lo, lo, Q2[0]); Q2[1] = ffma2(hi, hi, Q2[1]);
            continue;
          }
          const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&a.x)), a1 = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
          const float2 a2 = __half22float2(*reinterpret_cast<const __half2*>(&a.z)), a3 = __half22float2(*reinterpret_cast<const __half2*>(&a.w));
          const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&b.x)), b1 = __half22float2(*reinterpret_cast<const __half2*>(&b.y));
          const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&b.z)), b3 = __half22float2(*reinterpret_cast<const __half2*>(&b.w));
          // fw = (w00, w01, w10, w11): 4 multiply-adds per channel instead of 6, two channels per packed
          // instruction (rounding differs from the reference op order by an ulp or two; this variant feeds a
          // bf16 volume)
          const float2 w00 = make_float2(fw.x, fw.x), w01 = make_float2(fw.y, fw.y), w10 = make_float2(fw.z, fw.z),
                       w11 = make_float2(fw.w, fw.w);
          const float2 lo = ffma2(w11, b2, ffma2(w10, b0, ffma2(w01, a2, fmul2(w00, a0))));
          const float2 hi = ffma2(w11, b3, ffma2(w10, b1, ffma2(w01, a3, fmul2(w00, a1))));
          float2* S2 = reinterpret_cast<float2*>(&S[dd]);
          float2* Q2 = reinterpret_cast<float2*>(&Q[dd]);
          S2[0] = fadd2(S2[0], lo); S2[1] = fadd2(S2[1], hi);
          Q2[0] = ffma2(lo, lo, Q2[0]); Q2[1] = ffma2(hi, hi, Q2[1]);
          continue;
        } else {
          const float wxl = fw.x, wxr = fw.y, wyl = fw.z, wyr = fw.w;
          const float* img = feats + (size_t)(rd * VPR + j / kdc + 1) * plane + g * 4;
          const int x0 = fi.x & 0xffff, x1 = fi.x >> 16;
          const int r0 = (fi.y & 0xffff) * Wf, r1 = (fi.y >> 16) * Wf;
          const float4 p00 = ldg4(img + (size_t)(r0 + x0) * 32), p01 = ldg4(img + (size_t)(r0 + x1) * 32);
          const float4 p10 = ldg4(img + (size_t)(r1 + x0) * 32), p11 = ldg4(img + (size_t)(r1 + x1) * 32);
          w.x = wyl * (wxl * p00.x + wxr * p01.x) + wyr * (wxl * p10.x + wxr * p11.x);
          w.y = wyl * (wxl * p00.y + wxr * p01.y) + wyr * (wxl * p10.y + wxr * p11.y);
          w.z = wyl * (wxl * p00.z + wxr * p01.z) + wyr * (wxl * p10.z + wxr * p11.z);
          w.w = wyl * (wxl * p00.w + wxr * p01.w) + wyr * (wxl * p10.w + wxr * p11.w);
        }
        S[dd].x += w.x; S[dd].y += w.y; S[dd].z += w.z; S[dd].w += w.w;
        Q[dd].x += w.x * w.x; Q[dd].y += w.y * w.y; Q[dd].z += w.z * w.z; Q[dd].w += w.w * w.w;
      }
    }
    // variance with reciprocal multiplies (<= 1 ulp from the reference's divisions, model.py:458-461 / :330-332)
    float4 c[KDC];
#pragma unroll
    for (int dd = 0; dd < KDC; ++dd) {
      if (order == MVSB200_ORDER_MEM) {
        c[dd].x = Q[dd].x * inv_n - (S[dd].x * S[dd].x) * inv_nn;
        c[dd].y = Q[dd].y * inv_n - (S[dd].y * S[dd].y) * inv_nn;
        c[dd].z = Q[dd].z * inv_n - (S[dd].z * S[dd].z) * inv_nn;
        c[dd].w = Q[dd].w * inv_n - (S[dd].w * S[dd].w) * inv_nn;
      } else {
        const float mx = S[dd].x * inv_n, my = S[dd].y * inv_n, mz = S[dd].z * inv_n, mw = S[dd].w * inv_n;
        c[dd].x = Q[dd].x * inv_n - mx * mx;
        c[dd].y = Q[dd].y * inv_n - my * my;
        c[dd].z = Q[dd].z * inv_n - mz * mz;
        c[dd].w = Q[dd].w * inv_n - mw * mw;
      }
    }
    if (OUT < 2) {
      if (active) {
#pragma unroll
        for (int dd = 0; dd < KDC; ++dd)
          if (d0 + dd < Dloc && (unsigned)(d0g + d0 + dd) < (unsigned)D)
            store_cost4(out, (((size_t)(d0 + dd) * Hf + y) * Wf + x) * 32 + g * 4, c[dd], OUT == 1);
      }
    } else {
      // planar layouts: lanes g and g^1 hold the two halves of a 16-byte cell (chunk g>>1).  They swap halves of
      // neighbouring planes so that the even lane owns the whole cell of plane 2k and the odd lane that of 2k+1;
      // the cells go through shared memory so that the global stores run along x (whole 32-byte sectors in both
      // the chunk-planar and the parity-split copy).
      const bool odd = g & 1;
      __syncthreads();                       // the previous group's cells have been read
#pragma unroll
      for (int k = 0; k < KDC / 2; ++k) {
        const float4 mine = odd ? c[2 * k + 1] : c[2 * k], give = odd ? c[2 * k] : c[2 * k + 1];
        __nv_bfloat162 g0 = __floats2bfloat162_rn(give.x, give.y), g1 = __floats2bfloat162_rn(give.z, give.w);
        const uint32_t t0 = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<uint32_t*>(&g0), 1);
        const uint32_t t1 = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<uint32_t*>(&g1), 1);
        __nv_bfloat162 m0 = __floats2bfloat162_rn(mine.x, mine.y), m1 = __floats2bfloat162_rn(mine.z, mine.w);
        const uint32_t u0 = *reinterpret_cast<uint32_t*>(&m0), u1 = *reinterpret_cast<uint32_t*>(&m1);
        // channel order inside the cell: even lane's 4 channels first
        const int pc = (2 * k + (odd ? 1 : 0)) * 4 + (g >> 1);
        s_cells[pc * 32 + p + (pc >> 1)] = odd ? make_uint4(t0, t1, u0, u1) : make_uint4(u0, u1, t0, t1);
      }
      __syncthreads();
      const int Hs = (Hf + 1) >> 1, Ws = (Wf + 1) >> 1;
      // cell i of the group: (plane, chunk) = i / 32, pixel slot = i % 32 visited parity-major along x
#pragma unroll
      for (int k = 0; k < KDC / 2; ++k) {
        const int i = k * 256 + tid;
        const int pc = i >> 5, l = i & 31;
        int px;                                           // pixel index inside the TX x TY tile
        if (TX == 16) px = (l >> 4) * 16 + 2 * (l & 7) + ((l >> 3) & 1);
        else px = 2 * (l & 15) + (l >> 4);
        const int xx = blockIdx.x * TX + (px % TX), yy = blockIdx.y * TY + (px / TX);
        const int d = d0 + (pc >> 2);
        if (xx < Wf && yy < Hf && d < Dloc && (unsigned)(d0g + d) < (unsigned)D) {
          const uint4 cell = s_cells[pc * 32 + px + (pc >> 1)];
          const size_t zc = (size_t)d * 4 + (pc & 3);
          if (out) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + ((zc * Hf + yy) * Wf + xx) * 8) = cell;
          if (out2)
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out2) +
                                      (((zc * 4 + (yy & 1) * 2 + (xx & 1)) * Hs + (yy >> 1)) * Ws + (xx >> 1)) * 8) = cell;
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
cost_volume_generic_kernel(const float* __restrict__ feats, const float* __restrict__ homographies,
                           const float* __restrict__ coef, int n_views, int D, int Hf, int Wf, int C, int order,
                           void* __restrict__ out) {
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
      transform_coords(coef + (v * D + d) * 8, (float)x, (float)y, ix, iy);
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

size_t cost_volume_pair_bytes(int n_views, int hf, int wf);
size_t cost_volume_window_scratch_bytes(int n_views, int hf, int wf);
bool cost_volume_window_ok(int n_views, int hf, int wf, int channels, int sampler);
int launch_cost_volume_window(const float* feats, const float* coef_table, int n_views, int depth_num, int d0g, int dloc,
                              int hf, int wf, int order, void* cp8, void* ps8, void* feats16, int blend32,
                              unsigned long long* stats, cudaStream_t s, bool feats16_ready = false);

// product mode (planar output, coefficient table, fp16 taps): is it the shared-memory window kernel (cost_volume_win.cu)
// that runs, i.e. is its planar fp16 feature copy the one to prepare?  (Not when the round-1 gather kernel is asked for.)
bool cost_volume_uses_window(int n_views, int dloc, int hf, int wf, int channels, int sampler) {
  const bool fast_ok = sampler == MVSB200_SAMPLER_TRANSFORM && channels == 32 && n_views - 1 <= kMaxSrcViews &&
                       hf < 65536 && wf < 32768;
  return fast_ok && tuning().cv_kernel == 0 && cost_volume_window_ok(n_views, hf, wf, channels, sampler) &&
         (size_t)dloc * 16 * ((hf + 1) / 2) * ((wf + 1) / 2) < ((size_t)1 << 31);      // 32-bit cell indices
}

// planar_ps8 != NULL or planar != 0: write the bf16 planar layouts (out = CP8, planar_ps8 = PS8); fast path only
static int launch_cost_volume_any(const float* feats, const float* homographies, int n_views, int depth_num, int hf,
                                  int wf, int channels, int order, int sampler, int out_dtype, void* out, int variant,
                                  int planar, void* planar_ps8, void* feats16, const float* coef_table,
                                  cudaStream_t s, int d0g = 0, int dloc = -1, bool feats16_ready = false) {
  // (d0g, dloc): plane window of the D-slab mode; the coefficient table always covers all depth_num planes
  if (dloc < 0) dloc = depth_num;
  MVS_CHECK_ARG(feats && homographies && (out || planar_ps8), "cost_volume: NULL pointer");
  MVS_CHECK_ARG(n_views >= 2 && depth_num >= 1 && hf >= 1 && wf >= 1 && channels >= 1,
                "cost_volume: bad shape N=%d D=%d %dx%dx%d", n_views, depth_num, hf, wf, channels);
  MVS_CHECK_ARG(order == MVSB200_ORDER_MEM || order == MVSB200_ORDER_TRAIN, "cost_volume: bad order %d", order);
  MVS_CHECK_ARG(sampler == MVSB200_SAMPLER_TRANSFORM || sampler == MVSB200_SAMPLER_LEGACY,
                "cost_volume: bad sampler %d", sampler);
  MVS_CHECK_ARG(out_dtype == MVSB200_F32 || out_dtype == MVSB200_BF16, "cost_volume: bad out_dtype %d", out_dtype);
  const int rows = (n_views - 1) * depth_num;
  const bool bf16 = out_dtype == MVSB200_BF16;
  const float* coef = coef_table;
  // stand-alone call: the coefficient rows live in a stream-ordered temporary, released (in stream order) on every
  // exit path of this function
  struct Temp {
    cudaStream_t s; void* p = nullptr;
    ~Temp() { if (p) cudaFreeAsync(p, s); }
  } temp{s};
  if (sampler == MVSB200_SAMPLER_TRANSFORM && !coef) {
    MVS_CUDA(cudaMallocAsync(&temp.p, (size_t)rows * 8 * sizeof(float), s));
    prepare_table_kernel<<<ceil_div(rows, 128), 128, 0, s>>>(homographies, rows, (float*)temp.p);
    MVS_LAUNCH_CHECK("prepare_table_kernel");
    coef = (const float*)temp.p;
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
  // product mode: the shared-memory window kernel (cost_volume_win.cu) unless the round-1 gather kernel is asked for
  if (planar && feats16 && coef && variant == 3 && cost_volume_uses_window(n_views, dloc, hf, wf, channels, sampler))
    return launch_cost_volume_window(feats, coef, n_views, depth_num, d0g, dloc, hf, wf, order, out, planar_ps8, feats16,
                                     tuning().cv_fp32_blend, nullptr, s, feats16_ready);
  MVS_CHECK_ARG(!feats16_ready, "cost_volume: the prepared feature copy belongs to the window kernel");
  if (variant >= 2) {
    const bool wide = variant == 2;
    const int tx = wide ? 32 : 16, ty = wide ? 1 : 2;
    dim3 grid(ceil_div(wf, tx), ceil_div(hf, ty), ceil_div(dloc, kPlaneChunk));
    MVS_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "cost_volume: grid too large");
#define CV_FAST(O, TX_, TY_, K_)                                                                                  \
  do {                                                                                                            \
    if (O == 2 && feats16 && half_interp && minb_env == 3 && K_ == 2)                                             \
      cost_volume_c32_kernel<O, TX_, TY_, 2, true, true, 3><<<grid, 256, 0, s>>>(feats, (const __half*)feats16,   \
                                                                               coef, n_views, depth_num, d0g, dloc, hf, wf, \
                                                                               order, out, planar_ps8);           \
    else if (O == 2 && feats16 && half_interp && minb_env == 2 && K_ == 2)                                        \
      cost_volume_c32_kernel<O, TX_, TY_, 2, true, true, 2><<<grid, 256, 0, s>>>(feats, (const __half*)feats16,   \
                                                                               coef, n_views, depth_num, d0g, dloc, hf, wf, \
                                                                               order, out, planar_ps8);           \
    else if (O == 2 && feats16 && half_interp && rec8)                                                            \
      cost_volume_c32_kernel<O, TX_, TY_, K_, true, 2><<<grid, 256, 0, s>>>(feats, (const __half*)feats16,        \
                                                                            coef, n_views, depth_num, d0g, dloc, hf, wf, \
                                                                            order, out, planar_ps8);              \
    else if (O == 2 && feats16 && half_interp)                                                                    \
      cost_volume_c32_kernel<O, TX_, TY_, K_, true, true><<<grid, 256, 0, s>>>(feats, (const __half*)feats16,     \
                                                                               coef, n_views, depth_num, d0g, dloc, hf, wf, \
                                                                               order, out, planar_ps8);           \
    else if (O == 2 && feats16)                                                                                   \
      cost_volume_c32_kernel<O, TX_, TY_, K_, true, false><<<grid, 256, 0, s>>>(feats, (const __half*)feats16,    \
                                                                                coef, n_views, depth_num, d0g, dloc, hf, wf, \
                                                                                order, out, planar_ps8);          \
    else                                                                                                          \
      cost_volume_c32_kernel<O, TX_, TY_, K_, false, false><<<grid, 256, 0, s>>>(feats, nullptr, coef, n_views,   \
                                                                                 depth_num, d0g, dloc, hf, wf, order, out, \
                                                                                 planar_ps8);                     \
  } while (0)
#define CV_FAST_K(O, TX_, TY_)                                  \
  do {                                                          \
    if (kdc == 2) CV_FAST(O, TX_, TY_, 2);                      \
    else if (kdc == 4) CV_FAST(O, TX_, TY_, 4);                 \
    else CV_FAST(O, TX_, TY_, 8);                               \
  } while (0)
    if (planar && ((hf | wf) & 1) && planar_ps8)
      MVS_CUDA(cudaMemsetAsync(planar_ps8, 0, (size_t)dloc * 16 * ((hf + 1) / 2) * ((wf + 1) / 2) * 16, s));
    const Tuning& tn = tuning();
    const bool half_interp = tn.cv_fp32_blend == 0;     // fp16 blend unless asked otherwise
    const int minb_env = tn.cv_minb;                    // tuning
    // the paired copy is addressed with 32-bit byte offsets: past 4 GiB the taps come from the fp32 features
    if (feats16 && cost_volume_pair_bytes(n_views, hf, wf) >= ((size_t)1 << 32)) feats16 = nullptr;
    // 8-byte footprint records over the zero-padded paired copy (tuning CV_REC16=1: the 16-byte records)
    const bool rec8 = planar && feats16 && half_interp && minb_env == 0 && tn.cv_rec16 == 0 &&
                      (size_t)n_views * (hf + 3) * (wf + 1) < ((size_t)1 << 25);
    if (planar && feats16) {
      pair_features_kernel<<<sm_count_current() * 8, 256, 0, s>>>(feats, n_views, hf, wf, rec8 ? 1 : 0, (__half*)feats16);
      MVS_LAUNCH_CHECK("pair_features_kernel");
    }
    // planes per thread: the smallest of 2 / 4 / 8 whose footprints fill whole rounds of 8 lanes
    const int n_src = n_views - 1;
    int kdc = (2 * n_src) % 8 == 0 ? 2 : ((4 * n_src) % 8 == 0 ? 4 : 8);
    if (const int k = tn.cv_kdc)                              // tuning: 2 / 4 / 8 when it still fills whole rounds
      if ((k == 2 || k == 4 || k == 8) && (k * n_src) % 8 == 0) kdc = k;
    if (planar) { if (wide) CV_FAST_K(2, 32, 1); else CV_FAST_K(2, 16, 2); }
    else if (wide) { if (bf16) CV_FAST_K(1, 32, 1); else CV_FAST_K(0, 32, 1); }
    else           { if (bf16) CV_FAST_K(1, 16, 2); else CV_FAST_K(0, 16, 2); }
#undef CV_FAST_K
#undef CV_FAST
    MVS_LAUNCH_CHECK("cost_volume_c32_kernel");
    return MVSB200_OK;
  }
  const int vec = channels % 4 == 0 ? 4 : 1;
  size_t total = (size_t)depth_num * hf * wf * (channels / vec);
  MVS_CHECK_ARG((total + 255) / 256 <= 0x7fffffffu, "cost_volume: problem too large for one launch");
  unsigned blocks = (unsigned)((total + 255) / 256);
#define CV_GEN(V, SM, BF)                                                                                  \
  cost_volume_generic_kernel<V, SM, BF><<<blocks, 256, 0, s>>>(feats, homographies, coef, n_views, depth_num, \
                                                               hf, wf, channels, order, out)
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
                                out, variant, 0, nullptr, nullptr, nullptr, s);
}

// whole-path variant: `coef_table` [(n_views-1), depth_num, 8] written by launch_homographies into caller memory
int launch_cost_volume_coef(const float* feats, const float* homographies, const float* coef_table, int n_views,
                            int depth_num, int hf, int wf, int channels, int order, int sampler, int out_dtype,
                            void* out, cudaStream_t s) {
  return launch_cost_volume_any(feats, homographies, n_views, depth_num, hf, wf, channels, order, sampler, out_dtype,
                                out, 0, 0, nullptr, nullptr, coef_table, s);
}

bool cost_volume_planar_ok(int n_views, int hf, int wf, int channels, int sampler) {
  return sampler == MVSB200_SAMPLER_TRANSFORM && channels == 32 && n_views - 1 <= kMaxSrcViews && hf < 65536 &&
         wf < 32768;
}

// feats16: scratch of cost_volume_pair_bytes() for the x-paired fp16 copy of the features (NULL: fp32 taps)
size_t cost_volume_pair_bytes(int n_views, int hf, int wf) { return (size_t)n_views * (hf + 3) * (wf + 1) * 128; }

int launch_cost_volume_planar(const float* feats, const float* homographies, int n_views, int depth_num, int hf,
                              int wf, int channels, int order, int sampler, void* cp8, void* ps8, void* feats16,
                              const float* coef_table, cudaStream_t s, bool feats16_ready) {
  return launch_cost_volume_any(feats, homographies, n_views, depth_num, hf, wf, channels, order, sampler,
                                MVSB200_BF16, cp8, 0, 1, ps8, feats16, coef_table, s, 0, -1, feats16_ready);
}

// D-slab mode: local planes [0, dloc) = global planes [d0g, d0g + dloc) of the depth_num-plane sweep
int launch_cost_volume_slab(const float* feats, const float* homographies, int n_views, int depth_num, int d0g,
                            int dloc, int hf, int wf, int channels, int order, int sampler, void* cp8, void* ps8,
                            void* feats16, const float* coef_table, cudaStream_t s) {
  return launch_cost_volume_any(feats, homographies, n_views, depth_num, hf, wf, channels, order, sampler,
                                MVSB200_BF16, cp8, 0, 1, ps8, feats16, coef_table, s, d0g, dloc);
}

}  // namespace mvsb200

extern "C" int mvsb200_cost_volume(const float* feats, const float* homographies, int n_views, int depth_num,
                                   int hf, int wf, int channels, int order, int sampler, int out_dtype, void* out,
                                   int variant, void* stream) {
  return mvsb200::launch_cost_volume(feats, homographies, n_views, depth_num, hf, wf, channels, order, sampler,
                                     out_dtype, out, variant, (cudaStream_t)stream);
}
