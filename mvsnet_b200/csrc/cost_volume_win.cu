// Kernel 2, product mode: fused homography warp + N-view variance with the source views staged in shared
// memory by TMA.  Replaces the D x (N-1) tf_transform_homography launches (homography_warping.py:211-253) and
// the running-sum ops of model.py:423-463 (inference_mem) / model.py:315-334 (inference); the warped volume is
// never written.
//
// Work item = 32 x 4 reference pixels x 8 consecutive depth planes (1 024 voxels).  Consecutive planes of a pixel
// sample a source view a fraction of a pixel apart, so the taps of a whole work item fall into
// one small window of each source view: the bounding box of the 8 corner samples (the sample position is projective
// in (x, y) and a Moebius function of the depth, hence monotone along every edge of the box).  Warp 0
// computes that box per (work item, view) and has the TMA unit copy it from an fp16 chunk-planar copy of the
// features ([N][4][Hf][Wf] cells of 8 channels = 16 bytes) into a shared-memory ring -- cells outside the image are
// zero-filled by the TMA unit, which IS the reference's zero fill -- one 8-channel chunk at a time (stage = work
// item x chunk; 48 x 16 cells per view, loaded as one or two boxes of 8 rows).  the 512 threads own two voxels
// each: the bilinear footprint of a voxel in every view (two IEEE divisions, weights) is computed once per work
// item and kept in registers (3 words per view), then every stage costs 4 conflict-free 16-byte shared-memory loads
// and 16 packed-half FMAs per (voxel, view); running sum and squared sum stay in fp32 registers across the views;
// the variance goes out as whole 16-byte cells of the regularizer's two planar layouts (lanes run along x: 512
// contiguous bytes per warp, no staging).  A voxel whose footprint is not inside the staged window (wild geometry,
// rounding at a box edge, window larger than the buffer) reads its taps from global memory instead: correctness
// never depends on the window.
#include "geometry.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <mutex>
#include <string.h>

namespace mvsb200 {
using namespace umma;

namespace cvw {

constexpr int TXP = 32, TYP = 4, PL = 8;                 // reference pixels x planes of a work item
constexpr int kItems = TXP * TYP * PL;                   // 1 024 voxels
constexpr int kThreads = 512, kIPT = kItems / kThreads;  // 16 warps = 4 per scheduler: 128 registers per thread
constexpr int kWarps = kThreads / 32;
constexpr int WX = 48, WROWS = 8, WBOXES = 2, WY = WROWS * WBOXES;
constexpr int kBoxBytes = WX * WROWS * 16;               // one TMA box: 8 rows of 48 cells
constexpr int kViewBytes = kBoxBytes * WBOXES;           // window buffer of one view: 48 x 16 cells
constexpr int kRowBytes = WX * 16;
constexpr int kMaxStages = 4, kMaxSrc = 7;
constexpr size_t kSmemBudget = 200 * 1024;

struct Meta { int wx0, wy0, rows, pad; };                // window origin (source pixels) and rows landed (0 / 8 / 16)

struct Params {
  alignas(64) CUtensorMap tmap;      // fp16 chunk-planar features as (4*Wf, Hf, 4*N) fp32 elements, box (4*WX, 8, 1)
  const float* feats;                // [N,Hf,Wf,32] fp32 (the reference view is read from here)
  const uint4* feats16;              // [N][4][Hf][Wf] cells of 8 halves
  const float* coef;                 // [(N-1), D, 8] pixel-coordinate transform rows
  __nv_bfloat16* cp8; __nv_bfloat16* ps8;
  int n_src, D, d0g, Dloc, Hf, Wf, order;
  int tiles_x, tiles_y, nwork, nstages;
  unsigned long long* stats;         // optional: [0] (voxel, view) pairs served from global memory
};

__global__ void planar_half_features_kernel(const float* __restrict__ feats, int n_views, int Hf, int Wf,
                                            uint4* __restrict__ out) {
  const size_t npix = (size_t)n_views * Hf * Wf, plane = (size_t)Hf * Wf;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const size_t v = i / plane, pix = i - v * plane;
    const float4* src = reinterpret_cast<const float4*>(feats + i * 32);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 a = __ldg(src + 2 * c), b = __ldg(src + 2 * c + 1);
      __half2 h[4] = {__floats2half2_rn(a.x, a.y), __floats2half2_rn(a.z, a.w), __floats2half2_rn(b.x, b.y),
                      __floats2half2_rn(b.z, b.w)};
      out[(v * 4 + c) * plane + pix] = *reinterpret_cast<const uint4*>(h);
    }
  }
}

__device__ __forceinline__ unsigned long long f2_bits(float2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 bits_f2(unsigned long long v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
  return bits_f2(d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(d);
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ __half2 as_h2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }

// NV = number of source views; BLEND32: the 4-tap blend in fp32 (taps still fp16-rounded) instead of packed fp16
template <int NV, bool BLEND32>
__global__ void __launch_bounds__(kThreads, 1) cost_volume_window_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int stage_bytes = NV * kViewBytes;
  unsigned char* s_ring = smem;
  Meta* s_meta = reinterpret_cast<Meta*>(smem + (size_t)p.nstages * stage_bytes);      // [2][NV]
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(s_meta + 2 * kMaxSrc);
  uint64_t* bar_empty = bar_full + kMaxStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxStages; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], kWarps); }
    fence_mbar_init();
  }
  __syncthreads();
  const int ns = p.nstages;
  const int nmine = p.nwork > (int)blockIdx.x ? (p.nwork - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  // ---- producer duty of warp 0 (it is a consumer like the others): stage g of this CTA = (work item g / 4, chunk g % 4);
  // at chunk 0 lanes < NV first work out the window of the item in "their" view and publish it in s_meta
  int m_wx0 = 0, m_wy0 = 0, m_rows = 0;        // lane v of warp 0: window of view v of the item being produced
  uint32_t m_bytes = 0;
  auto produce = [&](int g) {
    if (g >= 4 * nmine) return;
    const int it = g >> 2, c = g & 3;
    if (c == 0) {
      const int wi = (int)blockIdx.x + it * (int)gridDim.x;
      const int tx = wi % p.tiles_x, ty = (wi / p.tiles_x) % p.tiles_y, dc = wi / (p.tiles_x * p.tiles_y);
      m_wx0 = m_wy0 = m_rows = 0;
      if (lane < NV) {
        // bounding box of the work item's samples in source view `lane`: the 8 corners of (x, y, plane)
        const int xl = tx * TXP, xh = min(xl + TXP - 1, p.Wf - 1);
        const int yl = ty * TYP, yh = min(yl + TYP - 1, p.Hf - 1);
        const int ll = dc * PL, lh = min(ll + PL - 1, p.Dloc - 1);
        float mnx = 3.0e38f, mxx = -3.0e38f, mny = 3.0e38f, mxy = -3.0e38f;
        bool finite = true;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int d = min(max(p.d0g + ((k & 4) ? lh : ll), 0), p.D - 1);
          float ix, iy;
          transform_coords(p.coef + ((size_t)lane * p.D + d) * 8, (float)((k & 1) ? xh : xl), (float)((k & 2) ? yh : yl),
                           ix, iy);
          finite = finite && fabsf(ix) <= 1.0e9f && fabsf(iy) <= 1.0e9f;      // false for NaN, inf and absurd values
          mnx = fminf(mnx, ix); mxx = fmaxf(mxx, ix); mny = fminf(mny, iy); mxy = fmaxf(mxy, iy);
        }
        if (finite) {
          const float fx0 = floorf(mnx), fx1 = floorf(mxx) + 1.0f, fy0 = floorf(mny), fy1 = floorf(mxy) + 1.0f;
          if (fx1 >= 0.0f && fx0 <= (float)(p.Wf - 1) && fy1 >= 0.0f && fy0 <= (float)(p.Hf - 1)) {
            m_wx0 = (int)fmaxf(fx0, (float)-WX);
            m_wy0 = (int)fmaxf(fy0, (float)-WY);
            m_rows = ((int)fminf(fy1, (float)p.Hf) - m_wy0 + 1 <= WROWS) ? WROWS : WY;
          }
        }
        s_meta[(it & 1) * kMaxSrc + lane] = Meta{m_wx0, m_wy0, m_rows, 0};
      }
      // bytes of one stage of this work item (sum over the views, the same for its four chunks)
      m_bytes = (uint32_t)(m_rows * kRowBytes);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m_bytes += __shfl_xor_sync(0xffffffffu, m_bytes, o);
    }
    const int stage = g % ns;
    const uint32_t round = (uint32_t)(g / ns);                // how many times the ring has wrapped
    if (round > 0) mbar_wait(&bar_empty[stage], (round - 1) & 1u);
    if (lane == 0) {
      if (m_bytes) mbar_arrive_expect_tx(&bar_full[stage], m_bytes);    // (orders the s_meta stores before the wait of the readers)
      else mbar_arrive(&bar_full[stage]);
    }
    __syncwarp();
    // lanes 2v and 2v+1 issue the one or two boxes of view v
    const int v = lane >> 1, b = lane & 1;
    const int rows_v = __shfl_sync(0xffffffffu, m_rows, v), wx0_v = __shfl_sync(0xffffffffu, m_wx0, v),
              wy0_v = __shfl_sync(0xffffffffu, m_wy0, v);
    if (v < NV && b * WROWS < rows_v)
      tma_load_3d(s_ring + (size_t)stage * stage_bytes + (size_t)v * kViewBytes + (size_t)b * kBoxBytes, &p.tmap,
                  wx0_v * 4, wy0_v + b * WROWS, (v + 1) * 4 + c, &bar_full[stage]);
  };
  if (warp == 0)
    for (int g = 0; g < ns - 1; ++g) produce(g);

  // ===================================== consumers =====================================
  const int t = threadIdx.x;
  const int px = t & (TXP - 1), py = (t >> 5) & (TYP - 1), pl0 = t >> 7;      // planes pl0 and pl0 + 4 of the work item
  const float inv_n = 1.0f / (float)(NV + 1), inv_nn = 1.0f / (float)((NV + 1) * (NV + 1));
  const size_t plane_cells = (size_t)p.Hf * p.Wf;
  const int Hs = (p.Hf + 1) >> 1, Ws = (p.Wf + 1) >> 1;
  const uint32_t ring_u32 = smem_u32(s_ring);
  int stage = 0;
  uint32_t round = 0;
  for (int it = 0; it < nmine; ++it) {
    const int wi = (int)blockIdx.x + it * (int)gridDim.x;
    const int tx = wi % p.tiles_x, ty = (wi / p.tiles_x) % p.tiles_y, dc = wi / (p.tiles_x * p.tiles_y);
    const int x = tx * TXP + px, y = ty * TYP + py;
    const int xc = min(x, p.Wf - 1), yc = min(y, p.Hf - 1);
    // footprints of this thread's two voxels in every source view, in image coordinates: top-left tap (x0, y0)
    // packed as (y0 + 2) << 15 | (x0 + 2), and the four tap weights
    uint32_t fo[kIPT][NV];
    uint32_t wa[kIPT][NV], wb[kIPT][NV];       // fp16 blend: half2 (w00, w01), (w10, w11); fp32 blend: wxr, wyr bits
    bool live[kIPT];
#pragma unroll
    for (int i = 0; i < kIPT; ++i) {
      const int l = dc * PL + pl0 + i * (PL / kIPT);
      const int dg = p.d0g + l;
      live[i] = x < p.Wf && y < p.Hf && l < p.Dloc && (unsigned)dg < (unsigned)p.D;
      const int d = min(max(dg, 0), p.D - 1);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4* row = reinterpret_cast<const float4*>(p.coef + ((size_t)v * p.D + d) * 8);
        const float4 c0 = __ldg(row), c1 = __ldg(row + 1);
        const float tc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        float ix, iy;
        transform_coords(tc, (float)xc, (float)yc, ix, iy);
        const Footprint f = make_footprint(ix, iy, p.Wf, p.Hf);
        fo[i][v] = (uint32_t)((f.y0 + 2) << 15) | (uint32_t)(f.x0 + 2);
        if (BLEND32) {
          wa[i][v] = __float_as_uint(f.wxr); wb[i][v] = __float_as_uint(f.wyr);
          if (f.wxl == 0.0f && f.wxr == 0.0f) { wa[i][v] = 0x7fc00000u; }      // non-finite sample: marks "all weights 0"
        } else {
          const __half2 h0 = __floats2half2_rn(f.wyl * f.wxl, f.wyl * f.wxr), h1 = __floats2half2_rn(f.wyr * f.wxl, f.wyr * f.wxr);
          wa[i][v] = *reinterpret_cast<const uint32_t*>(&h0); wb[i][v] = *reinterpret_cast<const uint32_t*>(&h1);
        }
      }
    }
    for (int c = 0; c < 4; ++c) {
      // reference view: S = r, Q = r^2 (model.py:436-437)
      float2 S[kIPT][4], Q[kIPT][4];
      {
        const float4* rp = reinterpret_cast<const float4*>(p.feats + ((size_t)yc * p.Wf + xc) * 32 + c * 8);
        const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
#pragma unroll
        for (int i = 0; i < kIPT; ++i) {
          S[i][0] = make_float2(r0.x, r0.y); S[i][1] = make_float2(r0.z, r0.w);
          S[i][2] = make_float2(r1.x, r1.y); S[i][3] = make_float2(r1.z, r1.w);
#pragma unroll
          for (int k = 0; k < 4; ++k) Q[i][k] = make_float2(S[i][k].x * S[i][k].x, S[i][k].y * S[i][k].y);
        }
      }
      if (warp == 0) produce(4 * it + c + ns - 1);       // keep ns - 1 stages in flight
      mbar_wait(&bar_full[stage], round & 1u);
      if (c == 0) {
        // window-relative tap offsets: bit 31 clear = byte offset of the top-left cell inside the view's buffer
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const Meta m = s_meta[(it & 1) * kMaxSrc + v];
#pragma unroll
          for (int i = 0; i < kIPT; ++i) {
            const int x0 = (int)(fo[i][v] & 0x7fffu) - 2, y0 = (int)(fo[i][v] >> 15) - 2;
            const int cx = x0 - m.wx0, cy = y0 - m.wy0;
            const unsigned ry = m.rows > 0 ? (unsigned)(m.rows - 1) : 0u;       // both tap rows must have landed
            if ((unsigned)cx < (unsigned)(WX - 1) && (unsigned)cy < ry) fo[i][v] = (uint32_t)((cy * WX + cx) * 16);
            else fo[i][v] |= 0x80000000u;
          }
        }
      }
      const uint32_t sbase = ring_u32 + (uint32_t)(stage * stage_bytes);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
#pragma unroll
        for (int i = 0; i < kIPT; ++i) {
          uint4 ta, tb, tc4, td;
          const uint32_t o = fo[i][v];
          if (!(o & 0x80000000u)) {
            const uint32_t a = sbase + (uint32_t)(v * kViewBytes) + o;
            ta = lds128(a); tb = lds128(a + 16); tc4 = lds128(a + kRowBytes); td = lds128(a + kRowBytes + 16);
          } else {
            // footprint outside the staged window: the same taps from global memory, zero outside the image
            const int x0 = (int)(o & 0x7fffu) - 2, y0 = (int)((o >> 15) & 0xffffu) - 2;
            const uint4* img = p.feats16 + ((size_t)(v + 1) * 4 + c) * plane_cells;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            const bool vx0 = (unsigned)x0 < (unsigned)p.Wf, vx1 = (unsigned)(x0 + 1) < (unsigned)p.Wf;
            const bool vy0 = (unsigned)y0 < (unsigned)p.Hf, vy1 = (unsigned)(y0 + 1) < (unsigned)p.Hf;
            ta = (vy0 && vx0) ? __ldg(img + (size_t)y0 * p.Wf + x0) : z;
            tb = (vy0 && vx1) ? __ldg(img + (size_t)y0 * p.Wf + x0 + 1) : z;
            tc4 = (vy1 && vx0) ? __ldg(img + (size_t)(y0 + 1) * p.Wf + x0) : z;
            td = (vy1 && vx1) ? __ldg(img + (size_t)(y0 + 1) * p.Wf + x0 + 1) : z;
            if (p.stats && c == 0) atomicAdd(p.stats, 1ull);
          }
          const uint32_t* A = reinterpret_cast<const uint32_t*>(&ta);
          const uint32_t* B = reinterpret_cast<const uint32_t*>(&tb);
          const uint32_t* C = reinterpret_cast<const uint32_t*>(&tc4);
          const uint32_t* Dd = reinterpret_cast<const uint32_t*>(&td);
          if (BLEND32) {
            float wxr = __uint_as_float(wa[i][v]), wyr = __uint_as_float(wb[i][v]);
            float wxl = 1.0f - wxr, wyl = 1.0f - wyr;
            if (wa[i][v] == 0x7fc00000u) { wxl = wxr = wyl = wyr = 0.0f; }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 a = __half22float2(as_h2(A[k])), b = __half22float2(as_h2(B[k]));
              const float2 cc = __half22float2(as_h2(C[k])), d = __half22float2(as_h2(Dd[k]));
              // the reference's association: wyl * (wxl*p00 + wxr*p01) + wyr * (wxl*p10 + wxr*p11)  (Appendix A.3)
              float2 w;
              w.x = wyl * (wxl * a.x + wxr * b.x) + wyr * (wxl * cc.x + wxr * d.x);
              w.y = wyl * (wxl * a.y + wxr * b.y) + wyr * (wxl * cc.y + wxr * d.y);
              S[i][k] = fadd2(S[i][k], w);
              Q[i][k] = ffma2(w, w, Q[i][k]);
            }
          } else {
            const __half2 h0 = as_h2(wa[i][v]), h1 = as_h2(wb[i][v]);
            const __half2 w00 = __low2half2(h0), w01 = __high2half2(h0), w10 = __low2half2(h1), w11 = __high2half2(h1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const __half2 h = __hfma2(w11, as_h2(Dd[k]), __hfma2(w10, as_h2(C[k]), __hfma2(w01, as_h2(B[k]), __hmul2(w00, as_h2(A[k])))));
              const float2 w = __half22float2(h);
              S[i][k] = fadd2(S[i][k], w);
              Q[i][k] = ffma2(w, w, Q[i][k]);
            }
          }
        }
      }
      // this warp is done with the stage: hand the buffer back to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_empty[stage]);
      if (++stage == ns) { stage = 0; ++round; }
      // variance with reciprocal multiplies (<= 1 ulp from the reference's divisions, model.py:458-461 / :330-332)
#pragma unroll
      for (int i = 0; i < kIPT; ++i) {
        if (!live[i]) continue;
        float cst[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (p.order == MVSB200_ORDER_MEM) {
            cst[2 * k] = Q[i][k].x * inv_n - (S[i][k].x * S[i][k].x) * inv_nn;
            cst[2 * k + 1] = Q[i][k].y * inv_n - (S[i][k].y * S[i][k].y) * inv_nn;
          } else {
            const float mx = S[i][k].x * inv_n, my = S[i][k].y * inv_n;
            cst[2 * k] = Q[i][k].x * inv_n - mx * mx;
            cst[2 * k + 1] = Q[i][k].y * inv_n - my * my;
          }
        }
        uint4 cell;
        {
          __nv_bfloat162 b0 = __floats2bfloat162_rn(cst[0], cst[1]), b1 = __floats2bfloat162_rn(cst[2], cst[3]);
          __nv_bfloat162 b2 = __floats2bfloat162_rn(cst[4], cst[5]), b3 = __floats2bfloat162_rn(cst[6], cst[7]);
          cell.x = *reinterpret_cast<uint32_t*>(&b0); cell.y = *reinterpret_cast<uint32_t*>(&b1);
          cell.z = *reinterpret_cast<uint32_t*>(&b2); cell.w = *reinterpret_cast<uint32_t*>(&b3);
        }
        const int l = dc * PL + pl0 + i * (PL / kIPT);
        const size_t zc = (size_t)l * 4 + c;
        if (p.cp8) *reinterpret_cast<uint4*>(p.cp8 + ((zc * p.Hf + y) * p.Wf + x) * 8) = cell;
        if (p.ps8)
          *reinterpret_cast<uint4*>(p.ps8 + (((zc * 4 + (y & 1) * 2 + (x & 1)) * Hs + (y >> 1)) * Ws + (x >> 1)) * 8) = cell;
      }
    }
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaDriverEntryPointQueryResult qres;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)f;
  });
  return fn;
}

using Kernel = void (*)(const Params);
template <bool B32>
static Kernel pick(int nv) {
  switch (nv) {
    case 1: return cost_volume_window_kernel<1, B32>;
    case 2: return cost_volume_window_kernel<2, B32>;
    case 3: return cost_volume_window_kernel<3, B32>;
    case 4: return cost_volume_window_kernel<4, B32>;
    case 5: return cost_volume_window_kernel<5, B32>;
    case 6: return cost_volume_window_kernel<6, B32>;
    default: return cost_volume_window_kernel<7, B32>;
  }
}

}  // namespace cvw

// development counter (tuning CV_STATS): (voxel, view) pairs the window kernel served from global memory
static unsigned long long* window_stats_buffer() {
  static unsigned long long* buf = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    if (cudaMalloc(&buf, sizeof(unsigned long long)) == cudaSuccess) cudaMemset(buf, 0, sizeof(unsigned long long));
    else buf = nullptr;
  });
  return buf;
}

size_t cost_volume_window_scratch_bytes(int n_views, int hf, int wf) { return (size_t)n_views * hf * wf * 64; }

bool cost_volume_window_ok(int n_views, int hf, int wf, int channels, int sampler) {
  return sampler == MVSB200_SAMPLER_TRANSFORM && channels == 32 && n_views >= 2 && n_views - 1 <= cvw::kMaxSrc &&
         hf < 32000 && wf < 32000 && cvw::get_encode() != nullptr;
}

// feats16: scratch of cost_volume_window_scratch_bytes() for the fp16 chunk-planar copy of the features.
// Local planes [0, dloc) = global planes [d0g, d0g + dloc) of the depth_num-plane sweep (D-slab mode; whole volume:
// d0g = 0, dloc = depth_num); planes whose global index falls outside [0, depth_num) are left alone.
int launch_cost_volume_window(const float* feats, const float* coef_table, int n_views, int depth_num, int d0g, int dloc,
                              int hf, int wf, int order, void* cp8, void* ps8, void* feats16, int blend32,
                              unsigned long long* stats, cudaStream_t s) {
  using namespace cvw;
  MVS_CHECK_ARG(feats && coef_table && feats16 && (cp8 || ps8), "cost_volume(window): NULL pointer");
  MVS_CHECK_ARG(cost_volume_window_ok(n_views, hf, wf, 32, MVSB200_SAMPLER_TRANSFORM),
                "cost_volume(window): needs 2..8 views, 32 channels, Hf, Wf < 32000 and a driver with TMA descriptors");
  const int nv = n_views - 1;
  const int sm = sm_count_current();
  {
    const size_t npix = (size_t)n_views * hf * wf;
    const size_t want = (npix + 255) / 256, cap = (size_t)sm * 16;
    planar_half_features_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(feats, n_views, hf, wf, (uint4*)feats16);
    MVS_LAUNCH_CHECK("planar_half_features_kernel");
  }
  if (ps8 && ((hf | wf) & 1)) MVS_CUDA(cudaMemsetAsync(ps8, 0, (size_t)dloc * 16 * ((hf + 1) / 2) * ((wf + 1) / 2) * 16, s));
  Params p;
  memset(&p, 0, sizeof(p));
  {
    cuuint64_t gdim[3] = {(cuuint64_t)wf * 4, (cuuint64_t)hf, (cuuint64_t)n_views * 4};
    cuuint64_t gstr[2] = {(cuuint64_t)wf * 16, (cuuint64_t)hf * wf * 16};
    cuuint32_t box[3] = {(cuuint32_t)WX * 4, (cuuint32_t)WROWS, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = get_encode()(&p.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, feats16, gdim, gstr, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cost_volume(window): cuTensorMapEncodeTiled failed (%d) for Hf=%d Wf=%d N=%d", (int)r, hf, wf, n_views);
      return MVSB200_ERR_CUDA;
    }
  }
  p.feats = feats; p.feats16 = (const uint4*)feats16; p.coef = coef_table;
  p.cp8 = (__nv_bfloat16*)cp8; p.ps8 = (__nv_bfloat16*)ps8;
  p.n_src = nv; p.D = depth_num; p.d0g = d0g; p.Dloc = dloc; p.Hf = hf; p.Wf = wf; p.order = order;
  p.tiles_x = ceil_div(wf, TXP); p.tiles_y = ceil_div(hf, TYP);
  p.nwork = p.tiles_x * p.tiles_y * ceil_div(dloc, PL);
  const size_t stage_bytes = (size_t)nv * kViewBytes;
  int nstages = (int)(kSmemBudget / stage_bytes);
  p.nstages = nstages > kMaxStages ? kMaxStages : nstages;
  MVS_CHECK_ARG(p.nstages >= 2, "cost_volume(window): shared-memory ring too small");
  p.stats = stats ? stats : (tuning().cv_stats ? window_stats_buffer() : nullptr);
  const size_t smem = (size_t)p.nstages * stage_bytes + 2 * kMaxSrc * sizeof(Meta) + 2 * kMaxStages * sizeof(uint64_t);
  Kernel k = blend32 ? pick<true>(nv) : pick<false>(nv);
  // the attribute is per device and per function: set it every time (cheap, and a set value is only re-set to itself)
  MVS_CUDA(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = p.nwork < sm ? p.nwork : sm;
  k<<<grid, kThreads, smem, s>>>(p);
  MVS_LAUNCH_CHECK("cost_volume_window_kernel");
  return MVSB200_OK;
}

}  // namespace mvsb200

// Development: (voxel, view) pairs the window kernel has served from global memory since the last reset (counted only
// while the tuning switch CV_STATS is on; single device).  Synchronises the device.
extern "C" int mvsb200_cost_volume_window_stats(unsigned long long* slow_pairs, int reset) {
  using namespace mvsb200;
  MVS_CHECK_ARG(slow_pairs != nullptr, "cost_volume_window_stats: NULL pointer");
  unsigned long long* buf = window_stats_buffer();
  MVS_CHECK_ARG(buf != nullptr, "cost_volume_window_stats: no counter");
  MVS_CUDA(cudaDeviceSynchronize());
  MVS_CUDA(cudaMemcpy(slow_pairs, buf, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) MVS_CUDA(cudaMemset(buf, 0, sizeof(unsigned long long)));
  return MVSB200_OK;
}
