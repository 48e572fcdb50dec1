// Kernel 2, product mode: fused homography warp + N-view variance with the source views staged in shared
// memory by TMA.  Replaces the D x (N-1) tf_transform_homography launches (homography_warping.py:211-253) and
// the running-sum ops of model.py:423-463 (inference_mem) / model.py:315-334 (inference); the warped volume is
// never written.
//
// Work item = 32 x 3 reference pixels x 10 consecutive depth planes (960 voxels).  Consecutive planes of a pixel
// sample a source view a fraction of a pixel apart, so the taps of a whole work item fall into
// one small window of each source view: the bounding box of the 8 corner samples (the sample position is projective
// in (x, y) and a Moebius function of the depth, hence monotone along every edge of the box).  A producer warp
// computes that box per (work item, view) and has the TMA unit copy it from an fp16 chunk-planar copy of the
// features ([N][4][Hf][Wf] cells of 8 channels = 16 bytes) into a shared-memory ring -- cells outside the image are
// zero-filled by the TMA unit, which IS the reference's zero fill -- one 8-channel chunk at a time (stage = work
// item x chunk; 48 x 16 cells per view, loaded as one or two boxes of 8 rows); it also stages the item's transform
// rows.  15 consumer warps own two voxels per lane: the bilinear footprint of a voxel in every view (one reciprocal,
// weights) is computed once per work item and kept in registers (3 words per view), then every stage costs 4
// conflict-free 16-byte shared-memory loads and 16 packed-half FMAs per (voxel, view); running sum and squared sum
// stay in fp32 registers across the views; the variance goes out as whole 16-byte cells of the regularizer's two
// planar layouts (lanes run along x: 512 contiguous bytes per warp, no staging).  A voxel whose footprint is not
// inside the staged window (wild geometry, window larger than the buffer) reads its taps from global memory instead:
// correctness never depends on the window.
#include "geometry.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <mutex>
#include <string.h>
#include <type_traits>

namespace mvsb200 {
using namespace umma;

namespace cvw {

constexpr int TXP = 32, TYP = 3, PL = 10;                // reference pixels x planes of a work item
constexpr int kItems = TXP * TYP * PL;                   // 960 voxels = 15 consumer warps x 2 voxels per lane
// 2 voxels per consumer thread: 15 consumer warps + the producer warp = 16 warps = 4 per scheduler at <= 128 registers
// (measured at config 2: 0.71 ms; 1 voxel per thread = 30 + 1 warps at <= 64 registers: 0.79 ms)
constexpr int kIPT = 2;
constexpr int kThreads = kItems / kIPT + 32;
constexpr int WX = 48, WROWS = 8, WBOXES = 2, WY = WROWS * WBOXES;
constexpr int kBoxBytes = WX * WROWS * 16;               // one TMA box: 8 rows of 48 cells
constexpr int kViewBytes = kBoxBytes * WBOXES;           // window buffer of one view: 48 x 16 cells
constexpr int kRowBytes = WX * 16;
constexpr int kMaxStages = 4, kMaxSrc = 7;
constexpr size_t kSmemBudget = 200 * 1024;

struct Meta { int wx0, wy0, rows, pad; };                // window origin (source pixels) and rows landed (8 / 16)
struct Item { int x0, y0, l0, pad; };                    // first reference pixel and first local plane of a work item

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

__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(d);
}
__device__ __forceinline__ float rcp_approx(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// Sample position of reference pixel (x, y) for product mode: the arithmetic of transform_coords (geometry.cuh) with
// ONE approximate reciprocal instead of two IEEE divisions (<= 2 ulp of the coordinate, ~1e-5 px: far below the fp16
// rounding of the taps this mode reads; the fp32 parity kernels keep the exact divisions)
__device__ __forceinline__ void fast_coords(const float4 c0, const float4 c1, float x, float y, float& ix, float& iy) {
  const float rp = rcp_approx(fmaf(c1.z, x, fmaf(c1.w, y, 1.0f)));
  ix = fmaf(c0.x, x, fmaf(c0.y, y, c0.z)) * rp;
  iy = fmaf(c0.w, x, fmaf(c1.x, y, c1.y)) * rp;
}

// NV = number of source views.  MODE selects the arithmetic of the blend and of the running sums:
//   0  4-tap blend in packed fp16 on the DIFFERENCE to the reference pixel (d_v = warped_v - r; the variance does not move
//      with a shift, and differences keep Q/n - (S/n)^2 away from cancellation), sums of d and d^2 in fp32;
//   1  blend in fp32 in the reference's association, sums of the warped values and their squares in fp32 (taps fp16);
//   2  as 0 with the sums of d and d^2 kept in packed fp16 as well (converted once per chunk).
template <int NV, int MODE>
__global__ void __launch_bounds__(kThreads, 1) cost_volume_window_kernel(const __grid_constant__ Params p) {
  constexpr bool BLEND32 = MODE == 1, ACC16 = MODE == 2;
  constexpr int IPT = kIPT, kConsumerWarps = kItems / IPT / 32;
  extern __shared__ __align__(128) unsigned char smem[];
  const int stage_bytes = NV * kViewBytes;
  unsigned char* s_ring = smem;
  Meta* s_meta = reinterpret_cast<Meta*>(smem + (size_t)p.nstages * stage_bytes);      // [2][kMaxSrc]
  float4* s_coef = reinterpret_cast<float4*>(s_meta + 2 * kMaxSrc);                     // [2][kMaxSrc][PL][2]: transform rows
  Item* s_item = reinterpret_cast<Item*>(s_coef + 2 * kMaxSrc * PL * 2);                // [2]
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(s_item + 2);
  uint64_t* bar_empty = bar_full + kMaxStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxStages; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], kConsumerWarps); }
    fence_mbar_init();
  }
  __syncthreads();
  const int ns = p.nstages;
  const int nmine = p.nwork > (int)blockIdx.x ? (p.nwork - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  // ---- producer warp: stage g of this CTA = (work item g / 4, chunk g % 4); at chunk 0 lanes < NV first work out the
  // window of the item in "their" view and publish it in s_meta.  (A consumer warp that also produced would be the
  // slowest warp and pace all the others: measured, 9 of 16 warps spinning on the full barrier.)
  if (warp == kConsumerWarps) {
    // window of work item `it` in view `lane` (lanes < NV) + the item's transform rows -> s_meta / s_coef [it & 1]
    int n_wx0 = 0, n_wy0 = 0, n_rows = 0;       // ... of the item prepared last
    auto prepare = [&](int it) {
      n_wx0 = n_wy0 = n_rows = 0;
      if (it >= nmine) return;
      const int wi = (int)blockIdx.x + it * (int)gridDim.x;
      const int tx = wi % p.tiles_x, ty = (wi / p.tiles_x) % p.tiles_y, dc = wi / (p.tiles_x * p.tiles_y);
      if (lane == 31) s_item[it & 1] = Item{tx * TXP, ty * TYP, dc * PL, 0};      // (the consumers skip the divisions)
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
          const float4* row = reinterpret_cast<const float4*>(p.coef + ((size_t)lane * p.D + d) * 8);
          float ix, iy;
          fast_coords(__ldg(row), __ldg(row + 1), (float)((k & 1) ? xh : xl), (float)((k & 2) ? yh : yl), ix, iy);
          finite = finite && fabsf(ix) <= 1.0e9f && fabsf(iy) <= 1.0e9f;      // false for NaN, inf and absurd values
          mnx = fminf(mnx, ix); mxx = fmaxf(mxx, ix); mny = fminf(mny, iy); mxy = fmaxf(mxy, iy);
        }
        // A window always lands (8 rows at least, at a position clamped to the image with its one-cell zero border):
        // a view none of whose taps touches the image gets zero weights on landed, finite cells, exactly like single
        // footprints outside the image -- the consumers carry no "skip this view" branch.  Rows above / left of the image
        // hold zeros only: the window starts at -1 at the earliest.
        n_rows = WROWS;
        if (finite) {
          n_wx0 = (int)fminf(fmaxf(floorf(mnx), -1.0f), (float)(p.Wf - 1));
          n_wy0 = (int)fminf(fmaxf(floorf(mny), -1.0f), (float)(p.Hf - 1));
          if ((int)fminf(floorf(mxy) + 1.0f, (float)p.Hf) - n_wy0 + 1 > WROWS) n_rows = WY;
        }
        s_meta[(it & 1) * kMaxSrc + lane] = Meta{n_wx0, n_wy0, n_rows, 0};
      }
      // the transform rows of the item's planes for every view: the consumers read them from shared memory
      for (int i = lane; i < NV * PL * 2; i += 32) {
        const int v = i / (PL * 2), r = i - v * (PL * 2);
        const int d = min(max(p.d0g + dc * PL + (r >> 1), 0), p.D - 1);
        s_coef[((it & 1) * kMaxSrc + v) * (PL * 2) + r] =
            __ldg(reinterpret_cast<const float4*>(p.coef + ((size_t)v * p.D + d) * 8) + (r & 1));
      }
      __syncwarp();        // every lane's stores precede lane 0's arrive on the item's first full barrier
    };
    prepare(0);
    for (int it = 0; it < nmine; ++it) {
      const int m_wx0 = n_wx0, m_wy0 = n_wy0, m_rows = n_rows;
      // bytes of one stage of this work item (sum over the views, the same for its four chunks)
      uint32_t m_bytes = (uint32_t)(m_rows * kRowBytes);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m_bytes += __shfl_xor_sync(0xffffffffu, m_bytes, o);
      // lanes 2v and 2v+1 issue the one or two boxes of view v
      const int v = lane >> 1, b = lane & 1;
      const int rows_v = __shfl_sync(0xffffffffu, m_rows, v), wx0_v = __shfl_sync(0xffffffffu, m_wx0, v),
                wy0_v = __shfl_sync(0xffffffffu, m_wy0, v);
      for (int c = 0; c < 4; ++c) {
        const int g = 4 * it + c;
        const int stage = g % ns;
        const uint32_t round = (uint32_t)(g / ns);              // how many times the ring has wrapped
        if (round > 0) mbar_wait(&bar_empty[stage], (round - 1) & 1u);
        if (lane == 0) mbar_arrive_expect_tx(&bar_full[stage], m_bytes);   // (orders the s_meta / s_coef stores before the readers' wait)
        __syncwarp();
        if (v < NV && b * WROWS < rows_v)
          tma_load_3d(s_ring + (size_t)stage * stage_bytes + (size_t)v * kViewBytes + (size_t)b * kBoxBytes, &p.tmap,
                      wx0_v * 4, wy0_v + b * WROWS, (v + 1) * 4 + c, &bar_full[stage]);
        // the next item's window while this item's boxes are in flight.  Its slot [(it + 1) & 1] was last read at chunk 0
        // of item it - 1, which every consumer has left: the stage just armed reuses a buffer they have all released
        // after it (ns <= 4 stages per item)
        if (c == 0) prepare(it + 1);
      }
    }
    return;
  }

  // ===================================== consumers =====================================
  const int t = threadIdx.x;
  const int px = t & (TXP - 1), py = (t >> 5) % TYP, pl0 = (t >> 5) / TYP;     // planes pl0 and pl0 + PL / 2 of the item
  const float inv_n = 1.0f / (float)(NV + 1), inv_nn = 1.0f / (float)((NV + 1) * (NV + 1));
  const float2 inv_n2 = make_float2(inv_n, inv_n), ninv_nn2 = make_float2(-inv_nn, -inv_nn);
  const uint32_t plane_cells = (uint32_t)(p.Hf * p.Wf);
  const int Hs = (p.Hf + 1) >> 1, Ws = (p.Wf + 1) >> 1;
  const uint32_t ps8_chunk = (uint32_t)(4 * Hs * Ws);
  const uint32_t ring_u32 = smem_u32(s_ring);
  int stage = 0;
  uint32_t round = 0;
  for (int it = 0; it < nmine; ++it) {
    uint32_t fo[IPT][NV];              // per (voxel, view): byte offset of the top-left tap in the view's window, or
                                       // bit 31 | (y0 + 2) << 15 | (x0 + 2) when the footprint is outside the window
    uint32_t wa[IPT][NV], wb[IPT][NV]; // fp16 blend: half2 (w00, w01), (w10, w11); fp32 blend: wxr, wyr bits
    bool live[IPT];
    uint32_t cell_cp8[IPT], cell_ps8[IPT];       // chunk-0 cell index of the voxel in the two output layouts
    bool slow_warp = false;            // some voxel of this warp reads a view from global memory

    // ---- chunk 0's stage carries the item's window and transform rows (published before the barrier was armed):
    // bilinear footprint of every (voxel, view), once per work item.  Sample position as in transform_coords
    // (geometry.cuh), clamped to [-2, W] x [-2, H]: beyond that every tap is outside the image anyway (and NaN / inf
    // land on the bound: "outside", as the reference's zero fill has it).  A footprint whose four taps sit inside
    // the landed rows of the window becomes a shared-memory offset.
    mbar_wait(&bar_full[stage], round & 1u);
    const Item item = s_item[it & 1];
    const int x = item.x0 + px, y = item.y0 + py;
    const int xc = min(x, p.Wf - 1), yc = min(y, p.Hf - 1);
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const int l = item.l0 + pl0 + i * (PL / IPT);
      live[i] = x < p.Wf && y < p.Hf && l < p.Dloc && (unsigned)(p.d0g + l) < (unsigned)p.D;
      cell_cp8[i] = ((uint32_t)l * 4u * (uint32_t)p.Hf + (uint32_t)y) * (uint32_t)p.Wf + (uint32_t)x;       // (< 2^31: launch check)
      cell_ps8[i] = (((uint32_t)l * 16u + (uint32_t)((y & 1) * 2 + (x & 1))) * (uint32_t)Hs + (uint32_t)(y >> 1)) * (uint32_t)Ws +
                    (uint32_t)(x >> 1);
    }
    // reference cell of chunk 0
    const uint4* rptr = p.feats16 + (size_t)yc * p.Wf + xc;
    uint4 rcell = __ldg(rptr);
    {
      bool slow = false;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const Meta m = s_meta[(it & 1) * kMaxSrc + v];
        const unsigned ry = (unsigned)(m.rows - 1);       // both tap rows must have landed
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
          const float4* row = s_coef + ((it & 1) * kMaxSrc + v) * (PL * 2) + (pl0 + i * (PL / IPT)) * 2;
          float ix, iy;
          fast_coords(row[0], row[1], (float)xc, (float)yc, ix, iy);
          ix = fminf(fmaxf(ix, -2.0f), (float)p.Wf);
          iy = fminf(fmaxf(iy, -2.0f), (float)p.Hf);
          const float xf = floorf(ix), yf = floorf(iy);
          const float wxr = ix - xf, wyr = iy - yf;
          const int x0 = (int)xf, y0 = (int)yf;
          if (BLEND32) {
            wa[i][v] = __float_as_uint(wxr); wb[i][v] = __float_as_uint(wyr);
          } else {
            const float wxl = 1.0f - wxr, wyl = 1.0f - wyr;
            const __half2 h0 = __floats2half2_rn(wyl * wxl, wyl * wxr), h1 = __floats2half2_rn(wyr * wxl, wyr * wxr);
            wa[i][v] = *reinterpret_cast<const uint32_t*>(&h0); wb[i][v] = *reinterpret_cast<const uint32_t*>(&h1);
          }
          const int cx = x0 - m.wx0, cy = y0 - m.wy0;
          if ((unsigned)cx < (unsigned)(WX - 1) && (unsigned)cy < ry) {
            fo[i][v] = (uint32_t)((cy * WX + cx) * 16);
          } else if (x0 + 1 < 0 || x0 >= p.Wf || y0 + 1 < 0 || y0 >= p.Hf) {
            // all four taps outside the image: zero weights on the first cells of the window (landed, finite)
            fo[i][v] = 0u;
            wa[i][v] = BLEND32 ? 0x7fc00000u : 0u; wb[i][v] = 0u;
          } else {
            fo[i][v] = 0x80000000u | (uint32_t)((y0 + 2) << 15) | (uint32_t)(x0 + 2);
            slow = true;
          }
        }
      }
      slow_warp = __any_sync(0xffffffffu, slow);
    }

    // ---- the four chunks of the item; instantiated twice (with / without the global-memory path) so that the
    // common case carries neither its branches nor its registers
    auto chunks = [&](auto check_tag) {
      constexpr bool CHECK = decltype(check_tag)::value;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        // the reference pixel's cell of this chunk, read like the source views from the fp16 copy (one coalesced
        // 16-byte load per thread, issued one chunk ahead; an fp32 NHWC pixel would cost a 128-byte line per lane)
        const uint4 rc = rcell;
        if (c < 3) rcell = __ldg(rptr + (size_t)(c + 1) * plane_cells);
        const uint32_t* R = reinterpret_cast<const uint32_t*>(&rc);
        float2 S[IPT][4], Q[IPT][4];           // fp32 sums (MODE 0, 1)
        __half2 S2[IPT][4], Q2[IPT][4];        // packed fp16 sums (MODE 2)
        if (BLEND32) {
          // S = r, Q = r^2 (model.py:436-437)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(as_h2(R[k]));
#pragma unroll
            for (int i = 0; i < IPT; ++i) { S[i][k] = f; Q[i][k] = fmul2(f, f); }
          }
        }
        if (c > 0) mbar_wait(&bar_full[stage], round & 1u);
        const uint32_t sbase = ring_u32 + (uint32_t)(stage * stage_bytes);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
#pragma unroll
          for (int i = 0; i < IPT; ++i) {
            uint4 ta, tb, tc4, td;
            const uint32_t o = fo[i][v];
            if (!CHECK || !(o & 0x80000000u)) {
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
              if (p.stats && c == 0 && (vx0 || vx1) && (vy0 || vy1)) atomicAdd(p.stats, 1ull);
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
                // d = warped - r: the chain starts from -r (sign bits flipped)
                const __half2 h = __hfma2(w11, as_h2(Dd[k]), __hfma2(w10, as_h2(C[k]), __hfma2(w01, as_h2(B[k]),
                                  __hfma2(w00, as_h2(A[k]), as_h2(R[k] ^ 0x80008000u)))));
                if (ACC16) {
                  if (v == 0) { S2[i][k] = h; Q2[i][k] = __hmul2(h, h); }
                  else { S2[i][k] = __hadd2(S2[i][k], h); Q2[i][k] = __hfma2(h, h, Q2[i][k]); }
                } else {
                  const float2 w = __half22float2(h);
                  if (v == 0) { S[i][k] = w; Q[i][k] = fmul2(w, w); }
                  else { S[i][k] = fadd2(S[i][k], w); Q[i][k] = ffma2(w, w, Q[i][k]); }
                }
              }
            }
          }
        }
        // this warp is done with the stage: hand the buffer back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[stage]);
        if (++stage == ns) { stage = 0; ++round; }
        // variance with reciprocal multiplies, two channels per instruction (within an ulp or two of the reference's
        // divisions, model.py:458-461 / :330-332; this mode stores bf16).  On differences both op orders of the
        // reference reduce to Q/n - (S/n)^2.
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
          uint4 cell;
          uint32_t* cw = reinterpret_cast<uint32_t*>(&cell);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float2 s, q;
            if (ACC16) { s = __half22float2(S2[i][k]); q = __half22float2(Q2[i][k]); }
            else { s = S[i][k]; q = Q[i][k]; }
            float2 cst;
            if (!BLEND32 || p.order == MVSB200_ORDER_MEM) {
              cst = ffma2(q, inv_n2, fmul2(fmul2(s, s), ninv_nn2));
            } else {
              const float2 m = fmul2(s, inv_n2);
              cst = ffma2(q, inv_n2, fmul2(fmul2(m, m), make_float2(-1.0f, -1.0f)));
            }
            const __nv_bfloat162 b = __floats2bfloat162_rn(cst.x, cst.y);
            cw[k] = *reinterpret_cast<const uint32_t*>(&b);
          }
          if (live[i]) {
            if (p.cp8) *reinterpret_cast<uint4*>(p.cp8 + (size_t)(cell_cp8[i] + (uint32_t)c * plane_cells) * 8) = cell;
            if (p.ps8) *reinterpret_cast<uint4*>(p.ps8 + (size_t)(cell_ps8[i] + (uint32_t)c * ps8_chunk) * 8) = cell;
          }
        }
      }
    };
    if (slow_warp) chunks(std::true_type{});
    else chunks(std::false_type{});
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
template <int B32>
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
// the window kernel's source: the feature maps as fp16, 8-channel groups planar per view (what its TMA boxes read)
int launch_planar_half_features(const float* feats, int n_views, int hf, int wf, void* feats16, cudaStream_t s) {
  using namespace cvw;
  MVS_CHECK_ARG(feats && feats16, "cost_volume(window): NULL pointer");
  const size_t npix = (size_t)n_views * hf * wf;
  const size_t want = (npix + 255) / 256, cap = (size_t)sm_count_current() * 16;
  planar_half_features_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(feats, n_views, hf, wf, (uint4*)feats16);
  MVS_LAUNCH_CHECK("planar_half_features_kernel");
  return MVSB200_OK;
}

// feats16_ready: launch_planar_half_features() has already run for these feature maps (ordered before `s` gets here)
int launch_cost_volume_window(const float* feats, const float* coef_table, int n_views, int depth_num, int d0g, int dloc,
                              int hf, int wf, int order, void* cp8, void* ps8, void* feats16, int blend32,
                              unsigned long long* stats, cudaStream_t s, bool feats16_ready) {
  using namespace cvw;
  MVS_CHECK_ARG(feats && coef_table && feats16 && (cp8 || ps8), "cost_volume(window): NULL pointer");
  MVS_CHECK_ARG(cost_volume_window_ok(n_views, hf, wf, 32, MVSB200_SAMPLER_TRANSFORM),
                "cost_volume(window): needs 2..8 views, 32 channels, Hf, Wf < 32000 and a driver with TMA descriptors");
  const int nv = n_views - 1;
  const int sm = sm_count_current();
  if (!feats16_ready) {
    const int rc = launch_planar_half_features(feats, n_views, hf, wf, feats16, s);
    if (rc) return rc;
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
  const size_t smem = (size_t)p.nstages * stage_bytes + 2 * kMaxSrc * sizeof(Meta) + 2 * kMaxSrc * PL * 2 * sizeof(float4) +
                      2 * sizeof(Item) + 2 * kMaxStages * sizeof(uint64_t);
  MVS_CHECK_ARG((size_t)dloc * 16 * ((hf + 1) / 2) * ((wf + 1) / 2) < ((size_t)1 << 31) && (size_t)dloc * 4 * hf * wf < ((size_t)1 << 31),
                "cost_volume(window): volume too large for 32-bit cell indices");
  Kernel k = blend32 == 1 ? pick<1>(nv) : (blend32 == 2 ? pick<2>(nv) : pick<0>(nv));
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
