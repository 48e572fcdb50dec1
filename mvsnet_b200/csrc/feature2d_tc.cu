// Image feature tower UNetDS2GN (cnn_wrapper/mvsnetworks.py:53-115, called per view at model.py:392-406) in bf16 on the
// 5th-gen tensor cores: every conv_gn / deconv_gn block (network.py:218-276, :349-409) is one launch of
// conv2d_tc_kernel, an implicit GEMM (tcgen05.mma, accumulators in TMEM) over a 128-row tile of one view.
//
// Activations: RAW (pre-normalisation) layer outputs, bf16, chunk-planar per view: [N][C/8][H][W][8] -- a (view,
// 8-channel chunk) plane is a dense 2-D array of 16-byte cells, so one TMA box of the haloed tile of all chunks lands in
// shared memory as [chunk][row][cell].  Next to them Sum / Sum^2 of every (view, group of 8 channels) in fp64.  No
// normalised tensor is ever written: the CONSUMER derives scale / shift per (view, channel) from its producers'
// statistics in its prologue (group normalisation: groups of 8 channels = one chunk, biased variance, eps inside the
// root; conv_gn has a ReLU, deconv_gn has none) and applies them while it moves the landed tile into the operand
// buffer -- the 31 group_norm passes of the fp32 implementation disappear, and so does the channel concatenation (the
// two sources are simply the first and the last chunks of the staged tile, each with its own statistics).
//
// One CTA (128 threads) = one (view, tile of 8 x TX positions): GEMM row m = yy * 16 + xx, the operand buffer holds
// per chunk one or four (stride 2: the four parities of the input) arrays of RY x 16 cells, and the A operand of ANY
// filter tap is a no-swizzle K-major descriptor over those bytes (start = tap-shifted cell, SBO = 128 B, LBO = chunk
// stride): im2col is never built.  Stride-2 convolutions (3x3 and 5x5) get their dense rows from the parity split the
// transform does on the fly; the transposed convolution runs its four output-parity classes as tap subsets into their
// own TMEM columns.  Weights: bf16 B images [tap][channel pair][2 x N x 8] per slice of <= 64 output channels, fetched by
// bulk async copies (the next slice's while this one drains).  Epilogue: tcgen05.ld -> group statistics (two floats
// per chunk and thread, one fp64 atomic pair per CTA and group) -> bf16 cells stored along x; the last layer
// (conv10_2: no normalisation) writes the fp32 NHWC feature maps the cost-volume kernel reads.
// Many small CTAs per SM (the full-resolution layers need ~12 KB of shared memory and 32 TMEM columns) overlap each
// other's load / transform / MMA / drain phases; programmatic dependent launch overlaps the launches.
#include "common.cuh"
#include "feature2d_plan.h"
#include "umma.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <mutex>
#include <string.h>

namespace mvsb200 {
using namespace umma;

namespace f2 {

constexpr int kThreads = 128;
constexpr int kMaxTaps = 25;
constexpr int kMaxOps = 72;                // MMAs per row block and slice: 9 taps x 8 channel pairs
constexpr int kMaxCin = 128;
constexpr int kMaxMB = 2;                  // 128-row MMA blocks per tile (8 tile rows of 16 cells each)
constexpr size_t kSmemMax = 220 * 1024;

struct Tap { short cls, a_off, widx, pad; };      // output class, cell offset inside a chunk of the operand buffer, filter tap
// One MMA (K = 16) of a row block: descriptor low words relative to the operand / weight buffers, accumulator columns.
// K halves = the two chunks of a channel pair of one tap -- or, with a single input chunk, two taps (LBO = their distance)
struct Op { uint32_t a_lo, b_lo, col, acc; };
struct OpSrc { short tap[2], cb[2]; };            // what the weight image of the op holds per K half: filter tap (-1: zeros), first channel

struct Params {
  alignas(64) CUtensorMap tmap_a;
  alignas(64) CUtensorMap tmap_b;
  int nch_a, nch_b;                        // 8-channel chunks of the two sources (nch_b = 0: one source)
  const double* stats_a; const double* stats_b;      // [N][nch][2] of the producing layers; NULL = not normalised (the images)
  const float *gamma_a, *beta_a, *gamma_b, *beta_b;
  int relu_a, relu_b;
  double count;                            // elements per (view, group) of the input: H * W * 8
  float eps;
  int H, W, Ho, Wo, Cout;
  int stride, transposed, ksize;
  int TX, TY, MB, tiles_x, tiles_y, n_views, grid_dbg, variant_dbg;   // tile = TY = 8 * MB rows of TX positions (MB 128-row MMA blocks)
  int px_shift;                            // log2(PXin)
  int org_mul, org_off;                    // staged window origin = org_mul * (x0, y0) + org_off
  int PXin, RYin;                          // staged window: RYin rows of PXin cells per chunk
  int nsub, sub_cells;                     // operand buffer: nsub arrays of sub_cells = RY * 16 cells per chunk
  int nchp;                                // chunks of the operand buffer (Cin / 8 rounded up to even: K = 16)
  int ntaps, ncls;
  int CS, nslices, N;                      // output channels per slice, slices, MMA N (CS rounded up to 16)
  int tmem_cols, w_slice_bytes;
  const unsigned char* wpacked;
  __nv_bfloat16* y_cp8; float* y_f32; double* stats_out;
  long long* prof;                         // development (tuning UNET_PROFILE=2): phase clocks of CTA 0, thread 0
  int nops, dbg;                           // dbg (tuning UNET_DBG): 1 = no MMAs
  int obuf;                                // operand buffers: 2 = the next item is transformed while this item's MMAs run
  int inplace;                             // the transform rewrites the staged window in place (no operand buffer of its own)
  Op ops[kMaxOps];
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__host__ __device__ inline int align128(int v) { return (v + 127) & ~127; }

// KIND = 0: every shape from the kernel parameters.  KIND = 1 (3x3 stride 1), 2 (3x3 stride 2), 3 (5x5 stride 2), 4 (transposed)
// with NCH input chunks, NCS output chunks (one slice) and MBT row blocks per tile: the same code with the shape as
// compile-time constants, for the layers that carry most of the pixels (the generic code spent 3 400 warp instructions per
// tile on run-time loop bounds, divisions and parameter reads: ncu, issue slots 58 %; 0.125 -> 0.07 ms per full-resolution
// layer).  The geometry below restates plan_layer(); the launch checks a plan against it before it picks a variant.
struct Shape { int MB, ncls, N, CS, PXin, pxs, RYin, nsub, subc, nchp, TX, TY, omul, ooff, trans, split; };
__host__ __device__ constexpr Shape shape_of(int kind, int nch, int ncs, int mb) {
  const bool s2 = kind == 2 || kind == 3;
  const int ryin = kind == 1 ? 8 * mb + 2 : (kind == 2 ? 16 * mb + 2 : (kind == 3 ? 16 * mb + 4 : 8 * mb + 1));
  return Shape{mb, kind == 4 ? 4 : 1, (ncs * 8 + 15) / 16 * 16, ncs * 8, s2 ? 32 : 16, s2 ? 5 : 4, ryin, s2 ? 4 : 1,
               (s2 ? ryin / 2 : ryin) * 16, nch == 1 ? 1 : nch, (kind == 1 || kind == 3) ? 14 : 15, 8 * mb, s2 ? 2 : 1,
               kind == 2 ? 0 : (kind == 3 ? -2 : -1), kind == 4 ? 1 : 0, s2 ? 1 : 0};
}

template <int KIND, int NCH, int NCS, int MBT, bool IP>
__global__ void __launch_bounds__(kThreads, 6) conv2d_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr bool H = KIND != 0;
  constexpr Shape SH = shape_of(H ? KIND : 1, H ? NCH : 1, H ? NCS : 1, H ? MBT : 1);
  const int nch = H ? NCH : p.nch_a + p.nch_b;
  const int c_MB = H ? SH.MB : p.MB, c_ncls = H ? SH.ncls : p.ncls, c_N = H ? SH.N : p.N, c_CS = H ? SH.CS : p.CS;
  const int c_PXin = H ? SH.PXin : p.PXin, c_pxs = H ? SH.pxs : p.px_shift, c_RYin = H ? SH.RYin : p.RYin;
  const int c_nsub = H ? SH.nsub : p.nsub, c_subc = H ? SH.subc : p.sub_cells, c_nchp = H ? SH.nchp : p.nchp;
  const int c_nsl = (H && !IP) ? 1 : p.nslices, c_obuf = H ? 1 : p.obuf, c_TX = H ? SH.TX : p.TX, c_TY = H ? SH.TY : p.TY;
  const bool c_trans = H ? SH.trans != 0 : p.transposed != 0;
  const int c_omul = H ? SH.omul : p.org_mul, c_ooff = H ? SH.ooff : p.org_off;
  const int win_cells = c_RYin * c_PXin;                       // cells of one chunk of the staged window
  const int s_bytes = nch * win_cells * 16;
  const int chunk_o = c_nsub * c_subc * 16;               // bytes of one chunk of the operand buffer
  unsigned char* s_S = smem;
  const bool c_inplace = H ? IP : p.inplace != 0;
  unsigned char* s_O = c_inplace ? s_S : s_S + align128(s_bytes);
  const int o_stride = align128(c_nchp * chunk_o + 128);      // bytes of one operand buffer
  unsigned char* s_W = c_inplace ? s_S + o_stride : s_O + (size_t)c_obuf * o_stride;
  float* s_aff = reinterpret_cast<float*>(s_W + align128(p.w_slice_bytes));     // [scale | shift][kMaxCin]
  float* s_red = s_aff + 2 * kMaxCin;                                           // [4 warps][sum(8) | sumsq(8)]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 64);                     // window landed, weights landed, MMAs done
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 3);
  uint4* s_ops = reinterpret_cast<uint4*>(bars + 4);                           // [kMaxOps] descriptor words with the buffer bases folded in
  uint64_t* bar_in = bars, *bar_w = bars + 1, *bar_mma = bars + 2;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(bar_in, 1); mbar_init(bar_w, 1); mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(s_tmem, (uint32_t)p.tmem_cols); tmem_relinquish(); }
  for (int i = tid; i < p.nops; i += kThreads)
    s_ops[i] = make_uint4(p.ops[i].a_lo + (smem_u32(s_O) >> 4), p.ops[i].b_lo + (smem_u32(s_W) >> 4), p.ops[i].col, p.ops[i].acc);
  // everything above overlapped the tail of the producing layer (programmatic dependent launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // the zero chunk of an odd Cin / 8 (K = 16 takes chunks in pairs) and the pad behind the buffer are written once
  for (int ob = 0; ob < c_obuf; ++ob)
    for (int i = tid; i < ((c_nchp - nch) * chunk_o + 128) / 16; i += kThreads)
      reinterpret_cast<uint4*>(s_O + (size_t)ob * o_stride + (size_t)nch * chunk_o)[i] = make_uint4(0u, 0u, 0u, 0u);
  // The CTA is persistent: items (view, tile) blockIdx.x, + gridDim.x, ...  The window of the NEXT item is fetched as
  // soon as the transform has emptied the staging buffer, i.e. under the MMAs and the drain of the current item;
  // TMEM, barriers, the weights of single-slice layers and the scale / shift table of a view are set up once.
  const int tiles = p.tiles_x * p.tiles_y, total = tiles * p.n_views;
  auto fetch = [&](int item) {           // one thread
    const int n = item / tiles, t = item - n * tiles;
    const int x0 = (t % p.tiles_x) * c_TX, y0 = (t / p.tiles_x) * c_TY;
    const int ix0 = c_omul * x0 + c_ooff, iy0 = c_omul * y0 + c_ooff;
    mbar_arrive_expect_tx(bar_in, (uint32_t)s_bytes);
    tma_load_4d(s_S, &p.tmap_a, ix0 * 4, iy0, 0, n, bar_in);
    if (p.nch_b) tma_load_4d(s_S + (size_t)p.nch_a * win_cells * 16, &p.tmap_b, ix0 * 4, iy0, 0, n, bar_in);
  };
  auto fetch_weights = [&](int sl) {     // one thread
    mbar_arrive_expect_tx(bar_w, (uint32_t)p.w_slice_bytes);
    const unsigned char* src = p.wpacked + (size_t)sl * p.w_slice_bytes;
    for (int off = 0; off < p.w_slice_bytes; off += 32768)
      bulk_g2s(s_W + off, src + off, (uint32_t)min(32768, p.w_slice_bytes - off), bar_w);
  };
  if (tid == 0 && (int)blockIdx.x < total) { fetch((int)blockIdx.x); fetch_weights(0); }

  const uint32_t idesc = make_idesc_bf16_f32(128, c_N);
  const uint64_t desc_hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;          // version 1, SBO = 128 B
  const int xx = lane & 15, yy0 = warp * 2 + (lane >> 4);      // GEMM row of block b: (yy0 + 8 b, xx)
  const bool row_ok = xx < c_TX;
  const int nck = c_CS >> 3, ncho = p.Cout >> 3;
  const bool split = H ? SH.split != 0 : (p.stride == 2 && !c_trans);
  const int cls_shift = c_ncls == 4 ? 2 : 0;                    // (1 or 4 output classes)
  const int lim_y = c_trans ? p.H : p.Ho, lim_x = c_trans ? p.W : p.Wo;
  const bool one_slice = c_nsl == 1;
  uint32_t ph_in = 0, ph_w = 0, ph_mma = 0;
  bool w_ready = false;                  // single-slice layers: the weights stay
  int cur_view = -1;
  float gs[8], gq[8];                    // group statistics of this thread: [chunk of the slice]
#pragma unroll
  for (int k = 0; k < 8; ++k) { gs[k] = 0.0f; gq[k] = 0.0f; }

  // group statistics: warp sums -> shared memory -> one fp64 atomic pair per CTA and group (of view n, slice sl)
  auto flush_stats = [&](int n, int sl) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float s_ = gs[k], q_ = gq[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { s_ += __shfl_xor_sync(0xffffffffu, s_, o); q_ += __shfl_xor_sync(0xffffffffu, q_, o); }
      if (lane == 0) { s_red[warp * 16 + k] = s_; s_red[warp * 16 + 8 + k] = q_; }
      gs[k] = 0.0f; gq[k] = 0.0f;
    }
    __syncthreads();
    if (tid < 2 * nck) {
      const int k = tid >> 1, which = tid & 1;
      const float v = s_red[which * 8 + k] + s_red[16 + which * 8 + k] + s_red[32 + which * 8 + k] + s_red[48 + which * 8 + k];
      atomicAdd(p.stats_out + ((size_t)n * ncho + sl * nck + k) * 2 + which, (double)v);
    }
    __syncthreads();
  };

  // scale / shift table of a view: group normalisation of the inputs per (view, channel) -- mean and biased variance of the
  // (view, group) from the producer's fp64 sums, rounded once (network.py:249-253), then gamma / beta (network.py:269)
  auto set_view = [&](int n) {
    if (tid < nch * 8) {
      const bool b = tid >= p.nch_a * 8;
      const int c = b ? tid - p.nch_a * 8 : tid;
      const double* st = b ? p.stats_b : p.stats_a;
      float sc = 1.0f, sh = 0.0f;
      if (st) {
        const int g = c >> 3, G = b ? p.nch_b : p.nch_a;
        const double sm = st[((size_t)n * G + g) * 2], sq = st[((size_t)n * G + g) * 2 + 1];
        const double mean = sm / p.count;
        double var = sq / p.count - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn((float)var, p.eps)));
        sc = __fmul_rn(rstd, (b ? p.gamma_b : p.gamma_a)[c]);
        sh = __fsub_rn((b ? p.beta_b : p.beta_a)[c], __fmul_rn((float)mean, sc));
      }
      s_aff[tid] = sc;
      s_aff[kMaxCin + tid] = sh;
    }
    __syncthreads();
  };

  // transform: staged window -> operand buffer (normalise, ReLU, zero outside the image = SAME padding, parity split for
  // the stride-2 convolutions).  A thread owns window positions and walks the chunks of each.
  auto transform = [&](int ix0, int iy0, int ob) {
    for (int pos = tid; pos < win_cells; pos += kThreads) {
      const int r = pos >> c_pxs, j = pos & (c_PXin - 1);
      const int ay = iy0 + r, ax = ix0 + j;
      const bool inside = ay >= 0 && ay < p.H && ax >= 0 && ax < p.W;
      int sub = 0, rr = r, jj = j;
      if (split) { sub = (r & 1) * 2 + (j & 1); rr = r >> 1; jj = j >> 1; }
      unsigned char* dst = s_O + (size_t)ob * o_stride + ((size_t)sub * c_subc + rr * 16 + jj) * 16;
      const unsigned char* src = s_S + (size_t)pos * 16;
      for (int ch = 0; ch < nch; ++ch) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (inside) {
          v = *reinterpret_cast<const uint4*>(src + (size_t)ch * win_cells * 16);
          const bool b = ch >= p.nch_a;
          if (b ? p.stats_b != nullptr : p.stats_a != nullptr) {
            const bool relu = (b ? p.relu_b : p.relu_a) != 0;
            // channel pairs: one shift / mask per half to unpack, one packed FMA per pair, ReLU folded into the rounding
            // (the kernel is bound by issue slots)
            const float4 s0 = *reinterpret_cast<const float4*>(s_aff + ch * 8), s1 = *reinterpret_cast<const float4*>(s_aff + ch * 8 + 4);
            const float4 h0 = *reinterpret_cast<const float4*>(s_aff + kMaxCin + ch * 8),
                         h1 = *reinterpret_cast<const float4*>(s_aff + kMaxCin + ch * 8 + 4);
            uint32_t* w = reinterpret_cast<uint32_t*>(&v);
            const float2 f0 = ffma2(unpack_bf16x2(w[0]), make_float2(s0.x, s0.y), make_float2(h0.x, h0.y));
            const float2 f1 = ffma2(unpack_bf16x2(w[1]), make_float2(s0.z, s0.w), make_float2(h0.z, h0.w));
            const float2 f2 = ffma2(unpack_bf16x2(w[2]), make_float2(s1.x, s1.y), make_float2(h1.x, h1.y));
            const float2 f3 = ffma2(unpack_bf16x2(w[3]), make_float2(s1.z, s1.w), make_float2(h1.z, h1.w));
            if (relu) {
              w[0] = pack_bf16x2_relu(f0.x, f0.y); w[1] = pack_bf16x2_relu(f1.x, f1.y);
              w[2] = pack_bf16x2_relu(f2.x, f2.y); w[3] = pack_bf16x2_relu(f3.x, f3.y);
            } else {
              w[0] = pack2(f0.x, f0.y); w[1] = pack2(f1.x, f1.y); w[2] = pack2(f2.x, f2.y); w[3] = pack2(f3.x, f3.y);
            }
          }
        }
        *reinterpret_cast<uint4*>(dst + (size_t)ch * chunk_o) = v;
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
  };

  // the MMAs of one slice into accumulator buffer `buf` (warp 0; one elected lane issues, then commits on bar_mma)
  const int acc_cols = c_MB * c_ncls * c_N;
  auto issue = [&](int buf, int ob) {
    if (warp == 0) {
      if (!w_ready) { mbar_wait(bar_w, ph_w); w_ready = one_slice; }
      tc_fence_after();
      if (elect_one()) {
        // one 16-byte shared-memory record per MMA (read ahead by the unrolled loop: an indexed read of the kernel
        // parameters per MMA cost ~250 clk each); row block b = the same descriptors 128 cells further down
        const uint32_t d0 = tmem_base + (uint32_t)(buf * acc_cols);
        const uint32_t oofs = (uint32_t)(ob * o_stride) >> 4;
#pragma unroll 4
        for (int o = 0; o < p.nops; ++o) {
          const uint4 e = s_ops[o];
          for (int b = 0; b < c_MB; ++b) {
            if (p.dbg & 1) continue;
            mma_bf16(d0 + e.z + (uint32_t)(b * c_ncls * c_N), desc_hi | (uint64_t)(e.x + oofs + (uint32_t)(b * 128)),
                     desc_hi | (uint64_t)e.y, idesc, e.w);
          }
        }
        mma_commit(bar_mma);
      }
      __syncwarp();
    }
    ph_w ^= one_slice ? 0u : 1u;
  };

  // drain of one slice of item (n, x0, y0): group statistics + stores
  auto drain = [&](int n, int x0, int y0, int sl, int buf) {
    const int gx = x0 + xx;                // GEMM-row position (output; input for the transposed conv)
    const int nbc = c_MB * c_ncls;
    if (c_CS == 8) {
      // 8-channel layers: one chunk per (row block, class) pair; two pairs per TMEM round trip (a round trip costs
      // ~1 000 clk while MMAs are queued)
      for (int bc0 = 0; bc0 < nbc; bc0 += 2) {
        uint32_t r[16];
        const uint32_t tb = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * acc_cols + bc0 * c_N);
        tmem_ld8(tb, r);
        if (bc0 + 1 < nbc) tmem_ld8(tb + (uint32_t)c_N, r + 8);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int bc = bc0 + u;
          if (bc >= nbc) break;
          const int b = bc >> cls_shift, cls = bc & (c_ncls - 1);
          const int gy = y0 + yy0 + 8 * b;
          int oy = gy, ox = gx;
          const bool ok = row_ok && gy < lim_y && gx < lim_x;
          if (c_trans) { oy = 2 * gy + (cls >> 1); ox = 2 * gx + (cls & 1); }
          if (!ok) continue;
          {
            const float2* f = reinterpret_cast<const float2*>(r + 8 * u);
            const float2 s2 = fadd2(fadd2(f[0], f[1]), fadd2(f[2], f[3]));
            const float2 q2 = ffma2(f[3], f[3], ffma2(f[2], f[2], ffma2(f[1], f[1], ffma2(f[0], f[0], make_float2(0.0f, 0.0f)))));
            gs[0] += s2.x + s2.y; gq[0] += q2.x + q2.y;
          }
          uint4 pk;
          pk.x = pack2(__uint_as_float(r[8 * u]), __uint_as_float(r[8 * u + 1]));
          pk.y = pack2(__uint_as_float(r[8 * u + 2]), __uint_as_float(r[8 * u + 3]));
          pk.z = pack2(__uint_as_float(r[8 * u + 4]), __uint_as_float(r[8 * u + 5]));
          pk.w = pack2(__uint_as_float(r[8 * u + 6]), __uint_as_float(r[8 * u + 7]));
          *reinterpret_cast<uint4*>(p.y_cp8 + ((((size_t)n * ncho + sl) * p.Ho + oy) * p.Wo + ox) * 8) = pk;
        }
      }
      tc_fence_before();
      return;
    }
    for (int bc = 0; bc < nbc; ++bc) {
      const int b = bc >> cls_shift, cls = bc & (c_ncls - 1);
      const int gy = y0 + yy0 + 8 * b;
      int oy = gy, ox = gx;
      const bool ok = row_ok && gy < lim_y && gx < lim_x;
      if (c_trans) { oy = 2 * gy + (cls >> 1); ox = 2 * gx + (cls & 1); }
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        if (c0 < c_N) {
          uint32_t r[16];
          tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * acc_cols + bc * c_N + c0), r);
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ck = (c0 >> 3) + h;                          // chunk of the slice (compile-time index)
            if (ck < nck && ok) {
              {
                const float2* f = reinterpret_cast<const float2*>(r + 8 * h);
                const float2 s2 = fadd2(fadd2(f[0], f[1]), fadd2(f[2], f[3]));
                const float2 q2 = ffma2(f[3], f[3], ffma2(f[2], f[2], ffma2(f[1], f[1], ffma2(f[0], f[0], make_float2(0.0f, 0.0f)))));
                gs[ck] += s2.x + s2.y; gq[ck] += q2.x + q2.y;
              }
              const int gch = sl * nck + ck;
              if (p.y_f32) {
                float4* yo = reinterpret_cast<float4*>(p.y_f32 + (((size_t)n * p.Ho + oy) * p.Wo + ox) * p.Cout + gch * 8);
                yo[0] = make_float4(__uint_as_float(r[8 * h]), __uint_as_float(r[8 * h + 1]), __uint_as_float(r[8 * h + 2]),
                                    __uint_as_float(r[8 * h + 3]));
                yo[1] = make_float4(__uint_as_float(r[8 * h + 4]), __uint_as_float(r[8 * h + 5]), __uint_as_float(r[8 * h + 6]),
                                    __uint_as_float(r[8 * h + 7]));
              } else {
                uint4 pk;
                pk.x = pack2(__uint_as_float(r[8 * h]), __uint_as_float(r[8 * h + 1]));
                pk.y = pack2(__uint_as_float(r[8 * h + 2]), __uint_as_float(r[8 * h + 3]));
                pk.z = pack2(__uint_as_float(r[8 * h + 4]), __uint_as_float(r[8 * h + 5]));
                pk.w = pack2(__uint_as_float(r[8 * h + 6]), __uint_as_float(r[8 * h + 7]));
                *reinterpret_cast<uint4*>(p.y_cp8 + ((((size_t)n * ncho + gch) * p.Ho + oy) * p.Wo + ox) * 8) = pk;
              }
            }
          }
        }
      }
    }
    tc_fence_before();
  };

  if (c_obuf == 2) {
    // Single-slice layers with room for two operand buffers: while the MMAs of item i run (issue + commit round trip
    // ~2 300 clk), the threads transform item i + 1 into the other buffer; then drain i.  The scale / shift table
    // follows the item being TRANSFORMED, the statistics registers the item being DRAINED.
    const int G = (int)gridDim.x;
    int item = (int)blockIdx.x, it = 0, stat_view = -1;
    // (view, tile row, tile column) of the item and of the next one, advanced by the grid stride without divisions
    const int step_x = G % p.tiles_x, step_y = (G / p.tiles_x) % p.tiles_y, step_n = G / tiles;
    int n = item / tiles, ty = (item - n * tiles) / p.tiles_x, tx = item - n * tiles - ty * p.tiles_x;
    auto advance = [&](int& vn, int& vy, int& vx) {
      vx += step_x;
      if (vx >= p.tiles_x) { vx -= p.tiles_x; ++vy; }
      vy += step_y;
      if (vy >= p.tiles_y) { vy -= p.tiles_y; ++vn; }
      vn += step_n;
    };
    int n1 = n, ty1 = ty, tx1 = tx;
    if (item < total) {
      set_view(n); cur_view = n;
      mbar_wait(bar_in, ph_in); ph_in ^= 1u;
      transform(c_omul * tx * c_TX + c_ooff, c_omul * ty * c_TY + c_ooff, 0);
      if (tid == 0 && item + G < total) fetch(item + G);
    }
    for (; item < total; item += G, ++it) {
      n = n1; ty = ty1; tx = tx1;
      const int x0 = tx * c_TX, y0 = ty * c_TY;
      tc_fence_after();
      issue(0, it & 1);
      if (item + G < total) {
        advance(n1, ty1, tx1);
        const int x1 = tx1 * c_TX, y1 = ty1 * c_TY;
        if (n1 != cur_view) { set_view(n1); cur_view = n1; }
        mbar_wait(bar_in, ph_in); ph_in ^= 1u;
        transform(c_omul * x1 + c_ooff, c_omul * y1 + c_ooff, (it + 1) & 1);
        if (tid == 0 && item + 2 * G < total) fetch(item + 2 * G);
      }
      mbar_wait(bar_mma, ph_mma); ph_mma ^= 1u;
      tc_fence_after();
      if (p.stats_out && n != stat_view) {
        if (stat_view >= 0) flush_stats(stat_view, 0);
        stat_view = n;
      }
      drain(n, x0, y0, 0, 0);
      __syncthreads();                     // accumulators drained before the next MMAs overwrite them
    }
    if (p.stats_out && stat_view >= 0) flush_stats(stat_view, 0);
  } else {
    for (int item = (int)blockIdx.x; item < total; item += (int)gridDim.x) {
      const int n = item / tiles, t = item - n * tiles;
      const int x0 = (t % p.tiles_x) * c_TX, y0 = (t / p.tiles_x) * c_TY;
      if (n != cur_view) {
        if (p.stats_out && one_slice && cur_view >= 0) flush_stats(cur_view, 0);
        set_view(n);
        cur_view = n;
      }
      long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
      if (p.prof) t0 = clock64();
      mbar_wait(bar_in, ph_in);
      ph_in ^= 1u;
      if (p.prof) t1 = clock64();
      transform(c_omul * x0 + c_ooff, c_omul * y0 + c_ooff, 0);
      if (p.prof) t2 = clock64();
      // the staging buffer is free: the next item's window lands under this item's MMAs and drain (in-place layers: the
      // window IS the operand buffer, so the fetch waits for the last MMAs)
      if (!c_inplace && tid == 0 && item + (int)gridDim.x < total) fetch(item + (int)gridDim.x);
      for (int sl = 0; sl < c_nsl; ++sl) {
        issue(0, 0);
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1u;
        tc_fence_after();
        if (p.prof) t3 = clock64();
        // multi-slice layers: the weight buffer is free, the images of the next (item, slice) land while this one drains
        if (!one_slice && tid == 0) {
          const bool more = sl + 1 < c_nsl || item + (int)gridDim.x < total;
          if (more) fetch_weights(sl + 1 < c_nsl ? sl + 1 : 0);
        }
        drain(n, x0, y0, sl, 0);
        __syncthreads();                   // accumulators drained: the next MMAs may overwrite them (and the operand buffer)
        tc_fence_after();
        if (p.stats_out && !one_slice) flush_stats(n, sl);
      }
      if (c_inplace && tid == 0 && item + (int)gridDim.x < total) fetch(item + (int)gridDim.x);
      if (p.prof && blockIdx.x == 0 && tid == 0) {
        t4 = clock64();
        p.prof[0] += t1 - t0; p.prof[1] += t2 - t1; p.prof[2] += t3 - t2; p.prof[3] += t4 - t3; p.prof[4] += 1;
      }
    }
    if (p.stats_out && one_slice && cur_view >= 0) flush_stats(cur_view, 0);
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---- images -> chunk-planar bf16 with the 3 channels padded to one cell ------------------------------------------
__global__ void image_to_cp8_kernel(const float* __restrict__ img, size_t npix, uint4* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const float a = img[i * 3], b = img[i * 3 + 1], c = img[i * 3 + 2];
    out[i] = make_uint4(pack2(a, b), pack2(c, 0.0f), 0u, 0u);
  }
}

// ---- weights -> bf16 B images, all layers in one launch --------------------------------------------------------
struct PackJob {
  const float* kernel_tf; uint16_t* out;
  int nops, N, CS, nslices, Cout, CinT, transposed;
  OpSrc src[kMaxOps];
};

// image (slice, op) = [2 halves][N rows][8]: row = output channel of the slice, half / k = input channel of the half's tap
__global__ void pack2d_all_kernel(const PackJob* __restrict__ jobs) {
  const PackJob& j = jobs[blockIdx.y];
  const int per_img = 2 * j.N * 8, total = j.nslices * j.nops * per_img;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k8 = i & 7, row = (i >> 3) % j.N, half = (i / (8 * j.N)) & 1;
    const int img = i / per_img, op = img % j.nops, sl = img / j.nops;
    const int tap = j.src[op].tap[half], ci = j.src[op].cb[half] + k8, co = sl * j.CS + row;
    float w = 0.0f;
    if (tap >= 0 && row < j.CS && co < j.Cout && ci < j.CinT)
      w = j.transposed ? j.kernel_tf[((size_t)tap * j.Cout + co) * j.CinT + ci]
                       : j.kernel_tf[((size_t)tap * j.CinT + ci) * j.Cout + co];
    const __nv_bfloat16 h = __float2bfloat16_rn(w);
    j.out[i] = *reinterpret_cast<const uint16_t*>(&h);
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

// chunk-planar tensor [N][nch][H][W] cells as (4 W, H, nch, N) fp32 elements; box = the staged window of all chunks
static bool make_map(CUtensorMap* tm, const void* base, int n, int nch, int H, int W, int PXin, int RYin) {
  cuuint64_t gdim[4] = {(cuuint64_t)W * 4, (cuuint64_t)H, (cuuint64_t)nch, (cuuint64_t)n};
  cuuint64_t gstr[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)nch * H * W * 16};
  cuuint32_t box[4] = {(cuuint32_t)PXin * 4, (cuuint32_t)RYin, (cuuint32_t)nch, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// geometry + tap table of one layer; cin = channels of the concatenated input as stored (image: 8)
static int plan_layer(int ksize, int stride, int transposed, int cin, int cout, int H, int W, int MB, Params* c, size_t* smem,
                      OpSrc* srcs, bool inplace = false) {
  memset(c, 0, sizeof(*c));
  Tap taps[kMaxTaps];
  const int kTY = 8 * MB;
  c->MB = MB; c->TY = kTY;
  c->H = H; c->W = W; c->Cout = cout; c->stride = stride; c->transposed = transposed; c->ksize = ksize;
  c->Ho = unet_out_extent(H, stride, transposed); c->Wo = unet_out_extent(W, stride, transposed);
  int RY = 0, nt = 0;
  c->ncls = 1; c->nsub = 1;
  if (transposed) {
    // out[2 i + k] += x[i] w[k] (network.py:327, TF SAME): class p = output parity; p = 0 reads x[i] (k = 0) and x[i-1]
    // (k = 2), p = 1 reads x[i] (k = 1).  GEMM rows = input positions; window origin (x0 - 1, y0 - 1).
    c->TX = 15; c->org_mul = 1; c->org_off = -1; c->PXin = 16; c->RYin = kTY + 1; RY = kTY + 1; c->ncls = 4;
    for (int cls = 0; cls < 4; ++cls) {
      const int py = cls >> 1, px = cls & 1;
      for (int sy = 0; sy <= (py ? 0 : 1); ++sy)
        for (int sx = 0; sx <= (px ? 0 : 1); ++sx)
          taps[nt++] = {(short)cls, (short)((1 - sy) * 16 + (1 - sx)), (short)((py + 2 * sy) * 3 + (px + 2 * sx)), 0};
    }
  } else if (stride == 1) {
    c->TX = 14; c->org_mul = 1; c->org_off = -1; c->PXin = 16; c->RYin = kTY + 2; RY = kTY + 2;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) taps[nt++] = {0, (short)(kh * 16 + kw), (short)(kh * 3 + kw), 0};
  } else {
    // TF SAME, even extents: 3x3 pads (0, 1): out o reads in 2 o + k; 5x5 pads (1, 2): out o reads in 2 o - 1 + k.
    // The window starts at an even input position, so cell (r, j) of it has parity (r & 1, j & 1).
    const int lead = ksize == 5 ? 2 : 0, e0 = ksize == 5 ? 1 : 0;       // window origin 2 o0 - lead; tap k -> 2 xx + e0 + k
    c->TX = ksize == 5 ? 14 : 15; c->org_mul = 2; c->org_off = -lead; c->PXin = 32;
    c->RYin = ksize == 5 ? 2 * kTY + 4 : 2 * kTY + 2; RY = c->RYin / 2; c->nsub = 4;
    for (int kh = 0; kh < ksize; ++kh)
      for (int kw = 0; kw < ksize; ++kw) {
        const int ey = e0 + kh, ex = e0 + kw;
        taps[nt++] = {0, (short)((((ey & 1) * 2 + (ex & 1)) * RY + (ey >> 1)) * 16 + (ex >> 1)), (short)(kh * ksize + kw), 0};
      }
  }
  c->ntaps = nt;
  c->sub_cells = RY * 16;
  const int nch = cin / 8;
  // ops: channel pairs of a tap -- or, with one input chunk, pairs of taps of the same class (K half 1 = the later tap,
  // LBO = their distance; an odd tap out runs with a zero second half)
  OpSrc src_local[kMaxOps];
  OpSrc* src = srcs ? srcs : src_local;
  int nops = 0;
  const int chunk_o_cells = c->nsub * c->sub_cells;
  unsigned seen = 0u;
  if (nch == 1) {
    c->nchp = 1;
    for (int i = 1; i < nt; ++i)           // by class, then by offset
      for (int k = i; k > 0 && (taps[k].cls < taps[k - 1].cls || (taps[k].cls == taps[k - 1].cls && taps[k].a_off < taps[k - 1].a_off)); --k) {
        const Tap t = taps[k]; taps[k] = taps[k - 1]; taps[k - 1] = t;
      }
    for (int i = 0; i < nt;) {
      const bool pair = i + 1 < nt && taps[i + 1].cls == taps[i].cls;
      const int lbo_cells = pair ? taps[i + 1].a_off - taps[i].a_off : 1;
      c->ops[nops] = {(uint32_t)taps[i].a_off | ((uint32_t)lbo_cells << 16), 0u, 0u, (seen >> taps[i].cls) & 1u};
      src[nops] = {{taps[i].widx, (short)(pair ? taps[i + 1].widx : -1)}, {0, 0}};
      c->ops[nops].col = (uint32_t)taps[i].cls;          // class for now; columns once N is known
      seen |= 1u << taps[i].cls;
      ++nops;
      i += pair ? 2 : 1;
    }
  } else {
    if (nch & 1) return MVSB200_ERR_UNSUPPORTED;          // (the tower has 1, 2, 4, 8 or 16 chunks)
    c->nchp = nch;
    if (nt * (nch / 2) > kMaxOps) return MVSB200_ERR_UNSUPPORTED;
    for (int i = 0; i < nt; ++i)
      for (int pr = 0; pr < nch / 2; ++pr) {
        c->ops[nops] = {(uint32_t)(2 * pr * chunk_o_cells + taps[i].a_off) | ((uint32_t)chunk_o_cells << 16), 0u,
                        (uint32_t)taps[i].cls, (seen >> taps[i].cls) & 1u};
        src[nops] = {{taps[i].widx, taps[i].widx}, {(short)(16 * pr), (short)(16 * pr + 8)}};
        seen |= 1u << taps[i].cls;
        ++nops;
      }
  }
  c->nops = nops;
  const int npairs = 1;
  nt = nops;                               // (weights: one image per op)
  // slice of output channels: weights of a slice <= 74 KB, accumulators of all classes <= 256 columns
  const size_t o_bytes = (size_t)align128(c->nchp * c->nsub * c->sub_cells * 16 + 128);
  // a second operand buffer where it is cheap (the big-resolution layers: a few KB): transform under the MMAs
  c->obuf = (o_bytes <= 20 * 1024 && tuning().unet_obuf == 2) ? 2 : 1;       // (measured: no gain at config 2; on request)
  // in place: the staged window is the operand buffer (same geometry: one array per chunk, an even chunk count)
  if (inplace && !(c->nsub == 1 && c->nchp == nch)) return MVSB200_ERR_UNSUPPORTED;
  c->inplace = inplace ? 1 : 0;
  if (inplace) c->obuf = 1;
  const size_t fixed = (inplace ? o_bytes : (size_t)align128(nch * c->RYin * c->PXin * 16) + c->obuf * o_bytes) +
                       2 * kMaxCin * sizeof(float) + 64 * sizeof(float) + 4 * sizeof(uint64_t) + kMaxOps * sizeof(uint4);
  int CS = cout < 64 ? cout : 64;
  for (;;) {
    const int N = (CS + 15) / 16 * 16;
    const size_t wb = (size_t)nt * npairs * 2 * N * 16;
    if (wb <= (inplace ? (size_t)160 * 1024 : (size_t)75776) && fixed + align128((int)wb) <= kSmemMax && MB * c->ncls * N <= 256) break;
    if (CS <= 8) return MVSB200_ERR_UNSUPPORTED;
    CS /= 2;
  }
  c->CS = CS; c->nslices = cout / CS; c->N = (CS + 15) / 16 * 16;
  c->w_slice_bytes = nt * npairs * 2 * c->N * 16;
  if (c->nslices > 1) c->obuf = 1;         // (multi-slice layers keep the sequential loop; their buffer is large anyway)
  for (int o = 0; o < nops; ++o) {
    c->ops[o].col *= (uint32_t)c->N;
    c->ops[o].b_lo = (uint32_t)((o * 2 * c->N * 16) >> 4) | ((uint32_t)((c->N * 16) >> 4) << 16);
  }
  int cols = 32;
  while (cols < MB * c->ncls * c->N) cols <<= 1;
  c->px_shift = c->PXin == 32 ? 5 : 4;
  c->tmem_cols = cols;
  c->tiles_x = ceil_div(transposed ? W : c->Wo, c->TX);
  c->tiles_y = ceil_div(transposed ? H : c->Ho, kTY);
  *smem = fixed + align128(c->w_slice_bytes);
  return MVSB200_OK;
}

struct TowerPlan {
  int h[MVSB200_UNET_LAYERS], w[MVSB200_UNET_LAYERS], c[MVSB200_UNET_LAYERS];
  size_t off[MVSB200_UNET_LAYERS], woff[MVSB200_UNET_LAYERS];
  size_t image_off, stats_off, stats_bytes, jobs_off, total;
  int gmax;
};

static int make_tower_plan(int n, int H, int W, int bf, TowerPlan* p) {
  MVS_CHECK_ARG(n >= 1 && n <= 65535 && H >= 16 && W >= 16 && bf >= 8 && bf % 8 == 0 && 16 * bf <= kMaxCin,
                "unet(bf16): bad shape N=%d %dx%d base_filter=%d (base_filter must be 8: at most %d channels)", n, H, W, bf,
                kMaxCin);
  MVS_CHECK_ARG(H % 16 == 0 && W % 16 == 0, "unet(bf16): H=%d and W=%d must be multiples of 16", H, W);
  size_t off = 0;
  p->image_off = off; off += align_up((size_t)n * H * W * 16, 256);
  for (int l = 0; l < MVSB200_UNET_LAYERS; ++l) {
    const F2Layer& L = kUnet[l];
    const int ih = L.src_a < 0 ? H : p->h[L.src_a], iw = L.src_a < 0 ? W : p->w[L.src_a];
    p->h[l] = unet_out_extent(ih, L.stride, L.transposed);
    p->w[l] = unet_out_extent(iw, L.stride, L.transposed);
    p->c[l] = bf * L.mult;
    p->off[l] = off;
    if (l != MVSB200_UNET_LAYERS - 1) off += align_up((size_t)n * p->h[l] * p->w[l] * p->c[l] * 2, 256);
  }
  p->gmax = 16 * bf / 8;
  p->stats_off = off;
  p->stats_bytes = (size_t)MVSB200_UNET_LAYERS * n * p->gmax * 2 * sizeof(double);
  off += align_up(p->stats_bytes, 256);
  for (int l = 0; l < MVSB200_UNET_LAYERS; ++l) {
    const F2Layer& L = kUnet[l];
    const int cin = (L.src_a < 0 ? 8 : p->c[L.src_a]) + (L.src_b >= 0 ? p->c[L.src_b] : 0);
    Params c;
    size_t smem;
    int rc = plan_layer(L.k, L.stride, L.transposed, cin, p->c[l], 16, 16, 1, &c, &smem, nullptr);
    if (rc) { set_error("unet(bf16): no plan for layer %s", L.name); return rc; }
    p->woff[l] = off;
    off += align_up((size_t)c.nslices * c.w_slice_bytes, 256);
  }
  p->jobs_off = off;
  off += align_up(sizeof(PackJob) * MVSB200_UNET_LAYERS, 256);
  p->total = off;
  return MVSB200_OK;
}

// Plan of layer l at this problem size: geometry, ops, slices -- with two 128-row blocks per tile where there are plenty of
// tiles (the per-item latency chain is paid once per tile; ops and weight images do not depend on it).
static int choose_plan(int l, int n_views, int height, int width, const TowerPlan& tp, Params* plan, size_t* smem, OpSrc* srcs) {
  const F2Layer& L = kUnet[l];
  const int ca = L.src_a < 0 ? 8 : tp.c[L.src_a], cb = L.src_b >= 0 ? tp.c[L.src_b] : 0;
  const int ih = L.src_a < 0 ? height : tp.h[L.src_a], iw = L.src_a < 0 ? width : tp.w[L.src_a];
  int rc = plan_layer(L.k, L.stride, L.transposed, ca + cb, tp.c[l], ih, iw, 1, plan, smem, srcs);
  if (rc) { set_error("unet(bf16): no plan for layer %s", L.name); return rc; }
  const int mb_forced = tuning().unet_mb;
  if ((mb_forced == 2 || (mb_forced == 0 && plan->tiles_x * plan->tiles_y * n_views >= 6000)) && plan->nslices == 1) {
    Params two;
    size_t smem2;
    if (plan_layer(L.k, L.stride, L.transposed, ca + cb, tp.c[l], ih, iw, 2, &two, &smem2, nullptr) == MVSB200_OK &&
        two.nslices == 1 && two.CS == plan->CS && two.nops == plan->nops) { *plan = two; *smem = smem2; }
  }
  // layers whose weights do not fit beside window + operand buffer: transform in place if that saves slices (each slice
  // re-fetches 74 KB of weights per tile)
  if (plan->nslices > 1 && tuning().unet_inplace != 0) {
    Params ip;
    size_t smem3;
    static thread_local OpSrc src3[kMaxOps];
    if (plan_layer(L.k, L.stride, L.transposed, ca + cb, tp.c[l], ih, iw, 1, &ip, &smem3, src3, true) == MVSB200_OK &&
        ip.nslices < plan->nslices && ip.nops == plan->nops) { *plan = ip; *smem = smem3; }
  }
  plan->nch_a = ca / 8; plan->nch_b = cb / 8;
  return MVSB200_OK;
}

}  // namespace f2
}  // namespace mvsb200

using namespace mvsb200;
using namespace mvsb200::f2;

/* Host-only view of the tower's launch plans (no device work; tests/test_planner.py): out[0..11] = {kind (1: 3x3 stride 1,
 * 2: 3x3 stride 2, 3: 5x5 stride 2, 4: transposed), input chunks, output channels per slice, slices, MMA N, row blocks per
 * tile, MMAs per row block and slice, shared memory bytes, TMEM columns, tiles per view, weight bytes per slice, operand
 * buffers} of layer `layer` at this problem size. */
extern "C" int mvsb200_unet_tc_plan(int n_views, int height, int width, int base_filter, int layer, int* out) {
  MVS_CHECK_ARG(out && layer >= 0 && layer < MVSB200_UNET_LAYERS, "unet_tc_plan: bad layer %d", layer);
  TowerPlan tp;
  int rc = make_tower_plan(n_views, height, width, base_filter, &tp);
  if (rc) return rc;
  static thread_local Params plan;
  size_t smem = 0;
  rc = choose_plan(layer, n_views, height, width, tp, &plan, &smem, nullptr);
  if (rc) return rc;
  const F2Layer& L = kUnet[layer];
  const int v[12] = {L.transposed ? 4 : (L.stride == 1 ? 1 : (L.k == 3 ? 2 : 3)), plan.nch_a + plan.nch_b, plan.CS, plan.nslices, plan.N,
                     plan.MB, plan.nops, (int)smem, plan.tmem_cols, plan.tiles_x * plan.tiles_y, plan.w_slice_bytes, plan.obuf};
  for (int i = 0; i < 12; ++i) out[i] = v[i];
  return MVSB200_OK;
}

extern "C" size_t mvsb200_unet_tc_workspace_bytes(int n_views, int height, int width, int base_filter) {
  TowerPlan p;
  if (make_tower_plan(n_views, height, width, base_filter, &p)) return 0;
  return p.total;
}

/* Raw (pre-normalisation) output of a layer after mvsb200_unet_tc_forward: byte offset in the workspace of the bf16
 * chunk-planar tensor [N][C/8][Ho][Wo][8], its dims {Ho, Wo, C}, and the offset of its [N][C/8][2] fp64 statistics. */
extern "C" int mvsb200_unet_tc_layer_raw(int n_views, int height, int width, int base_filter, int layer, size_t* offset,
                                         int* dims, size_t* stats_offset) {
  TowerPlan p;
  int rc = make_tower_plan(n_views, height, width, base_filter, &p);
  if (rc) return rc;
  MVS_CHECK_ARG(layer >= 0 && layer < MVSB200_UNET_LAYERS - 1 && offset && dims && stats_offset, "unet_tc_layer_raw: bad layer %d", layer);
  *offset = p.off[layer];
  dims[0] = p.h[layer]; dims[1] = p.w[layer]; dims[2] = p.c[layer];
  *stats_offset = p.stats_off + (size_t)layer * n_views * p.gmax * 2 * sizeof(double);
  return MVSB200_OK;
}

extern "C" int mvsb200_unet_tc_forward(const float* images, const mvsb200_unet_params* params, int n_views, int height,
                                       int width, int base_filter, float gn_eps, float* feats, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  MVS_CHECK_ARG(images && params && feats && workspace, "unet_tc_forward: NULL pointer");
  MVS_CHECK_ARG(get_encode() != nullptr, "unet_tc_forward: cuTensorMapEncodeTiled is not available from the driver");
  TowerPlan tp;
  int rc = make_tower_plan(n_views, height, width, base_filter, &tp);
  if (rc) return rc;
  if (workspace_bytes < tp.total) {
    set_error("unet_tc_forward: workspace %zu < required %zu bytes", workspace_bytes, tp.total);
    return MVSB200_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  double* stats = (double*)(ws + tp.stats_off);
  MVS_CUDA(cudaMemsetAsync(stats, 0, tp.stats_bytes, s));
  {
    const size_t npix = (size_t)n_views * height * width;
    const size_t want = (npix + 255) / 256, cap = (size_t)sm_count_current() * 16;
    image_to_cp8_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(images, npix, (uint4*)(ws + tp.image_off));
    MVS_LAUNCH_CHECK("image_to_cp8_kernel");
  }
  // plans + packed weights (the job table travels through the workspace: it is larger than a kernel's parameters)
  static thread_local Params plans[MVSB200_UNET_LAYERS];
  static thread_local PackJob jobs[MVSB200_UNET_LAYERS];
  size_t smems[MVSB200_UNET_LAYERS];
  for (int l = 0; l < MVSB200_UNET_LAYERS; ++l) {
    const F2Layer& L = kUnet[l];
    MVS_CHECK_ARG(params->kernel[l] != nullptr, "unet_tc_forward: kernel[%d] (%s) is NULL", l, L.name);
    if (L.gn) MVS_CHECK_ARG(params->gamma[l] && params->beta[l], "unet_tc_forward: gamma/beta[%d] (%s) is NULL", l, L.name);
    const int ca = L.src_a < 0 ? 8 : tp.c[L.src_a], cb = L.src_b >= 0 ? tp.c[L.src_b] : 0;
    const int ih = L.src_a < 0 ? height : tp.h[L.src_a], iw = L.src_a < 0 ? width : tp.w[L.src_a];
    if (L.src_b >= 0)
      MVS_CHECK_ARG(tp.h[L.src_b] == ih && tp.w[L.src_b] == iw, "unet_tc_forward: concat extents differ at %s", L.name);
    PackJob& j = jobs[l];
    rc = choose_plan(l, n_views, height, width, tp, &plans[l], &smems[l], j.src);
    if (rc) return rc;
    j.kernel_tf = params->kernel[l]; j.out = (uint16_t*)(ws + tp.woff[l]);
    j.nops = plans[l].nops; j.N = plans[l].N; j.CS = plans[l].CS;
    j.nslices = plans[l].nslices; j.Cout = tp.c[l]; j.CinT = L.src_a < 0 ? 3 : ca + cb; j.transposed = L.transposed;
  }
  MVS_CUDA(cudaMemcpyAsync(ws + tp.jobs_off, jobs, sizeof(jobs), cudaMemcpyHostToDevice, s));
  pack2d_all_kernel<<<dim3(32, MVSB200_UNET_LAYERS), 256, 0, s>>>((const PackJob*)(ws + tp.jobs_off));
  MVS_LAUNCH_CHECK("pack2d_all_kernel");
  const bool profile = tuning().unet_profile != 0;
  cudaEvent_t pev[MVSB200_UNET_LAYERS + 1];
  if (profile) {
    for (int i = 0; i <= MVSB200_UNET_LAYERS; ++i) cudaEventCreate(&pev[i]);
    cudaEventRecord(pev[0], s);
  }
  using Kernel = void (*)(const Params);
  struct Variant { int kind, nch, ncs, mb, ip; Kernel fn; };
#define F2V(K, C, S, M) {K, C, S, M, 0, conv2d_tc_kernel<K, C, S, M, false>}
#define F2I(K, C, S, M) {K, C, S, M, 1, conv2d_tc_kernel<K, C, S, M, true>}
  static const Variant kVariants[] = {
      {0, 0, 0, 0, 0, conv2d_tc_kernel<0, 0, 0, 0, false>},
      F2V(1, 1, 1, 2), F2V(1, 2, 1, 2), F2V(1, 2, 2, 2), F2V(1, 4, 2, 2),          // full resolution and level 1, 3x3 stride 1
      F2V(1, 4, 4, 1), F2V(1, 8, 4, 1), F2V(1, 8, 8, 1),                           // levels 2 and 3
      F2V(2, 1, 2, 2), F2V(2, 2, 4, 1), F2V(2, 4, 8, 1),                           // 3x3 stride 2
      F2V(3, 1, 2, 2), F2V(3, 2, 4, 1),                                            // 5x5 stride 2
      F2V(4, 2, 1, 1), F2V(4, 2, 1, 2), F2V(4, 4, 2, 1), F2V(4, 8, 4, 1),          // transposed
      F2I(1, 16, 8, 1), F2I(4, 16, 8, 1),                                          // deep layers, transform in place
  };
#undef F2V
#undef F2I
  constexpr int kNumVariants = (int)(sizeof(kVariants) / sizeof(kVariants[0]));
  static std::atomic<uint64_t> attr_done{0};
  static int kernel_regs_of[kNumVariants];
  int dev = 0;
  MVS_CUDA(cudaGetDevice(&dev));
  if (!(attr_done.load(std::memory_order_acquire) >> (dev & 63) & 1u)) {
    for (int k = 0; k < kNumVariants; ++k) {
      MVS_CUDA(cudaFuncSetAttribute((const void*)kVariants[k].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
      cudaFuncAttributes fa;
      MVS_CUDA(cudaFuncGetAttributes(&fa, (const void*)kVariants[k].fn));
      kernel_regs_of[k] = fa.numRegs;
    }
    attr_done.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  for (int l = 0; l < MVSB200_UNET_LAYERS; ++l) {
    const F2Layer& L = kUnet[l];
    Params& c = plans[l];
    const bool last = l == MVSB200_UNET_LAYERS - 1;
    const int ca = L.src_a < 0 ? 8 : tp.c[L.src_a], cb = L.src_b >= 0 ? tp.c[L.src_b] : 0;
    c.nch_a = ca / 8; c.nch_b = cb / 8;
    const void* xa = L.src_a < 0 ? (const void*)(ws + tp.image_off) : (const void*)(ws + tp.off[L.src_a]);
    MVS_CHECK_ARG(make_map(&c.tmap_a, xa, n_views, c.nch_a, c.H, c.W, c.PXin, c.RYin), "unet(bf16): tensor map of %s failed", L.name);
    if (cb)
      MVS_CHECK_ARG(make_map(&c.tmap_b, ws + tp.off[L.src_b], n_views, c.nch_b, c.H, c.W, c.PXin, c.RYin),
                    "unet(bf16): tensor map of %s failed", L.name);
    c.stats_a = L.src_a < 0 ? nullptr : stats + (size_t)L.src_a * n_views * tp.gmax * 2;
    c.gamma_a = L.src_a < 0 ? nullptr : params->gamma[L.src_a]; c.beta_a = L.src_a < 0 ? nullptr : params->beta[L.src_a];
    c.relu_a = L.src_a < 0 ? 0 : kUnet[L.src_a].relu;
    c.stats_b = cb ? stats + (size_t)L.src_b * n_views * tp.gmax * 2 : nullptr;
    c.gamma_b = cb ? params->gamma[L.src_b] : nullptr; c.beta_b = cb ? params->beta[L.src_b] : nullptr;
    c.relu_b = cb ? kUnet[L.src_b].relu : 0;
    c.count = (double)c.H * c.W * 8.0;
    c.eps = gn_eps;
    c.wpacked = (const unsigned char*)(ws + tp.woff[l]);
    c.y_cp8 = last ? nullptr : (__nv_bfloat16*)(ws + tp.off[l]);
    c.y_f32 = last ? feats : nullptr;
    c.stats_out = L.gn ? stats + (size_t)l * n_views * tp.gmax * 2 : nullptr;
    // NOTE the statistics of a layer are laid out [N][C/8][2] with ITS OWN group count
    cudaLaunchConfig_t cfg = {};
    // a compile-time variant when there is one whose geometry IS the plan's (one slice, one operand buffer)
    int variant = 0;
    if (tuning().unet_hot != 0 && c.obuf == 1 && (c.nslices == 1 || c.inplace)) {       // (fixed-shape code: one slice unless in place)
      const int kind = L.transposed ? 4 : (L.stride == 1 ? 1 : (L.k == 3 ? 2 : 3));
      const int nchl = c.nch_a + c.nch_b;
      for (int k = 1; k < kNumVariants; ++k) {
        const Variant& v = kVariants[k];
        if (v.kind != kind || v.nch != nchl || v.ncs * 8 != c.CS || v.mb != c.MB || v.ip != c.inplace) continue;
        const Shape sh = shape_of(v.kind, v.nch, v.ncs, v.mb);
        if (sh.ncls == c.ncls && sh.N == c.N && sh.PXin == c.PXin && sh.pxs == c.px_shift && sh.RYin == c.RYin && sh.nsub == c.nsub &&
            sh.subc == c.sub_cells && sh.nchp == c.nchp && sh.TX == c.TX && sh.TY == c.TY && sh.omul == c.org_mul &&
            sh.ooff == c.org_off && (sh.trans != 0) == (c.transposed != 0))
          variant = k;
      }
    }
    const int kernel_regs = kernel_regs_of[variant];
    // persistent CTAs: as many as fit an SM (shared memory, 512 TMEM columns, 16 x 128 threads), each walks its items
    c.n_views = n_views;
    // (registers, shared memory with its 1 KB per-block reserve, 512 TMEM columns, 16 x 128 threads)
    int per_sm = 65536 / (kThreads * ((kernel_regs + 7) / 8 * 8));
    if (per_sm > (int)((227 * 1024) / (smems[l] + 1024))) per_sm = (int)((227 * 1024) / (smems[l] + 1024));
    if (per_sm > 512 / c.tmem_cols) per_sm = 512 / c.tmem_cols;
    if (per_sm > 16) per_sm = 16;
    if (per_sm < 1) per_sm = 1;
    const int items = c.tiles_x * c.tiles_y * n_views, cap = per_sm * sm_count_current();
    cfg.gridDim = dim3((unsigned)(items < cap ? items : cap));
    cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smems[l]; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = tuning().tc_no_pdl ? 0 : 1;
    c.grid_dbg = (int)cfg.gridDim.x;
    c.variant_dbg = variant;
    c.dbg = tuning().unet_dbg;
    static long long* prof_buf = nullptr;
    c.prof = nullptr;
    if (tuning().unet_profile == 2) {
      if (!prof_buf) MVS_CUDA(cudaMalloc(&prof_buf, 64));
      MVS_CUDA(cudaMemsetAsync(prof_buf, 0, 64, s));
      c.prof = prof_buf;
    }
    const cudaError_t lerr = cudaLaunchKernelEx(&cfg, kVariants[variant].fn, c);
    if (lerr != cudaSuccess) {
      set_error("launch of conv2d_tc_kernel (%s) failed: %s", L.name, cudaGetErrorString(lerr));
      return MVSB200_ERR_CUDA;
    }
    if (c.prof) {
      long long h[5];
      cudaStreamSynchronize(s);
      cudaMemcpy(h, c.prof, sizeof(h), cudaMemcpyDeviceToHost);
      if (h[4] > 0)
        fprintf(stderr, "[unet-tc-prof] %-10s CTA 0: %lld items; clk per item: wait window %lld, transform %lld, MMAs %lld, drain %lld\n",
                L.name, h[4], h[0] / h[4], h[1] / h[4], h[2] / h[4], h[3] / h[4]);
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
      fprintf(stderr, "[unet-tc] %-10s %4dx%-4d C=%-3d items %5d grid %4d smem %6zu slices %d N %3d var %d  %.3f ms\n", kUnet[l].name, tp.h[l],
              tp.w[l], tp.c[l], plans[l].tiles_x * plans[l].tiles_y * n_views, plans[l].grid_dbg, smems[l], plans[l].nslices, plans[l].N, plans[l].variant_dbg, ms);
    }
    fprintf(stderr, "[unet-tc] total %.3f ms\n", total);
    for (int i = 0; i <= MVSB200_UNET_LAYERS; ++i) cudaEventDestroy(pev[i]);
  }
  return MVSB200_OK;
}
