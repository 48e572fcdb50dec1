// Kernel 3: regularizer layers as bf16 implicit GEMM on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// One persistent CTA owns an (y, x) tile of the GEMM-row space and marches along z.  Input planes
// live in a shared-memory ring in "chunk-planar position" layout: for every 8-channel chunk a dense
// array of 16-byte cells, one per (row, col) position of the haloed tile.  In that layout the
// 128 x 16 A operand of ANY filter tap is a plain no-swizzle K-major UMMA descriptor over the same
// bytes: start = cell of the tap-shifted first row, SBO = 128 B (8 consecutive cells), LBO = chunk
// plane stride -- so the im2col matrix is never built and each input voxel is staged exactly once
// per (tile, z).  GEMM rows run over the linearised padded tile (row pitch PX), rows that fall
// into the halo columns are computed and dropped.
//
//   warps 0-3  epilogue: tcgen05.ld accumulators -> bf16/fp32 store + per-channel batch statistics
//   warp  4    TMEM allocation; one thread issues every tcgen05.mma and the commits
//   warps 5-12 loaders: global -> (BN scale/shift + ReLU + skip add of the producers) -> bf16 cells
//
// The three layer kinds of RegNetUS0 are all "tap GEMMs" over such planes:
//   conv s=1 (network.py:210)   27 taps, plane z-1..z+1, cell offset kh*PX+kw
//   conv s=2 (TF SAME)          input split into 4 (y,x)-parity sub-arrays so stride-2 rows are dense
//   deconv s=2 (network.py:327) 8 output-parity classes, each with its 1/2/4/8 taps, own TMEM columns
// Weights are pre-packed (pack kernel below) into the matching K-major B images and fetched with
// one bulk async copy (TMA engine) per CTA.
#include "common.cuh"
#include "umma.cuh"
#include <stdlib.h>

namespace mvsb200 {
using namespace umma;

constexpr int kMaxOps = 108;            // 27 taps x (64 channels / 16)
constexpr int kEpiThreads = 128, kLoadThreads = 256;
constexpr int kThreads = kEpiThreads + 32 + kLoadThreads;
constexpr int kMaxRing = 8;
constexpr int kLoadBatch = 4;            // 16-byte loads in flight per loader thread

enum { MODE_CONV1 = 0, MODE_CONV2 = 1, MODE_DECONV = 2 };

// Pre-baked descriptor words of one MMA (read from the constant bank with a uniform index):
struct UmmaOp {
  uint32_t a_lo;   // A descriptor low word relative to the slot: [0,14) a_off>>4 | [16,30) a_lbo>>4
  uint32_t b_lo;   // B descriptor low word relative to the B image: [0,14) b_off>>4 | [16,30) b_lbo>>4
  uint32_t meta;   // [0,16) tmem column offset in the block | [16] first (overwrite) | [20,22) dz
  uint32_t pad;
};

// host-side description of the two K halves of an op, consumed by the weight pack kernel
struct PackOp { int16_t tap[2]; int16_t cbase[2]; };

struct ConvParams {
  const __nv_bfloat16* x; const __nv_bfloat16* skip;
  const float *xs, *xb, *ss, *sb;
  const uint4* wpacked;
  void* y; double* stats;
  int mode, y_is_f32;
  int D, H, W, Cin;              // input volume
  int Do, Ho, Wo, Cout;          // output volume (all channels)
  int cout_base, cout_n;         // channel slice handled by this launch
  int Mz, My, Mx;                // GEMM-row space (output voxels; input voxels for deconv)
  int TX, TY, tiles_x, tiles_y, zsplit;
  int PX, RY, nsub, SUBP;        // slot geometry: nsub sub-arrays of RY x PX cells
  int xstep, xoff, yoff, zstep, zoff, span;
  int NCH, PS, slot_bytes, R;
  int MB, NB, CP;                // blocks per step, TMEM columns per block, padded channels per MMA
  int nops, b_bytes, tmem_cols;
  int dbg;                       // development switches (env MVSB200_UMMA_DBG): 1 no global loads, 2 no MMA, 4 no stores
  UmmaOp ops[kMaxOps];
};

// ---------------------------------------------------------------------------------------------
// weight packing: TF fp32 kernel -> bf16 B images, one [2 halves][CP rows][8] block per op
// ---------------------------------------------------------------------------------------------
struct PackParams {
  const float* kernel_tf; uint16_t* out; int Cin, Cout, cout_base, cout_n, CP, transposed, nops;
  PackOp ops[kMaxOps];
};

__global__ void pack_weights_kernel(const __grid_constant__ PackParams p) {
  const int total = p.nops * 2 * p.CP * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int k8 = i & 7, n = (i >> 3) % p.CP, half = (i / (8 * p.CP)) & 1, op = i / (16 * p.CP);
    int tap = p.ops[op].tap[half], ci = p.ops[op].cbase[half] + k8;
    float w = 0.0f;
    if (tap >= 0 && n < p.cout_n && ci < p.Cin) {
      int co = p.cout_base + n;
      w = p.transposed ? p.kernel_tf[((size_t)tap * p.Cout + co) * p.Cin + ci]
                       : p.kernel_tf[((size_t)tap * p.Cin + ci) * p.Cout + co];
    }
    __nv_bfloat16 h = __float2bfloat16_rn(w);
    p.out[i] = *reinterpret_cast<uint16_t*>(&h);
  }
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(h);
}

template <int CP>
__global__ void __launch_bounds__(kThreads, 1) conv3d_umma_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  // layout: [B image][R slots][cell table][op table][4*Cin floats][barriers][tmem ptr]
  unsigned char* s_b = smem;
  unsigned char* s_slots = smem + p.b_bytes;
  const int ncells = p.nsub * p.RY * p.PX;
  int2* s_cells = reinterpret_cast<int2*>(s_slots + (size_t)p.R * p.slot_bytes);   // {offset in plane | -1, cell}
  float* s_aff = reinterpret_cast<float*>(s_cells + ((ncells + 1) & ~1));
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_aff + 4 * p.Cin);
  uint64_t* bar_full = bars;                    // [R]   loaders -> MMA
  uint64_t* bar_empty = bars + kMaxRing;        // [R]   MMA (commit) -> loaders
  uint64_t* bar_acc_full = bars + 2 * kMaxRing; // [2]   MMA (commit) -> epilogue
  uint64_t* bar_acc_empty = bar_acc_full + 2;   // [2]   epilogue -> MMA
  uint64_t* bar_b = bar_acc_empty + 2;          // [1]   weights landed
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_b + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile coordinates
  int bid = blockIdx.x;
  const int tx = bid % p.tiles_x; bid /= p.tiles_x;
  const int ty = bid % p.tiles_y; bid /= p.tiles_y;
  const int zs = bid;
  const int x0 = tx * p.TX, y0 = ty * p.TY;
  const int zseg = (p.Mz + p.zsplit - 1) / p.zsplit;
  const int zb = zs * zseg, ze = min(p.Mz, zb + zseg);
  const int nsteps = ze - zb;
  const int TXe = min(p.TX, p.Mx - x0), TYe = min(p.TY, p.My - y0);
  const int nplanes = nsteps > 0 ? p.zstep * (nsteps - 1) + p.span : 0;

  for (int i = threadIdx.x; i < p.Cin; i += blockDim.x) {
    s_aff[i] = p.xs ? p.xs[i] : 1.0f;
    s_aff[p.Cin + i] = p.xb ? p.xb[i] : 0.0f;
    s_aff[2 * p.Cin + i] = p.ss ? p.ss[i] : 1.0f;
    s_aff[3 * p.Cin + i] = p.sb ? p.sb[i] : 0.0f;
  }
  // loop-invariant loader addressing: one entry per cell of a slot
  for (int i = threadIdx.x; i < ncells; i += blockDim.x) {
    const int c = i % p.PX;
    int rest = i / p.PX;
    const int r = rest % p.RY, sub = rest / p.RY;
    const int ix = p.xstep * (x0 + c) + p.xoff + (sub & 1);
    const int iy = p.xstep * (y0 + r) + p.yoff + (sub >> 1);
    const bool ok = ix >= 0 && ix < p.W && iy >= 0 && iy < p.H;
    s_cells[i] = make_int2(ok ? (iy * p.W + ix) * p.Cin : -1, sub * p.SUBP + r * p.PX + c);
  }
  // Every cell of the ring starts finite: halo rows of the GEMM read a few cells past the loaded
  // area (their results are dropped, but 0 * NaN from stale shared memory must not reach a zero-
  // weighted K half of a valid row).
  for (int i = threadIdx.x; i < p.R * p.slot_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(s_slots)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.R; ++i) { mbar_init(&bar_full[i], kLoadThreads); mbar_init(&bar_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], kEpiThreads / 32); }
    mbar_init(bar_b, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(s_tmem, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (nsteps > 0) {
    if (warp >= 5) {
      // ===================================== loaders =====================================
      const int lt = threadIdx.x - (kEpiThreads + 32);
      const bool x_act = p.xs != nullptr, has_skip = p.skip != nullptr, s_act = p.ss != nullptr;
      const bool transform = x_act || has_skip;
      const int items = ncells * p.NCH;
      const int nch_shift = (p.NCH & (p.NCH - 1)) == 0 ? __ffs(p.NCH) - 1 : -1;
      const size_t plane_elems = (size_t)p.H * p.W * p.Cin;
      for (int seq = 0; seq < nplanes; ++seq) {
        const int slot = seq % p.R;
        if (seq >= p.R) mbar_wait(&bar_empty[slot], (uint32_t)((seq / p.R) - 1) & 1u);
        unsigned char* sl = s_slots + (size_t)slot * p.slot_bytes;
        const int iz = p.zstep * zb + p.zoff + seq;
        const bool zok = iz >= 0 && iz < p.D;
        const __nv_bfloat16* xp = p.x + (size_t)(zok ? iz : 0) * plane_elems;
        const __nv_bfloat16* kp = has_skip ? p.skip + (size_t)(zok ? iz : 0) * plane_elems : nullptr;
        for (int base = 0; base < items; base += kLoadThreads * kLoadBatch) {
          uint4 v[kLoadBatch], sv[kLoadBatch];
          int dst[kLoadBatch], chn[kLoadBatch];
          bool live[kLoadBatch];
#pragma unroll
          for (int u = 0; u < kLoadBatch; ++u) {
            const int i = base + u * kLoadThreads + lt;
            const bool in = i < items;
            const int cell = nch_shift >= 0 ? (i >> nch_shift) : (i / p.NCH);
            const int ch = i - cell * p.NCH;
            const int2 e = in ? s_cells[cell] : make_int2(-1, 0);
            chn[u] = ch;
            dst[u] = in ? ch * p.PS + e.y * 16 : -1;
            live[u] = in && zok && e.x >= 0 && !(p.dbg & 1);
            v[u] = make_uint4(0u, 0u, 0u, 0u);
            sv[u] = make_uint4(0u, 0u, 0u, 0u);
            if (live[u]) {
              v[u] = __ldg(reinterpret_cast<const uint4*>(xp + e.x + ch * 8));
              if (has_skip) sv[u] = __ldg(reinterpret_cast<const uint4*>(kp + e.x + ch * 8));
            }
          }
#pragma unroll
          for (int u = 0; u < kLoadBatch; ++u) {
            if (dst[u] < 0) continue;
            if (live[u] && transform) {
              uint32_t* vw = reinterpret_cast<uint32_t*>(&v[u]);
              const uint32_t* sw = reinterpret_cast<const uint32_t*>(&sv[u]);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int cc = chn[u] * 8 + 2 * k;
                float2 f = unpack_bf16x2(vw[k]);
                if (x_act) {
                  f.x = fmaxf(fmaf(f.x, s_aff[cc], s_aff[p.Cin + cc]), 0.0f);
                  f.y = fmaxf(fmaf(f.y, s_aff[cc + 1], s_aff[p.Cin + cc + 1]), 0.0f);
                }
                if (has_skip) {
                  float2 g = unpack_bf16x2(sw[k]);
                  if (s_act) {
                    g.x = fmaxf(fmaf(g.x, s_aff[2 * p.Cin + cc], s_aff[3 * p.Cin + cc]), 0.0f);
                    g.y = fmaxf(fmaf(g.y, s_aff[2 * p.Cin + cc + 1], s_aff[3 * p.Cin + cc + 1]), 0.0f);
                  }
                  f.x += g.x; f.y += g.y;
                }
                vw[k] = pack_bf16x2(f.x, f.y);
              }
            }
            *reinterpret_cast<uint4*>(sl + dst[u]) = v[u];
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&bar_full[slot]);
      }
    } else if (warp == 4) {
      // ===================================== MMA issuer =====================================
      // The whole warp walks the loop with uniform control flow (descriptor words come from the
      // constant bank and loop counters, i.e. the uniform datapath); one elected lane issues.
      if (elect_one()) {
        // weights: one bulk async copy (TMA engine) in <= 32 KB pieces
        mbar_arrive_expect_tx(bar_b, (uint32_t)p.b_bytes);
        for (int off = 0; off < p.b_bytes; off += 32768) {
          int n = min(32768, p.b_bytes - off);
          bulk_g2s(s_b + off, reinterpret_cast<const unsigned char*>(p.wpacked) + off, (uint32_t)n, bar_b);
        }
      }
      __syncwarp();
      mbar_wait(bar_b, 0);
      const uint32_t idesc = make_idesc_bf16_f32(128, CP);
      const uint32_t slots16 = smem_u32(s_slots) >> 4, slot16 = (uint32_t)p.slot_bytes >> 4;
      const uint32_t b16 = smem_u32(s_b) >> 4;
      const uint64_t desc_hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;   // version 1, SBO = 128 B
      int waited = 0, wslot = 0;
      uint32_t wphase = 0;
      int slot_lo = 0;                                  // ring slot of plane seq_lo
      for (int t = 0; t < nsteps; ++t) {
        const int seq_lo = p.zstep * t, seq_hi = seq_lo + p.span - 1;
        while (waited <= seq_hi) {
          mbar_wait(&bar_full[wslot], wphase);
          ++waited;
          if (++wslot == p.R) { wslot = 0; wphase ^= 1u; }
        }
        const int stage = t & 1;
        mbar_wait(&bar_acc_empty[stage], ((uint32_t)(t >> 1) & 1u) ^ 1u);
        tc_fence_after();
        int s1 = slot_lo + 1; if (s1 >= p.R) s1 -= p.R;
        int s2 = s1 + 1; if (s2 >= p.R) s2 -= p.R;
        const uint32_t sl0 = slots16 + (uint32_t)slot_lo * slot16;
        const uint32_t sl1 = slots16 + (uint32_t)s1 * slot16;
        const uint32_t sl2 = slots16 + (uint32_t)s2 * slot16;
        if (elect_one()) {
          if (!(p.dbg & 2)) {
            // op-major order: the descriptor words of an op are formed once and reused for all MB
            // row blocks (start address + 2 KB, next TMEM column group)
            const uint32_t d_base = tmem_base + (uint32_t)(stage * p.MB * p.NB);
#pragma unroll 2
            for (int o = 0; o < p.nops; ++o) {
              const UmmaOp e = p.ops[o];
              const uint32_t dz = e.meta >> 20;
              const uint32_t sl = dz == 0 ? sl0 : (dz == 1 ? sl1 : sl2);
              uint32_t a_lo = e.a_lo + sl;
              uint32_t d_col = d_base + (e.meta & 0xFFFFu);
              const uint64_t db = desc_hi | (uint64_t)(e.b_lo + b16);
              const uint32_t acc = ((e.meta >> 16) & 1u) ^ 1u;
#pragma unroll 4
              for (int b = 0; b < p.MB; ++b) {
                mma_bf16(d_col, desc_hi | (uint64_t)a_lo, db, idesc, acc);
                a_lo += 2048u >> 4;
                d_col += (uint32_t)p.NB;
              }
            }
          }
          mma_commit(&bar_acc_full[stage]);
          // planes no later step needs go back to the loaders
          int rs = slot_lo;
          for (int s = seq_lo; s < min(p.zstep * (t + 1), nplanes); ++s) {
            mma_commit(&bar_empty[rs]);
            if (++rs == p.R) rs = 0;
          }
        }
        __syncwarp();
        slot_lo += p.zstep;
        if (slot_lo >= p.R) slot_lo -= p.R;
      }
    } else {
      // ===================================== epilogue =====================================
      float sum[CP], sq[CP];
#pragma unroll
      for (int k = 0; k < CP; ++k) { sum[k] = 0.0f; sq[k] = 0.0f; }
      const int ncls = p.mode == MODE_DECONV ? 8 : 1;
      for (int t = 0; t < nsteps; ++t) {
        const int stage = t & 1;
        mbar_wait(&bar_acc_full[stage], (uint32_t)(t >> 1) & 1u);
        tc_fence_after();
        const int mz = zb + t;
        for (int b = 0; b < p.MB; ++b) {
          const int m = b * 128 + warp * 32 + lane;
          const int yy = m / p.PX, xx = m - yy * p.PX;
          const bool valid = xx < TXe && yy < TYe && !(p.dbg & 4);
          for (int cls = 0; cls < ncls; ++cls) {
            uint32_t r[CP];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) +
                                   (uint32_t)((stage * p.MB + b) * p.NB + cls * CP);
#pragma unroll
            for (int c0 = 0; c0 < CP; c0 += 16) tmem_ld16(taddr + c0, r + c0);
            tmem_ld_wait();
            if (valid) {
              int oz, oy, ox;
              if (p.mode == MODE_DECONV) {
                oz = 2 * mz + (cls >> 2); oy = 2 * (y0 + yy) + ((cls >> 1) & 1); ox = 2 * (x0 + xx) + (cls & 1);
              } else { oz = mz; oy = y0 + yy; ox = x0 + xx; }
              const size_t vox = ((size_t)oz * p.Ho + oy) * p.Wo + ox;
#pragma unroll
              for (int k = 0; k < CP; ++k) {
                const float v = __uint_as_float(r[k]);
                sum[k] += v; sq[k] = fmaf(v, v, sq[k]);
              }
              if (p.y_is_f32) {
                float* yo = reinterpret_cast<float*>(p.y) + vox * p.Cout + p.cout_base;
#pragma unroll
                for (int k = 0; k < CP; ++k)
                  if (k < p.cout_n) yo[k] = __uint_as_float(r[k]);
              } else {
                __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(p.y) + vox * p.Cout + p.cout_base;
                if ((p.cout_n & 7) == 0) {
#pragma unroll
                  for (int k = 0; k < CP; k += 8) {
                    if (k < p.cout_n) {
                      uint4 pk;
                      pk.x = pack_bf16x2(__uint_as_float(r[k]), __uint_as_float(r[k + 1]));
                      pk.y = pack_bf16x2(__uint_as_float(r[k + 2]), __uint_as_float(r[k + 3]));
                      pk.z = pack_bf16x2(__uint_as_float(r[k + 4]), __uint_as_float(r[k + 5]));
                      pk.w = pack_bf16x2(__uint_as_float(r[k + 6]), __uint_as_float(r[k + 7]));
                      *reinterpret_cast<uint4*>(yo + k) = pk;
                    }
                  }
                } else {
#pragma unroll
                  for (int k = 0; k < CP; ++k)
                    if (k < p.cout_n) yo[k] = __float2bfloat16_rn(__uint_as_float(r[k]));
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_acc_empty[stage]);
      }
      // batch statistics: per-thread partials -> warp reduce -> one double atomic per channel and warp
      if (p.stats) {
#pragma unroll
        for (int k = 0; k < CP; ++k) {
          float s = sum[k], q = sq[k];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
          }
          if (lane == 0 && k < p.cout_n) {
            atomicAdd(p.stats + p.cout_base + k, (double)s);
            atomicAdd(p.stats + p.Cout + p.cout_base + k, (double)q);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host-side planning
// ---------------------------------------------------------------------------------------------
namespace {

constexpr size_t kSmemBudget = 225 * 1024;

struct Plan {
  ConvParams cp;
  PackParams pp;
  size_t smem;
};

int pow2_at_least(int v) { int p = 32; while (p < v) p <<= 1; return p; }

// Build the op table for one (mode, Cin, CP) and the slot geometry for tile (TX, TY).
bool build_plan(int mode, int D, int H, int W, int cin, int cout, int cout_base, int cout_n, int TX, int TY,
                Plan* pl) {
  ConvParams& c = pl->cp;
  PackParams& pk = pl->pp;
  const int CP = cout_n <= 16 ? 16 : 32;
  c.CP = CP;
  c.mode = mode; c.D = D; c.H = H; c.W = W; c.Cin = cin; c.Cout = cout; c.cout_base = cout_base; c.cout_n = cout_n;
  int pbd = 0, pbh = 0, pbw = 0;
  if (mode == MODE_CONV1) {
    c.Do = D; c.Ho = H; c.Wo = W; c.Mz = D; c.My = H; c.Mx = W;
    c.PX = TX + 2; c.RY = TY + 2; c.nsub = 1; c.xstep = 1; c.xoff = -1; c.yoff = -1; c.zstep = 1; c.zoff = -1;
    c.span = 3; c.R = 4; c.NB = CP;
  } else if (mode == MODE_CONV2) {
    c.Do = ceil_div(D, 2); c.Ho = ceil_div(H, 2); c.Wo = ceil_div(W, 2); c.Mz = c.Do; c.My = c.Ho; c.Mx = c.Wo;
    pbd = tf_same_pad_before(D, 3, 2); pbh = tf_same_pad_before(H, 3, 2); pbw = tf_same_pad_before(W, 3, 2);
    c.PX = TX + 1 + pbw; c.RY = TY + 1 + pbh; c.nsub = 4; c.xstep = 2; c.xoff = -2 * pbw; c.yoff = -2 * pbh;
    c.zstep = 2; c.zoff = -pbd; c.span = 3; c.R = 5; c.NB = CP;
  } else {
    c.Do = 2 * D; c.Ho = 2 * H; c.Wo = 2 * W; c.Mz = D; c.My = H; c.Mx = W;
    c.PX = TX + 1; c.RY = TY + 1; c.nsub = 1; c.xstep = 1; c.xoff = -1; c.yoff = -1; c.zstep = 1; c.zoff = -1;
    c.span = 2; c.R = 3; c.NB = 8 * CP;
  }
  c.TX = TX; c.TY = TY;
  c.tiles_x = ceil_div(c.Mx, TX); c.tiles_y = ceil_div(c.My, TY);
  c.SUBP = c.RY * c.PX;
  c.NCH = cin / 8;
  c.MB = ceil_div(TY * c.PX, 128);
  if (2 * c.MB * c.NB > 512) return false;
  c.tmem_cols = pow2_at_least(2 * c.MB * c.NB);

  // ---- op table -----------------------------------------------------------------------------------
  struct Tap { int dz, pos, widx, cls; };
  Tap taps[27];
  int ntaps = 0;
  if (mode == MODE_DECONV) {
    for (int cls = 0; cls < 8; ++cls) {
      const int pz = cls >> 2, py = (cls >> 1) & 1, px = cls & 1;
      for (int sz = 0; sz >= (pz ? 0 : -1); --sz)
        for (int sy = 0; sy >= (py ? 0 : -1); --sy)
          for (int sx = 0; sx >= (px ? 0 : -1); --sx) {
            const int kd = pz - 2 * sz, kh = py - 2 * sy, kw = px - 2 * sx;
            taps[ntaps++] = {1 + sz, (1 + sy) * c.PX + (1 + sx), (kd * 3 + kh) * 3 + kw, cls};
          }
    }
  } else {
    for (int kd = 0; kd < 3; ++kd)
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          int pos;
          if (mode == MODE_CONV1) pos = kh * c.PX + kw;
          else {
            const int ey = kh - pbh, ex = kw - pbw;
            const int qy = ey & 1, qx = ex & 1;
            const int fy = (ey - qy) / 2, fx = (ex - qx) / 2;
            pos = ((qy << 1) | qx) * c.SUBP + (fy + pbh) * c.PX + (fx + pbw);
          }
          taps[ntaps++] = {kd, pos, (kd * 3 + kh) * 3 + kw, 0};
        }
  }
  if (ntaps != 27) return false;
  int max_pos = 0;
  for (int i = 0; i < 27; ++i) max_pos = taps[i].pos > max_pos ? taps[i].pos : max_pos;
  const int sp_cells = max_pos + c.MB * 128 + 8;
  int ps = sp_cells * 16;
  if (c.NCH >= 2) {                                  // conflict-free chunk-interleaved loader stores
    const int want = 128 / c.NCH;
    ps = (ps + 127) / 128 * 128 + want;
  } else {
    ps = (ps + 127) / 128 * 128;
  }
  c.PS = ps;
  c.slot_bytes = c.NCH * c.PS;

  int nops = 0;
  const int b_op_bytes = 2 * CP * 16;
  auto add_op = [&](int dz, uint32_t a_off, uint32_t a_lbo, int col, bool first, int tap0, int cb0, int tap1, int cb1) {
    UmmaOp& o = c.ops[nops];
    o.a_lo = (a_off >> 4) | ((a_lbo >> 4) << 16);
    o.b_lo = (uint32_t)((nops * b_op_bytes) >> 4) | (((uint32_t)CP * 16u >> 4) << 16);
    o.meta = (uint32_t)col | ((first ? 1u : 0u) << 16) | ((uint32_t)dz << 20);
    o.pad = 0;
    pk.ops[nops].tap[0] = (int16_t)tap0; pk.ops[nops].cbase[0] = (int16_t)cb0;
    pk.ops[nops].tap[1] = (int16_t)tap1; pk.ops[nops].cbase[1] = (int16_t)cb1;
    ++nops;
  };
  bool seen_cls[8] = {false, false, false, false, false, false, false, false};
  if (cin >= 16) {
    for (int i = 0; i < 27; ++i)
      for (int j = 0; j < cin / 16; ++j) {
        const bool first = !seen_cls[taps[i].cls];
        seen_cls[taps[i].cls] = true;
        add_op(taps[i].dz, (uint32_t)(2 * j * c.PS + taps[i].pos * 16), (uint32_t)c.PS, taps[i].cls * CP, first,
               taps[i].widx, 16 * j, taps[i].widx, 16 * j + 8);
      }
  } else {
    // Cin == 8: K = 16 pairs two taps of the same plane and class (second half = first shifted by LBO)
    bool used[27] = {false};
    for (int i = 0; i < 27; ++i) {
      if (used[i]) continue;
      used[i] = true;
      int mate = -1;
      for (int j = i + 1; j < 27; ++j)
        if (!used[j] && taps[j].dz == taps[i].dz && taps[j].cls == taps[i].cls && taps[j].pos > taps[i].pos) {
          mate = j; break;
        }
      const bool first = !seen_cls[taps[i].cls];
      seen_cls[taps[i].cls] = true;
      if (mate >= 0) {
        used[mate] = true;
        add_op(taps[i].dz, (uint32_t)(taps[i].pos * 16), (uint32_t)((taps[mate].pos - taps[i].pos) * 16),
               taps[i].cls * CP, first, taps[i].widx, 0, taps[mate].widx, 0);
      } else {
        add_op(taps[i].dz, (uint32_t)(taps[i].pos * 16), 16u, taps[i].cls * CP, first, taps[i].widx, 0, -1, 0);
      }
    }
  }
  c.nops = nops;
  c.b_bytes = nops * b_op_bytes;
  pk.nops = nops; pk.Cin = cin; pk.Cout = cout; pk.cout_base = cout_base; pk.cout_n = cout_n; pk.CP = CP;
  pk.transposed = mode == MODE_DECONV;
  const int ncells = c.nsub * c.RY * c.PX;
  const size_t fixed = (size_t)c.b_bytes + (size_t)((ncells + 1) & ~1) * sizeof(int2) +
                       (size_t)4 * cin * sizeof(float) + (2 * kMaxRing + 5) * sizeof(uint64_t) + 16;
  // deepen the ring while shared memory allows: more planes in flight hide the L2 / HBM latency
  while (c.R < kMaxRing && fixed + (size_t)(c.R + 1) * c.slot_bytes <= kSmemBudget) ++c.R;
  pl->smem = fixed + (size_t)c.R * c.slot_bytes;
  return c.slot_bytes < (1 << 18) && (size_t)c.PS < (1u << 18);
}

}  // namespace

size_t conv3d_umma_scratch_bytes(int cin, int cout, int transposed) {
  (void)transposed;
  const int nops_max = cin >= 16 ? 27 * (cin / 16) : 27;
  return align_up((size_t)nops_max * 2 * 32 * 16, 256) * 2;
}

int launch_conv3d_umma(const void* x, int x_dtype, const float* xs, const float* xb, const void* skip,
                       const float* ss, const float* sb, const float* kernel_tf, int D, int H, int W, int cin,
                       int cout, int stride, int transposed, void* y, int y_dtype, double* stats, void* scratch,
                       cudaStream_t s) {
  if (x_dtype != MVSB200_BF16) {
    set_error("conv3d(bf16/tcgen05): input must be bf16");
    return MVSB200_ERR_UNSUPPORTED;
  }
  if (cin % 8 != 0 || cin < 8 || cin > 64 || (cin > 8 && cin % 16 != 0)) {
    set_error("conv3d(bf16/tcgen05): Cin=%d unsupported (need 8, 16, 32, 48 or 64)", cin);
    return MVSB200_ERR_UNSUPPORTED;
  }
  const int mode = transposed ? MODE_DECONV : (stride == 2 ? MODE_CONV2 : MODE_CONV1);
  // temporary scratch when the caller has none (single-layer entry point): a static device buffer
  static void* s_scratch = nullptr;
  if (!scratch) {
    if (!s_scratch) MVS_CUDA(cudaMalloc(&s_scratch, 1 << 20));
    scratch = s_scratch;
  }
  int sm_count = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  static bool attr_done = false;
  if (!attr_done) {
    MVS_CUDA(cudaFuncSetAttribute(conv3d_umma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    MVS_CUDA(cudaFuncSetAttribute(conv3d_umma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    attr_done = true;
  }
  int launch_idx = 0;
  for (int cb = 0; cb < cout; cb += 32, ++launch_idx) {
    const int cn = cout - cb < 32 ? cout - cb : 32;
    // tile search: prefer wide x tiles, the most rows that fit TMEM and shared memory
    Plan best;
    bool found = false;
    double best_score = -1.0;
    const int Mx = mode == MODE_CONV2 ? ceil_div(W, 2) : W, My = mode == MODE_CONV2 ? ceil_div(H, 2) : H,
              Mz = mode == MODE_CONV2 ? ceil_div(D, 2) : D;
    const int tx_cands[] = {72, 48, 36, 32, 24, 18, 16, 12, 8};
    for (int ti = 0; ti < 9; ++ti) {
      int TX = tx_cands[ti] < Mx ? tx_cands[ti] : Mx;
      for (int TY = 1; TY <= 32 && TY <= My; ++TY) {
        Plan pl;
        if (!build_plan(mode, D, H, W, cin, cout, cb, cn, TX, TY, &pl)) continue;
        if (pl.smem > kSmemBudget) continue;
        const ConvParams& c = pl.cp;
        // useful fraction of MMA rows x SM utilisation x halo efficiency
        const double row_eff = (double)(c.tiles_x * c.tiles_y ? (double)Mx * My / ((double)c.tiles_x * c.tiles_y) : 0) /
                               (c.MB * 128.0);
        const int tiles = c.tiles_x * c.tiles_y;
        int zsplit = 1;
        while (tiles * zsplit < sm_count && Mz / (zsplit + 1) >= 4) ++zsplit;
        const int ctas = tiles * zsplit;
        const double waves = (double)ctas / sm_count;
        const double sm_eff = waves / (double)((ctas + sm_count - 1) / sm_count);
        const double halo = (double)(TX * TY) / ((double)c.PX * c.RY * (mode == MODE_CONV2 ? 1.0 : 1.0));
        const double score = row_eff * sm_eff * (0.5 + 0.5 * halo);
        if (score > best_score) { best_score = score; best = pl; best.cp.zsplit = zsplit; found = true; }
      }
    }
    if (!found) {
      set_error("conv3d(bf16/tcgen05): no tile fits (Cin=%d Cout=%d mode=%d)", cin, cout, mode);
      return MVSB200_ERR_UNSUPPORTED;
    }
    ConvParams& c = best.cp;
    c.x = (const __nv_bfloat16*)x; c.skip = (const __nv_bfloat16*)skip;
    c.xs = xs; c.xb = xb; c.ss = ss; c.sb = sb;
    c.y = y; c.stats = stats; c.y_is_f32 = y_dtype == MVSB200_F32;
    {
      const char* dbg = getenv("MVSB200_UMMA_DBG");
      c.dbg = dbg ? atoi(dbg) : 0;
      if (getenv("MVSB200_UMMA_VERBOSE"))
        fprintf(stderr, "[umma] mode=%d Cin=%d Cout=%d(+%d) tile %dx%d PX=%d MB=%d R=%d zsplit=%d grid=%d smem=%zu nops=%d\n", mode,
                cin, cn, cb, c.TX, c.TY, c.PX, c.MB, c.R, c.zsplit, c.tiles_x * c.tiles_y * c.zsplit, best.smem, c.nops);
    }
    unsigned char* wp = (unsigned char*)scratch + (size_t)(launch_idx & 1) * align_up((size_t)kMaxOps * 2 * 32 * 16, 256);
    c.wpacked = (const uint4*)wp;
    best.pp.kernel_tf = kernel_tf;
    best.pp.out = (uint16_t*)wp;
    pack_weights_kernel<<<ceil_div(c.nops * 2 * c.CP * 8, 256), 256, 0, s>>>(best.pp);
    MVS_LAUNCH_CHECK("pack_weights_kernel");
    const int grid = c.tiles_x * c.tiles_y * c.zsplit;
    if (c.CP == 16) conv3d_umma_kernel<16><<<grid, kThreads, best.smem, s>>>(c);
    else conv3d_umma_kernel<32><<<grid, kThreads, best.smem, s>>>(c);
    MVS_LAUNCH_CHECK("conv3d_umma_kernel");
  }
  return MVSB200_OK;
}

}  // namespace mvsb200
