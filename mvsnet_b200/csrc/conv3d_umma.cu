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
//
// z-fold (stride-1 convs): zf consecutive output planes share one step; their channels sit side by
// side in the MMA N dimension (N = zf * Cout), so an input plane is read once from shared memory
// for all the output planes it feeds (the A-operand reads, not the tensor pipe, bound these skinny
// GEMMs).  The B image of (input plane dz, tap kh,kw) holds W[kd = dz - j] in the columns of output
// plane j and zeros where kd falls outside the filter.
//
// TMA loader: a layer whose input needs no transform (the two layers reading the raw cost volume)
// gets its planes by cp.async.bulk.tensor: one 5-D tiled tensor map (8 ch, chunk, W, H, D) whose box
// (8, 1, PX, RY, 1) lands exactly as one chunk plane of cells; out-of-range halo cells are zero
// filled by the TMA unit, stride-2 layers use element strides (2, 2) per parity sub-array.
#include "common.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace mvsb200 {
using namespace umma;

constexpr int kMaxOps = 108;            // 27 taps x (64 channels / 16)
constexpr int kEpiThreads = 128, kLoadThreads = 256;
constexpr int kThreads = kEpiThreads + 32 + kLoadThreads;
constexpr int kMaxRing = 12;
constexpr int kLoadBatch = 4;            // 16-byte loads in flight per loader thread
constexpr int kMaxMB = 4;                // 128-row blocks per step

enum { MODE_CONV1 = 0, MODE_CONV2 = 1, MODE_DECONV = 2 };

// Pre-baked descriptor words of one MMA (read from the constant bank with a uniform index):
struct UmmaOp {
  uint32_t a_lo;   // A descriptor low word relative to the slot: [0,14) a_off>>4 | [16,30) a_lbo>>4
  uint32_t b_lo;   // B descriptor low word relative to the B image: [0,14) b_off>>4 | [16,30) b_lbo>>4
  uint32_t meta;   // [0,16) tmem column offset in the block | [16] first (overwrite) | [20,24) dz
  uint32_t pad;
};

// host-side description of the two K halves of an op, consumed by the weight pack kernel
struct PackOp { int16_t tap[2]; int16_t cbase[2]; };

constexpr int kMaxSpan = 6;     // input planes a step may touch (zf + 2 with zf <= 4)

struct ConvParams {
  alignas(64) CUtensorMap tmap;  // input tensor map (use_tma only)
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
  int xstep, xoff, yoff, zstep, zoff, span;   // zstep = input planes advanced per step (= xstep * zf)
  int NCH, PS, slot_bytes, R;
  int MB, NB, CP;                // blocks per step, TMEM columns per block, padded channels per MMA
 int nops, b_bytes, tmem_cols;
  int zf;                        // output planes per step (z-fold), 1 unless stride-1 conv
  int cn_shift;                  // log2(cout_n) when the z-fold is on (cout_n is a power of two then)
  int use_tma;                   // planes arrive by cp.async.bulk.tensor instead of the loader warps
  int dz_begin[kMaxSpan + 1];    // ops [dz_begin[d], dz_begin[d+1]) read input plane d of the step
  int dbg;                       // development switches (env MVSB200_UMMA_DBG): 1 no global loads, 2 no MMA, 4 no stores
  UmmaOp ops[kMaxOps];
};

// ---------------------------------------------------------------------------------------------
// weight packing: TF fp32 kernel -> bf16 B images, one [2 halves][CP rows][8] block per op
// ---------------------------------------------------------------------------------------------
struct PackParams {
  const float* kernel_tf; uint16_t* out; int Cin, Cout, cout_base, cout_n, CP, transposed, nops, zf;
  PackOp ops[kMaxOps];   // tap code = dz*9 + kh*3 + kw; column group j (output plane j of the step) uses tap - 9*j
};

__global__ void pack_weights_kernel(const __grid_constant__ PackParams p) {
  const int total = p.nops * 2 * p.CP * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int k8 = i & 7, n = (i >> 3) % p.CP, half = (i / (8 * p.CP)) & 1, op = i / (16 * p.CP);
    const int j = n / p.cout_n, cn = n - j * p.cout_n;
    int tap = p.ops[op].tap[half], ci = p.ops[op].cbase[half] + k8;
    if (tap >= 0) tap -= 9 * j;
    float w = 0.0f;
    if (tap >= 0 && tap < 27 && j < p.zf && ci < p.Cin) {
      int co = p.cout_base + cn;
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
  // layout: [B image][R slots][cell table][4*Cin floats][op table][plane op ranges][barriers][tmem ptr]
  unsigned char* s_b = smem;
  unsigned char* s_slots = smem + p.b_bytes;
  const int ncells = p.nsub * p.RY * p.PX;
  int2* s_cells = reinterpret_cast<int2*>(s_slots + (size_t)p.R * p.slot_bytes);   // {offset in plane | -1, cell}
  float* s_aff = reinterpret_cast<float*>(s_cells + ((ncells + 1) & ~1));
  uint4* s_ops = reinterpret_cast<uint4*>(s_aff + 4 * p.Cin);                      // {a_lo, b_lo (absolute), tmem col, accumulate}
  int* s_dzb = reinterpret_cast<int*>(s_ops + kMaxOps);                            // op range of every input plane of a step
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_dzb + 8);
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
  const int nsteps = ze > zb ? (ze - zb + p.zf - 1) / p.zf : 0;
  const int TXe = min(p.TX, p.Mx - x0), TYe = min(p.TY, p.My - y0);
  const int nplanes = nsteps > 0 ? p.zstep * (nsteps - 1) + p.span : 0;

  for (int i = threadIdx.x; i < p.Cin; i += blockDim.x) {
    s_aff[i] = p.xs ? p.xs[i] : 1.0f;
    s_aff[p.Cin + i] = p.xb ? p.xb[i] : 0.0f;
    s_aff[2 * p.Cin + i] = p.ss ? p.ss[i] : 1.0f;
    s_aff[3 * p.Cin + i] = p.sb ? p.sb[i] : 0.0f;
  }
  // MMA op table with the B image address folded in: the issuing thread reads one 16-byte record per op
  {
    const uint32_t b16 = smem_u32(s_b) >> 4;
    for (int i = threadIdx.x; i < p.nops; i += blockDim.x) {
      const UmmaOp e = p.ops[i];
      s_ops[i] = make_uint4(e.a_lo, e.b_lo + b16, e.meta & 0xFFFFu, ((e.meta >> 16) & 1u) ^ 1u);
    }
    if (threadIdx.x <= kMaxSpan) s_dzb[threadIdx.x] = p.dz_begin[threadIdx.x];
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
    for (int i = 0; i < p.R; ++i) { mbar_init(&bar_full[i], p.use_tma ? 1 : kLoadThreads); mbar_init(&bar_empty[i], 1); }
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
      if (p.use_tma) {
        // one elected thread streams the planes with cp.async.bulk.tensor (zero-filled halo, no thread work)
        if (warp == 5 && elect_one()) {
          const uint32_t plane_bytes = (uint32_t)(p.nsub * p.NCH * p.RY * p.PX * 16);
          for (int seq = 0; seq < nplanes; ++seq) {
            const int slot = seq % p.R;
            if (seq >= p.R) mbar_wait(&bar_empty[slot], (uint32_t)((seq / p.R) - 1) & 1u);
            unsigned char* sl = s_slots + (size_t)slot * p.slot_bytes;
            const int iz = p.xstep * zb + p.zoff + seq;      // xstep = input voxels per GEMM-row-space voxel (1 or 2)
            mbar_arrive_expect_tx(&bar_full[slot], plane_bytes);
            for (int sub = 0; sub < p.nsub; ++sub)
              for (int ch = 0; ch < p.NCH; ++ch)
                tma_load_5d(sl + (size_t)ch * p.PS + (size_t)sub * p.SUBP * 16, &p.tmap, 0, ch,
                            p.xstep * x0 + p.xoff + (sub & 1), p.xstep * y0 + p.yoff + (sub >> 1), iz,
                            &bar_full[slot]);
          }
        }
      } else {
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
        const int iz = p.xstep * zb + p.zoff + seq;      // xstep = input voxels per GEMM-row-space voxel (1 or 2)
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
        if (elect_one()) {
          if (!(p.dbg & 2)) {
            // plane-major, op-major order.  One 16-byte shared-memory record per op (prefetched by the
            // unrolled loop); the slot base of a plane is added once per op, the MB row blocks of an op
            // reuse its descriptors (A start + 2 KB, next TMEM column group).
            const uint32_t d_base = tmem_base + (uint32_t)(stage * p.MB * p.NB);
            int sl_idx = slot_lo;
            for (int dz = 0; dz < p.span; ++dz) {
              const uint32_t sl = slots16 + (uint32_t)sl_idx * slot16;
              if (++sl_idx == p.R) sl_idx = 0;
              const int ob = s_dzb[dz], oe = s_dzb[dz + 1];
              if (p.MB == 1) {
#pragma unroll 6
                for (int o = ob; o < oe; ++o) {
                  const uint4 e = s_ops[o];
                  mma_bf16(d_base + e.z, desc_hi | (uint64_t)(e.x + sl), desc_hi | (uint64_t)e.y, idesc, e.w);
                }
              } else {
#pragma unroll 2
                for (int o = ob; o < oe; ++o) {
                  const uint4 e = s_ops[o];
                  uint32_t a_lo = e.x + sl;
                  uint32_t d_col = d_base + e.z;
                  const uint64_t db = desc_hi | (uint64_t)e.y;
                  for (int b = 0; b < p.MB; ++b) {
                    mma_bf16(d_col, desc_hi | (uint64_t)a_lo, db, idesc, e.w);
                    a_lo += 2048u >> 4;
                    d_col += (uint32_t)p.NB;
                  }
                }
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
      const bool deconv = p.mode == MODE_DECONV;
      const int ncls = deconv ? 8 : 1;
      const int ncol = p.zf * p.cout_n;
      const size_t zpitch = (size_t)p.Ho * p.Wo;
      for (int t = 0; t < nsteps; ++t) {
        const int stage = t & 1;
        mbar_wait(&bar_acc_full[stage], (uint32_t)(t >> 1) & 1u);
        tc_fence_after();
        const int mz = zb + t * p.zf;
        const int nlive = min(p.zf, ze - mz);          // output planes of this step inside the segment
        for (int b = 0; b < p.MB; ++b) {
          const int m = b * 128 + warp * 32 + lane;
          const int yy = m / p.PX, xx = m - yy * p.PX;
          const bool valid = xx < TXe && yy < TYe && !(p.dbg & 4);
          const int voff = deconv ? 2 * (y0 + yy) * p.Wo + 2 * (x0 + xx) : (y0 + yy) * p.Wo + x0 + xx;
          for (int cls = 0; cls < ncls; ++cls) {
            uint32_t r[CP];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) +
                                   (uint32_t)((stage * p.MB + b) * p.NB + cls * CP);
#pragma unroll
            for (int c0 = 0; c0 < CP; c0 += 16) tmem_ld16(taddr + c0, r + c0);
            tmem_ld_wait();
            if (!valid) continue;
            // columns [j*cout_n, (j+1)*cout_n) belong to output plane mz + j (z-fold); planes past the
            // segment end are computed but neither stored nor counted.  Unused columns are exact zeros.
            if (nlive == p.zf) {
#pragma unroll
              for (int k = 0; k < CP; ++k) {
                const float v = __uint_as_float(r[k]);
                sum[k] += v; sq[k] = fmaf(v, v, sq[k]);
              }
            } else {
#pragma unroll
              for (int k = 0; k < CP; ++k) {
                const float v = (k >> p.cn_shift) < nlive ? __uint_as_float(r[k]) : 0.0f;
                sum[k] += v; sq[k] = fmaf(v, v, sq[k]);
              }
            }
            size_t vox;
            if (deconv) vox = (size_t)(2 * mz + (cls >> 2)) * zpitch + voff + ((cls >> 1) & 1) * p.Wo + (cls & 1);
            else vox = (size_t)mz * zpitch + voff;
            if (p.y_is_f32) {
              float* yo = reinterpret_cast<float*>(p.y) + vox * p.Cout + p.cout_base;
#pragma unroll
              for (int k = 0; k < CP; ++k) {
                const int j = p.zf == 1 ? 0 : (k >> p.cn_shift), cn = k - (p.zf == 1 ? 0 : (j << p.cn_shift));
                if (k < ncol && j < nlive) yo[(size_t)j * zpitch * p.Cout + cn] = __uint_as_float(r[k]);
              }
            } else {
              __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(p.y) + vox * p.Cout + p.cout_base;
              if ((p.cout_n & 7) == 0) {
#pragma unroll
                for (int k = 0; k < CP; k += 8) {
                  const int j = p.zf == 1 ? 0 : (k >> p.cn_shift), cn = k - (p.zf == 1 ? 0 : (j << p.cn_shift));
                  if (k < ncol && j < nlive) {
                    uint4 pk;
                    pk.x = pack_bf16x2(__uint_as_float(r[k]), __uint_as_float(r[k + 1]));
                    pk.y = pack_bf16x2(__uint_as_float(r[k + 2]), __uint_as_float(r[k + 3]));
                    pk.z = pack_bf16x2(__uint_as_float(r[k + 4]), __uint_as_float(r[k + 5]));
                    pk.w = pack_bf16x2(__uint_as_float(r[k + 6]), __uint_as_float(r[k + 7]));
                    *reinterpret_cast<uint4*>(yo + (size_t)j * zpitch * p.Cout + cn) = pk;
                  }
                }
              } else {
#pragma unroll
                for (int k = 0; k < CP; ++k) {
                  const int j = p.zf == 1 ? 0 : (k >> p.cn_shift), cn = k - (p.zf == 1 ? 0 : (j << p.cn_shift));
                  if (k < ncol && j < nlive) yo[(size_t)j * zpitch * p.Cout + cn] = __float2bfloat16_rn(__uint_as_float(r[k]));
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
          if (lane == 0 && k < ncol) {
            const int cn = p.zf == 1 ? k : (k & (p.cout_n - 1));
            atomicAdd(p.stats + p.cout_base + cn, (double)s);
            atomicAdd(p.stats + p.Cout + p.cout_base + cn, (double)q);
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
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Plan {
  ConvParams cp;
  PackParams pp;
  size_t smem;
};

int pow2_at_least(int v) { int p = 32; while (p < v) p <<= 1; return p; }

// Build the op table for one (mode, Cin, CP) and the slot geometry for tile (TX, TY).
bool build_plan(int mode, int D, int H, int W, int cin, int cout, int cout_base, int cout_n, int TX, int TY, int zf,
                bool use_tma, Plan* pl) {
  ConvParams& c = pl->cp;
  PackParams& pk = pl->pp;
  if (mode != MODE_CONV1) zf = 1;
  if (zf * cout_n > 32 || zf + 2 > kMaxSpan) return false;
  const int CP = zf * cout_n <= 16 ? 16 : 32;
  c.CP = CP;
  if (zf > 1 && (cout_n & (cout_n - 1)) != 0) return false;   // the epilogue splits the folded columns by shift
  c.zf = zf; c.use_tma = use_tma ? 1 : 0;
  c.cn_shift = 0;
  while ((1 << c.cn_shift) < cout_n) ++c.cn_shift;
  c.mode = mode; c.D = D; c.H = H; c.W = W; c.Cin = cin; c.Cout = cout; c.cout_base = cout_base; c.cout_n = cout_n;
  int pbd = 0, pbh = 0, pbw = 0;
  if (mode == MODE_CONV1) {
    c.Do = D; c.Ho = H; c.Wo = W; c.Mz = D; c.My = H; c.Mx = W;
    c.PX = TX + 2; c.RY = TY + 2; c.nsub = 1; c.xstep = 1; c.xoff = -1; c.yoff = -1; c.zstep = zf; c.zoff = -1;
    c.span = zf + 2; c.R = c.span + 1; c.NB = CP;
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
  c.SUBP = (c.RY * c.PX + 7) / 8 * 8;          // sub-arrays start 128-byte aligned (TMA destination)
  c.NCH = cin / 8;
  c.MB = ceil_div(TY * c.PX, 128);
  if (2 * c.MB * c.NB > 512 || c.MB > kMaxMB) return false;
  c.tmem_cols = pow2_at_least(2 * c.MB * c.NB);

  // ---- op table -----------------------------------------------------------------------------------
  struct Tap { int dz, pos, widx, cls; };
  Tap taps[9 * kMaxSpan];
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
    // CONV1: one tap per input plane dz of the step and (kh, kw); its tap code dz*9+kh*3+kw selects
    // W[kd = dz - j] for output plane j in the pack kernel.  CONV2: dz = kd.
    const int ndz = mode == MODE_CONV1 ? zf + 2 : 3;
    for (int kd = 0; kd < ndz; ++kd)
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
  // plane-major order (the issue loop walks the input planes of a step and their op ranges)
  for (int i = 1; i < ntaps; ++i)
    for (int j = i; j > 0 && taps[j].dz < taps[j - 1].dz; --j) { Tap t = taps[j]; taps[j] = taps[j - 1]; taps[j - 1] = t; }
  int max_pos = 0;
  for (int i = 0; i < ntaps; ++i) max_pos = taps[i].pos > max_pos ? taps[i].pos : max_pos;
  const int sp_cells = max_pos + c.MB * 128 + 8;
  int ps = sp_cells * 16;
  if (use_tma) {
    ps = (ps + 127) / 128 * 128;                     // TMA destinations are 128-byte aligned
  } else if (c.NCH >= 2) {                           // conflict-free chunk-interleaved loader stores
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
  if (ntaps * (cin >= 16 ? cin / 16 : 1) > kMaxOps) return false;
  if (cin >= 16) {
    for (int i = 0; i < ntaps; ++i)
      for (int j = 0; j < cin / 16; ++j) {
        const bool first = !seen_cls[taps[i].cls];
        seen_cls[taps[i].cls] = true;
        add_op(taps[i].dz, (uint32_t)(2 * j * c.PS + taps[i].pos * 16), (uint32_t)c.PS, taps[i].cls * CP, first,
               taps[i].widx, 16 * j, taps[i].widx, 16 * j + 8);
      }
  } else {
    // Cin == 8: K = 16 pairs two taps of the same plane and class (second half = first shifted by LBO)
    bool used[9 * kMaxSpan] = {false};
    for (int i = 0; i < ntaps; ++i) {
      if (used[i]) continue;
      used[i] = true;
      int mate = -1;
      for (int j = i + 1; j < ntaps; ++j)
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
  // op ranges per input plane of the step (ops are in plane-major order)
  for (int d = 0; d <= kMaxSpan; ++d) c.dz_begin[d] = nops;
  for (int o = nops - 1; o >= 0; --o) c.dz_begin[(c.ops[o].meta >> 20) & 15] = o;
  for (int d = kMaxSpan - 1; d >= 0; --d) if (c.dz_begin[d] > c.dz_begin[d + 1]) c.dz_begin[d] = c.dz_begin[d + 1];
  pk.zf = zf;
  pk.nops = nops; pk.Cin = cin; pk.Cout = cout; pk.cout_base = cout_base; pk.cout_n = cout_n; pk.CP = CP;
  pk.transposed = mode == MODE_DECONV;
  const int ncells = c.nsub * c.RY * c.PX;
  const size_t fixed = (size_t)c.b_bytes + (size_t)((ncells + 1) & ~1) * sizeof(int2) +
                       (size_t)4 * cin * sizeof(float) + (size_t)kMaxOps * 16 + 32 + (2 * kMaxRing + 5) * sizeof(uint64_t) + 16;
  // deepen the ring while shared memory allows: more planes in flight hide the L2 / HBM latency
  while (c.R < kMaxRing && c.R < c.span + 2 * c.zstep &&
         fixed + (size_t)(c.R + 1) * c.slot_bytes <= kSmemBudget) ++c.R;
  pl->smem = fixed + (size_t)c.R * c.slot_bytes;
  return c.slot_bytes < (1 << 18) && (size_t)c.PS < (1u << 18);
}

}  // namespace

size_t conv3d_umma_scratch_bytes(int cin, int cout, int transposed) {
  (void)transposed; (void)cin; (void)cout;
  return align_up((size_t)kMaxOps * 2 * 32 * 16, 256) * 2;
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
  // planes of a layer that needs no input transform are fetched by the TMA unit
  const bool use_tma = !xs && !skip && !getenv("MVSB200_UMMA_NO_TMA");
  static PFN_encodeTiled encode_tiled = nullptr;
  if (use_tma && !encode_tiled) {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    MVS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      set_error("conv3d(bf16/tcgen05): cuTensorMapEncodeTiled is not available from the driver");
      return MVSB200_ERR_CUDA;
    }
    encode_tiled = (PFN_encodeTiled)fn;
  }
  const char* zf_env = getenv("MVSB200_UMMA_ZF");
  int launch_idx = 0;
  for (int cb = 0; cb < cout; cb += 32, ++launch_idx) {
    const int cn = cout - cb < 32 ? cout - cb : 32;
    // tile search: the z-fold, tile width and rows that maximise useful MMA rows x SM utilisation
    Plan best;
    bool found = false;
    double best_score = -1.0;
    const int Mx = mode == MODE_CONV2 ? ceil_div(W, 2) : W, My = mode == MODE_CONV2 ? ceil_div(H, 2) : H,
              Mz = mode == MODE_CONV2 ? ceil_div(D, 2) : D;
    const int tx_cands[] = {72, 48, 36, 32, 24, 18, 16, 12, 8};
    const int zf_cands[] = {4, 2, 1};
    for (int zi = 0; zi < 3; ++zi) {
      const int zf = zf_cands[zi];
      if (mode != MODE_CONV1 && zf != 1) continue;
      if (zf_env && atoi(zf_env) != zf && mode == MODE_CONV1) continue;
      if (zf > 1 && !zf_env && (Mz < 4 * zf)) continue;
      for (int ti = 0; ti < 9; ++ti) {
        int TX = tx_cands[ti] < Mx ? tx_cands[ti] : Mx;
        if (mode == MODE_CONV2 && use_tma && 2 * (TX + 2) > 256) continue;     // TMA box limit with element stride 2
        for (int TY = 1; TY <= 32 && TY <= My; ++TY) {
          Plan pl;
          if (!build_plan(mode, D, H, W, cin, cout, cb, cn, TX, TY, zf, use_tma, &pl)) continue;
          if (pl.smem > kSmemBudget) continue;
          const ConvParams& c = pl.cp;
          if (use_tma && (c.PX * c.xstep > 256 || c.RY * c.xstep > 256)) continue;
          // useful fraction of MMA rows x SM utilisation x halo efficiency x A-operand reads saved by the z-fold
          const double row_eff = ((double)Mx * My / ((double)c.tiles_x * c.tiles_y)) / (c.MB * 128.0);
          const int tiles = c.tiles_x * c.tiles_y;
          int zsplit = 1;
          while (tiles * zsplit < sm_count && Mz / (zsplit + 1) >= 4 * zf) ++zsplit;
          const int ctas = tiles * zsplit;
          const double waves = (double)ctas / sm_count;
          const double sm_eff = waves / (double)((ctas + sm_count - 1) / sm_count);
          const double halo = (double)(TX * TY) / ((double)c.PX * c.RY);
          const double fold = (double)(3 * zf) / (zf + 2);                     // A reads per output plane saved
          const double ring = c.R >= c.span + c.zstep ? 1.0 : 0.7;             // loads overlap the MMAs only then
          const double score = row_eff * sm_eff * (0.5 + 0.5 * halo) * (0.5 + 0.5 * fold / 2.0) * ring;
          if (score > best_score) { best_score = score; best = pl; best.cp.zsplit = zsplit; found = true; }
        }
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
    if (use_tma) {
      // tensor (8 ch, chunk, W, H, D) bf16; box = one chunk plane of PX x RY cells (element stride xstep in W, H)
      cuuint64_t gdim[5] = {8, (cuuint64_t)c.NCH, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D};
      cuuint64_t gstr[4] = {16, (cuuint64_t)cin * 2, (cuuint64_t)W * cin * 2, (cuuint64_t)H * W * cin * 2};
      cuuint32_t box[5] = {8, 1, (cuuint32_t)(c.PX * c.xstep), (cuuint32_t)(c.RY * c.xstep), 1};
      cuuint32_t estr[5] = {1, 1, (cuuint32_t)c.xstep, (cuuint32_t)c.xstep, 1};
      CUresult cr = encode_tiled(&c.tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) {
        set_error("conv3d(bf16/tcgen05): cuTensorMapEncodeTiled failed with %d (PX=%d RY=%d step=%d)", (int)cr, c.PX, c.RY,
                  c.xstep);
        return MVSB200_ERR_CUDA;
      }
    }
    {
      const char* dbg = getenv("MVSB200_UMMA_DBG");
      c.dbg = dbg ? atoi(dbg) : 0;
      if (getenv("MVSB200_UMMA_VERBOSE"))
        fprintf(stderr, "[umma] mode=%d Cin=%d Cout=%d(+%d) tile %dx%d PX=%d MB=%d R=%d zf=%d tma=%d zsplit=%d grid=%d smem=%zu nops=%d\n",
                mode, cin, cn, cb, c.TX, c.TY, c.PX, c.MB, c.R, c.zf, c.use_tma, c.zsplit,
                c.tiles_x * c.tiles_y * c.zsplit, best.smem, c.nops);
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
