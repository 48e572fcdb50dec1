// Kernel 3: regularizer layers as bf16 implicit GEMM on the 5th-gen tensor cores (tcgen05 + TMEM),
// fed by the TMA unit.
//
// Activation layouts in HBM (bf16, written by the producer's epilogue, read by TMA boxes):
//   CP8  "chunk planar"   [D][C/8][H][W][8]          input of stride-1 convs and transposed convs
//   PS8  "parity split"   [D][C/8][4][Hs][Ws][8]     input of stride-2 convs; sub-plane q = (y&1)*2+(x&1)
//                                                    holds voxel (2*ys+(q>>1), 2*xs+(q&1)), Hs=ceil(H/2)
// In both a (z, chunk[, parity]) plane is a dense 2-D array of 16-byte cells, so one TMA box of
// PX x RY cells lands in shared memory exactly as the "cell plane" the MMA descriptors want, halo
// cells outside the volume zero-filled by the TMA unit (= SAME padding).
//
// One persistent CTA owns an (y, x) tile of the GEMM-row space and marches along z.  Input planes
// live in a shared-memory ring, per 8-channel chunk a dense array of 16-byte cells, one per position
// of the haloed tile.  In that layout the 128 x 16 A operand of ANY filter tap is a plain no-swizzle
// K-major UMMA descriptor over the same bytes (start = cell of the tap-shifted first row, SBO = 128 B,
// LBO = chunk plane stride): im2col is never built, each input voxel is staged once per (tile, z).
// GEMM rows run over the linearised padded tile (row pitch PX); rows in the halo columns are dropped.
//
//   warps 0-3   epilogue: tcgen05.ld accumulators -> bf16/fp32 stores + per-channel batch statistics
//   warps 4-11  transform (layers whose input needs it): relu(x*scale+shift) [+ relu(skip*..+..)]
//               applied IN PLACE on the landed plane, halo cells left at zero
//   warp  12    producer: one thread issues the TMA box loads (input planes, skip planes, weights)
//   warp  13    TMEM allocation; one thread issues every tcgen05.mma and the commits (the highest warp id:
//               the issue arbiter favours it over the busy transform warps)
//
// Layer kinds (all "tap GEMMs" over such planes):
//   conv s=1 (network.py:210)   taps (kh,kw) of plane dz, cell offset kh*PX+kw
//   conv s=2 (TF SAME)          PS8 input: the 4 parity sub-arrays make stride-2 rows dense
//   deconv s=2 (network.py:327) 8 output-parity classes, each with its 1/2/4/8 taps, own TMEM columns
// x-fold (stride-1 convs): the three kw taps of a filter row are folded into the MMA N dimension as well.  The
// padded tile is exactly 8 / 16 / 32 cells wide, GEMM row (yy, xx) = input column x0 + xx - 1, and column
// group kw of that row holds the partial sum that belongs to output column xx - (kw - 1); the epilogue adds
// the three partials with two intra-warp shuffles (a warp's 32 TMEM lanes are whole tile rows, so the
// neighbours are always in the warp).  One MMA per (plane, kh, 16 channels) instead of three: the A operand is
// read from shared memory a third as often, which is what bounds these skinny GEMMs.
// z-fold (stride-1 convs): zf consecutive output planes share one step, their channels side by side
// in the MMA N dimension (N = zf*Cout): an input plane is read from shared memory once for all the
// output planes it feeds.  Measured on B200: a 128xNx16 MMA from shared memory costs
// max(32 + N/4, N/2) cycles (operand reads at 128 B/clk), so small-N GEMMs are bound by the A reads
// and the fold is free.  The B images of the zf+2 input planes are shifted windows of one master
// image per (kh, kw, 16-channel pair): rows [g*Cout, (g+1)*Cout) hold W[kd = zf+1-g] (zero outside
// the filter), plane dz starts at g = zf+1-dz.
#include "common.cuh"
#include "umma.cuh"
#include "conv3d_tc.h"
#include <cuda.h>
#include <stdlib.h>
#include <array>
#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

namespace mvsb200 {
using namespace umma;

// Batch-norm source of an input tensor: the producer's channel statistics (sum | sum of squares over `count`
// voxels) and its gamma / beta.  The consumer derives scale / shift itself (no bn_finalize launch in between).
namespace tc {

constexpr int kMaxOps = 108;            // 27 taps x (64 channels / 16)
// Per-role cycle counters of MVSB200_TC_PROF live behind a build switch (-DMVSB200_TC_PROF_BUILD=1): even as a
// not-taken branch per plane they were ~5 % of the transform warps' instructions.  Without it MVSB200_TC_PROF=1 still
// prints every launch's stand-alone time.
#ifndef MVSB200_TC_PROF_BUILD
#define MVSB200_TC_PROF_BUILD 0
#endif
constexpr bool kProf = MVSB200_TC_PROF_BUILD != 0;
constexpr int kEpiWarps = 4, kXfWarps = 8;
constexpr int kXfThreads = kXfWarps * 32;
constexpr int kThreads = (kEpiWarps + 2 + kXfWarps) * 32;
constexpr int kXfWarp0 = kEpiWarps, kProdWarp = kEpiWarps + kXfWarps, kMmaWarp = kProdWarp + 1;
constexpr int kMaxRing = 12, kMaxSkipRing = 8, kMinSkipRing = 2, kMaxSpan = 6, kMaxMB = 4, kMaxK = 12;

enum { MODE_CONV1 = 0, MODE_CONV2 = 1, MODE_DECONV = 2 };

// Descriptor words of one MMA, relative to the slot / B image:
struct UmmaOp {
  uint32_t a_lo;   // [0,14) a_off>>4 | [16,30) a_lbo>>4
  uint32_t b_lo;   // [0,14) b_off>>4 | [16,30) b_lbo>>4
  uint32_t meta;   // [0,16) tmem column offset in the block | [16] first (overwrite) | [20,24) dz |
                   // [24,30) (N of the launch - N of this op) / 8: an op spans only the column blocks it has weights for
  uint32_t pad;
};

struct Params {
  alignas(64) CUtensorMap tmap_x;   // input planes
  alignas(64) CUtensorMap tmap_s;   // skip planes (has_skip)
  const float *xs, *xb, *ss, *sb;
  TcBnSrc xbn, sbn;                 // used instead of xs/xb, ss/sb when .stats is set
  const uint4* wpacked;
  __nv_bfloat16* y_cp8; __nv_bfloat16* y_ps8; float* y_f32; double* stats;
  int stats_reps, stats_rep_stride;   // CTAs spread their atomics over `stats_reps` partial copies of the statistics
  // D-slab mode with peer memory (NVLink): boundary planes of the output are ALSO stored into the neighbours'
  // halo planes, [0] = chunk-planar, [1] = parity-split copy (NULL: no neighbour / not in that mode)
  __nv_bfloat16* mir_prev[2];       // previous rank's AFTER-halo plane of this output tensor
  __nv_bfloat16* mir_next[2];       // next rank's BEFORE-halo plane
  // ... and the kernel waits, before it reads anything, until every rank has published the layers it consumes
  const unsigned* wait_flags[2];    // per producer layer: `wait_n` words, one per rank (NULL: nothing to wait for)
  int wait_n; unsigned wait_seq; unsigned* err_flag;
  int mode, has_skip, transform;
  int D, H, W, Cin;              // input volume
  int Do, Ho, Wo, Cout;          // output volume (all channels)
  int Hso, Wso;                  // PS8 output sub-plane extents
  int cout_base, cout_n;         // channel slice handled by this launch
  int Mz, My, Mx;                // GEMM-row space (output voxels; input voxels for deconv)
  int TX, TY, tiles_x, tiles_y, zsplit;
  int PX, RY, nsub, SUBP;        // slot geometry: nsub sub-arrays of RY x PX cells, SUBP cells apart
  int vstep, cx_off, cy_off;     // cell (sub, r, c) = voxel (vstep*(y0+r+cy_off)+(sub>>1), vstep*(x0+c+cx_off)+(sub&1))
  int zv_lo, zv_hi;              // input planes outside [zv_lo, zv_hi) are padding: the transform leaves them at zero
  int zmul, zoff, zstep, span;   // plane seq of a segment = input z (zmul*zb + zoff + seq); zstep planes per step
  int NCH, PS, slot_bytes, R, RS;    // RS = skip ring depth (has_skip)
  int MB, NB, CP;                // row blocks per step, TMEM columns per block, MMA N
  int nops, b_bytes, tmem_cols;
  int zf, cn_shift;              // output planes per step; log2(cout_n) when zf > 1
  int xfold;                     // kw taps folded into N: columns [j][kw][co], see the epilogue
  int mma_n;                     // N of every MMA of this launch
  int cw, dmerge;                // transposed conv: TMEM columns per output-parity class; classes merged into one N
  int one_box;                   // chunk planes are dense in the slot: one TMA box loads all chunks of a plane
  int ring_pad;                  // zeroed bytes after the last slot (rows of the last block may read past their plane)
  int xf_k;                      // cells per transform thread and plane
  int xf_groups;                 // transform warp groups, each taking every xf_groups-th plane (1 or 2)
  int dz_begin[kMaxSpan + 1];    // ops [dz_begin[d], dz_begin[d+1]) read input plane d of the step
  float* rg_out; float rg_start, rg_step;   // fused soft-argmin partials of the last layer (NULL: off)
  // "rider" (3dconv0_1 + 3dconv1_0 in one pass over the cost volume): a stride-2 conv with the same input is the
  // stride-1 conv of its filter evaluated at the odd (z, y, x) positions only (TF SAME padding, even extents), so its
  // n2 output channels ride along as extra MMA columns of the odd output plane of a 2-plane step
  int n2, Cout2, Ho2, Wo2, Hso2, Wso2;      // rider channels of this launch (0: off), its whole output volume
  __nv_bfloat16* y2_cp8; __nv_bfloat16* y2_ps8; double* stats2;
  int pdl;                       // programmatic dependent launch: 1 = let the next layer start at CTA start, 2 = at CTA end
  int dbg;                       // development switches (env MVSB200_TC_DBG): 1 no loads, 2 no MMA, 4 no stores
  long long* prof;               // development: per-role cycle counters of CTA 0 (env MVSB200_TC_PROF)
  UmmaOp ops[kMaxOps];
};

// ---------------------------------------------------------------------------------------------
// weight packing: TF fp32 kernel -> bf16 B images
// ---------------------------------------------------------------------------------------------
struct PackOp { int16_t tap[2]; int16_t cbase[2]; };
struct PackParams {
  const float* kernel_tf; uint16_t* out;
  int Cin, Cout, cout_base, cout_n, CP, transposed, nops, zf, master, xfold, dmerge, cw;
  int CinT, CoutT;       // channel counts of kernel_tf itself (<= Cin / Cout: channels padded to whole cells hold zeros)
  const float* kernel_tf2; int n2, CoutT2;   // rider: filter [3,3,3,CinT,CoutT2] of the stride-2 conv, n2 columns per kw
  PackOp ops[kMaxOps];   // per-op images: tap = kd*9+kh*3+kw per K half (-1 = zero half);
                         // master images: tap = kh*3+kw (kd comes from the row group), one per (kh,kw,pair)
};

// Rider launch (3dconv0_1 + 3dconv1_0): image of an op = [2 halves][CP rows][8]; row n = column n of the step:
// [even plane: kw x n1 channels of the stride-1 filter][odd plane: kw x (n1 channels of the stride-1 filter | n2 of the
// stride-2 filter)]; the op's tap code is dz*9 + kh*3 of its input plane dz (kw lives in the column, kd = dz - j).
__device__ __forceinline__ void pack_rider(const PackParams& p) {
  const int n1 = p.cout_n, g0 = 3 * n1, g1 = 3 * (n1 + p.n2);
  const int total = p.nops * 2 * p.CP * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k8 = i & 7, n = (i >> 3) % p.CP, half = (i / (8 * p.CP)) & 1, op = i / (16 * p.CP);
    const int ci = p.ops[op].cbase[half] + k8;
    int j = 0, kw = 0, c = -1;
    if (n < g0) { kw = n / n1; c = n - kw * n1; }
    else if (n < g0 + g1) { j = 1; const int r = n - g0; kw = r / (n1 + p.n2); c = r - kw * (n1 + p.n2); }
    const int tap = p.ops[op].tap[half] + kw - 9 * j;
    float w = 0.0f;
    if (c >= 0 && tap >= 0 && tap < 27 && ci < p.CinT) {
      if (c < n1) { if (p.cout_base + c < p.CoutT) w = p.kernel_tf[((size_t)tap * p.CinT + ci) * p.CoutT + p.cout_base + c]; }
      else if (c - n1 < p.CoutT2) w = p.kernel_tf2[((size_t)tap * p.CinT + ci) * p.CoutT2 + (c - n1)];
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(w);
    p.out[i] = *reinterpret_cast<const uint16_t*>(&h);
  }
}

__global__ void pack_weights_kernel(const __grid_constant__ PackParams p) {
  if (p.n2) { pack_rider(p); return; }
  if (p.dmerge) {
    // transposed conv, classes merged: image = [2 halves][8 classes x cw rows][8]; the op's tap code is its input
    // shift (bit 2: z-1, bit 1: y-1, bit 0: x-1); class (pz,py,px) takes filter tap k = parity + 2 on a shifted axis
    const int rows = 8 * p.cw;
    const int total = p.nops * 2 * rows * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int k8 = i & 7, n = (i >> 3) % rows, half = (i / (8 * rows)) & 1, op = i / (16 * rows);
      const int cls = n / p.cw, cn = n - cls * p.cw;
      const int sh = p.ops[op].tap[half], ci = p.ops[op].cbase[half] + k8;
      const int pz = cls >> 2, py = (cls >> 1) & 1, px = cls & 1;
      const int sz = (sh >> 2) & 1, sy = (sh >> 1) & 1, sx = sh & 1;
      float w = 0.0f;
      if ((!sz || !pz) && (!sy || !py) && (!sx || !px) && cn < p.cout_n && ci < p.CinT && p.cout_base + cn < p.CoutT) {
        const int tap = ((pz + 2 * sz) * 3 + (py + 2 * sy)) * 3 + (px + 2 * sx);
        w = p.kernel_tf[((size_t)tap * p.CoutT + p.cout_base + cn) * p.CinT + ci];
      }
      const __nv_bfloat16 h = __float2bfloat16_rn(w);
      p.out[i] = *reinterpret_cast<const uint16_t*>(&h);
    }
    return;
  }
  // column n of a B image: [output plane j of the step][kw when the x-fold is on][output channel]
  const int kwn = p.xfold ? 3 : 1, grp = kwn * p.cout_n;
  if (!p.master) {
    const int total = p.nops * 2 * p.CP * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int k8 = i & 7, n = (i >> 3) % p.CP, half = (i / (8 * p.CP)) & 1, op = i / (16 * p.CP);
      // column group j = output plane j of the step (z-fold): its filter plane is kd = dz - j
      const int j = n / grp, rem = n - j * grp, kw = rem / p.cout_n, cn = rem - kw * p.cout_n;
      int tap = p.ops[op].tap[half];
      const int ci = p.ops[op].cbase[half] + k8;
      if (tap >= 0) tap += (p.xfold ? kw : 0) - 9 * j;       // x-fold: the op's tap code has kw = 0
      float w = 0.0f;
      if (tap >= 0 && tap < 27 && j < p.zf && ci < p.CinT && p.cout_base + cn < p.CoutT) {
        const int co = p.cout_base + cn;
        w = p.transposed ? p.kernel_tf[((size_t)tap * p.CoutT + co) * p.CinT + ci]
                         : p.kernel_tf[((size_t)tap * p.CinT + ci) * p.CoutT + co];
      }
      const __nv_bfloat16 h = __float2bfloat16_rn(w);
      p.out[i] = *reinterpret_cast<const uint16_t*>(&h);
    }
  } else {
    // image m = [2 halves][(2*zf+1) groups][grp rows][8]; group g holds W[kd = zf+1-g]
    const int groups = 2 * p.zf + 1, rows = groups * grp;
    const int total = p.nops * 2 * rows * 8;      // nops = number of master images here
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int k8 = i & 7, row = (i >> 3) % rows, half = (i / (8 * rows)) & 1, m = i / (16 * rows);
      const int g = row / grp, rem = row - g * grp, kw = rem / p.cout_n, cn = rem - kw * p.cout_n, kd = p.zf + 1 - g;
      const int ci = p.ops[m].cbase[half] + k8;
      float w = 0.0f;
      if (kd >= 0 && kd < 3 && ci < p.CinT && p.cout_base + cn < p.CoutT) {
        const int tap = kd * 9 + p.ops[m].tap[0] + (p.xfold ? kw : 0);
        w = p.kernel_tf[((size_t)tap * p.CinT + ci) * p.CoutT + p.cout_base + cn];
      }
      const __nv_bfloat16 h = __float2bfloat16_rn(w);
      p.out[i] = *reinterpret_cast<const uint16_t*>(&h);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// scale / shift of channel c from batch statistics: the arithmetic of bn_finalize_kernel (conv3d_direct.cu), i.e.
// fp64 moments rounded once, then tf.nn.batch_normalization in fp32 (network.py:496-506, Appendix A.6)
__device__ __forceinline__ void bn_scale_shift(const TcBnSrc& b, int c, float& scale, float& shift) {
  if (b.channels_true > 0 && c >= b.channels_true) { scale = 0.0f; shift = 0.0f; return; }     // padding channel: stays 0
  // all loads of a batch of partial copies are issued before the first add (one L2 round trip per batch instead of one
  // per copy: measured 1.2 - 4.4 us of every launch's prologue); the order of the fp64 adds is unchanged
  const float gam = b.gamma[c], bet = b.beta[c];
  double sm = 0.0, sq = 0.0;
  for (int r0 = 0; r0 < b.reps; r0 += 8) {
    double a[8], q[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const bool in = r0 + u < b.reps;
      a[u] = in ? b.stats[(size_t)(r0 + u) * b.rep_stride + c] : 0.0;
      q[u] = in ? b.stats[(size_t)(r0 + u) * b.rep_stride + b.channels + c] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) { sm += a[u]; sq += q[u]; }
  }
  const double mean = sm / b.count;
  double var = sq / b.count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float meanf = (float)mean, varf = (float)var;
  const float inv = __fmul_rn(__fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(varf, b.eps))), gam);
  scale = inv;
  shift = __fsub_rn(bet, __fmul_rn(meanf, inv));
}

// Batch statistics of a CTA: per-thread partials -> warp reduce -> the 4 epilogue warps combine in shared memory ->
// one double atomic per channel and CTA, spread over partial copies (thousands of atomics on the same 2*C
// addresses serialise in L2 and used to cost more than the small layers themselves).
template <int NV>
__device__ __forceinline__ void flush_stats(const Params& p, float (&sum)[NV], float (&sq)[NV], int ncol, bool fold,
                                            float* s_red, int warp, int lane) {
  if (lane < 32) { s_red[(warp * 2 + 0) * 32 + lane] = 0.0f; s_red[(warp * 2 + 1) * 32 + lane] = 0.0f; }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float s = sum[k], q = sq[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0 && k < ncol) {
      const int cn = fold ? (k & (p.cout_n - 1)) : k;       // z-folded columns of the same channel
      s_red[(warp * 2 + 0) * 32 + cn] += s;
      s_red[(warp * 2 + 1) * 32 + cn] += q;
    }
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");        // the four epilogue warps
  if (warp == 0 && lane < p.cout_n) {
    float s = 0.0f, q = 0.0f;
#pragma unroll
    for (int w = 0; w < 4; ++w) { s += s_red[(w * 2 + 0) * 32 + lane]; q += s_red[(w * 2 + 1) * 32 + lane]; }
    double* st = p.stats + (size_t)(blockIdx.x % p.stats_reps) * p.stats_rep_stride;
    atomicAdd(st + p.cout_base + lane, (double)s);
    atomicAdd(st + p.Cout + p.cout_base + lane, (double)q);
  }
}

// Rider launch: three groups of four epilogue warps.  Groups 0 and 1 hold the stride-1 layer's cout_n (8) channels of
// the even / odd planes, group 2 the rider's n2 (<= 16) channels; each group reduces in its own 128 floats of shared
// memory behind its own named barrier.
template <int N2>
__device__ __forceinline__ void flush_stats_rider(const Params& p, float (&sum)[8], float (&sq)[8], float (&sum2)[N2],
                                                  float (&sq2)[N2], int grp, float* s_red, int ewarp, int lane) {
  const int ncol = grp == 2 ? p.n2 : p.cout_n;
  if (lane < 16) { s_red[(ewarp * 2 + 0) * 16 + lane] = 0.0f; s_red[(ewarp * 2 + 1) * 16 + lane] = 0.0f; }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 8 + N2; ++k) {
    float s = k < 8 ? sum[k < 8 ? k : 0] : sum2[k < 8 ? 0 : (k - 8) % N2];
    float q = k < 8 ? sq[k < 8 ? k : 0] : sq2[k < 8 ? 0 : (k - 8) % N2];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0 && k < ncol) { s_red[(ewarp * 2 + 0) * 16 + k] = s; s_red[(ewarp * 2 + 1) * 16 + k] = q; }
  }
  asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
  if (ewarp == 0 && lane < ncol) {
    float s = 0.0f, q = 0.0f;
#pragma unroll
    for (int w = 0; w < 4; ++w) { s += s_red[(w * 2 + 0) * 16 + lane]; q += s_red[(w * 2 + 1) * 16 + lane]; }
    double* st = (grp == 2 ? p.stats2 : p.stats);
    if (st) {
      st += (size_t)(blockIdx.x % p.stats_reps) * p.stats_rep_stride;
      const int cbase = grp == 2 ? 0 : p.cout_base, ctot = grp == 2 ? p.Cout2 : p.Cout;
      atomicAdd(st + cbase + lane, (double)s);
      atomicAdd(st + ctot + cbase + lane, (double)q);
    }
  }
}

// One 16-byte cell of the chunk-planar (cell index inside the plane, which = 0) or parity-split (which = 1) output;
// boundary planes go to the neighbouring slabs' halo planes as well (peer memory over NVLink).
template <bool PEER>
__device__ __forceinline__ void store_cell(const Params& p, int which, int oz, int chunk, size_t plane_cells, size_t cell,
                                           const uint4& pk) {
  __nv_bfloat16* y = which ? p.y_ps8 : p.y_cp8;
  const int ncho = p.Cout >> 3;
  *reinterpret_cast<uint4*>(y + (((size_t)oz * ncho + chunk) * plane_cells + cell) * 8) = pk;
  if (PEER) {                           // peer-memory D-slab mode only (separate instantiation)
    if (oz == 0 && p.mir_prev[which])
      *reinterpret_cast<uint4*>(p.mir_prev[which] + ((size_t)chunk * plane_cells + cell) * 8) = pk;
    if (oz == p.Do - 1 && p.mir_next[which])
      *reinterpret_cast<uint4*>(p.mir_next[which] + ((size_t)chunk * plane_cells + cell) * 8) = pk;
  }
}

// Issue the MMAs of ops [ob, oe) of one input plane: one 16-byte shared-memory record per op (the unrolled loop
// prefetches them), the MB row blocks of an op reuse its descriptors (A start + 2 KB, next TMEM column group).
template <int MB>
__device__ __forceinline__ void issue_ops(const uint4* s_ops, int ob, int oe, uint32_t sl, uint32_t d_base, uint32_t nb,
                                          uint64_t desc_hi) {
#pragma unroll(MB == 1 ? 6 : (MB == 2 ? 3 : 2))
  for (int o = ob; o < oe; ++o) {
    const uint4 e = s_ops[o];
    // record: A start | B start | accumulator column [0,16) + accumulate flag [16] | instruction descriptor (N of the op)
    const uint32_t a_lo = e.x + sl, d_col = d_base + (e.z & 0xFFFFu);
    const uint64_t db = desc_hi | (uint64_t)e.y;
#pragma unroll
    for (int b = 0; b < MB; ++b)
      mma_bf16(d_col + (uint32_t)b * nb, desc_hi | (uint64_t)(a_lo + (uint32_t)b * (2048u >> 4)), db, e.w, e.z >> 16);
  }
}

// XFC = 0: classic epilogue (CP accumulator columns per row block); XFC = 1 / 2 / 4: x-fold epilogue for launches of
// 8 * XFC output channels (Cout = 1 uses XFC = 1)
// FLAGS bit 0: D-slab mode over peer memory (flag wait in the prologue, boundary planes mirrored to the neighbours);
// bit 1: the single-channel layer (Cout = 1, fp32 output, optional fused soft-argmin); bit 2: the rider launch
// (3dconv0_1 with 3dconv1_0's channels riding on the odd planes; XFC = 2: up to 16 channels of statistics per warp).  Separate
// instantiations, so that none of them costs the common path registers or instructions.
template <int CP, int XFC, int FLAGS>
__global__ void __launch_bounds__(kThreads, 1) conv3d_tc_kernel(const __grid_constant__ Params p) {
  constexpr bool XF = XFC != 0;
  constexpr bool PEER = (FLAGS & 1) != 0, C1 = (FLAGS & 2) != 0, RIDER = (FLAGS & 4) != 0;
  extern __shared__ __align__(128) unsigned char smem[];
  // layout: [B image][R slots][skip slots][op table][plane op ranges][barriers][tmem ptr]
  unsigned char* s_b = smem;
  unsigned char* s_slots = smem + p.b_bytes;
  unsigned char* s_skip = s_slots + (size_t)p.R * p.slot_bytes + p.ring_pad;
  uint4* s_ops = reinterpret_cast<uint4*>(s_skip + (p.has_skip ? (size_t)p.RS * p.slot_bytes : 0));
  int* s_dzb = reinterpret_cast<int*>(s_ops + kMaxOps);
  float* s_red = reinterpret_cast<float*>(s_dzb + 8);      // [4 warps][sum | sumsq][32 channels]
  float* s_aff = s_red + 256;                              // [x scale | x shift | skip scale | skip shift][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_aff + 256);
  uint64_t* bar_land = bars;                        // [R]  TMA -> transform / MMA
  uint64_t* bar_ready = bars + kMaxRing;            // [R]  transform -> MMA
  uint64_t* bar_empty = bars + 2 * kMaxRing;        // [R]  MMA (commit) -> producer
  uint64_t* bar_sland = bars + 3 * kMaxRing;        // [RS] TMA -> transform
  uint64_t* bar_sempty = bar_sland + kMaxSkipRing;  // [RS] transform -> producer
  uint64_t* bar_acc_full = bar_sempty + kMaxSkipRing;  // [2]  MMA (commit) -> epilogue
  uint64_t* bar_acc_empty = bar_acc_full + 2;       // [2]  epilogue -> MMA
  uint64_t* bar_b = bar_acc_empty + 2;              // [1]  weights landed
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_b + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long pr_entry = 0;
  if (kProf && p.prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pr_entry));
  int bid = blockIdx.x;
  const int tx = bid % p.tiles_x; bid /= p.tiles_x;
  const int ty = bid % p.tiles_y; bid /= p.tiles_y;
  const int zs = bid;
  const int x0 = tx * p.TX, y0 = ty * p.TY;
  // z segments in units of steps so that only the last segment can end on a partial step
  const int steps_all = (p.Mz + p.zf - 1) / p.zf;
  const int sseg = (steps_all + p.zsplit - 1) / p.zsplit;
  const int zb = zs * sseg * p.zf, ze = min(p.Mz, zb + sseg * p.zf);
  const int nsteps = ze > zb ? (ze - zb + p.zf - 1) / p.zf : 0;
  const int TXe = min(p.TX, p.Mx - x0), TYe = min(p.TY, p.My - y0);
  const int nplanes = nsteps > 0 ? p.zstep * (nsteps - 1) + p.span : 0;

  {
    const uint32_t b16 = smem_u32(s_b) >> 4;
    for (int i = threadIdx.x; i < p.nops; i += blockDim.x) {
      const UmmaOp e = p.ops[i];
      s_ops[i] = make_uint4(e.a_lo, e.b_lo + b16, (e.meta & 0xFFFFu) | ((((e.meta >> 16) & 1u) ^ 1u) << 16),
                            make_idesc_bf16_f32(128, p.mma_n - (int)((e.meta >> 24) << 3)));
    }
    if (threadIdx.x <= kMaxSpan) s_dzb[threadIdx.x] = p.dz_begin[threadIdx.x];
  }
  // Every cell of the ring starts finite: halo rows of the GEMM read a few cells past the landed boxes
  // (their results are dropped, but 0 * NaN from stale shared memory must not reach a zero-weighted
  // K half of a valid row).
  // With dense chunk planes every byte of a slot is rewritten by the plane's TMA box before it is read (out-of-image
  // cells arrive as zeros), and a valid row never reads past its plane (x-fold, or whole K = 16 cells): only the pad
  // behind the ring has to be cleared (the clear of ~200 KB was 1 us of every CTA's prologue).
  const int clear_from = (p.one_box && (p.xfold || p.Cin >= 16)) ? p.R * p.slot_bytes / 16 : 0;
  for (int i = clear_from + threadIdx.x; i < (p.R * p.slot_bytes + p.ring_pad) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(s_slots)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  auto stamp = [&](int slot) {
    if (kProf && p.prof && blockIdx.x == 0 && threadIdx.x == 0) {
      long long g;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
      p.prof[slot] = g - pr_entry;
    }
  };
  stamp(17);                                        // op table + ring clear
  if (threadIdx.x == 0) {
    const int xfw = kXfWarps / p.xf_groups;         // transform warps per plane
    for (int i = 0; i < p.R; ++i) { mbar_init(&bar_land[i], 1); mbar_init(&bar_ready[i], xfw); mbar_init(&bar_empty[i], 1); }
    for (int i = 0; i < kMaxSkipRing; ++i) { mbar_init(&bar_sland[i], 1); mbar_init(&bar_sempty[i], xfw); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], RIDER ? kEpiWarps + kXfWarps : kEpiWarps); }
    mbar_init(bar_b, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(s_tmem, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  // Programmatic dependent launch: everything above overlapped the tail of the previous kernel in the stream; from
  // here on its outputs (statistics, activations) are read.  A single-wave grid lets the next kernel's CTAs take
  // over SMs as they free up; a multi-wave grid only triggers at CTA end (its own waiting CTAs must get the SMs).
  stamp(18);                                        // barriers initialised
  if (p.pdl) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (p.pdl == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  stamp(19);                                        // previous grid complete
  // D-slab mode over peer memory: the layers this one consumes must have been published by EVERY rank (their
  // statistics) -- which covers the two neighbours whose boundary planes landed in our halo planes.  One flag word
  // per (layer, rank) holds the sequence number of the last published inference.  The spin is bounded.
  if (PEER) {
    if (p.wait_n > 0 && threadIdx.x < 2 * p.wait_n) {
      const unsigned* f = p.wait_flags[threadIdx.x / p.wait_n];
      if (f) {
        f += threadIdx.x % p.wait_n;
        const long long t0 = clock64();
        for (;;) {
          unsigned v;
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
          if ((int)(v - p.wait_seq) >= 0) break;
          // give up when the host raised the abort word (the word after the error word: a watchdog that lost a rank
          // releases every kernel waiting for it at once) or after ~2 s
          unsigned stop = 0u;
          if (p.err_flag) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(stop) : "l"(p.err_flag + 1) : "memory");
          if (stop || clock64() - t0 > 4000000000LL) { if (p.err_flag) atomicExch(p.err_flag, 1u); break; }
          __nanosleep(200);
        }
        asm volatile("fence.proxy.async;" ::: "memory");      // the halo planes are read by the TMA unit (async proxy)
      }
    }
    __syncthreads();
  }
  {
    // BN scale / shift of the input (and skip) channels, one thread per channel (fp64 moments are slow: not per
    // transform thread)
    if (threadIdx.x < 2 * p.Cin) {
      const int c = threadIdx.x % p.Cin, which = threadIdx.x / p.Cin;
      float sc = 1.0f, sh = 0.0f;
      if (which == 0) {
        if (p.xbn.stats) bn_scale_shift(p.xbn, c, sc, sh);
        else if (p.xs) { sc = p.xs[c]; sh = p.xb[c]; }
      } else {
        if (p.sbn.stats) bn_scale_shift(p.sbn, c, sc, sh);
        else if (p.ss) { sc = p.ss[c]; sh = p.sb[c]; }
      }
      s_aff[which * 128 + c] = sc;
      s_aff[which * 128 + 64 + c] = sh;
    }
  }
  stamp(20);                                        // statistics -> scale / shift (thread 0's channel)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  stamp(21);                                        // every warp through the prologue
  const uint32_t tmem_base = *s_tmem;

  if (nsteps > 0) {
    if (warp == kProdWarp) {
      // ===================================== producer =====================================
      // The whole warp walks the planes; lane 0 arms the barriers, then the lanes issue the plane's TMA boxes
      // (chunks x parity sub-arrays, skip chunks) in parallel.
      if (lane == 0) {
        // weights: bulk async copies in <= 32 KB pieces
        mbar_arrive_expect_tx(bar_b, (uint32_t)p.b_bytes);
        for (int off = 0; off < p.b_bytes; off += 32768)
          bulk_g2s(s_b + off, reinterpret_cast<const unsigned char*>(p.wpacked) + off,
                   (uint32_t)min(32768, p.b_bytes - off), bar_b);
      }
      const uint32_t plane_bytes = (uint32_t)(p.nsub * p.NCH * p.RY * p.PX * 16);
      const int cx = (x0 + p.cx_off) * 8, cy = y0 + p.cy_off;
      const int nbox = p.nsub * p.NCH;
      long long pw_empty = 0, pw_sempty = 0, pw_t0 = 0, pw_issue = 0;
      if (kProf && p.prof) pw_t0 = clock64();
      int slot = 0, ss = 0;                 // ring positions; phases of the "empty" barriers = (lap - 1) & 1
      uint32_t ph = 1u, sph = 1u;
      for (int seq = 0; seq < nplanes; ++seq) {
        long long pa = 0;
        if (kProf && p.prof) pa = clock64();
        if (seq >= p.R) mbar_wait(&bar_empty[slot], ph);
        long long pc0 = 0;
        if (kProf && p.prof) { pc0 = clock64(); pw_empty += pc0 - pa; }
        const int iz = p.zmul * zb + p.zoff + seq;
        unsigned char* sl = s_slots + (size_t)slot * p.slot_bytes;
        if (!(p.dbg & 1)) {
          if (lane == 0) mbar_arrive_expect_tx(&bar_land[slot], plane_bytes);
          __syncwarp();
          if (p.one_box) {
            // dense chunk planes (x-fold): one box covers every channel chunk of the plane
            if (lane == 0) tma_load_5d(sl, &p.tmap_x, cx, cy, 0, 0, iz, &bar_land[slot]);      // box spans parities and chunks
          } else {
            for (int i = lane; i < nbox; i += 32) {
              const int sub = i / p.NCH, ch = i - sub * p.NCH;
              tma_load_5d(sl + (size_t)ch * p.PS + (size_t)sub * p.SUBP * 16, &p.tmap_x, cx, cy, sub, ch, iz, &bar_land[slot]);
            }
          }
        } else if (lane == 0) {
          mbar_arrive(&bar_land[slot]);
        }
        if (kProf && p.prof) pw_issue += clock64() - pc0;
        if (p.has_skip) {
          long long pb = 0;
          if (kProf && p.prof) pb = clock64();
          if (seq >= p.RS) mbar_wait(&bar_sempty[ss], sph);
          if (kProf && p.prof) pw_sempty += clock64() - pb;
          unsigned char* sk = s_skip + (size_t)ss * p.slot_bytes;
          if (!(p.dbg & 1)) {
            if (lane == 0) mbar_arrive_expect_tx(&bar_sland[ss], plane_bytes);
            __syncwarp();
            if (p.one_box) {
              if (lane == 0) tma_load_5d(sk, &p.tmap_s, cx, cy, 0, 0, iz, &bar_sland[ss]);
            } else if (lane < p.NCH) {
              tma_load_5d(sk + (size_t)lane * p.PS, &p.tmap_s, cx, cy, 0, lane, iz, &bar_sland[ss]);
            }
          } else if (lane == 0) {
            mbar_arrive(&bar_sland[ss]);
          }
          if (++ss == p.RS) { ss = 0; sph ^= 1u; }
        }
        if (++slot == p.R) { slot = 0; ph ^= 1u; }
      }
      if (kProf && p.prof && blockIdx.x == 0 && lane == 0) { p.prof[9] = clock64() - pw_t0; p.prof[10] = pw_empty; p.prof[11] = pw_sempty; p.prof[5 + 11] = pw_issue; }
    } else if (!RIDER && warp >= kXfWarp0 && warp < kProdWarp) {
      // ===================================== transform =====================================
      if (p.transform) {
        // the planes of a step are independent: with two groups of four warps, each taking every other plane, two
        // planes are in flight (a plane is a latency chain: barrier wait, loads, ~60 dependent instructions per cell,
        // stores, proxy fence, arrive -- measured 1 570 clk per plane of 3dconv6_2 for ~260 issue slots of work)
        const int xt_all = threadIdx.x - kXfWarp0 * 32;
        const int gthreads = kXfThreads / p.xf_groups;
        const int xgrp = xt_all / gthreads, xt = xt_all - xgrp * gthreads;
        const int tpc = gthreads / p.NCH;             // threads per channel chunk
        const int ch = xt / tpc, ti = xt - ch * tpc;
        const int npos = p.nsub * p.RY * p.PX;
        // loop-invariant cell list of this thread: byte offset in the slot, -1 = outside the volume / none
        int off[kMaxK];
#pragma unroll
        for (int k = 0; k < kMaxK; ++k) {
          const int q = ti + k * tpc;
          off[k] = -1;
          if (k < p.xf_k && q < npos) {
            const int c = q % p.PX;
            const int rest = q / p.PX;
            const int r = rest % p.RY, sub = rest / p.RY;
            const int ix = p.vstep * (x0 + c + p.cx_off) + (sub & 1);
            const int iy = p.vstep * (y0 + r + p.cy_off) + (sub >> 1);
            if (ix >= 0 && ix < p.W && iy >= 0 && iy < p.H) off[k] = ch * p.PS + (sub * p.SUBP + r * p.PX + c) * 16;
          }
        }
        float2 xsc[4], xsh[4], ssc[4], ssh[4];      // channel pairs
        const bool x_act = p.xs != nullptr || p.xbn.stats != nullptr, s_act = p.ss != nullptr || p.sbn.stats != nullptr;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          xsc[k] = *reinterpret_cast<const float2*>(&s_aff[ch * 8 + 2 * k]); xsh[k] = *reinterpret_cast<const float2*>(&s_aff[64 + ch * 8 + 2 * k]);
          ssc[k] = *reinterpret_cast<const float2*>(&s_aff[128 + ch * 8 + 2 * k]); ssh[k] = *reinterpret_cast<const float2*>(&s_aff[192 + ch * 8 + 2 * k]);
        }
        long long xw = 0, xt0 = 0;
        if (kProf && p.prof) xt0 = clock64();
        // ring positions and barrier phases are carried along (a modulo by a run-time ring depth costs ~50
        // instructions per plane and warp: measured a third of this loop at 3dconv6_2)
        int slot = xgrp, ss = p.has_skip ? xgrp % p.RS : 0;
        uint32_t ph = 0u, sph = p.has_skip ? (uint32_t)(xgrp / p.RS) & 1u : 0u;
        const int iz0 = p.zmul * zb + p.zoff;
        for (int seq = xgrp; seq < nplanes; seq += p.xf_groups) {
          long long xa = 0;
          if (kProf && p.prof) xa = clock64();
          mbar_wait(&bar_land[slot], ph);
          if (p.has_skip) mbar_wait(&bar_sland[ss], sph);
          if (kProf && p.prof) xw += clock64() - xa;
          const int iz = iz0 + seq;
          if (iz >= p.zv_lo && iz < p.zv_hi) {
            unsigned char* sl = s_slots + (size_t)slot * p.slot_bytes;
            const unsigned char* sk = s_skip + (size_t)ss * p.slot_bytes;
            // groups of 4 cells: all loads of a group are issued before its first store (the in-place stores would
            // otherwise serialise every load behind the previous cell's store)
#pragma unroll
            for (int k0 = 0; k0 < kMaxK; k0 += 4) {
              if (k0 >= p.xf_k) break;
              uint4 v[4], s4[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                v[u] = make_uint4(0u, 0u, 0u, 0u); s4[u] = v[u];
                if (off[k0 + u] >= 0) {
                  v[u] = *reinterpret_cast<const uint4*>(sl + off[k0 + u]);
                  if (p.has_skip) s4[u] = *reinterpret_cast<const uint4*>(sk + off[k0 + u]);
                }
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                if (off[k0 + u] < 0) continue;
                uint32_t* vw = reinterpret_cast<uint32_t*>(&v[u]);
                const uint32_t* sw = reinterpret_cast<const uint32_t*>(&s4[u]);
                if (p.has_skip) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    float2 f = unpack_bf16x2(vw[j]), g = unpack_bf16x2(sw[j]);
                    if (x_act) {
                      f = ffma2(f, xsc[j], xsh[j]);
                      f.x = fmaxf(f.x, 0.0f); f.y = fmaxf(f.y, 0.0f);
                    }
                    if (s_act) {
                      g = ffma2(g, ssc[j], ssh[j]);
                      g.x = fmaxf(g.x, 0.0f); g.y = fmaxf(g.y, 0.0f);
                    }
                    f = fadd2(f, g);
                    vw[j] = pack_bf16x2(f.x, f.y);
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 f = ffma2(unpack_bf16x2(vw[j]), xsc[j], xsh[j]);
                    vw[j] = pack_bf16x2_relu(f.x, f.y);
                  }
                }
                *reinterpret_cast<uint4*>(sl + off[k0 + u]) = v[u];
              }
            }
          }
          // every writer fences its generic-proxy stores, then ONE lane per warp arrives (256 arrivals on one
          // mbarrier word serialise for ~1 us per plane)
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&bar_ready[slot]);
            if (p.has_skip) mbar_arrive(&bar_sempty[ss]);
          }
          slot += p.xf_groups;
          if (slot >= p.R) { slot -= p.R; ph ^= 1u; }
          if (p.has_skip) {
            ss += p.xf_groups;
            if (ss >= p.RS) { ss -= p.RS; sph ^= 1u; }
          }
        }
        if (kProf && p.prof && blockIdx.x == 0 && xt_all == 0) { p.prof[12] = clock64() - xt0; p.prof[13] = xw; }
      }
    } else if (warp == kMmaWarp) {
      // ===================================== MMA issuer =====================================
      mbar_wait(bar_b, 0);
      const uint32_t slots16 = smem_u32(s_slots) >> 4, slot16 = (uint32_t)p.slot_bytes >> 4;
      const uint64_t desc_hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;   // version 1, SBO = 128 B
      uint64_t* bar_in = p.transform ? bar_ready : bar_land;
      int waited = 0, wslot = 0;
      uint32_t wphase = 0;
      int slot_lo = 0;                                  // ring slot of the first plane of the step
      long long pr_t0 = 0, pr_in = 0, pr_acc = 0, pr_issue = 0, pr_g0 = 0;
      if (kProf && p.prof) { pr_t0 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pr_g0)); }
      for (int t = 0; t < nsteps; ++t) {
        const int seq_lo = p.zstep * t, seq_hi = seq_lo + p.span - 1;
        long long pr_a = 0;
        if (kProf && p.prof) pr_a = clock64();
        while (waited <= seq_hi) {
          mbar_wait(&bar_in[wslot], wphase);
          ++waited;
          if (++wslot == p.R) { wslot = 0; wphase ^= 1u; }
        }
        const int stage = t & 1;
        long long pr_b = 0;
        if (kProf && p.prof) pr_b = clock64();
        if (kProf && p.prof && t == 0 && blockIdx.x == 0 && lane == 0) {
          long long g;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
          p.prof[22] = g - pr_entry;                 // input planes of the first step ready
        }
        mbar_wait(&bar_acc_empty[stage], ((uint32_t)(t >> 1) & 1u) ^ 1u);
        tc_fence_after();
        long long pr_c = 0;
        if (kProf && p.prof) { pr_c = clock64(); pr_in += pr_b - pr_a; pr_acc += pr_c - pr_b; }
        if (elect_one()) {
          if (!(p.dbg & 2)) {
            // plane-major, op-major order.  One 16-byte shared-memory record per op (prefetched by the
            // unrolled loop); the MB row blocks of an op reuse its descriptors (A start + 2 KB, next
            // TMEM column group).
            const uint32_t d_base = tmem_base + (uint32_t)(stage * p.MB * p.NB);
            int sl_idx = slot_lo;
            for (int dz = 0; dz < p.span; ++dz) {
              const uint32_t sl = slots16 + (uint32_t)sl_idx * slot16;
              if (++sl_idx == p.R) sl_idx = 0;
              const int ob = s_dzb[dz], oe = s_dzb[dz + 1];
              switch (p.MB) {
                case 1: issue_ops<1>(s_ops, ob, oe, sl, d_base, (uint32_t)p.NB, desc_hi); break;
                case 2: issue_ops<2>(s_ops, ob, oe, sl, d_base, (uint32_t)p.NB, desc_hi); break;
                case 3: issue_ops<3>(s_ops, ob, oe, sl, d_base, (uint32_t)p.NB, desc_hi); break;
                default: issue_ops<4>(s_ops, ob, oe, sl, d_base, (uint32_t)p.NB, desc_hi); break;
              }
            }
          }
          mma_commit(&bar_acc_full[stage]);
          // planes no later step needs go back to the producer
          int rs = slot_lo;
          for (int s = seq_lo; s < min(p.zstep * (t + 1), nplanes); ++s) {
            mma_commit(&bar_empty[rs]);
            if (++rs == p.R) rs = 0;
          }
        }
        __syncwarp();
        if (kProf && p.prof) pr_issue += clock64() - pr_c;
        slot_lo += p.zstep;
        if (slot_lo >= p.R) slot_lo -= p.R;
      }
      if (kProf && p.prof && blockIdx.x == 0 && lane == 0) {
        long long g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        p.prof[0] = clock64() - pr_t0; p.prof[1] = g1 - pr_g0; p.prof[2] = pr_in; p.prof[3] = pr_acc; p.prof[4] = pr_issue;
        p.prof[5] = (long long)nsteps * p.nops * p.MB;
        p.prof[6] = pr_g0 - pr_entry;      // prologue: kernel entry -> MMA warp past the weights barrier
        p.prof[7] = g1 - pr_entry;         // kernel entry -> MMA warp done issuing
      }
    } else {
      // ===================================== epilogue =====================================
      if (XF) {
        // rider launch: the input needs no transform, so the 8 transform warps are epilogue warps too.  Warp w reads
        // TMEM sub-partition w % 4; group w / 4 takes one kind of item: 0 = stride-1 layer, even plane, 1 = stride-1
        // layer, odd plane, 2 = the rider's chunks (odd plane) -- two items per row block each, 8 / 8 / 16 channels of
        // statistics, so every group can keep its next TMEM load in flight.
        const int ewarp = RIDER ? (warp & 3) : warp, egrp = RIDER ? (warp >> 2) : 0;
        // x-fold: columns [j][kw][co]; output (yy, xx-1) = P[kw=0] of lane-1 + P[kw=1] + P[kw=2] of lane+1.
        // A warp's 32 TMEM lanes are 32 / PX whole tile rows, so the shuffles (width PX) never leave a row; the
        // halo columns xx = 0 and PX-1 only supply partial sums.
        // (rider launch: the rider's second chunk has arrays of its own -- with one array the compiler turned the
        // chunk selection into a run-time index and moved the statistics to local memory)
        constexpr int NST = XF ? (RIDER ? 8 : XFC * 8) : 8;
        float sum[NST], sq[NST], sum2[RIDER ? 8 : 1], sq2[RIDER ? 8 : 1];
#pragma unroll
        for (int k = 0; k < NST; ++k) { sum[k] = 0.0f; sq[k] = 0.0f; }
#pragma unroll
        for (int k = 0; k < (RIDER ? 8 : 1); ++k) { sum2[k] = 0.0f; sq2[k] = 0.0f; }
        const int grp = 3 * p.cout_n, nchunk = p.cout_n >> 3;
        const size_t zpitch = (size_t)p.Ho * p.Wo;
        const int chunk0 = p.cout_base >> 3;
        const int px_shift = 31 - __clz(p.PX);
        long long ew = 0, et0 = 0;
        if (kProf && p.prof) et0 = clock64();
        float rg_m[kMaxMB], rg_s[kMaxMB], rg_w[kMaxMB];     // fused soft-argmin state of this thread's pixels
#pragma unroll
        for (int b = 0; b < kMaxMB; ++b) { rg_m[b] = -INFINITY; rg_s[b] = 0.0f; rg_w[b] = 0.0f; }
        for (int t = 0; t < nsteps; ++t) {
          const int stage = t & 1;
          long long ea = 0;
          if (kProf && p.prof) ea = clock64();
          mbar_wait(&bar_acc_full[stage], (uint32_t)(t >> 1) & 1u);
          tc_fence_after();
          if (kProf && p.prof) ew += clock64() - ea;
          const int mz = zb + t * p.zf;
          const int nlive = min(p.zf, ze - mz);
          if (C1) {
            // Cout = 1 (3dconv6_2): N = 3*zf <= 12 columns, fp32 [D,H,W] output.  When the CTA covers the whole depth
            // range the soft-argmin of the regression (model.py:472-495) is folded in: every thread keeps the running
            // (max of -F, sum of exp, depth-weighted sum) of its pixels, rescaled when the maximum moves.
            // two register buffers of 3 * zf <= 12 columns: the next row block's accumulators are on their way while this
            // one is shuffled, stored and folded into the soft-argmin (a TMEM round trip per block was most of this loop)
            const uint32_t tb0 = tmem_base + ((uint32_t)(ewarp * 32) << 16) + (uint32_t)(stage * p.MB * p.NB);
            uint32_t rbuf[2][12];
            tmem_ld8(tb0, rbuf[0]);
            tmem_ld4(tb0 + 8, rbuf[0] + 8);
#pragma unroll
            for (int b = 0; b < kMaxMB; ++b) {
              if (b >= p.MB) break;
              const int m = b * 128 + ewarp * 32 + lane;
              const int yy = m >> px_shift, xx = m & (p.PX - 1);
              const bool valid = xx >= 1 && xx <= TXe && yy < TYe && !(p.dbg & 4);
              const int oy = y0 + yy, ox = x0 + xx - 1;
              uint32_t* r = rbuf[b & 1];
              tmem_ld_wait();
              if (b + 1 < p.MB) {
                tmem_ld8(tb0 + (uint32_t)((b + 1) * p.NB), rbuf[(b + 1) & 1]);
                tmem_ld4(tb0 + (uint32_t)((b + 1) * p.NB) + 8, rbuf[(b + 1) & 1] + 8);
              }
              float xj[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                xj[j] = -INFINITY;
                if (j < p.zf) {
                  const float lft = __shfl_up_sync(0xffffffffu, __uint_as_float(r[3 * j]), 1, p.PX);
                  const float rgt = __shfl_down_sync(0xffffffffu, __uint_as_float(r[3 * j + 2]), 1, p.PX);
                  const float v = lft + __uint_as_float(r[3 * j + 1]) + rgt;
                  if (valid && j < nlive) {
                    sum[0] += v; sq[0] = fmaf(v, v, sq[0]);
                    if (p.y_f32) p.y_f32[((size_t)(mz + j) * p.Ho + oy) * p.Wo + ox] = v;
                    xj[j] = -v;
                  }
                }
              }
              if (p.rg_out && valid) {
                // one rescale per step: new maximum over the old one and the step's planes, then the planes' terms
                const float mx = fmaxf(fmaxf(rg_m[b], fmaxf(xj[0], xj[1])), fmaxf(xj[2], xj[3]));
                const float sc = __expf(rg_m[b] - mx);          // 0 on the first step (rg_m starts at -inf)
                float s_ = rg_s[b] * sc, w_ = rg_w[b] * sc;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float e = __expf(xj[j] - mx);           // 0 for planes outside the volume (-inf)
                  const float dsample = __fadd_rn(p.rg_start, __fmul_rn(p.rg_step, (float)(mz + j)));   // tf.linspace
                  s_ += e;
                  w_ = fmaf(dsample, e, w_);
                }
                rg_m[b] = mx; rg_s[b] = s_; rg_w[b] = w_;
              }
            }
          }
          if (!C1) {
            // items (row block, output plane j, 8-channel chunk): the three TMEM loads of the next item are in flight
            // while this one is shuffled, reduced and stored
            // rider launch: plane j = 0 (even) has the stride-1 layer's chunk only, plane j = 1 (odd) that chunk and
            // the rider's n2 / 8 chunks; its columns start after the 3 * cout_n columns of plane 0
            const int nrid = p.n2 >> 3;
            const int nitems = RIDER ? (egrp == 2 ? p.MB * nrid : p.MB) : p.MB * p.zf * nchunk;
            uint32_t r[24];
            const uint32_t tstage = tmem_base + ((uint32_t)(ewarp * 32) << 16) + (uint32_t)(stage * p.MB * p.NB);
            auto issue = [&](int b, int j, int ck, uint32_t* r) {
              const uint32_t kws = RIDER ? (uint32_t)(j ? p.cout_n + p.n2 : p.cout_n) : (uint32_t)p.cout_n;
              const uint32_t col = tstage + (uint32_t)(b * p.NB + j * grp + ck * 8);
              tmem_ld8(col, r);
              tmem_ld8(col + kws, r + 8);
              tmem_ld8(col + 2u * kws, r + 16);
            };
            // ONE register buffer: the three partial sums of an item are folded into v[8] by the shuffles first, then
            // the next item's TMEM loads are issued into the same registers and fly while this item is reduced, packed
            // and stored (a second buffer cost 24 registers and pushed the statistics into local memory)
            int b = 0, j = RIDER && egrp > 0 ? 1 : 0, ck = RIDER && egrp == 2 ? 1 : 0;      // current item; (nb, nj, nck) = the next one
            if (nitems > 0) issue(0, j, ck, r);
            for (int it = 0; it < nitems; ++it) {
              {
                int nck = ck + 1, nj = j, nb = b;
                if (RIDER) {
                  if (egrp < 2) { nck = 0; ++nb; }
                  else if (nck > nrid) { nck = 1; ++nb; }
                } else if (nck == nchunk) { nck = 0; if (++nj == p.zf) { nj = 0; ++nb; } }
                tmem_ld_wait();
                const int m = b * 128 + ewarp * 32 + lane;
                const int yy = m >> px_shift, xx = m & (p.PX - 1);
                const bool valid = xx >= 1 && xx <= TXe && yy < TYe && !(p.dbg & 4);
                const int oy = y0 + yy, ox = x0 + xx - 1;
                float v[8];
                float2* v2 = reinterpret_cast<float2*>(v);
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                  float2 lft, rgt;
                  lft.x = __shfl_up_sync(0xffffffffu, __uint_as_float(r[k]), 1, p.PX);
                  lft.y = __shfl_up_sync(0xffffffffu, __uint_as_float(r[k + 1]), 1, p.PX);
                  rgt.x = __shfl_down_sync(0xffffffffu, __uint_as_float(r[16 + k]), 1, p.PX);
                  rgt.y = __shfl_down_sync(0xffffffffu, __uint_as_float(r[16 + k + 1]), 1, p.PX);
                  v2[k >> 1] = fadd2(fadd2(lft, make_float2(__uint_as_float(r[8 + k]), __uint_as_float(r[8 + k + 1]))), rgt);
                }
                if (it + 1 < nitems) issue(nb, nj, nck, r);
                if (RIDER && ck > 0) {
                  // rider chunk ck - 1 (plane j = 1 only): the stride-2 conv's output (oz, oy, ox) >> 1 lives at the odd
                  // positions; its own statistics (registers 8 ..) and its own output tensors
                  if (valid && (oy & ox & 1)) {
                    if (RIDER) {
                      if (ck == 1) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) { sum[k] += v[k]; sq[k] = fmaf(v[k], v[k], sq[k]); }
                      } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) { sum2[k] += v[k]; sq2[k] = fmaf(v[k], v[k], sq2[k]); }
                      }
                    }
                    uint4 pk;
                    pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
                    pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
                    const int oz2 = (mz + 1) >> 1, oy2 = oy >> 1, ox2 = ox >> 1;
                    const size_t zc = (size_t)oz2 * (p.Cout2 >> 3) + (ck - 1);
                    if (p.y2_cp8 && !(p.dbg & 16))
                      *reinterpret_cast<uint4*>(p.y2_cp8 + ((zc * p.Ho2 + oy2) * p.Wo2 + ox2) * 8) = pk;
                    if (p.y2_ps8 && !(p.dbg & 16))
                      *reinterpret_cast<uint4*>(p.y2_ps8 + ((((zc * 4 + (oy2 & 1) * 2 + (ox2 & 1)) * p.Hso2) + (oy2 >> 1)) * p.Wso2 +
                                                            (ox2 >> 1)) * 8) = pk;
                  }
                } else if (valid && j < nlive) {
                  // channel chunk ck of this launch: statistics registers are indexed at compile time
#pragma unroll
                  for (int c4 = 0; c4 < (XF ? (RIDER ? 1 : XFC) : 1); ++c4)
                    if (c4 == ck && !(p.dbg & 64)) {
#pragma unroll
                      for (int k = 0; k < 4; ++k) {
                        float2* s2 = reinterpret_cast<float2*>(sum + c4 * 8) + k;
                        float2* q2 = reinterpret_cast<float2*>(sq + c4 * 8) + k;
                        *s2 = fadd2(*s2, v2[k]);
                        *q2 = ffma2(v2[k], v2[k], *q2);
                      }
                    }
                  const int oz = mz + j;
                  if (p.y_f32) {
                    float4* yo = reinterpret_cast<float4*>(p.y_f32 + (((size_t)oz * p.Ho + oy) * p.Wo + ox) * p.Cout +
                                                           p.cout_base + ck * 8);
                    yo[0] = make_float4(v[0], v[1], v[2], v[3]);
                    yo[1] = make_float4(v[4], v[5], v[6], v[7]);
                  } else {
                    uint4 pk;
                    pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
                    pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
                    if (p.y_cp8 && !(p.dbg & 32)) store_cell<PEER>(p, 0, oz, chunk0 + ck, zpitch, (size_t)oy * p.Wo + ox, pk);
                    if (p.y_ps8 && !(p.dbg & 32)) {
                      const size_t pcell = ((size_t)((oy & 1) * 2 + (ox & 1)) * p.Hso + (oy >> 1)) * p.Wso + (ox >> 1);
                      store_cell<PEER>(p, 1, oz, chunk0 + ck, 4 * (size_t)p.Hso * p.Wso, pcell, pk);
                    }
                  }
                }
                b = nb; j = nj; ck = nck;
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_acc_empty[stage]);
        }
        if (C1 && p.rg_out) {
          // (m, s, w) per pixel, [3][Ho*Wo]: combined and turned into depth + probability by regress_combine_kernel
          const size_t npix = (size_t)p.Ho * p.Wo;
#pragma unroll
          for (int b = 0; b < kMaxMB; ++b) {
            if (b >= p.MB) break;
            const int m = b * 128 + warp * 32 + lane;
            const int yy = m >> px_shift, xx = m & (p.PX - 1);
            if (xx >= 1 && xx <= TXe && yy < TYe) {
              const size_t pix = (size_t)(y0 + yy) * p.Wo + (x0 + xx - 1);
              p.rg_out[pix] = rg_m[b]; p.rg_out[npix + pix] = rg_s[b]; p.rg_out[2 * npix + pix] = rg_w[b];
            }
          }
        }
        if (kProf && p.prof && blockIdx.x == 0 && threadIdx.x == 0) { p.prof[14] = clock64() - et0; p.prof[15] = ew; }
        if constexpr (RIDER) {
          flush_stats_rider(p, sum, sq, sum2, sq2, egrp, s_red + egrp * 128, ewarp, lane);
        } else {
          if (p.stats && !(p.dbg & 8)) flush_stats<NST>(p, sum, sq, p.cout_n, false, s_red, warp, lane);
        }
      } else {
      float sum[CP], sq[CP];
#pragma unroll
      for (int k = 0; k < CP; ++k) { sum[k] = 0.0f; sq[k] = 0.0f; }
      const bool deconv = p.mode == MODE_DECONV;
      const int ncls = deconv ? 8 : 1;
      const int ncol = p.zf * p.cout_n;
      const size_t zpitch = (size_t)p.Ho * p.Wo;          // voxels per output plane
      const int chunk0 = p.cout_base >> 3;
      const int ncho = p.Cout >> 3;
      long long ew = 0, et0 = 0;
      if (kProf && p.prof) et0 = clock64();
      for (int t = 0; t < nsteps; ++t) {
        const int stage = t & 1;
        long long ea = 0;
        if (kProf && p.prof) ea = clock64();
        mbar_wait(&bar_acc_full[stage], (uint32_t)(t >> 1) & 1u);
        tc_fence_after();
        if (kProf && p.prof) ew += clock64() - ea;
        const int mz = zb + t * p.zf;
        const int nlive = min(p.zf, ze - mz);          // output planes of this step inside the volume
        if (CP == 16 && deconv && !p.y_f32) {
          // transposed conv, bf16 out: the two x-parity classes of a row are adjacent cells of the output, so they
          // are drained together (2*cw columns per wait, 32 contiguous bytes per thread and channel chunk); the
          // TMEM load of the next pair is in flight while this one is reduced and stored.
          const int npair = p.MB * 4;
          const uint32_t tbase = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(stage * p.MB * p.NB);
          const bool narrow = p.cw == 8;               // 8 columns per class: one 16-column load holds the pair
          uint32_t ra[32], rb[32];
          tmem_ld16(tbase, ra);
          if (!narrow) tmem_ld16(tbase + 16, ra + 16);
          // (position and output address once per row block: the four class pairs of a block differ by one output
          // plane / one output row -- the division and the 64-bit address chain per pair were half of this loop)
          const size_t pr_z = (size_t)ncho * zpitch * 8, pr_y = (size_t)p.Wo * 8;      // elements per output plane / row
          for (int b = 0; b < p.MB; ++b) {
            const int m = b * 128 + warp * 32 + lane;
            const int yy = m / p.PX, xx = m - yy * p.PX;
            const bool valid = xx < TXe && yy < TYe && !(p.dbg & 4);
            const int oy0 = 2 * (y0 + yy), ox = 2 * (x0 + xx);
            const size_t cell0 = (size_t)oy0 * p.Wo + ox;
            __nv_bfloat16* const base = p.y_cp8 + (((size_t)(2 * mz) * ncho + chunk0) * zpitch + cell0) * 8;
#pragma unroll
            for (int pr = 0; pr < 4; ++pr) {                    // pair pr = classes 2*pr, 2*pr+1 (pz, py fixed)
              uint32_t* r = (pr & 1) ? rb : ra;
              uint32_t* rn = (pr & 1) ? ra : rb;
              tmem_ld_wait();
              if (b * 4 + pr + 1 < npair) {
                const uint32_t ta = tbase + (uint32_t)((pr == 3 ? b + 1 : b) * p.NB + ((pr + 1) & 3) * 2 * p.cw);
                tmem_ld16(ta, rn);
                if (!narrow) tmem_ld16(ta + 16, rn + 16);
              }
              if (!valid) continue;
              const int oz = 2 * mz + (pr >> 1);
              const size_t cell = cell0 + (size_t)(pr & 1) * p.Wo;
              __nv_bfloat16* const dst0 = base + (pr >> 1) * pr_z + (pr & 1) * pr_y;
              const float2* f = reinterpret_cast<const float2*>(r);
              float2* sum2 = reinterpret_cast<float2*>(sum);
              float2* sq2 = reinterpret_cast<float2*>(sq);
              if (narrow) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  sum2[k] = fadd2(sum2[k], fadd2(f[k], f[4 + k]));
                  sq2[k] = ffma2(f[4 + k], f[4 + k], ffma2(f[k], f[k], sq2[k]));
                }
                uint4 c0, c1;
                c0.x = pack_bf16x2(__uint_as_float(r[0]), __uint_as_float(r[1]));
                c0.y = pack_bf16x2(__uint_as_float(r[2]), __uint_as_float(r[3]));
                c0.z = pack_bf16x2(__uint_as_float(r[4]), __uint_as_float(r[5]));
                c0.w = pack_bf16x2(__uint_as_float(r[6]), __uint_as_float(r[7]));
                c1.x = pack_bf16x2(__uint_as_float(r[8]), __uint_as_float(r[9]));
                c1.y = pack_bf16x2(__uint_as_float(r[10]), __uint_as_float(r[11]));
                c1.z = pack_bf16x2(__uint_as_float(r[12]), __uint_as_float(r[13]));
                c1.w = pack_bf16x2(__uint_as_float(r[14]), __uint_as_float(r[15]));
                if (PEER) {
                  store_cell<true>(p, 0, oz, chunk0, zpitch, cell, c0);
                  store_cell<true>(p, 0, oz, chunk0, zpitch, cell + 1, c1);
                } else {
                  uint4* dst = reinterpret_cast<uint4*>(dst0);
                  dst[0] = c0; dst[1] = c1;
                }
              } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  sum2[k] = fadd2(sum2[k], fadd2(f[k], f[8 + k]));
                  sq2[k] = ffma2(f[8 + k], f[8 + k], ffma2(f[k], f[k], sq2[k]));
                }
#pragma unroll
                for (int k = 0; k < 16; k += 8) {
                  if (k < p.cout_n) {
                    uint4 c0, c1;
                    c0.x = pack_bf16x2(__uint_as_float(r[k]), __uint_as_float(r[k + 1]));
                    c0.y = pack_bf16x2(__uint_as_float(r[k + 2]), __uint_as_float(r[k + 3]));
                    c0.z = pack_bf16x2(__uint_as_float(r[k + 4]), __uint_as_float(r[k + 5]));
                    c0.w = pack_bf16x2(__uint_as_float(r[k + 6]), __uint_as_float(r[k + 7]));
                    c1.x = pack_bf16x2(__uint_as_float(r[16 + k]), __uint_as_float(r[16 + k + 1]));
                    c1.y = pack_bf16x2(__uint_as_float(r[16 + k + 2]), __uint_as_float(r[16 + k + 3]));
                    c1.z = pack_bf16x2(__uint_as_float(r[16 + k + 4]), __uint_as_float(r[16 + k + 5]));
                    c1.w = pack_bf16x2(__uint_as_float(r[16 + k + 6]), __uint_as_float(r[16 + k + 7]));
                    if (PEER) {
                      store_cell<true>(p, 0, oz, chunk0 + (k >> 3), zpitch, cell, c0);
                      store_cell<true>(p, 0, oz, chunk0 + (k >> 3), zpitch, cell + 1, c1);
                    } else {
                      uint4* dst = reinterpret_cast<uint4*>(dst0 + (size_t)(k >> 3) * zpitch * 8);
                      dst[0] = c0; dst[1] = c1;
                    }
                  }
                }
              }
            }
          }
        } else
        for (int b = 0; b < p.MB; ++b) {
          const int m = b * 128 + warp * 32 + lane;
          const int yy = m / p.PX, xx = m - yy * p.PX;
          const bool valid = xx < TXe && yy < TYe && !(p.dbg & 4);
          for (int cls = 0; cls < ncls; ++cls) {
            uint32_t r[CP];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) +
                                   (uint32_t)((stage * p.MB + b) * p.NB + cls * p.cw);
            if (p.cw == 8) {
              // merged transposed conv with 8 columns per class: the other CP - 8 registers count as zero columns
              tmem_ld8(taddr, r);
#pragma unroll
              for (int k = 8; k < CP; ++k) r[k] = 0u;
            } else {
#pragma unroll
              for (int c0 = 0; c0 < CP; c0 += 16) tmem_ld16(taddr + c0, r + c0);
            }
            tmem_ld_wait();
            if (!valid) continue;
            // columns [j*cout_n, (j+1)*cout_n) belong to output plane mz + j (z-fold); planes past the
            // end of the volume are computed but neither stored nor counted.  Unused columns are exact zeros.
            if (nlive == p.zf) {
              // packed fp32 pairs: the accumulator columns of a TMEM load sit in consecutive registers
              const float2* f = reinterpret_cast<const float2*>(r);
              float2* sum2 = reinterpret_cast<float2*>(sum);
              float2* sq2 = reinterpret_cast<float2*>(sq);
#pragma unroll
              for (int k = 0; k < CP / 2; ++k) {
                sum2[k] = fadd2(sum2[k], f[k]);
                sq2[k] = ffma2(f[k], f[k], sq2[k]);
              }
            } else {
#pragma unroll
              for (int k = 0; k < CP; ++k) {
                const float v = (k >> p.cn_shift) < nlive ? __uint_as_float(r[k]) : 0.0f;
                sum[k] += v; sq[k] = fmaf(v, v, sq[k]);
              }
            }
            int oz, oy, ox;
            if (deconv) { oz = 2 * mz + (cls >> 2); oy = 2 * (y0 + yy) + ((cls >> 1) & 1); ox = 2 * (x0 + xx) + (cls & 1); }
            else { oz = mz; oy = y0 + yy; ox = x0 + xx; }
            if (p.y_f32) {
              float* yo = p.y_f32 + (((size_t)oz * p.Ho + oy) * p.Wo + ox) * p.Cout + p.cout_base;
#pragma unroll
              for (int k = 0; k < CP; ++k) {
                const int j = p.zf == 1 ? 0 : (k >> p.cn_shift), cn = k - (p.zf == 1 ? 0 : (j << p.cn_shift));
                if (k < ncol && j < nlive) yo[(size_t)j * zpitch * p.Cout + cn] = __uint_as_float(r[k]);
              }
            } else {
              // CP8 (and optionally PS8) bf16: one 16-byte cell per 8 channels
              const size_t cell = ((size_t)oy * p.Wo + ox);
              const size_t pcell = ((size_t)((oy & 1) * 2 + (ox & 1)) * p.Hso + (oy >> 1)) * p.Wso + (ox >> 1);
#pragma unroll
              for (int k = 0; k < CP; k += 8) {
                const int j = p.zf == 1 ? 0 : (k >> p.cn_shift), cn = k - (p.zf == 1 ? 0 : (j << p.cn_shift));
                if (k < ncol && j < nlive) {
                  uint4 pk;
                  pk.x = pack_bf16x2(__uint_as_float(r[k]), __uint_as_float(r[k + 1]));
                  pk.y = pack_bf16x2(__uint_as_float(r[k + 2]), __uint_as_float(r[k + 3]));
                  pk.z = pack_bf16x2(__uint_as_float(r[k + 4]), __uint_as_float(r[k + 5]));
                  pk.w = pack_bf16x2(__uint_as_float(r[k + 6]), __uint_as_float(r[k + 7]));
                  if (p.y_cp8) store_cell<PEER>(p, 0, oz + j, chunk0 + (cn >> 3), zpitch, cell, pk);
                  if (p.y_ps8) store_cell<PEER>(p, 1, oz + j, chunk0 + (cn >> 3), 4 * (size_t)p.Hso * p.Wso, pcell, pk);
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_acc_empty[stage]);
      }
      if (kProf && p.prof && blockIdx.x == 0 && threadIdx.x == 0) { p.prof[14] = clock64() - et0; p.prof[15] = ew; }
      if (p.stats && !(p.dbg & 8)) flush_stats<CP>(p, sum, sq, ncol, p.zf != 1, s_red, warp, lane);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.pdl == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  if (kProf && p.prof && threadIdx.x == 0) {
    long long g2;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g2));
    if (blockIdx.x == 0) p.prof[8] = g2 - pr_entry;             // whole CTA
    p.prof[32 + 2 * blockIdx.x] = pr_entry;
    p.prof[33 + 2 * blockIdx.x] = g2;
  }
}

// ---------------------------------------------------------------------------------------------
// layout conversion (stand-alone entry points; the fused path writes CP8 / PS8 directly)
// ---------------------------------------------------------------------------------------------
__global__ void ndhwc_to_planar_kernel(const __nv_bfloat16* __restrict__ x, int D, int H, int W, int C,
                                       __nv_bfloat16* __restrict__ cp8, __nv_bfloat16* __restrict__ ps8) {
  const int nch = C >> 3, Hs = (H + 1) >> 1, Ws = (W + 1) >> 1;
  const size_t total = (size_t)D * H * W * nch;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % nch);
    size_t v = i / nch;
    const int xx = (int)(v % W); v /= W;
    const int yy = (int)(v % H);
    const int z = (int)(v / H);
    const uint4 cell = *reinterpret_cast<const uint4*>(x + i * 8);
    const size_t zc = (size_t)z * nch + ch;
    if (cp8) *reinterpret_cast<uint4*>(cp8 + ((zc * H + yy) * W + xx) * 8) = cell;
    if (ps8)
      *reinterpret_cast<uint4*>(ps8 + (((zc * 4 + (yy & 1) * 2 + (xx & 1)) * Hs + (yy >> 1)) * Ws + (xx >> 1)) * 8) = cell;
  }
}

__global__ void planar_to_ndhwc_kernel(const __nv_bfloat16* __restrict__ cp8, int D, int H, int W, int C,
                                       __nv_bfloat16* __restrict__ y) {
  const int nch = C >> 3;
  const size_t total = (size_t)D * H * W * nch;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % nch);
    size_t v = i / nch;
    const int xx = (int)(v % W); v /= W;
    const int yy = (int)(v % H);
    const int z = (int)(v / H);
    *reinterpret_cast<uint4*>(y + i * 8) =
        *reinterpret_cast<const uint4*>(cp8 + ((((size_t)z * nch + ch) * H + yy) * W + xx) * 8);
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ x, size_t n, __nv_bfloat16* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i]);
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
  Params cp;
  PackParams pp;
  size_t smem;
  double est_clk;
};

int pow2_at_least(int v) { int p = 32; while (p < v) p <<= 1; return p; }
double mma_clk(int n) { const double a = 32.0 + n / 4.0, b = n / 2.0; return (a > b ? a : b) + 2.0; }

// Build the op table for one (mode, Cin, cout slice) and the slot geometry for tile (TX, TY), z-fold zf.
bool build_plan(int mode, int D, int H, int W, int cin, int cout, int cout_base, int cout_n, int TX, int TY, int zf,
                bool xfold, bool has_skip, bool transform, int n2, Plan* pl) {
  Params& c = pl->cp;
  PackParams& pk = pl->pp;
  if (mode != MODE_CONV1) zf = 1;
  // rider launch: x-folded stride-1 conv, steps of one even and one odd output plane, even extents (the stride-2 conv's
  // SAME padding then has nothing before the volume: its outputs sit at the odd positions)
  if (n2 && (mode != MODE_CONV1 || !xfold || zf != 2 || cin < 16 || (n2 & 7) || (cout_n & 7) || ((D | H | W) & 1) || has_skip ||
             transform)) return false;
  if (xfold) {
    // x-fold: the padded tile is 8 / 16 / 32 cells wide and fills whole 128-row blocks
    if (mode != MODE_CONV1 || !(cout_n == 1 || cout_n % 8 == 0)) return false;
    const int px = TX + 2;
    if ((px != 8 && px != 16 && px != 32) || (TY * px) % 128 != 0) return false;
  }
  if (zf > 1 && !xfold && ((cout_n & (cout_n - 1)) != 0)) return false;     // the epilogue splits folded columns by shift
  const int kwn = xfold ? 3 : 1, grp = kwn * cout_n, ncols = n2 ? grp + 3 * (cout_n + n2) : zf * grp;
  if (ncols > (xfold ? 128 : 32) || zf + 2 > kMaxSpan) return false;
  const int CP = xfold ? (ncols + 15) / 16 * 16 : (ncols <= 16 ? 16 : 32);
  const bool master = zf > 1 && ncols == CP && cin >= 16 && !n2;
  c.n2 = n2; c.Cout2 = n2; c.Ho2 = H / 2; c.Wo2 = W / 2; c.Hso2 = (H / 2 + 1) / 2; c.Wso2 = (W / 2 + 1) / 2;
  c.y2_cp8 = nullptr; c.y2_ps8 = nullptr; c.stats2 = nullptr;
  c.CP = CP;
  c.zf = zf;
  c.xfold = xfold ? 1 : 0;
  c.cn_shift = 0;
  while ((1 << c.cn_shift) < cout_n) ++c.cn_shift;
  c.mode = mode; c.D = D; c.H = H; c.W = W; c.Cin = cin; c.Cout = cout; c.cout_base = cout_base; c.cout_n = cout_n;
  c.has_skip = has_skip ? 1 : 0; c.transform = transform ? 1 : 0;
  int pbd = 0, pbh = 0, pbw = 0;
  if (mode == MODE_CONV1) {
    c.Do = D; c.Ho = H; c.Wo = W; c.Mz = D; c.My = H; c.Mx = W;
    c.PX = TX + 2; c.RY = TY + 2; c.nsub = 1; c.vstep = 1; c.cx_off = -1; c.cy_off = -1;
    c.zmul = 1; c.zoff = -1; c.zstep = zf; c.span = zf + 2; c.NB = CP;
  } else if (mode == MODE_CONV2) {
    c.Do = ceil_div(D, 2); c.Ho = ceil_div(H, 2); c.Wo = ceil_div(W, 2); c.Mz = c.Do; c.My = c.Ho; c.Mx = c.Wo;
    pbd = tf_same_pad_before(D, 3, 2); pbh = tf_same_pad_before(H, 3, 2); pbw = tf_same_pad_before(W, 3, 2);
    c.PX = TX + 1 + pbw; c.RY = TY + 1 + pbh; c.nsub = 4; c.vstep = 2; c.cx_off = -pbw; c.cy_off = -pbh;
    c.zmul = 2; c.zoff = -pbd; c.zstep = 2; c.span = 3; c.NB = CP;
  } else {
    c.Do = 2 * D; c.Ho = 2 * H; c.Wo = 2 * W; c.Mz = D; c.My = H; c.Mx = W;
    c.PX = TX + 1; c.RY = TY + 1; c.nsub = 1; c.vstep = 1; c.cx_off = -1; c.cy_off = -1;
    c.zmul = 1; c.zoff = -1; c.zstep = 1; c.span = 2; c.NB = 8 * CP;
  }
  // transposed conv with <= 16 output channels: the 8 output-parity classes that read the same shifted input
  // are merged into ONE MMA (N = 8 classes x cw columns, zero rows for the classes that do not use the shift):
  // 8 MMAs per 16 input channels instead of 27
  c.cw = CP; c.dmerge = 0; c.mma_n = CP;
  if (mode == MODE_DECONV && cout_n <= 16 && cin >= 16) {
    c.dmerge = 1; c.cw = cout_n <= 8 ? 8 : 16; c.NB = 8 * c.cw; c.mma_n = 8 * c.cw;
  }
  c.Hso = (c.Ho + 1) / 2; c.Wso = (c.Wo + 1) / 2;
  if (c.PX * 8 > 256 || c.RY > 256) return false;     // TMA box limits
  c.TX = TX; c.TY = TY;
  c.tiles_x = ceil_div(c.Mx, TX); c.tiles_y = ceil_div(c.My, TY);
  c.SUBP = (c.RY * c.PX + 7) / 8 * 8;          // sub-arrays start 128-byte aligned (TMA destination)
  const bool dense = (c.RY * c.PX) % 8 == 0;   // sub-arrays and chunk planes back to back: one TMA box per plane
  c.NCH = cin / 8;
  c.MB = ceil_div(TY * c.PX, 128);
  if (2 * c.MB * c.NB > 512 || c.MB > kMaxMB) return false;
  c.tmem_cols = pow2_at_least(2 * c.MB * c.NB);
  c.xf_k = ceil_div(c.nsub * c.RY * c.PX, kXfThreads / c.NCH);
  c.xf_groups = 1;
  if (transform && c.xf_k > kMaxK) return false;
  // (measured at config 2: no gain -- 3dconv6_2 0.218 -> 0.214 ms, the coarse layers slightly slower -- so only on request)
  if (transform && tuning().tc_xf_groups == 2 && 2 * c.NCH <= kXfThreads / 16) {
    const int k2 = ceil_div(c.nsub * c.RY * c.PX, (kXfThreads / 2) / c.NCH);
    if (k2 <= kMaxK) { c.xf_groups = 2; c.xf_k = k2; }
  }

  // ---- taps ---------------------------------------------------------------------------------------
  struct Tap { int dz, pos, widx, cls; };
  Tap taps[9 * kMaxSpan];
  int ntaps = 0;
  if (mode == MODE_DECONV && c.dmerge) {
    for (int sh = 0; sh < 8; ++sh) {
      const int sz = -((sh >> 2) & 1), sy = -((sh >> 1) & 1), sx = -(sh & 1);
      taps[ntaps++] = {1 + sz, (1 + sy) * c.PX + (1 + sx), sh, 0};
    }
  } else if (mode == MODE_DECONV) {
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
    // CONV1: one tap per input plane dz of the step and (kh, kw) (widx = kh*3+kw when the B images are
    // windows of a master image, else dz = kd).  CONV2: dz = kd.
    const int ndz = mode == MODE_CONV1 ? zf + 2 : 3;
    for (int kd = 0; kd < ndz; ++kd)
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < (xfold ? 1 : 3); ++kw) {      // x-fold: one tap per filter row, kw lives in N
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
  const int sp_cells = max_pos + c.MB * 128 + 8;            // cells an A descriptor may touch from the plane start
  c.one_box = 0;
  c.ring_pad = 0;
  if (dense) {
    // rows of the last 128-row block read past their chunk plane into the next one, the next slot or the zeroed
    // pad behind the ring: finite data, and those rows are dropped
    c.PS = c.nsub * c.SUBP * 16;
    c.one_box = 1;
    if (sp_cells * 16 > c.PS) c.ring_pad = (sp_cells * 16 - c.PS + 127) / 128 * 128;
  } else {
    c.PS = (sp_cells * 16 + 127) / 128 * 128;          // chunk planes start 128-byte aligned (TMA destination)
    if (c.nsub * c.SUBP * 16 > c.PS) c.PS = c.nsub * c.SUBP * 16;
  }
  c.slot_bytes = c.NCH * c.PS;

  // ---- ops ----------------------------------------------------------------------------------------
  int nops = 0, nimg = 0;
  const int b_op_bytes = 2 * CP * 16;
  const int m_rows = (2 * zf + 1) * grp, m_bytes = 2 * m_rows * 16;      // master image
  // [c_lo, c_hi): the columns of the image this op has weights for (the others are zero rows: the op skips them,
  // except the first op of a step, which has to overwrite every accumulator column).  N stays a multiple of 16 and
  // the first column a multiple of 16 (the B window then starts on a whole 8-row core matrix).
  const bool trim = tuning().tc_trim != 0;
  auto add_op = [&](int dz, uint32_t a_off, uint32_t a_lbo, uint32_t b_off, uint32_t b_lbo, int col, bool first,
                    int c_lo = 0, int c_hi = 1 << 20) {
    UmmaOp& o = c.ops[nops++];
    int lo = 0, n_op = c.mma_n;
    if (trim && !first) {
      lo = max(0, c_lo) / 16 * 16;
      n_op = (min(c.mma_n, c_hi) - lo + 15) / 16 * 16;
      if (lo + n_op > c.mma_n) { lo = 0; n_op = c.mma_n; }
    }
    o.a_lo = (a_off >> 4) | ((a_lbo >> 4) << 16);
    o.b_lo = ((b_off + (uint32_t)lo * 16u) >> 4) | ((b_lbo >> 4) << 16);
    o.meta = (uint32_t)(col + lo) | ((first ? 1u : 0u) << 16) | ((uint32_t)dz << 20) | ((uint32_t)((c.mma_n - n_op) >> 3) << 24);
    o.pad = 0;
  };
  auto add_img = [&](int tap0, int cb0, int tap1, int cb1) {
    pk.ops[nimg].tap[0] = (int16_t)tap0; pk.ops[nimg].cbase[0] = (int16_t)cb0;
    pk.ops[nimg].tap[1] = (int16_t)tap1; pk.ops[nimg].cbase[1] = (int16_t)cb1;
    return nimg++;
  };
  bool seen_cls[8] = {false, false, false, false, false, false, false, false};
  if (ntaps * (cin >= 16 ? cin / 16 : 1) > kMaxOps) return false;
  if (master) {
    // images: one per (kh, kw, channel pair); op of plane dz = window starting at row group zf+1-dz
    for (int t = 0; t < 9; t += kwn)
      for (int j = 0; j < cin / 16; ++j) add_img(t, 16 * j, t, 16 * j + 8);
    for (int i = 0; i < ntaps; ++i)
      for (int j = 0; j < cin / 16; ++j) {
        const bool first = !seen_cls[0];
        seen_cls[0] = true;
        const int img = ((taps[i].widx % 9) / kwn) * (cin / 16) + j;
        // input plane dz of the step feeds the output planes max(0, dz - 2) .. min(zf - 1, dz) only
        const int j_lo = taps[i].dz > 2 ? taps[i].dz - 2 : 0, j_hi = taps[i].dz < zf - 1 ? taps[i].dz : zf - 1;
        add_op(taps[i].dz, (uint32_t)(2 * j * c.PS + taps[i].pos * 16), (uint32_t)c.PS,
               (uint32_t)(img * m_bytes + (zf + 1 - taps[i].dz) * grp * 16), (uint32_t)(m_rows * 16), 0, first,
               j_lo * grp, (j_hi + 1) * grp);
      }
    c.b_bytes = nimg * m_bytes;
  } else if (c.dmerge) {
    const int img_bytes = 2 * c.mma_n * 16;
    for (int i = 0; i < ntaps; ++i)
      for (int j = 0; j < cin / 16; ++j) {
        const bool first = nops == 0;
        const int img = add_img(taps[i].widx, 16 * j, taps[i].widx, 16 * j + 8);
        int cls_hi = 0;                              // classes that use input shift sh: cls & sh == 0
        for (int cls = 0; cls < 8; ++cls) if ((cls & taps[i].widx) == 0) cls_hi = cls + 1;
        add_op(taps[i].dz, (uint32_t)(2 * j * c.PS + taps[i].pos * 16), (uint32_t)c.PS, (uint32_t)(img * img_bytes),
               (uint32_t)(c.mma_n * 16), 0, first, 0, cls_hi * c.cw);
      }
    c.b_bytes = nimg * img_bytes;
  } else if (cin >= 16) {
    for (int i = 0; i < ntaps; ++i)
      for (int j = 0; j < cin / 16; ++j) {
        const bool first = !seen_cls[taps[i].cls];
        seen_cls[taps[i].cls] = true;
        const int img = add_img(taps[i].widx, 16 * j, taps[i].widx, 16 * j + 8);
        // rider launch: input plane 0 of the step only feeds the even output plane (columns [0, grp)), plane 3 only the
        // odd one (columns [grp, ncols))
        const int c_lo = n2 && taps[i].dz == 3 ? grp : 0, c_hi = n2 && taps[i].dz == 0 ? grp : 1 << 20;
        add_op(taps[i].dz, (uint32_t)(2 * j * c.PS + taps[i].pos * 16), (uint32_t)c.PS, (uint32_t)(img * b_op_bytes),
               (uint32_t)(CP * 16), taps[i].cls * CP, first, c_lo, c_hi);
      }
    c.b_bytes = nimg * b_op_bytes;
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
        const int img = add_img(taps[i].widx, 0, taps[mate].widx, 0);
        add_op(taps[i].dz, (uint32_t)(taps[i].pos * 16), (uint32_t)((taps[mate].pos - taps[i].pos) * 16),
               (uint32_t)(img * b_op_bytes), (uint32_t)(CP * 16), taps[i].cls * CP, first);
      } else {
        const int img = add_img(taps[i].widx, 0, -1, 0);
        add_op(taps[i].dz, (uint32_t)(taps[i].pos * 16), 16u, (uint32_t)(img * b_op_bytes), (uint32_t)(CP * 16),
               taps[i].cls * CP, first);
      }
    }
    c.b_bytes = nimg * b_op_bytes;
  }
  if (c.b_bytes >= (1 << 18)) return false;
  c.nops = nops;
  // op ranges per input plane of the step (ops are in plane-major order)
  for (int d = 0; d <= kMaxSpan; ++d) c.dz_begin[d] = nops;
  for (int o = nops - 1; o >= 0; --o) c.dz_begin[(c.ops[o].meta >> 20) & 15] = o;
  for (int d = kMaxSpan - 1; d >= 0; --d) if (c.dz_begin[d] > c.dz_begin[d + 1]) c.dz_begin[d] = c.dz_begin[d + 1];
  pk.zf = zf; pk.master = master ? 1 : 0; pk.xfold = xfold ? 1 : 0; pk.dmerge = c.dmerge; pk.cw = c.cw;
  pk.nops = nimg; pk.Cin = cin; pk.Cout = cout; pk.cout_base = cout_base; pk.cout_n = cout_n; pk.CP = CP;
  pk.CinT = cin; pk.CoutT = cout;
  pk.kernel_tf2 = nullptr; pk.n2 = n2; pk.CoutT2 = n2;
  pk.transposed = mode == MODE_DECONV;
  const size_t fixed = (size_t)c.b_bytes + (size_t)c.ring_pad + (size_t)kMaxOps * 16 + 32 + 2048 + (3 * kMaxRing + 2 * kMaxSkipRing + 5) * sizeof(uint64_t) + 16;
  c.RS = has_skip ? kMinSkipRing : 0;
  size_t skip_bytes = (size_t)c.RS * c.slot_bytes;
  c.R = c.span + c.zstep;                         // the planes of the next step land while this one computes
  if (fixed + skip_bytes + (size_t)c.R * c.slot_bytes > kSmemBudget) {
    c.R = c.span + 1;
    if (fixed + skip_bytes + (size_t)c.R * c.slot_bytes > kSmemBudget) return false;
  }
  // deepen the rings while shared memory allows: more planes in flight hide the L2 / HBM latency.  A skip plane
  // is held until its input plane is transformed, so the skip ring bounds the producer's lead.
  for (;;) {
    const bool grow_skip = has_skip && c.RS < kMaxSkipRing && c.RS < c.R - c.span + c.zstep + 1;
    const bool grow_ring = c.R < kMaxRing && c.R < c.span + 3 * c.zstep;
    if (grow_skip && fixed + skip_bytes + c.slot_bytes + (size_t)c.R * c.slot_bytes <= kSmemBudget) {
      ++c.RS; skip_bytes += c.slot_bytes;
    } else if (grow_ring && fixed + skip_bytes + (size_t)(c.R + 1) * c.slot_bytes <= kSmemBudget) {
      ++c.R;
    } else {
      break;
    }
  }
  pl->smem = fixed + skip_bytes + (size_t)c.R * c.slot_bytes;
  return c.slot_bytes < (1 << 18) && (size_t)c.PS < (1u << 18);
}

// Cycle model of one launch (per SM): the roles run concurrently, a step costs the slowest of them.
double estimate_clk(const Params& c, int sm_count) {
  const double mma = (double)c.nops * c.MB * mma_clk(c.mma_n);
  const double plane_bytes = (double)c.nsub * c.NCH * c.RY * c.PX * 16.0 * (c.has_skip ? 2.0 : 1.0);
  const double load = plane_bytes * c.zstep / 18.0;                       // ~HBM share of one SM, B/clk
  const int ncls = c.mode == MODE_DECONV ? 8 : 1;
  const double epi = c.xfold ? (c.cout_n == 1 ? (double)c.MB * 300.0
                                              : (double)c.MB * (c.n2 ? 2 : c.zf * (c.cout_n / 8)) * 420.0)      // rider: three warp groups side by side
                             : (double)c.MB * ncls * (c.CP * 4.0 * 128.0 / 110.0 + 12.0 * c.CP + 80.0);
  const double tma_issue = (c.one_box ? 1.0 : (double)c.nsub * c.NCH) * (c.has_skip ? 2.0 : 1.0) * 250.0 * c.zstep;
  const double xf = c.transform ? (double)c.xf_k / c.xf_groups * c.zstep * (c.has_skip ? 70.0 : 45.0) : 0.0;
  double step = mma;
  if (load > step) step = load;
  if (epi > step) step = epi;
  if (xf > step) step = xf;
  if (tma_issue > step) step = tma_issue;
  step += 120.0;
  const int steps_all = ceil_div(c.Mz, c.zf);
  const int sseg = ceil_div(steps_all, c.zsplit);
  const double fill = plane_bytes * c.span / 18.0 + 1500.0;
  const double cta = sseg * step + fill + 3000.0 + c.b_bytes / 32.0;
  const int ctas = c.tiles_x * c.tiles_y * c.zsplit;
  const int waves = ceil_div(ctas, sm_count);
  return waves * cta;
}

std::map<std::array<int, 15>, Plan> g_plan_cache;
std::mutex g_plan_mutex;

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult qres;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)f;
  }
  return fn;
}

// 5-D map over a CP8 (subs = 1) or PS8 (subs = 4) tensor: (8*Wp, Hp, subs, NCH, D), box (8*PX, RY, 1, 1, 1)
bool make_tmap(CUtensorMap* tm, const void* base, int Wp, int Hp, int subs, int nch, int D, int PX, int RY,
               int box_ch = 1, int box_sub = 1) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[5] = {(cuuint64_t)Wp * 8, (cuuint64_t)Hp, (cuuint64_t)subs, (cuuint64_t)nch, (cuuint64_t)D};
  cuuint64_t gstr[4] = {(cuuint64_t)Wp * 16, (cuuint64_t)Hp * Wp * 16, (cuuint64_t)subs * Hp * Wp * 16,
                        (cuuint64_t)nch * subs * Hp * Wp * 16};
  cuuint32_t box[5] = {(cuuint32_t)PX * 8, (cuuint32_t)RY, (cuuint32_t)box_sub, (cuuint32_t)box_ch, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace
}  // namespace tc

using namespace tc;

// Plan (tile, folds, z split, op table) of one launch: output channels [cb, cb+cn) of a layer.  Cached per shape.
static bool find_plan(int mode, int D, int H, int W, int cin, int cout, int cb, int cn, bool has_skip, bool transform,
                      int sm_count, Plan* out, int* dbg_out, int n2 = 0) {
  // tuning / debugging switches (mvsb200_set_tuning); TC_LAYER="cin,cout,mode" restricts them to one layer shape
  const Tuning& tn = tuning();
  const bool mine = !tn.tc_layer_set || (tn.tc_layer[0] == cin && tn.tc_layer[1] == cout && tn.tc_layer[2] == mode);
  const int zf_forced = mine ? tn.tc_zf : -1;                    // -1: planner's choice
  const int force_tx = mine ? tn.tc_tile_x : 0, force_ty = mine ? tn.tc_tile_y : 0;
  const int force_zs = mine ? tn.tc_zsplit : 0;
  const bool no_xfold = mine && tn.tc_xfold == 0;
  const int dbg_forced = mine ? tn.tc_dbg : 0;
  const int rank = mine && tn.tc_rank > 0 ? tn.tc_rank : 0;      // development: the rank-th best plan of the cycle model
    const int Mx = mode == MODE_CONV2 ? ceil_div(W, 2) : W, My = mode == MODE_CONV2 ? ceil_div(H, 2) : H,
              Mz = mode == MODE_CONV2 ? ceil_div(D, 2) : D;
    const std::array<int, 15> key = {mode, D, H, W, cin, cout, cb, has_skip, transform, zf_forced > 0 ? zf_forced : 0,
                                     force_tx, force_ty, sm_count, force_zs * 2 + (no_xfold ? 1 : 0) + 64 * rank + (mine && tn.tc_xfold == 2 ? 32 : 0) + (tn.tc_trim == 0 ? 1 << 20 : 0), n2};
    Plan best;
    bool found = false;
    {
      std::lock_guard<std::mutex> lock(g_plan_mutex);
      auto it = g_plan_cache.find(key);
      if (it != g_plan_cache.end()) { best = it->second; found = true; }
    }
    std::vector<Plan> ranked;
    if (!found) {
      const int zf_cands[] = {4, 2, 1};
      for (int zi = 0; zi < 3; ++zi) {
        const int zf = zf_cands[zi];
        if (mode != MODE_CONV1 && zf != 1) continue;
        if (zf_forced > 0 && zf_forced != zf && mode == MODE_CONV1 && !n2) continue;
        if (zf > 1 && zf_forced <= 0 && Mz < 2 * zf) continue;
        if (n2 && zf != 2) continue;
        for (int xf = (mode == MODE_CONV1 && !no_xfold) ? 1 : 0; xf >= (mine && tn.tc_xfold == 2 && mode == MODE_CONV1 ? 1 : 0); --xf)
        for (int TX = 4; TX <= 30; ++TX) {
          if (xf && TX != 6 && TX != 14 && TX != 30) continue;
          if (!xf && TX > Mx && TX != 4 && TX - 1 >= Mx) break;        // one clipped candidate is enough
          const int tx_eff = xf ? TX : (TX < Mx ? TX : Mx);
          if (force_tx && tx_eff != (xf ? force_tx : (force_tx < Mx ? force_tx : Mx))) continue;
          for (int TY = 1; TY <= 64 && (TY <= My || xf); ++TY) {
            if (force_ty && TY != (force_ty < My || xf ? force_ty : My)) continue;
            Plan pl;
            if (!build_plan(mode, D, H, W, cin, cout, cb, cn, tx_eff, TY, zf, xf != 0, has_skip, transform, n2, &pl)) continue;
            if (pl.smem > kSmemBudget) continue;
            Params& c = pl.cp;
            const int tiles = c.tiles_x * c.tiles_y;
            const int steps_all = ceil_div(Mz, zf);
            // z split: candidates that fill whole waves of SMs
            int zcands[6] = {1, 0, 0, 0, 0, 0};
            int nz = 1;
            for (int waves = 1; waves <= 4 && nz < 6; ++waves) {
              int z = (waves * sm_count) / tiles;
              if (z < 1) z = 1;
              if (z > steps_all) z = steps_all;
              zcands[nz++] = z;
            }
            if (force_zs) { zcands[0] = force_zs < steps_all ? force_zs : steps_all; nz = 1; }
            for (int zi2 = 0; zi2 < nz; ++zi2) {
              const int sseg = ceil_div(steps_all, zcands[zi2]);
              c.zsplit = ceil_div(steps_all, sseg);           // drops empty trailing segments
              pl.est_clk = estimate_clk(c, sm_count);
              if (rank > 0) ranked.push_back(pl);
              if (!found || pl.est_clk < best.est_clk) { best = pl; found = true; }
            }
          }
        }
      }
      if (found && rank > 0) {
        // distinct plans in the model's order (z-split candidates repeat)
        std::stable_sort(ranked.begin(), ranked.end(), [](const Plan& a, const Plan& b) { return a.est_clk < b.est_clk; });
        std::vector<Plan> uniq;
        for (const Plan& q : ranked) {
          bool dup = false;
          for (const Plan& u : uniq)
            dup = dup || (u.cp.TX == q.cp.TX && u.cp.TY == q.cp.TY && u.cp.zf == q.cp.zf && u.cp.xfold == q.cp.xfold && u.cp.zsplit == q.cp.zsplit);
          if (!dup) uniq.push_back(q);
          if ((int)uniq.size() > rank) break;
        }
        best = uniq[(size_t)rank < uniq.size() ? rank : uniq.size() - 1];
      }
      if (found) {
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        g_plan_cache[key] = best;
      }
    }
    if (found) *out = best;
    if (dbg_out) *dbg_out = dbg_forced;
    return found;
}

size_t conv3d_tc_pack_slot_bytes() { return align_up((size_t)kMaxOps * 2 * 32 * 16, 256); }
size_t conv3d_tc_scratch_bytes() { return conv3d_tc_pack_slot_bytes() * 2; }

// ---------------------------------------------------------------------------------------------
// all weights of a network in one launch
// ---------------------------------------------------------------------------------------------
constexpr int kMaxPackJobs = 24;
struct PackAll { int n; PackParams job[kMaxPackJobs]; };

__global__ void pack_all_kernel(const __grid_constant__ PackAll a) {
  const PackParams& p = a.job[blockIdx.y];
  if (p.n2) { pack_rider(p); return; }
  if (p.dmerge) {
    // transposed conv, classes merged: image = [2 halves][8 classes x cw rows][8]; the op's tap code is its input
    // shift (bit 2: z-1, bit 1: y-1, bit 0: x-1); class (pz,py,px) takes filter tap k = parity + 2 on a shifted axis
    const int rows = 8 * p.cw;
    const int total = p.nops * 2 * rows * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int k8 = i & 7, n = (i >> 3) % rows, half = (i / (8 * rows)) & 1, op = i / (16 * rows);
      const int cls = n / p.cw, cn = n - cls * p.cw;
      const int sh = p.ops[op].tap[half], ci = p.ops[op].cbase[half] + k8;
      const int pz = cls >> 2, py = (cls >> 1) & 1, px = cls & 1;
      const int sz = (sh >> 2) & 1, sy = (sh >> 1) & 1, sx = sh & 1;
      float w = 0.0f;
      if ((!sz || !pz) && (!sy || !py) && (!sx || !px) && cn < p.cout_n && ci < p.CinT && p.cout_base + cn < p.CoutT) {
        const int tap = ((pz + 2 * sz) * 3 + (py + 2 * sy)) * 3 + (px + 2 * sx);
        w = p.kernel_tf[((size_t)tap * p.CoutT + p.cout_base + cn) * p.CinT + ci];
      }
      const __nv_bfloat16 h = __float2bfloat16_rn(w);
      p.out[i] = *reinterpret_cast<const uint16_t*>(&h);
    }
    return;
  }
  const int kwn = p.xfold ? 3 : 1, grp = kwn * p.cout_n;
  if (!p.master) {
    const int total = p.nops * 2 * p.CP * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int k8 = i & 7, n = (i >> 3) % p.CP, half = (i / (8 * p.CP)) & 1, op = i / (16 * p.CP);
      const int j = n / grp, rem = n - j * grp, kw = rem / p.cout_n, cn = rem - kw * p.cout_n;
      int tap = p.ops[op].tap[half];
      const int ci = p.ops[op].cbase[half] + k8;
      if (tap >= 0) tap += (p.xfold ? kw : 0) - 9 * j;
      float w = 0.0f;
      if (tap >= 0 && tap < 27 && j < p.zf && ci < p.CinT && p.cout_base + cn < p.CoutT) {
        const int co = p.cout_base + cn;
        w = p.transposed ? p.kernel_tf[((size_t)tap * p.CoutT + co) * p.CinT + ci]
                         : p.kernel_tf[((size_t)tap * p.CinT + ci) * p.CoutT + co];
      }
      const __nv_bfloat16 h = __float2bfloat16_rn(w);
      p.out[i] = *reinterpret_cast<const uint16_t*>(&h);
    }
  } else {
    const int groups = 2 * p.zf + 1, rows = groups * grp;
    const int total = p.nops * 2 * rows * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int k8 = i & 7, row = (i >> 3) % rows, half = (i / (8 * rows)) & 1, m = i / (16 * rows);
      const int g = row / grp, rem = row - g * grp, kw = rem / p.cout_n, cn = rem - kw * p.cout_n, kd = p.zf + 1 - g;
      const int ci = p.ops[m].cbase[half] + k8;
      float w = 0.0f;
      if (kd >= 0 && kd < 3 && ci < p.CinT && p.cout_base + cn < p.CoutT) {
        const int tap = kd * 9 + p.ops[m].tap[0] + (p.xfold ? kw : 0);
        w = p.kernel_tf[((size_t)tap * p.CinT + ci) * p.CoutT + p.cout_base + cn];
      }
      const __nv_bfloat16 h = __float2bfloat16_rn(w);
      p.out[i] = *reinterpret_cast<const uint16_t*>(&h);
    }
  }
}

int conv3d_tc_pack_all(const TcPackJob* jobs, int njobs, void* dst_base, cudaStream_t s) {
  static PackAll a;      // too large for the stack of some callers; filled and consumed under the launch below
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  const int sm_count = sm_count_current();
  a.n = 0;
  for (int j = 0; j < njobs; ++j) {
    const TcPackJob& jb = jobs[j];
    const int mode = jb.transposed ? MODE_DECONV : (jb.stride == 2 ? MODE_CONV2 : MODE_CONV1);
    int slot = jb.slot0;
    for (int cb = 0; cb < jb.cout; cb += 32, ++slot) {
      const int cn = jb.cout - cb < 32 ? jb.cout - cb : 32;
      Plan pl;
      if (!find_plan(mode, jb.D, jb.H, jb.W, jb.cin, jb.cout, cb, cn, jb.has_skip != 0, jb.transform != 0, sm_count, &pl,
                     nullptr, jb.cout2)) {
        set_error("conv3d(bf16/tcgen05): no tile fits (Cin=%d Cout=%d mode=%d)", jb.cin, jb.cout, mode);
        return MVSB200_ERR_UNSUPPORTED;
      }
      MVS_CHECK_ARG(a.n < kMaxPackJobs, "conv3d_tc_pack_all: too many launches");
      a.job[a.n] = pl.pp;
      a.job[a.n].kernel_tf = jb.kernel_tf;
      if (jb.cin_true > 0) a.job[a.n].CinT = jb.cin_true;
      if (jb.cout_true > 0) a.job[a.n].CoutT = jb.cout_true;
      if (jb.cout2 > 0) {
        a.job[a.n].kernel_tf2 = jb.kernel_tf2;
        a.job[a.n].CoutT2 = jb.cout2_true > 0 ? jb.cout2_true : jb.cout2;
      }
      a.job[a.n].out = (uint16_t*)((unsigned char*)dst_base + (size_t)slot * conv3d_tc_pack_slot_bytes());
      ++a.n;
    }
  }
  pack_all_kernel<<<dim3(16, a.n), 256, 0, s>>>(a);
  MVS_LAUNCH_CHECK("pack_all_kernel");
  return MVSB200_OK;
}

using TcKernel = void (*)(const Params);
// 0: <16 classic>, 1: <32 classic>, 2 / 3 / 4: x-fold with 8 / 16 / 32 channels per launch, 5: x-fold, Cout = 1
static TcKernel kernel_variant(int k, bool peer_mode) {
  switch (k) {
    case 0: return peer_mode ? conv3d_tc_kernel<16, 0, 1> : conv3d_tc_kernel<16, 0, 0>;
    case 1: return peer_mode ? conv3d_tc_kernel<32, 0, 1> : conv3d_tc_kernel<32, 0, 0>;
    case 2: return peer_mode ? conv3d_tc_kernel<32, 1, 1> : conv3d_tc_kernel<32, 1, 0>;
    case 3: return peer_mode ? conv3d_tc_kernel<32, 2, 1> : conv3d_tc_kernel<32, 2, 0>;
    case 4: return peer_mode ? conv3d_tc_kernel<32, 4, 1> : conv3d_tc_kernel<32, 4, 0>;
    case 5: return peer_mode ? conv3d_tc_kernel<32, 1, 3> : conv3d_tc_kernel<32, 1, 2>;
    default: return conv3d_tc_kernel<32, 2, 4>;      // rider launch (no peer-memory flavour)
  }
}

// x: CP8 for stride-1 convs and transposed convs, PS8 for stride-2 convs.  skip: CP8.
// Outputs: y_cp8 and / or y_ps8 (bf16, Cout % 8 == 0), or y_f32 (NDHWC fp32, any Cout).
int launch_conv3d_tc(const void* x, const float* xs, const float* xb, const void* skip, const float* ss,
                     const float* sb, const float* kernel_tf, int D, int H, int W, int cin, int cout, int stride,
                     int transposed, void* y_cp8, void* y_ps8, float* y_f32, double* stats, void* scratch,
                     const TcBnSrc* x_bn, const TcBnSrc* s_bn, const void* prepacked, int stats_reps,
                     int stats_rep_stride, const TcSlab* slab, const TcPeer* peer, TcRegress* regress,
                     cudaStream_t s, const TcRider* rider) {
  if (rider && (peer || (slab && slab->halo) || transposed || stride != 1 || cout > 32 || y_f32 || skip || xs || (x_bn && x_bn->stats))) {
    set_error("conv3d(bf16/tcgen05): a rider needs a plain stride-1 conv on a raw input, one launch, one GPU");
    return MVSB200_ERR_UNSUPPORTED;
  }
  if (cin != 8 && cin != 16 && cin != 32 && cin != 64) {
    set_error("conv3d(bf16/tcgen05): Cin=%d unsupported (need 8, 16, 32 or 64)", cin);
    return MVSB200_ERR_UNSUPPORTED;
  }
  if (!y_f32 && (cout % 8 != 0 || (!y_cp8 && !y_ps8))) {
    set_error("conv3d(bf16/tcgen05): bf16 output needs Cout %% 8 == 0 (got %d)", cout);
    return MVSB200_ERR_UNSUPPORTED;
  }
  const int mode = transposed ? MODE_DECONV : (stride == 2 ? MODE_CONV2 : MODE_CONV1);
  if (skip && mode == MODE_CONV2) {
    set_error("conv3d(bf16/tcgen05): skip input on a stride-2 conv is not supported");
    return MVSB200_ERR_UNSUPPORTED;
  }
  if (!get_encode()) {
    set_error("conv3d(bf16/tcgen05): cuTensorMapEncodeTiled is not available from the driver");
    return MVSB200_ERR_CUDA;
  }
  const int sm_count = sm_count_current();
  int dev = 0;
  MVS_CUDA(cudaGetDevice(&dev));
  // function attributes are per device: once per device of this process (a set bit is only ever re-set to itself)
  static std::atomic<uint64_t> attr_done{0};
  if (!(attr_done.load(std::memory_order_acquire) >> (dev & 63) & 1u)) {
    for (int peer_mode = 0; peer_mode < 2; ++peer_mode)
      for (int k = 0; k < 7; ++k)
        MVS_CUDA(cudaFuncSetAttribute((const void*)kernel_variant(k, peer_mode), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kSmemBudget));
    attr_done.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  const bool has_skip = skip != nullptr, transform = xs != nullptr || has_skip || (x_bn && x_bn->stats);
  int launch_idx = 0;
  for (int cb = 0; cb < cout; cb += 32, ++launch_idx) {
    const int cn = cout - cb < 32 ? cout - cb : 32;
    Plan best;
    int dbg_flags = 0;
    if (!find_plan(mode, D, H, W, cin, cout, cb, cn, has_skip, transform, sm_count, &best, &dbg_flags, rider ? rider->cout2 : 0)) {
      set_error("conv3d(bf16/tcgen05): no tile fits (Cin=%d Cout=%d mode=%d)", cin, cout, mode);
      return MVSB200_ERR_UNSUPPORTED;
    }
    Params& c = best.cp;
    c.xs = xs; c.xb = xb; c.ss = ss; c.sb = sb;
    c.xbn = x_bn ? *x_bn : TcBnSrc{nullptr, nullptr, nullptr, 1.0, 0.f, 0, 1, 0};
    c.sbn = s_bn ? *s_bn : TcBnSrc{nullptr, nullptr, nullptr, 1.0, 0.f, 0, 1, 0};
    c.stats_reps = stats_reps > 0 ? stats_reps : 1; c.stats_rep_stride = stats_rep_stride;
    for (int t = 0; t < 2; ++t) {
      c.mir_prev[t] = peer ? (__nv_bfloat16*)peer->mir_prev[t] : nullptr;
      c.mir_next[t] = peer ? (__nv_bfloat16*)peer->mir_next[t] : nullptr;
      c.wait_flags[t] = peer ? peer->wait_flags[t] : nullptr;
    }
    c.wait_n = peer ? peer->wait_n : 0; c.wait_seq = peer ? peer->wait_seq : 0u; c.err_flag = peer ? peer->err_flag : nullptr;
    // soft-argmin folded into the epilogue: only the x-folded single-channel layer whose CTAs walk the whole depth
    c.rg_out = nullptr; c.rg_start = 0.f; c.rg_step = 0.f;
    if (regress) {
      regress->fused = 0;
      if (regress->partial && c.xfold && cout == 1 && c.zsplit == 1 && y_f32 && !(slab && slab->halo)) {
        c.rg_out = regress->partial; c.rg_start = regress->start; c.rg_step = regress->step;
        regress->fused = 1;
      }
    }
    c.y_cp8 = (__nv_bfloat16*)y_cp8; c.y_ps8 = (__nv_bfloat16*)y_ps8; c.y_f32 = y_f32; c.stats = stats;
    if (rider) {
      c.y2_cp8 = (__nv_bfloat16*)rider->y2_cp8; c.y2_ps8 = (__nv_bfloat16*)rider->y2_ps8; c.stats2 = rider->stats2;
      best.pp.kernel_tf2 = rider->kernel_tf2;
    }
    // D-slab mode: the tensors hold D + 2 planes (halo before / after), local plane l is extended plane l + 1
    const int Dt = slab && slab->halo ? D + 2 : D;
    c.zv_lo = 0; c.zv_hi = D;
    if (slab && slab->halo) { c.zoff += 1; c.zv_lo = slab->zv_lo; c.zv_hi = slab->zv_hi; }
    const bool ok_x = mode == MODE_CONV2
                          ? make_tmap(&c.tmap_x, x, (W + 1) / 2, (H + 1) / 2, 4, c.NCH, Dt, c.PX, c.RY, c.one_box ? c.NCH : 1,
                                      c.one_box ? 4 : 1)
                          : make_tmap(&c.tmap_x, x, W, H, 1, c.NCH, Dt, c.PX, c.RY, c.one_box ? c.NCH : 1);
    const bool ok_s = !has_skip || make_tmap(&c.tmap_s, skip, W, H, 1, c.NCH, Dt, c.PX, c.RY, c.one_box ? c.NCH : 1);
    if (!ok_x || !ok_s) {
      set_error("conv3d(bf16/tcgen05): cuTensorMapEncodeTiled failed (PX=%d RY=%d W=%d H=%d)", c.PX, c.RY, W, H);
      return MVSB200_ERR_CUDA;
    }
    {
      c.dbg = dbg_flags;
      if (tuning().tc_verbose)
        fprintf(stderr, "[tc] mode=%d Cin=%d Cout=%d(+%d) tile %dx%d PX=%d RY=%d MB=%d N=%d R=%d zf=%d xf=%d zsplit=%d grid=%d smem=%zu "
                "RS=%d nops=%d b=%dB xf=%d/%d est=%.0f clk rider=%d\n",
                mode, cin, cn, cb, c.TX, c.TY, c.PX, c.RY, c.MB, c.CP, c.R, c.zf, c.xfold, c.zsplit, c.tiles_x * c.tiles_y * c.zsplit,
                best.smem, c.RS, c.nops, c.b_bytes, c.transform, c.xf_k, best.est_clk, c.n2);
    }
    if (prepacked) {
      // weights were packed for the whole network in one launch (conv3d_tc_pack_all), one slot per launch
      c.wpacked = (const uint4*)((const unsigned char*)prepacked + (size_t)launch_idx * conv3d_tc_pack_slot_bytes());
    } else {
      unsigned char* wp = (unsigned char*)scratch + (size_t)(launch_idx & 1) * conv3d_tc_pack_slot_bytes();
      c.wpacked = (const uint4*)wp;
      best.pp.kernel_tf = kernel_tf;
      best.pp.out = (uint16_t*)wp;
      pack_weights_kernel<<<ceil_div(c.b_bytes / 2, 256), 256, 0, s>>>(best.pp);
      MVS_LAUNCH_CHECK("pack_weights_kernel");
    }
    const int grid = c.tiles_x * c.tiles_y * c.zsplit;
    static long long* prof_buf = nullptr;
    c.prof = nullptr;
    const bool prof = tuning().tc_prof != 0;
    if (prof && kProf) {
      if (!prof_buf) MVS_CUDA(cudaMalloc(&prof_buf, 256 + 16 * 4096));
      c.prof = prof_buf;
    }
    // every launch asks for the same (maximum) dynamic shared memory: a different carve-out per layer makes the
    // SMs re-partition L1 / shared memory between launches
    const bool exact_smem = tuning().tc_exact_smem != 0;
    const size_t smem_launch = exact_smem ? best.smem : kSmemBudget;
    // programmatic dependent launch (prologue of this layer under the tail of the previous kernel)
    const bool no_pdl = tuning().tc_no_pdl != 0;
    c.pdl = no_pdl ? 0 : (grid <= sm_count ? 1 : 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem_launch; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = c.pdl ? 1 : 0;
    cudaError_t lerr = cudaSuccess;
    // instantiation: epilogue shape x (peer-memory D-slab mode or not)
    const int variant = c.n2 ? 6 : c.xfold ? (c.cout_n == 1 ? 5 : c.cout_n <= 8 ? 2 : c.cout_n <= 16 ? 3 : 4) : (c.CP == 16 ? 0 : 1);
    auto launch = [&]() { lerr = cudaLaunchKernelEx(&cfg, kernel_variant(variant, peer != nullptr), c); };
    launch();
    if (lerr != cudaSuccess) {
      set_error("launch of conv3d_tc_kernel failed: %s", cudaGetErrorString(lerr));
      return MVSB200_ERR_CUDA;
    }
    MVS_LAUNCH_CHECK("conv3d_tc_kernel");
    if (prof) {
      long long h[24];
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaStreamSynchronize(s);
      // second, timed launch of the same kernel (the first one may have overlapped the pack kernel's tail)
      cudaEventRecord(e0, s);
      launch();
      cudaEventRecord(e1, s);
      cudaStreamSynchronize(s);
      float kms = 0.f;
      cudaEventElapsedTime(&kms, e0, e1);
      cudaEventDestroy(e0); cudaEventDestroy(e1);
      fprintf(stderr, "[tc-prof] kernel alone %.1f us (grid %d); ", kms * 1e3, grid);
      if (!kProf) {
        fprintf(stderr, "mode=%d Cin=%d Cout=%d (per-role counters: build with -DMVSB200_TC_PROF_BUILD=1)\n", mode, cin, cn);
        continue;
      }
      cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "CTA 0: prologue %.1f us, MMA warp done at %.1f us, CTA end %.1f us\n", h[6] * 1e-3, h[7] * 1e-3, h[8] * 1e-3);
      fprintf(stderr, "[tc-prof] producer %lld clk (wait empty %lld, wait skip-empty %lld); transform %lld clk (wait landed %lld); "
              "epilogue %lld clk (wait acc %lld)\n", h[9], h[10], h[11], h[12], h[13], h[14], h[15]);
      fprintf(stderr, "[tc-prof] producer: %lld clk in expect_tx + TMA issue\n", h[16]);
      fprintf(stderr, "[tc-prof] CTA 0 timeline (us from entry): ring cleared %.2f, barriers %.2f, grid dependency %.2f, scale/shift %.2f, "
              "prologue barrier %.2f, weights %.2f, first step's planes %.2f\n", h[17] * 1e-3, h[18] * 1e-3, h[19] * 1e-3,
              h[20] * 1e-3, h[21] * 1e-3, h[6] * 1e-3, h[22] * 1e-3);
      if (grid <= 4096) {
        static long long hh[2 * 4096];
        cudaMemcpy(hh, prof_buf + 32, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
        long long t0 = hh[0], t1 = hh[1], smax = hh[0], emin = hh[1];
        double avg = 0;
        for (int i = 0; i < grid; ++i) {
          if (hh[2 * i] < t0) t0 = hh[2 * i];
          if (hh[2 * i] > smax) smax = hh[2 * i];
          if (hh[2 * i + 1] > t1) t1 = hh[2 * i + 1];
          if (hh[2 * i + 1] < emin) emin = hh[2 * i + 1];
          avg += (double)(hh[2 * i + 1] - hh[2 * i]);
        }
        fprintf(stderr, "[tc-prof] CTAs: first start 0, last start %.1f us, first end %.1f us, last end %.1f us, mean life %.1f us\n",
                (smax - t0) * 1e-3, (emin - t0) * 1e-3, (t1 - t0) * 1e-3, avg / grid * 1e-3);
      }
      fprintf(stderr, "[tc-prof] mode=%d Cin=%d Cout=%d: MMA warp of CTA 0: %lld clk in %lld ns (%.0f MHz); wait input %lld, "
              "wait acc %lld, issue %lld clk; %lld MMAs -> %.1f clk/MMA issued, %.1f clk/MMA overall\n", mode, cin, cn,
              h[0], h[1], h[1] ? 1e3 * h[0] / h[1] : 0.0, h[2], h[3], h[4], h[5], h[5] ? (double)h[4] / h[5] : 0.0,
              h[5] ? (double)h[0] / h[5] : 0.0);
    }
  }
  return MVSB200_OK;
}

// Host-only view of the planner (no device work): the plan the bf16 path would use for a layer shape, as text, and
// its invariants as numbers (tests/test_planner.py).  out[0..11] = {launches, TX, TY, PX, RY, MB, mma_n, zf, xfold,
// ring, smem_bytes, tmem_cols} of the first launch.
int conv3d_tc_describe(int D, int H, int W, int cin, int cout, int stride, int transposed, int has_skip, int transform,
                       int sm_count, int* out, char* text, int text_len) {
  const int mode = transposed ? MODE_DECONV : (stride == 2 ? MODE_CONV2 : MODE_CONV1);
  if (cin != 8 && cin != 16 && cin != 32 && cin != 64) {
    set_error("conv3d(bf16/tcgen05): Cin=%d unsupported (need 8, 16, 32 or 64)", cin);
    return MVSB200_ERR_UNSUPPORTED;
  }
  int launches = 0, written = 0;
  for (int cb = 0; cb < cout; cb += 32, ++launches) {
    const int cn = cout - cb < 32 ? cout - cb : 32;
    Plan pl;
    if (!find_plan(mode, D, H, W, cin, cout, cb, cn, has_skip != 0, transform != 0, sm_count > 0 ? sm_count : 148 /* B200: host-only planning without a device */, &pl,
                   nullptr)) {
      set_error("conv3d(bf16/tcgen05): no tile fits (Cin=%d Cout=%d mode=%d)", cin, cout, mode);
      return MVSB200_ERR_UNSUPPORTED;
    }
    const Params& c = pl.cp;
    if (launches == 0 && out) {
      const int v[12] = {0, c.TX, c.TY, c.PX, c.RY, c.MB, c.mma_n, c.zf, c.xfold, c.R, (int)pl.smem, c.tmem_cols};
      for (int i = 0; i < 12; ++i) out[i] = v[i];
    }
    // columns the ops of one step span (ops skip their all-zero column blocks; first op full width) / nops * N; ok = every
    // op keeps N a multiple of 16 inside [0, N), starts on a multiple of 16, and the step's first op is full width
    int cols = 0, ok = 1;
    for (int o = 0; o < c.nops; ++o) {
      const int n_op = c.mma_n - (int)((c.ops[o].meta >> 24) << 3), first = (int)((c.ops[o].meta >> 16) & 1u);
      const int col = (int)(c.ops[o].meta & 0xFFFFu) % (c.NB > 0 ? c.NB : 1);
      cols += n_op;
      if (n_op < 16 || n_op % 16 || n_op > c.mma_n || (n_op != c.mma_n && (col % 16 || first))) ok = 0;
    }
    if (c.nops > 0 && !((c.ops[0].meta >> 16) & 1u)) ok = 0;
    if (text && written < text_len)
      written += snprintf(text + written, (size_t)(text_len - written),
                          "cout[%d:%d] tile %dx%d cells %dx%d blocks %d N %d zfold %d xfold %d ring %d+%d grid %dx%dx%d ops %d "
                          "smem %zu tmem %d one_box %d cols %d/%d ok %d\n", cb, cb + cn, c.TX, c.TY, c.PX, c.RY, c.MB, c.mma_n,
                          c.zf, c.xfold, c.R, c.RS, c.tiles_x, c.tiles_y, c.zsplit, c.nops, pl.smem, c.tmem_cols, c.one_box,
                          cols, c.nops * c.mma_n, ok);
  }
  if (out) out[0] = launches;
  return MVSB200_OK;
}

// ---------------------------------------------------------------------------------------------
// layout helpers for callers
// ---------------------------------------------------------------------------------------------
size_t planar_bytes(int D, int H, int W, int C, int parity_split) {
  if (!parity_split) return (size_t)D * H * W * C * 2;
  return (size_t)D * (C / 8) * 4 * ((H + 1) / 2) * ((W + 1) / 2) * 16;
}

static unsigned grid_for(size_t items) {       // 256-thread blocks, at most 32 per multiprocessor
  const size_t cap = (size_t)sm_count_current() * 32, want = (items + 255) / 256;
  return (unsigned)(want < cap ? want : cap);
}

int launch_ndhwc_to_planar(const void* x_ndhwc, int D, int H, int W, int C, void* cp8, void* ps8, cudaStream_t s) {
  MVS_CHECK_ARG(C % 8 == 0, "layout: C must be a multiple of 8 (got %d)", C);
  if (ps8 && ((H | W) & 1)) MVS_CUDA(cudaMemsetAsync(ps8, 0, planar_bytes(D, H, W, C, 1), s));
  const size_t total = (size_t)D * H * W * (C / 8);
  const unsigned blocks = grid_for(total);
  ndhwc_to_planar_kernel<<<blocks, 256, 0, s>>>((const __nv_bfloat16*)x_ndhwc, D, H, W, C, (__nv_bfloat16*)cp8,
                                                (__nv_bfloat16*)ps8);
  MVS_LAUNCH_CHECK("ndhwc_to_planar_kernel");
  return MVSB200_OK;
}

int launch_planar_to_ndhwc(const void* cp8, int D, int H, int W, int C, void* y_ndhwc, cudaStream_t s) {
  MVS_CHECK_ARG(C % 8 == 0, "layout: C must be a multiple of 8 (got %d)", C);
  const size_t total = (size_t)D * H * W * (C / 8);
  const unsigned blocks = grid_for(total);
  planar_to_ndhwc_kernel<<<blocks, 256, 0, s>>>((const __nv_bfloat16*)cp8, D, H, W, C, (__nv_bfloat16*)y_ndhwc);
  MVS_LAUNCH_CHECK("planar_to_ndhwc_kernel");
  return MVSB200_OK;
}

int launch_f32_to_bf16(const float* x, size_t n, void* y, cudaStream_t s) {
  const unsigned blocks = grid_for(n);
  f32_to_bf16_kernel<<<blocks, 256, 0, s>>>(x, n, (__nv_bfloat16*)y);
  MVS_LAUNCH_CHECK("f32_to_bf16_kernel");
  return MVSB200_OK;
}

namespace {
// stream-ordered temporaries of the stand-alone entry point, released on every exit path
struct StreamTemps {
  cudaStream_t s;
  void* ptr[8];
  int n = 0;
  explicit StreamTemps(cudaStream_t stream) : s(stream) {}
  ~StreamTemps() { for (int i = 0; i < n; ++i) cudaFreeAsync(ptr[i], s); }
  cudaError_t alloc(void** p, size_t bytes) {
    cudaError_t e = cudaMallocAsync(p, bytes, s);
    if (e == cudaSuccess) ptr[n++] = *p;
    return e;
  }
};
}  // namespace

// Stand-alone layer on NDHWC tensors (mvsb200_conv3d_layer, tests): converts to the planar layouts in
// stream-ordered temporaries, runs the layer, converts back.  The fused path never comes through here.
int launch_conv3d_tc_ndhwc(const void* x, const float* xs, const float* xb, const void* skip, const float* ss,
                           const float* sb, const float* kernel_tf, int D, int H, int W, int cin, int cout,
                           int stride, int transposed, void* y, int y_dtype, double* stats, cudaStream_t s) {
  const bool s2 = !transposed && stride == 2;
  const int Do = transposed ? 2 * D : ceil_div(D, stride), Ho = transposed ? 2 * H : ceil_div(H, stride),
            Wo = transposed ? 2 * W : ceil_div(W, stride);
  MVS_CHECK_ARG(cin % 8 == 0, "conv3d(bf16/tcgen05): Cin must be a multiple of 8 (got %d)", cin);
  StreamTemps tmp(s);
  void *xp = nullptr, *kp = nullptr, *yp = nullptr, *scratch = nullptr;
  float* yf = nullptr;
  MVS_CUDA(tmp.alloc(&xp, planar_bytes(D, H, W, cin, s2)));
  MVS_CUDA(tmp.alloc(&scratch, conv3d_tc_scratch_bytes()));
  int rc = launch_ndhwc_to_planar(x, D, H, W, cin, s2 ? nullptr : xp, s2 ? xp : nullptr, s);
  if (rc) return rc;
  if (skip) {
    MVS_CUDA(tmp.alloc(&kp, planar_bytes(D, H, W, cin, 0)));
    if ((rc = launch_ndhwc_to_planar(skip, D, H, W, cin, kp, nullptr, s))) return rc;
  }
  const bool direct_f32 = y_dtype == MVSB200_F32;
  const bool via_f32 = !direct_f32 && cout % 8 != 0;        // bf16 NDHWC with a ragged channel count: via fp32
  if (direct_f32) yf = (float*)y;
  else if (via_f32) MVS_CUDA(tmp.alloc((void**)&yf, (size_t)Do * Ho * Wo * cout * sizeof(float)));
  else MVS_CUDA(tmp.alloc(&yp, planar_bytes(Do, Ho, Wo, cout, 0)));
  rc = launch_conv3d_tc(xp, xs, xb, kp, ss, sb, kernel_tf, D, H, W, cin, cout, stride, transposed, yp, nullptr, yf, stats,
                        scratch, nullptr, nullptr, nullptr, 1, 0, nullptr, nullptr, nullptr, s, nullptr);
  if (rc) return rc;
  if (yp) rc = launch_planar_to_ndhwc(yp, Do, Ho, Wo, cout, y, s);
  else if (via_f32) rc = launch_f32_to_bf16(yf, (size_t)Do * Ho * Wo * cout, y, s);
  return rc;
}

}  // namespace mvsb200
