// Internal interface of the bf16 / tcgen05 regularizer layers (conv3d_tc.cu) used by the network code (regnet.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace mvsb200 {

// Batch-norm source of an input tensor: the producer's channel statistics (sum | sum of squares over `count`
// voxels) and its gamma / beta.  The consumer derives scale / shift itself (no bn_finalize launch in between).
// The statistics are the sum of `reps` partial copies, rep_stride doubles apart.
struct TcBnSrc {
  const double* stats; const float* gamma; const float* beta; double count; float eps; int channels;
  int reps; int rep_stride;
  int channels_true;      // > 0: gamma / beta hold this many entries; channels beyond are padding (scale = shift = 0)
};

// D-slab mode (one volume split along z over several GPUs): the input (and skip) tensor carries one halo plane
// before and after the D local planes; planes outside [zv_lo, zv_hi) (extended coordinates) are SAME padding.
struct TcSlab { int halo; int zv_lo; int zv_hi; };

// D-slab mode over peer memory (NVLink): where the boundary planes of the output are mirrored (the neighbours' halo
// planes, [0] chunk-planar / [1] parity-split, NULL = none) and which publication flags the kernel waits for before
// it reads its inputs (per consumed layer `wait_n` words, one per rank, each >= wait_seq when published).
struct TcPeer {
  void* mir_prev[2]; void* mir_next[2];
  const unsigned* wait_flags[2]; int wait_n; unsigned wait_seq; unsigned* err_flag;
};

// Soft-argmin of the regression folded into the last layer's epilogue (linear depth sampling): partial [3][Ho*Wo]
// = per-pixel (max of -F, sum of exp, depth-weighted sum) with depth_i = start + step * i (tf.linspace); the launch
// sets `fused` when the plan allowed it (x-fold, one output channel, CTAs covering the whole depth range).
struct TcRegress { float* partial; float start, step; int fused; };

// One job per layer of a network for conv3d_tc_pack_all; its launches (output-channel slices of 32) take
// consecutive weight slots starting at slot0.
// cin_true / cout_true > 0: channel counts of kernel_tf itself when the layer runs with channels padded to whole cells.
// kernel_tf2 / cout2 (> 0): the stride-2 conv riding on this stride-1 conv's launch (TcRider).
struct TcPackJob { const float* kernel_tf; int D, H, W, cin, cout, stride, transposed, has_skip, transform, slot0, cin_true, cout_true;
                   const float* kernel_tf2; int cout2, cout2_true; };

// Rider of a stride-1 conv on a raw input (3dconv0_1 + 3dconv1_0: one pass over the cost volume): the stride-2 conv
// with filter kernel_tf2 [3,3,3,cin,cout2] on the same input is the stride-1 conv of that filter at the odd (z, y, x)
// positions (TF SAME padding, even extents); its raw output goes to y2_cp8 / y2_ps8 ([D/2,H/2,W/2,cout2] in the planar
// layouts), its statistics to stats2.
struct TcRider { const float* kernel_tf2; int cout2; void* y2_cp8; void* y2_ps8; double* stats2; };

// x: CP8 for stride-1 convs and transposed convs, PS8 for stride-2 convs.  skip: CP8.
// Outputs: y_cp8 and / or y_ps8 (bf16, Cout % 8 == 0), or y_f32 (NDHWC fp32, any Cout).
int launch_conv3d_tc(const void* x, const float* xs, const float* xb, const void* skip, const float* ss,
                     const float* sb, const float* kernel_tf, int D, int H, int W, int cin, int cout, int stride,
                     int transposed, void* y_cp8, void* y_ps8, float* y_f32, double* stats, void* scratch,
                     const TcBnSrc* x_bn, const TcBnSrc* s_bn, const void* prepacked, int stats_reps,
                     int stats_rep_stride, const TcSlab* slab, const TcPeer* peer, TcRegress* regress,
                     cudaStream_t s, const TcRider* rider = nullptr);
int launch_conv3d_tc_ndhwc(const void* x, const float* xs, const float* xb, const void* skip, const float* ss,
                           const float* sb, const float* kernel_tf, int D, int H, int W, int cin, int cout,
                           int stride, int transposed, void* y, int y_dtype, double* stats, cudaStream_t s);
int conv3d_tc_pack_all(const TcPackJob* jobs, int njobs, void* dst_base, cudaStream_t s);
int conv3d_tc_describe(int D, int H, int W, int cin, int cout, int stride, int transposed, int has_skip, int transform,
                       int sm_count, int* out, char* text, int text_len);
size_t conv3d_tc_pack_slot_bytes();
size_t conv3d_tc_scratch_bytes();
size_t planar_bytes(int D, int H, int W, int C, int parity_split);
int launch_ndhwc_to_planar(const void* x_ndhwc, int D, int H, int W, int C, void* cp8, void* ps8, cudaStream_t s);
int launch_planar_to_ndhwc(const void* cp8, int D, int H, int W, int C, void* y_ndhwc, cudaStream_t s);

}  // namespace mvsb200
