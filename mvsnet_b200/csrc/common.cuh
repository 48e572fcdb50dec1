// Shared host/device helpers for the mvsnet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/mvsnet_b200.h"

namespace mvsb200 {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launch_count;

inline void count_launch(int n = 1) { g_launch_count.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// Development / tuning switches.  One immutable snapshot, read with a single atomic load per use: it is filled from
// the MVSB200_* environment variables ONCE (first use) and replaced as a whole by mvsb200_set_tuning(); nothing on
// the hot path calls getenv().  -1 = not set for the fields that distinguish "unset" from 0.
struct Tuning {
  int no_fused_regress = 0;   // NO_FUSED_REGRESS: stand-alone regression kernel instead of the 3dconv6_2 epilogue
  int cv_fp32_taps = 0;       // CV_FP32_TAPS: product-mode cost volume from fp32 features (north_star arithmetic)
  int cv_kernel = 0;          // CV_KERNEL: 0 = shared-memory window kernel (TMA), 1 = the round-1 gather kernel
  int cv_fp32_blend = 0, cv_minb = 0, cv_rec16 = 0, cv_kdc = 0;      // round-1 gather kernel variants
  int cv_planes = 0, cv_stats = 0, cv_dbg = 0;                                   // window kernel: planes per block; fit counters
  int tc_zf = -1, tc_xfold = -1, tc_zsplit = 0, tc_dbg = 0, tc_verbose = 0, tc_prof = 0, tc_exact_smem = 0, tc_no_pdl = 0;
  int tc_tile_x = 0, tc_tile_y = 0;
  int tc_xf_groups = 0;       // TC_XF_GROUPS: 2 = two groups of transform warps on alternate planes (default: all warps on one plane)
  int tc_rank = 0;            // TC_RANK: take the rank-th best plan of the cycle model (development: is the model's order right?)
  int tc_trim = -1;           // TC_TRIM: 0 = every MMA of a launch spans all N columns (default: ops skip their all-zero column blocks)
  int infer_side = -1;        // INFER_SIDE: 0 = statistics clear + weight packing on the compute stream (default: side stream beside K2)
  int tc_fuse01 = -1;         // TC_FUSE01: 0 = 3dconv0_1 and 3dconv1_0 as two launches (and a parity-split cost volume)
  int tc_layer_set = 0, tc_layer[3] = {0, 0, 0};                     // TC_LAYER="cin,cout,mode": restrict the tc_* switches
  int regnet_profile = 0, unet_no_tile = 0, unet_profile = 0, unet_fp32 = 0;
  int unet_dbg = 0;
  int unet_obuf = 0;          // UNET_OBUF: 2 = a second operand buffer where it is cheap (transform under the MMAs)
  int unet_inplace = -1;      // UNET_INPLACE: 0 = never transform in place
  int unet_hot = -1;          // UNET_HOT: 0 = no compile-time variants of the tower kernel
  int unet_mb = 0;            // UNET_MB: 1 / 2 = 128-row blocks per tile of the tensor-core tower (0: by tile count)
};
const Tuning& tuning();

// multiprocessor count of the current device (cached per device)
int sm_count_current();

#define MVS_CHECK_ARG(cond, ...)                         \
  do {                                                   \
    if (!(cond)) {                                       \
      ::mvsb200::set_error(__VA_ARGS__);                 \
      return MVSB200_ERR_INVALID;                        \
    }                                                    \
  } while (0)

#define MVS_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      ::mvsb200::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,              \
                           cudaGetErrorString(e__));                                         \
      return MVSB200_ERR_CUDA;                                                               \
    }                                                                                        \
  } while (0)

#define MVS_LAUNCH_CHECK(name)                                                               \
  do {                                                                                       \
    ::mvsb200::count_launch();                                                               \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess) {                                                                \
      ::mvsb200::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));        \
      return MVSB200_ERR_CUDA;                                                               \
    }                                                                                        \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// TF SAME padding before the first element: out=ceil(in/s); total=max((out-1)s+k-in,0); before=total/2.
static inline int tf_same_pad_before(int size, int k, int s) {
  int out = (size + s - 1) / s;
  int total = (out - 1) * s + k - size;
  if (total < 0) total = 0;
  return total / 2;
}

// internal launchers shared between translation units -------------------------------------------
int launch_homographies(const float* cams, int n_views, int depth_num, float depth_start, float depth_step,
                        int inverse_depth, float* homographies, float* transforms, cudaStream_t s);
int launch_cost_volume(const float* feats, const float* homographies, int n_views, int depth_num, int hf,
                       int wf, int channels, int order, int sampler, int out_dtype, void* out, int variant,
                       cudaStream_t s);
int launch_conv3d_layer(const void* x, int x_dtype, const float* x_scale, const float* x_shift,
                        const void* skip, const float* skip_scale, const float* skip_shift,
                        const float* kernel_tf, int depth, int height, int width, int cin, int cout,
                        int stride, int transposed, int precision, void* y_raw, int y_dtype, double* stats,
                        cudaStream_t s);
int launch_bn_finalize(const double* stats, const float* gamma, const float* beta, int channels,
                       double count, float eps, float* scale, float* shift, cudaStream_t s);
int launch_depth_regress(const float* filtered, int depth_num, int hf, int wf, float depth_start,
                         float depth_interval, int inverse_depth, int num_buckets, float* depth_map,
                         float* prob_map, float* prob_volume, cudaStream_t s);

}  // namespace mvsb200
