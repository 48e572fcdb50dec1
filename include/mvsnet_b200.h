/*
 * mvsnet_b200.h -- C ABI of the B200-native MVSNet cost-volume hot path.
 *
 * The reference (ubiquity6/MVSNet) has no FFI of its own: the path is a set of
 * Python functions in mvsnet/homography_warping.py and mvsnet/model.py that build
 * TensorFlow ops.  Each entry point below replaces the TF ops behind one of those
 * functions (cited as file:line under the reference root).  The Python mirror in
 * mvsnet_b200/ re-creates the reference names on top of this ABI via ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - tensors are dense, channels-last, fp32 unless a *_dtype argument says otherwise;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*, NULL =
 *     legacy default stream) except the *_host entry points, which synchronise;
 *   - return 0 on success, a negative MVSB200_ERR_* otherwise; the message of the last
 *     error on the calling thread is mvsb200_last_error();
 *   - no hidden allocation on the hot path: workspaces are caller-owned and sized by the
 *     *_workspace_bytes queries.  There is no CPU fallback anywhere behind this ABI.
 */
#ifndef MVSNET_B200_H_
#define MVSNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVSB200_OK 0
#define MVSB200_ERR_INVALID (-1)     /* bad argument (shape, enum, null pointer) */
#define MVSB200_ERR_CUDA (-2)        /* CUDA runtime error; see mvsb200_last_error() */
#define MVSB200_ERR_UNSUPPORTED (-3) /* valid in the reference but not built here */
#define MVSB200_ERR_WORKSPACE (-4)   /* workspace too small */

/* variance op order: model.py:458-461 (inference_mem) vs model.py:330-332 (inference) */
#define MVSB200_ORDER_MEM 0
#define MVSB200_ORDER_TRAIN 1
/* sampler: tf.contrib.image.transform zero-fill (homography_warping.py:251) vs the
 * legacy clamp-gather `interpolate` (homography_warping.py:131-174) */
#define MVSB200_SAMPLER_TRANSFORM 0
#define MVSB200_SAMPLER_LEGACY 1
/* storage types */
#define MVSB200_F32 0
#define MVSB200_BF16 1
/* regularizer arithmetic: fp32 CUDA-core direct convolution (parity mode) or bf16 operands
 * with fp32 accumulation on the tcgen05 tensor cores (product mode) */
#define MVSB200_PRECISION_FP32 0
#define MVSB200_PRECISION_BF16 1

/* RegNetUS0 layers in execution order (mvsnetworks.py:131-158). */
#define MVSB200_REGNET_LAYERS 11
enum {
  MVSB200_L_3DCONV1_0 = 0, MVSB200_L_3DCONV2_0 = 1, MVSB200_L_3DCONV3_0 = 2,
  MVSB200_L_3DCONV0_1 = 3, MVSB200_L_3DCONV1_1 = 4, MVSB200_L_3DCONV2_1 = 5,
  MVSB200_L_3DCONV3_1 = 6, MVSB200_L_3DCONV4_0 = 7, MVSB200_L_3DCONV5_0 = 8,
  MVSB200_L_3DCONV6_0 = 9, MVSB200_L_3DCONV6_2 = 10
};

/* Weights of RegNetUS0 in TF variable layout, fp32, device memory.
 * kernel[l]: conv `<layer>/kernel` [3,3,3,Cin,Cout]; deconv (layers 7..9) [3,3,3,Cout,Cin].
 * gamma/beta[l]: `<layer>/bn/{gamma,beta}` [Cout]; NULL for 3dconv6_2 (no BN, mvsnetworks.py:158). */
typedef struct mvsb200_regnet_params {
  const float* kernel[MVSB200_REGNET_LAYERS];
  const float* gamma[MVSB200_REGNET_LAYERS];
  const float* beta[MVSB200_REGNET_LAYERS];
} mvsb200_regnet_params;

const char* mvsb200_last_error(void);
int mvsb200_version(void);
/* sm count / compute capability of the current device; fails unless it is sm_100. */
int mvsb200_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* get_homographies (homography_warping.py:10-58) and get_homographies_inv_depth (:60-106)
 * for every source view of one reference view.
 *   cams            [n_views,2,4,4] (view 0 = reference; layout mvs_cluster.py:103-111)
 *   depth_step      depth_interval, or depth_end when inverse_depth != 0
 *   homographies    out [(n_views-1), depth_num, 9]  image-coordinate H (may be NULL)
 *   transforms      out [(n_views-1), depth_num, 8]  pixel-coordinate coefficients of
 *                   tf_transform_homography (homography_warping.py:216-250) (may be NULL) */
int mvsb200_homographies(const float* cams, int n_views, int depth_num, float depth_start,
                         float depth_step, int inverse_depth, float* homographies,
                         float* transforms, void* stream);

/* Coefficient conversion alone (homography_warping.py:216-250): H [count,9] -> T [count,8]. */
int mvsb200_transform_coefs(const float* homographies, int count, float* transforms, void* stream);

/* tf_transform_homography (homography_warping.py:211-253, sampler TRANSFORM) or the legacy
 * homography_warping (:176-210, sampler LEGACY).
 *   image [image_count,H,W,C]; homographies [hom_count,9]; out [hom_count,H,W,C].
 *   image_count must equal hom_count, or be 1 (the one image is warped by every homography). */
int mvsb200_warp(const float* image, int image_count, const float* homographies, int hom_count,
                 int height, int width, int channels, int sampler, float* out, void* stream);

/* Legacy interpolate (homography_warping.py:131-174) on caller-supplied image coordinates:
 * image [B,H,W,C], xs/ys [B*H*W] -> out [B*H*W, C]. */
int mvsb200_interpolate(const float* image, const float* xs, const float* ys, int batch, int height,
                        int width, int channels, float* out, void* stream);

/* get_pixel_grids (homography_warping.py:108-117): out [3*H*W] = concat(x, y, 1) at pixel centres. */
int mvsb200_pixel_grids(int height, int width, float* out, void* stream);

/* Sample coordinates only (parity gate "sample coordinates bit-exact"): for sampler TRANSFORM the
 * pixel coordinates (ix,iy) of the contrib kernel, for LEGACY the image coordinates (x,y) fed to
 * interpolate().  out [hom_count,H,W,2]. */
int mvsb200_sample_coords(const float* homographies, int hom_count, int height, int width,
                          int sampler, float* out, void* stream);

/* Fused warp + N-view variance cost volume (model.py:423-463 / :315-334).  The warped
 * volume is never materialised.
 *   feats [n_views,Hf,Wf,C]; homographies [(n_views-1),depth_num,9];
 *   out   [depth_num,Hf,Wf,C] of out_dtype (MVSB200_F32 / MVSB200_BF16).
 *   variant: 0 = automatic; other values select a specific kernel (tests / tuning). */
int mvsb200_cost_volume(const float* feats, const float* homographies, int n_views, int depth_num,
                        int hf, int wf, int channels, int order, int sampler, int out_dtype,
                        void* out, int variant, void* stream);

/* One layer of the regularizer: y_raw = conv3d(act(x) [+ act(skip)], kernel) with TF SAME padding
 * (network.py:210 conv, :327 transposed), where act(t) = relu(t*scale + shift) per channel when
 * scale/shift are given and identity when NULL; plus per-channel sum / sum-of-squares of the fp32
 * result (network.py:496 batch statistics) accumulated into stats[2*Cout] (double, must be zeroed
 * by the caller).  stats may be NULL.  Output extents: conv ceil(in/stride); deconv 2*in. */
int mvsb200_conv3d_layer(const void* x, int x_dtype, const float* x_scale, const float* x_shift,
                         const void* skip, const float* skip_scale, const float* skip_shift,
                         const float* kernel_tf, int depth, int height, int width, int cin, int cout,
                         int stride, int transposed, int precision, void* y_raw, int y_dtype,
                         double* stats, void* stream);

/* Host-only: the launch plan the bf16 / tcgen05 path uses for one layer shape (tile, folds, ring depth, shared and
 * tensor memory), without touching the device.  numbers[12] = {launches, tile_x, tile_y, cells_x, cells_y, row
 * blocks, mma_n, z_fold, x_fold, ring, smem_bytes, tmem_columns} of the first launch; text (may be NULL) gets one
 * line per launch.  sm_count <= 0 means 148. */
int mvsb200_conv3d_plan(int depth, int height, int width, int cin, int cout, int stride, int transposed,
                        int has_skip, int transform, int sm_count, int* numbers, char* text, int text_len);

/* Batch-norm finalise: stats[2*C] (sum, sumsq over `count` voxels) -> scale = gamma*rsqrt(var+eps),
 * shift = beta - mean*scale (network.py:496-506, Appendix A.6). */
int mvsb200_bn_finalize(const double* stats, const float* gamma, const float* beta, int channels,
                        double count, float eps, float* scale, float* shift, void* stream);

size_t mvsb200_regnet_workspace_bytes(int depth, int hf, int wf, int in_channels, int base_filter,
                                      int precision);

/* RegNetUS0 forward (mvsnetworks.py:122-158): cost [D,Hf,Wf,Cin] -> filtered [D,Hf,Wf] fp32.
 * D, Hf, Wf must be multiples of 8 (the reference graph does not close otherwise). */
int mvsb200_regnet_forward(const void* cost, int cost_dtype, const mvsb200_regnet_params* params,
                           int depth, int hf, int wf, int in_channels, int base_filter, float bn_eps,
                           int precision, float* filtered, void* workspace, size_t workspace_bytes,
                           void* stream);

/* After mvsb200_regnet_forward: address of a layer's raw (pre-BN) output inside the workspace and
 * of its BN scale/shift (tests, layer-wise parity).  Returns NULL for an invalid layer. */
const void* mvsb200_regnet_layer_raw(const void* workspace, int depth, int hf, int wf, int in_channels,
                                     int base_filter, int precision, int layer, const float** scale,
                                     const float** shift);

/* softmax(-F) over depth, soft-argmin and the 4-neighbour probability map
 * (model.py:472-498, 45-144).  filtered [D,Hf,Wf] -> depth_map [Hf,Wf], prob_map [Hf,Wf];
 * prob_volume [D,Hf,Wf] is written when non-NULL.  depth_interval is the plane spacing. */
int mvsb200_depth_regress(const float* filtered, int depth_num, int hf, int wf, float depth_start,
                          float depth_interval, int inverse_depth, int num_buckets, float* depth_map,
                          float* prob_map, float* prob_volume, void* stream);

/* get_probability_map_slice alone (model.py:45-144): prob_volume [D,H,W], depth_map [H,W]. */
int mvsb200_probability_map(const float* prob_volume, const float* depth_map, int depth_num, int height,
                            int width, float depth_start, float depth_interval, int inverse_depth,
                            int num_buckets, float* prob_map, void* stream);

size_t mvsb200_infer_workspace_bytes(int n_views, int depth_num, int hf, int wf, int channels,
                                     int base_filter, int precision);

/* Whole hot path for one reference view, device buffers (model.py:407-502 after the feature
 * towers): feats [n_views,Hf,Wf,C], cams [n_views,2,4,4] -> depth_map, prob_map [Hf,Wf]. */
int mvsb200_infer(const float* feats, const float* cams, int n_views, int depth_num, int hf, int wf,
                  int channels, float depth_start, float depth_interval, int inverse_depth, int order,
                  int sampler, const mvsb200_regnet_params* params, int base_filter, float bn_eps,
                  int precision, float* depth_map, float* prob_map, void* workspace,
                  size_t workspace_bytes, void* stream);

/* Byte offsets, inside the mvsb200_infer workspace, of the cost volume in the regularizer's planar layouts
 * (bf16 mode): chunk-planar [D][C/8][Hf][Wf][8] and parity-split [D][C/8][4][Hf/2][Wf/2][8] bf16. */
int mvsb200_infer_cost_offsets(int n_views, int depth_num, int hf, int wf, int channels, int base_filter,
                               int precision, size_t* cp8_offset, size_t* ps8_offset);
/* Byte offset, inside the mvsb200_infer workspace, of the filtered cost volume [D,Hf,Wf] fp32 (the squeezed output of
 * RegNetUS0, model.py:468-469) the last call left there: lets a test run the reference's regression (model.py:472-498)
 * on exactly the volume the fused soft-argmin saw. */
int mvsb200_infer_filtered_offset(int n_views, int depth_num, int hf, int wf, int channels, int base_filter,
                                  int precision, size_t* offset);

/* Optional instrumentation: five cudaEvent_t handles recorded by mvsb200_infer on its stream at the
 * stage boundaries (start, after homographies, after cost volume, after regularizer, after regression);
 * NULL switches it off.  Per calling thread. */
int mvsb200_infer_set_stage_events(void* const* events);

/* Same with HOST buffers for feats / cams / outputs (the sess.run feed/fetch boundary of
 * inference.py:105-112): copies in, runs, copies out, synchronises.  `params` and `workspace`
 * stay device-resident (weights are loaded once per model, predictlib.py:69-76).  When
 * staging_dev is non-NULL it must hold mvsb200_infer_host_staging_bytes() bytes. */
size_t mvsb200_infer_host_staging_bytes(int n_views, int hf, int wf, int channels);
int mvsb200_infer_host(const float* feats_host, const float* cams_host, int n_views, int depth_num,
                       int hf, int wf, int channels, float depth_start, float depth_interval,
                       int inverse_depth, int order, int sampler, const mvsb200_regnet_params* params,
                       int base_filter, float bn_eps, int precision, float* depth_map_host,
                       float* prob_map_host, void* staging_dev, void* workspace, size_t workspace_bytes,
                       void* stream);

/* The same without the final synchronisation: the copies and kernels are enqueued on `stream` and the
 * host outputs are valid once the caller has synchronised that stream.  Host buffers must be pinned
 * for the copies to overlap.  Two reference views in flight on two streams (each with its own staging
 * and workspace) hide the host->device feed of one behind the kernels of the other, the way a
 * prefetching input pipeline feeds sess.run (inference.py:105-112). */
int mvsb200_infer_host_async(const float* feats_host, const float* cams_host, int n_views, int depth_num,
                             int hf, int wf, int channels, float depth_start, float depth_interval,
                             int inverse_depth, int order, int sampler, const mvsb200_regnet_params* params,
                             int base_filter, float bn_eps, int precision, float* depth_map_host,
                             float* prob_map_host, void* staging_dev, void* workspace, size_t workspace_bytes,
                             void* stream);
/* Same, with the copies on `copy_stream` and the kernels on `compute_stream`, chained by events (feed -> kernels ->
 * fetch).  A caller that alternates two (staging buffer, workspace, copy stream) sets over ONE compute stream keeps the
 * kernels of consecutive reference views back to back while the 40 MB feed of the next view and the fetch of the
 * previous one run beside them (two compute streams would let the next view's first kernels take SMs from the current
 * view's single-wave layers).  The copy stream also orders the re-use of its staging buffer, so use one copy stream per
 * staging buffer.  Host buffers must be pinned; outputs are valid once copy_stream is synchronised. */
int mvsb200_infer_host_pipelined(const float* feats_host, const float* cams_host, int n_views, int depth_num,
                                 int hf, int wf, int channels, float depth_start, float depth_interval,
                                 int inverse_depth, int order, int sampler, const mvsb200_regnet_params* params,
                                 int base_filter, float bn_eps, int precision, float* depth_map_host,
                                 float* prob_map_host, void* staging_dev, void* workspace, size_t workspace_bytes,
                                 void* compute_stream, void* copy_stream);

/* ---- D-slab mode (BASELINE config 5): ONE volume split along depth over `slabs` GPUs ------------
 * Rank `slab` runs the bf16 path on depth_num/slabs consecutive planes (a multiple of 8).  Tensors in the slab
 * workspace carry one halo plane before and after the local planes.  Between the calls below the HOST exchanges
 * data between ranks (mvsnet_b200/dslab.py does it with torch.distributed / NCCL over NVLink):
 *   after mvsb200_slab_layer(l): all-reduce (SUM, fp64) the statistics region of l and swap the boundary planes
 *   of l's output tensors with both neighbours; after the last layer all-gather the filtered slabs and call
 *   mvsb200_depth_regress.  Every rank holds all feature maps, so the cost volume needs no exchange. */
size_t mvsb200_slab_workspace_bytes(int n_views, int depth_num, int slabs, int hf, int wf, int channels,
                                    int base_filter);
/* homographies, the slab's cost-volume planes (and its two halo planes), weight packing, cleared statistics */
int mvsb200_slab_begin(const float* feats, const float* cams, int n_views, int depth_num, int slab, int slabs,
                       int hf, int wf, int channels, float depth_start, float depth_interval,
                       int inverse_depth, int order, const mvsb200_regnet_params* params, int base_filter,
                       void* workspace, size_t workspace_bytes, void* stream);
/* RegNetUS0 layer `layer` (MVSB200_L_*) on the local slab */
int mvsb200_slab_layer(int layer, int n_views, int depth_num, int slab, int slabs, int hf, int wf, int channels,
                       const mvsb200_regnet_params* params, int base_filter, float bn_eps, void* workspace,
                       void* stream);
/* Host-only: byte offsets into the slab workspace of what is exchanged after `layer`.
 *   out[0..1]   statistics: offset, bytes
 *   out[2+5t..6+5t], t = 0 (chunk-planar output) / 1 (parity-split copy): plane bytes (0 = absent), first local
 *               plane (-> previous rank's AFTER halo), last local plane (-> next rank's BEFORE halo), own BEFORE
 *               halo, own AFTER halo
 *   out[12..13] filtered slab [depth_num/slabs, hf, wf] fp32: offset, bytes */
int mvsb200_slab_regions(int layer, int n_views, int depth_num, int slabs, int hf, int wf, int channels,
                         int base_filter, unsigned long long* out);

/* D-slab mode: the softmax over depth split over the ranks.  mvsb200_regress_partial reduces planes [d0, d0+dl) of
 * the filtered volume to per-pixel (max of -F, sum of exp, depth-weighted sum) [3, npix]; after an all-gather of the
 * partials mvsb200_regress_combine gives the depth map (identical on every rank) and this rank's share of the
 * probability map (buckets of model.py:113-140 that fall into its planes; sum the shares over the ranks). */
int mvsb200_regress_partial(const float* filtered, int dl, int d0, int depth_num, int npix, float depth_start,
                            float depth_interval, int inverse_depth, float* partial, void* stream);
int mvsb200_regress_combine(const float* partials, int slabs, const float* filtered, int dl, int d0, int depth_num,
                            int npix, float depth_start, float depth_interval, int inverse_depth, int num_buckets,
                            float* depth_map, float* prob_partial, void* stream);

/* D-slab mode with the exchange fused into the kernels (peer memory over NVLink, no collective between layers).
 * The slab workspaces live in IPC-exportable memory (mvsb200_ipc_*) and every rank maps all of them:
 * peers_dev / peers_host = the same `slabs` base addresses as a device array and a host array (own workspace at
 * index `slab`).  The layer's epilogue also stores its boundary planes into the neighbours' halo planes, a one-block
 * kernel publishes its statistics into every rank's per-source table and raises flag (layer, slab) = seq on every
 * rank, and a consuming kernel first waits (bounded spin) until all ranks have published the layers it reads.
 * seq = 1, 2, 3, ... per inference; the caller puts a cross-rank barrier (e.g. the final all-gather) between
 * inferences.  mvsb200_slab_p2p_error returns 1 (and clears it) if a wait timed out. */
int mvsb200_slab_layer_p2p(int layer, int n_views, int depth_num, int slab, int slabs, int hf, int wf, int channels,
                           const mvsb200_regnet_params* params, int base_filter, float bn_eps, void* workspace,
                           void* const* peers_dev, void* const* peers_host, unsigned seq, void* stream);
int mvsb200_slab_p2p_error(int n_views, int depth_num, int slabs, int hf, int wf, int channels, int base_filter,
                           void* workspace, void* stream);
/* Release every kernel of this rank that is waiting for another rank's publication flag: it gives up at once and raises
 * the error word (mvsb200_slab_p2p_error then returns 1 and clears both).  For a watchdog that has lost a rank; the write
 * uses a stream of its own, the compute stream being the one that is stuck. */
int mvsb200_slab_p2p_abort(int n_views, int depth_num, int slabs, int hf, int wf, int channels, int base_filter,
                           void* workspace);
/* cudaMalloc'ed, zeroed device memory and its CUDA-IPC handle (64 bytes) / mapping in another process */
int mvsb200_ipc_alloc(size_t bytes, void** ptr);
int mvsb200_ipc_free(void* ptr);
int mvsb200_ipc_export(void* ptr, unsigned char* handle64);
int mvsb200_ipc_open(const unsigned char* handle64, void** ptr);
int mvsb200_ipc_close(void* ptr);

/* ---- image feature tower (the step before the hot path) --------------------------------
 * UNetDS2GN (cnn_wrapper/mvsnetworks.py:53-115), run per view with shared weights (model.py:392-406).  fp32, NHWC.
 * Layer order = the order the reference builds them: 2dconv1_0 2_0 3_0 4_0 0_1 0_2 1_1 1_2 2_1 2_2 3_1 3_2 4_1 4_2
 * 5_0 5_1 5_2 6_0 6_1 6_2 7_0 7_1 7_2 8_0 8_1 8_2 conv9_0 9_1 9_2 conv10_0 10_1 10_2.
 * kernel[l]: `<layer>/kernel` [k,k,Cin,Cout] (conv) or [k,k,Cout,Cin] (the four deconvs 2dconv{5,6,7,8}_0);
 * gamma/beta[l]: `<layer>/gn/{gamma,beta}` [Cout]; NULL for conv10_2 (no normalisation, mvsnetworks.py:113-115). */
#define MVSB200_UNET_LAYERS 32
typedef struct mvsb200_unet_params {
  const float* kernel[MVSB200_UNET_LAYERS];
  const float* gamma[MVSB200_UNET_LAYERS];
  const float* beta[MVSB200_UNET_LAYERS];
} mvsb200_unet_params;

/* One conv_gn / deconv_gn building block, split in its two halves (network.py:218-276, :349-409):
 * mvsb200_conv2d_layer: tf.layers.conv2d / conv2d_transpose, SAME, no bias, of the channel concatenation of xa
 *   [N,H,W,ca] and xb [N,H,W,cb] (cb = 0: xa alone): y [N,Ho,Wo,cout] raw; ksize 3 or 5, stride 1 or 2, transposed
 *   only 3x3 stride 2 (Ho = 2H).  stats (may be NULL): [N][cout/8][2] doubles, the kernel ADDS the sum and the sum
 *   of squares of every (view, group of 8 channels) of y (the caller zeroes it).
 * mvsb200_group_norm: y [N,pixels,channels] in place: ((y - mean) / sqrt(var + eps)) * gamma + beta, ReLU if relu. */
int mvsb200_conv2d_layer(const float* xa, int ca, const float* xb, int cb, const float* kernel_tf, int n_views,
                         int height, int width, int cout, int ksize, int stride, int transposed, float* y,
                         double* stats, void* stream);
int mvsb200_group_norm(float* y, const double* stats, const float* gamma, const float* beta, int n_views,
                       int pixels, int channels, float eps, int relu, void* stream);

/* Whole tower: images [N,H,W,3] (centred, mvs_data_generation/utils.py:33-38) -> feats [N,H/4,W/4,4*base_filter].
 * H and W must be multiples of 16 and base_filter a multiple of 8 (network mode "normal").  0 bytes = bad shape. */
size_t mvsb200_unet_workspace_bytes(int n_views, int height, int width, int base_filter);
int mvsb200_unet_forward(const float* images, const mvsb200_unet_params* params, int n_views, int height,
                         int width, int base_filter, float gn_eps, float* feats, void* workspace,
                         size_t workspace_bytes, void* stream);
/* After mvsb200_unet_forward: byte offset inside the workspace and {Ho, Wo, C} of a layer's (normalised) output
 * (layer-wise parity tests).  The last layer is written to `feats`, not to the workspace. */
int mvsb200_unet_layer_output(int n_views, int height, int width, int base_filter, int layer, size_t* offset,
                              int* dims);

/* The same tower in bf16 on the tensor cores (csrc/feature2d_tc.cu: tcgen05 implicit GEMM per layer, raw activations
 * bf16 chunk-planar [N][C/8][H][W][8], group normalisation applied by the consuming layer from fp64 statistics).  Same
 * arguments and result as mvsb200_unet_forward; base_filter must be 8 (network mode "normal"); its own workspace size. */
size_t mvsb200_unet_tc_workspace_bytes(int n_views, int height, int width, int base_filter);
int mvsb200_unet_tc_forward(const float* images, const mvsb200_unet_params* params, int n_views, int height,
                            int width, int base_filter, float gn_eps, float* feats, void* workspace,
                            size_t workspace_bytes, void* stream);
/* After mvsb200_unet_tc_forward: byte offset of a layer's RAW (pre-normalisation) bf16 chunk-planar output in the
 * workspace, its dims {Ho, Wo, C}, and the byte offset of its statistics [N][C/8][2] (sum, sum of squares; fp64).
 * Layers 0 .. 30 (the last layer is written to `feats`). */
int mvsb200_unet_tc_layer_raw(int n_views, int height, int width, int base_filter, int layer, size_t* offset,
                              int* dims, size_t* stats_offset);

/* Host-only view of the tensor-core tower's launch plan of one layer (no device work): out[0..11] = {kind (1: 3x3 stride 1,
 * 2: 3x3 stride 2, 3: 5x5 stride 2, 4: transposed), input chunks of 8 channels, output channels per slice, slices, MMA N,
 * 128-row blocks per tile, MMAs per row block and slice, shared-memory bytes, TMEM columns, tiles per view, weight bytes
 * per slice, operand buffers}. */
int mvsb200_unet_tc_plan(int n_views, int height, int width, int base_filter, int layer, int* out);

/* ---- training step of the path (BASELINE config 4; train.py:314-315 `inference` inside get_loss, loss.py:190-220
 * mvsnet_regression_loss with loss_type 'original', train.py:429 opt.compute_gradients) ------------------------------
 * Gradient buffers in the layout of the variables they belong to (fp32, device memory); gamma/beta[10] unused. */
typedef struct mvsb200_regnet_grads {
  float* kernel[MVSB200_REGNET_LAYERS];
  float* gamma[MVSB200_REGNET_LAYERS];
  float* beta[MVSB200_REGNET_LAYERS];
} mvsb200_regnet_grads;
size_t mvsb200_train_workspace_bytes(int n_views, int depth_num, int hf, int wf, int channels, int base_filter);
/* Forward in the fp32 parity mode (variance order `order`: MVSB200_ORDER_TRAIN is model.py:330-332), loss against
 * gt_depth [Hf,Wf] (0 = invalid pixel, loss.py:21), backward.
 *   grads     d loss / d every RegNetUS0 variable
 *   dfeats    d loss / d feats [n_views,Hf,Wf,32]; the warp's gradient is the exact adjoint (scatter) of the bilinear
 *             zero-fill gather -- TF 1.12 registers a resampling with the inverse transform instead
 *   depth_map out [Hf,Wf]; metrics out [3] (device): loss, less_one_accuracy, less_three_accuracy (loss.py:161-187) */
int mvsb200_train_step(const float* feats, const float* cams, const float* gt_depth, int n_views, int depth_num, int hf,
                       int wf, int channels, float depth_start, float depth_interval, int order,
                       const mvsb200_regnet_params* params, int base_filter, float bn_eps,
                       const mvsb200_regnet_grads* grads, float* dfeats, float* depth_map, float* metrics, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---- refinement glue after the path (model.py:753-811 `depth_refine`) ------------------------------
 * tf.image.resize_bilinear of TF 1.x (align_corners = False): x [n,height,width,channels] fp32 -> y [n,out_height,
 * out_width,channels], then y = (y - subtract) * multiply (the depth normalisation of model.py:763-765 in the same pass;
 * pass 0 and 1 for a plain resize; equal sizes make it a pure affine). */
int mvsb200_resize_bilinear(const float* x, int n, int height, int width, int channels, float* y, int out_height,
                            int out_width, float subtract, float multiply, void* stream);
/* scaled = x * multiply; sum = scaled + add (either output may be NULL; add NULL: sum = scaled): the residual back in
 * millimetres and the refined depth map, model.py:803-809. */
int mvsb200_scale_add(const float* x, float multiply, const float* add, size_t count, float* scaled, float* sum,
                      void* stream);
/* tf.layers.conv2d(3x3, SAME, stride 1, use_bias=True) of concat(xa, xb) with optional ReLU: the layers of RefineNetConv
 * (mvsnetworks.py:178-193, network.py:171-206).  kernel_tf [3,3,ca+cb,cout], bias [cout] or NULL. */
int mvsb200_conv2d_bias(const float* xa, int ca, const float* xb, int cb, const float* kernel_tf, const float* bias, int n,
                        int height, int width, int cout, int relu, float* y, void* stream);

/* Diagnostic (not on the product path): one 128 x n x (16*kblocks) tcgen05.mma tile computed from
 * caller-built shared-memory images of the A and B operands (no-swizzle K-major core-matrix
 * layout).  Pins the descriptor semantics conv3d_tc.cu relies on.  d_out [128*n] fp32. */
int mvsb200_umma_probe(const void* a_image, int a_bytes, const void* b_image, int b_bytes, int n,
                       int kblocks, int a_kblock_stride, int a_start, int a_lbo, int a_sbo,
                       int b_kblock_stride, int b_lbo, int b_sbo, float* d_out, void* stream);

/* Number of kernel launches issued through this library by the calling process so far. */
uint64_t mvsb200_launch_count(void);

/* Development counter of the fused warp + variance kernel (cost_volume_win.cu): (voxel, view) pairs whose taps came
 * from global memory because their footprint was not inside the shared-memory window, since the last reset.  Counted
 * only while the tuning switch CV_STATS is on.  Synchronises the device.  No counterpart in the reference. */
int mvsb200_cost_volume_window_stats(unsigned long long* slow_pairs, int reset);

/* Development / tuning switch by name, e.g. ("NO_FUSED_REGRESS", "1"), ("CV_KERNEL", "1"), ("TC_TILE", "14x8").
 * The MVSB200_<NAME> environment variables are only the initial values, read once at first use; nothing on the hot
 * path reads the environment.  value NULL restores the default.  No counterpart in the reference. */
int mvsb200_set_tuning(const char* name, const char* value);

#ifdef __cplusplus
}
#endif
#endif /* MVSNET_B200_H_ */
