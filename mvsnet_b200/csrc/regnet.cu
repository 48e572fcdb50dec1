// RegNetUS0 forward (mvsnetworks.py:122-158) and the whole-path entry points.
//
// Data flow: every conv / deconv layer writes its RAW (pre-BN) output plus per-channel
// sum / sum-of-squares; the consumer turns those into (scale, shift) (fp32 mode: a bn_finalize launch;
// bf16 mode: inside the consumer's prologue) and applies relu(raw*scale+shift) (and the skip add) while
// reading its input.  No normalised tensor is ever written back to HBM.  BN uses batch statistics, as the
// reference does at inference (network.py:54,64; model.py:337-338; SURVEY.md "facts").
#include "common.cuh"
#include "conv3d_tc.h"
#include "regnet_plan.h"
#include <stdlib.h>
#include <string.h>

namespace mvsb200 {

int launch_conv3d_direct(const void* x, int x_dtype, const float* xs, const float* xb, const void* skip,
                         const float* ss, const float* sb, const float* kernel_tf, int D, int H, int W, int cin,
                         int cout, int stride, int transposed, void* y, int y_dtype, double* stats, cudaStream_t s,
                         int accumulate = 0);
int launch_bn_finalize_all(const double* stats, const float* const* gamma, const float* const* beta, const int* channels,
                           const int* channels_true, const double* counts, int layers, int cpad, int reps, float eps,
                           float* scale, float* shift, cudaStream_t s);

int launch_conv3d_layer(const void* x, int x_dtype, const float* x_scale, const float* x_shift, const void* skip,
                        const float* skip_scale, const float* skip_shift, const float* kernel_tf, int depth,
                        int height, int width, int cin, int cout, int stride, int transposed, int precision,
                        void* y_raw, int y_dtype, double* stats, cudaStream_t s) {
  MVS_CHECK_ARG(x && kernel_tf && y_raw, "conv3d_layer: NULL pointer");
  MVS_CHECK_ARG(depth > 0 && height > 0 && width > 0 && cin > 0 && cout > 0, "conv3d_layer: bad shape");
  MVS_CHECK_ARG(x_dtype == MVSB200_F32 || x_dtype == MVSB200_BF16, "conv3d_layer: bad x_dtype %d", x_dtype);
  MVS_CHECK_ARG(y_dtype == MVSB200_F32 || y_dtype == MVSB200_BF16, "conv3d_layer: bad y_dtype %d", y_dtype);
  MVS_CHECK_ARG((x_scale == nullptr) == (x_shift == nullptr), "conv3d_layer: x_scale/x_shift must come together");
  MVS_CHECK_ARG(!skip || ((skip_scale == nullptr) == (skip_shift == nullptr)),
                "conv3d_layer: skip_scale/skip_shift must come together");
  if (transposed) MVS_CHECK_ARG(stride == 2, "conv3d_layer: transposed conv supports stride 2 only (got %d)", stride);
  else MVS_CHECK_ARG(stride == 1 || stride == 2, "conv3d_layer: stride must be 1 or 2 (got %d)", stride);
  if (precision == MVSB200_PRECISION_FP32)
    return launch_conv3d_direct(x, x_dtype, x_scale, x_shift, skip, skip_scale, skip_shift, kernel_tf, depth, height,
                                width, cin, cout, stride, transposed, y_raw, y_dtype, stats, s);
  if (precision == MVSB200_PRECISION_BF16) {
    if (x_dtype != MVSB200_BF16) {
      set_error("conv3d(bf16/tcgen05): input must be bf16");
      return MVSB200_ERR_UNSUPPORTED;
    }
    return launch_conv3d_tc_ndhwc(x, x_scale, x_shift, skip, skip_scale, skip_shift, kernel_tf, depth, height, width,
                                  cin, cout, stride, transposed, y_raw, y_dtype, stats, s);
  }
  set_error("conv3d_layer: bad precision %d", precision);
  return MVSB200_ERR_INVALID;
}

// bf16 mode: 3dconv0_1 and 3dconv1_0 both read the cost volume; when the shapes allow it they run as ONE launch (the
// stride-2 layer rides on the stride-1 layer's MMAs, conv3d_tc.h TcRider): the volume is read once, and its parity-split
// copy is never needed (the cost-volume kernel then writes the chunk-planar copy only).  Tuning TC_FUSE01=0 turns it off.
bool regnet_fuse01(int D, int H, int W, int cin, int b) {
  if (tuning().tc_fuse01 == 0) return false;
  RegnetPlan p;
  make_plan(D, H, W, cin, b, MVSB200_PRECISION_BF16, &p);
  const LayerDesc& a = p.layer[MVSB200_L_3DCONV0_1];
  const LayerDesc& r = p.layer[MVSB200_L_3DCONV1_0];
  return a.cin == r.cin && a.cin % 16 == 0 && a.cin <= 64 && a.cout == 8 && (r.cout == 8 || r.cout == 16) && !((D | H | W) & 1);
}

// What a forward pass needs that does not depend on the cost volume: cleared statistics and (bf16 mode) every layer's
// weights as bf16 B images, one launch (slot 2*i, 2*i+1 = the <= 2 output-channel slices of layer i).  mvsb200_infer
// runs it on a side stream beside the cost-volume kernel.
static int regnet_prepare(const RegnetPlan& p, const mvsb200_regnet_params* params, int D, int H, int W, int cin, int b,
                          bool bf16, char* ws, cudaStream_t s) {
  for (int i = 0; i < MVSB200_REGNET_LAYERS; ++i) {
    MVS_CHECK_ARG(params->kernel[i] != nullptr, "regnet_forward: kernel[%d] is NULL", i);
    if (i != MVSB200_L_3DCONV6_2)
      MVS_CHECK_ARG(params->gamma[i] && params->beta[i], "regnet_forward: gamma/beta[%d] is NULL", i);
  }
  MVS_CUDA(cudaMemsetAsync(ws + p.stats_off, 0, p.stats_bytes, s));
  if (!bf16) return MVSB200_OK;
  TcPackJob jobs[MVSB200_REGNET_LAYERS];
  for (int i = 0; i < MVSB200_REGNET_LAYERS; ++i) {
    const LayerDesc& L = p.layer[i];
    const int* d = p.dims[L.in_level];
    jobs[i] = {params->kernel[i], d[0], d[1], d[2], L.cin, L.cout, L.stride, L.transposed, L.skip >= 0 ? 1 : 0,
               (L.src >= 0 || L.skip >= 0) ? 1 : 0, 2 * i, L.cin_true, L.cout_true};
  }
  if (regnet_fuse01(D, H, W, cin, b)) {
    // slot of 3dconv1_0 = the rider launch: 3dconv0_1's filter with 3dconv1_0's riding on it
    const LayerDesc& A = p.layer[MVSB200_L_3DCONV0_1];
    const LayerDesc& R = p.layer[MVSB200_L_3DCONV1_0];
    jobs[MVSB200_L_3DCONV1_0] = {params->kernel[MVSB200_L_3DCONV0_1], D, H, W, A.cin, A.cout, 1, 0, 0, 0,
                                 2 * MVSB200_L_3DCONV1_0, A.cin_true, A.cout_true, params->kernel[MVSB200_L_3DCONV1_0],
                                 R.cout, R.cout_true};
  }
  return conv3d_tc_pack_all(jobs, MVSB200_REGNET_LAYERS, ws + p.scratch_off, s);
}

// cost_planar != 0 (bf16 mode only): the caller has already written the cost volume in the planar layouts
// (chunk-planar always; parity-split unless regnet_fuse01()) at regnet_cost_planar() inside the workspace and `cost`
// is ignored.
int regnet_forward_impl(const void* cost, int cost_dtype, int cost_planar, const mvsb200_regnet_params* params, int D,
                        int H, int W, int cin, int b, float eps, int precision, float* filtered, void* workspace,
                        size_t workspace_bytes, cudaStream_t s, TcRegress* regress = nullptr, bool inspect = true,
                        bool prepared = false) {
  // prepared: regnet_prepare() has already run for this pass (ordered before `s` reaches this call)
  // inspect: also materialise every layer's BN scale / shift for mvsb200_regnet_layer_raw (the whole-path entry
  // points do not need them: one launch less on the critical path)
  MVS_CHECK_ARG((cost || cost_planar) && params && filtered && workspace, "regnet_forward: NULL pointer");
  int rc = check_regnet_shape(D, H, W, cin, b);
  if (rc) return rc;
  MVS_CHECK_ARG(precision == MVSB200_PRECISION_FP32 || precision == MVSB200_PRECISION_BF16,
                "regnet_forward: bad precision %d", precision);
  const bool bf16 = precision == MVSB200_PRECISION_BF16;
  if (bf16) MVS_CHECK_ARG(cin % 8 == 0, "regnet_forward(bf16): in_channels must be a multiple of 8 (got %d)", cin);
  RegnetPlan p;
  make_plan(D, H, W, cin, b, precision, &p);
  if (workspace_bytes < p.total) {
    set_error("regnet_forward: workspace %zu < required %zu bytes", workspace_bytes, p.total);
    return MVSB200_ERR_WORKSPACE;
  }
  char* ws = (char*)workspace;
  const int cpad = plan_cpad(p);
  double* stats = (double*)(ws + p.stats_off);
  float* scale = (float*)(ws + p.scale_off);
  float* shift = (float*)(ws + p.shift_off);
  const int act_dtype = bf16 ? MVSB200_BF16 : MVSB200_F32;
  if (!prepared) {
    rc = regnet_prepare(p, params, D, H, W, cin, b, bf16, ws, s);
    if (rc) return rc;
  }
  const bool fuse01 = bf16 && regnet_fuse01(D, H, W, cin, b);
  if (bf16 && !cost_planar) {
    MVS_CHECK_ARG(cost_dtype == MVSB200_BF16, "regnet_forward: precision bf16 needs a bf16 cost volume");
    rc = launch_ndhwc_to_planar(cost, D, H, W, cin, ws + p.cost_cp8_off, fuse01 ? nullptr : ws + p.cost_ps8_off, s);
    if (rc) return rc;
  }
  // development aid: MVSB200_REGNET_PROFILE=1 prints per-layer device times (synchronises; not for timed runs)
  const bool profile = tuning().regnet_profile != 0;
  cudaEvent_t pev[MVSB200_REGNET_LAYERS + 1];
  if (profile) {
    for (int i = 0; i <= MVSB200_REGNET_LAYERS; ++i) cudaEventCreate(&pev[i]);
    cudaEventRecord(pev[0], s);
  }
  for (int i = 0; i < MVSB200_REGNET_LAYERS; ++i) {
    const LayerDesc& L = p.layer[i];
    const bool last = i == MVSB200_L_3DCONV6_2;
    if (fuse01 && (i == MVSB200_L_3DCONV1_0 || i == MVSB200_L_3DCONV0_1)) {
      if (i == MVSB200_L_3DCONV1_0) {
        const LayerDesc& A = p.layer[MVSB200_L_3DCONV0_1];
        const int* o = p.dims[L.out_level];
        if ((o[1] | o[2]) & 1) MVS_CUDA(cudaMemsetAsync(ws + p.ps8_off[i], 0, planar_bytes(o[0], o[1], o[2], L.cout, 1), s));
        const int rep_stride = MVSB200_REGNET_LAYERS * 2 * cpad;
        const TcRider rider = {params->kernel[i], L.cout, ws + p.raw_off[i], p.has_ps8[i] ? ws + p.ps8_off[i] : nullptr,
                               stats + (size_t)i * 2 * cpad};
        rc = launch_conv3d_tc(ws + p.cost_cp8_off, nullptr, nullptr, nullptr, nullptr, nullptr, params->kernel[MVSB200_L_3DCONV0_1],
                              D, H, W, A.cin, A.cout, 1, 0, ws + p.raw_off[MVSB200_L_3DCONV0_1], nullptr, nullptr,
                              stats + (size_t)MVSB200_L_3DCONV0_1 * 2 * cpad, nullptr, nullptr, nullptr,
                              ws + p.scratch_off + (size_t)2 * i * conv3d_tc_pack_slot_bytes(), kStatsReps, rep_stride, nullptr,
                              nullptr, nullptr, s, &rider);
        if (rc) return rc;
      }
      if (profile) cudaEventRecord(pev[i + 1], s);
      continue;
    }
    const float* xs = L.src < 0 ? nullptr : scale + (size_t)L.src * cpad;
    const float* xb = L.src < 0 ? nullptr : shift + (size_t)L.src * cpad;
    const void* sk = L.skip < 0 ? nullptr : (const void*)(ws + p.raw_off[L.skip]);
    const float* ss = L.skip < 0 ? nullptr : scale + (size_t)L.skip * cpad;
    const float* sb = L.skip < 0 ? nullptr : shift + (size_t)L.skip * cpad;
    double* st = last ? nullptr : stats + (size_t)i * 2 * cpad;
    const int* d = p.dims[L.in_level];
    if (!bf16) {
      const void* x = L.src < 0 ? cost : (const void*)(ws + p.raw_off[L.src]);
      const int x_dtype = L.src < 0 ? cost_dtype : act_dtype;
      void* y = last ? (void*)filtered : (void*)(ws + p.raw_off[i]);
      rc = launch_conv3d_direct(x, x_dtype, xs, xb, sk, ss, sb, params->kernel[i], d[0], d[1], d[2], L.cin, L.cout,
                                L.stride, L.transposed, y, MVSB200_F32, st, s);
      if (rc) return rc;
      if (!last) {
        // stats hold sum over L.cout channels laid out [sum(cout) | sumsq(cout)]
        rc = launch_bn_finalize(st, params->gamma[i], params->beta[i], L.cout, (double)p.vox[L.out_level], eps,
                                scale + (size_t)i * cpad, shift + (size_t)i * cpad, s);
        if (rc) return rc;
      }
    } else {
      // stride-2 convs read the parity-split copy of their input, everything else the chunk-planar one; the BN of
      // the producer(s) is derived from their statistics inside the consumer (no launch in between)
      const bool s2 = L.stride == 2 && !L.transposed;
      const void* x = L.src < 0 ? (const void*)(ws + (s2 ? p.cost_ps8_off : p.cost_cp8_off))
                                : (const void*)(ws + (s2 ? p.ps8_off[L.src] : p.raw_off[L.src]));
      if (!last && p.has_ps8[i] && ((p.dims[L.out_level][1] | p.dims[L.out_level][2]) & 1)) {
        const int* o = p.dims[L.out_level];
        MVS_CUDA(cudaMemsetAsync(ws + p.ps8_off[i], 0, planar_bytes(o[0], o[1], o[2], L.cout, 1), s));
      }
      // statistics: kStatsReps partial copies [rep][layer][2*cpad]; CTAs pick a copy, consumers add them up
      const int rep_stride = MVSB200_REGNET_LAYERS * 2 * cpad;
      TcBnSrc xbn = {nullptr, nullptr, nullptr, 1.0, eps, 0, 1, 0}, sbn = xbn;
      if (L.src >= 0)
        xbn = {stats + (size_t)L.src * 2 * cpad, params->gamma[L.src], params->beta[L.src],
               (double)p.vox[p.layer[L.src].out_level], eps, p.layer[L.src].cout, kStatsReps, rep_stride,
               p.layer[L.src].cout_true};
      if (L.skip >= 0)
        sbn = {stats + (size_t)L.skip * 2 * cpad, params->gamma[L.skip], params->beta[L.skip],
               (double)p.vox[p.layer[L.skip].out_level], eps, p.layer[L.skip].cout, kStatsReps, rep_stride,
               p.layer[L.skip].cout_true};
      rc = launch_conv3d_tc(x, nullptr, nullptr, sk, nullptr, nullptr, params->kernel[i], d[0], d[1], d[2], L.cin, L.cout,
                            L.stride, L.transposed, last ? nullptr : ws + p.raw_off[i],
                            (!last && p.has_ps8[i]) ? ws + p.ps8_off[i] : nullptr, last ? filtered : nullptr, st,
                            nullptr, L.src >= 0 ? &xbn : nullptr, L.skip >= 0 ? &sbn : nullptr,
                            ws + p.scratch_off + (size_t)2 * i * conv3d_tc_pack_slot_bytes(), kStatsReps, rep_stride, nullptr,
                            nullptr, last ? regress : nullptr, s);
      if (rc) return rc;
    }
    if (profile) cudaEventRecord(pev[i + 1], s);
  }
  if (bf16 && inspect) {
    // scale / shift of every layer for mvsb200_regnet_layer_raw (inspection only; the layers above do not read them)
    const float* gam[MVSB200_REGNET_LAYERS];
    const float* bet[MVSB200_REGNET_LAYERS];
    int chans[MVSB200_REGNET_LAYERS], chans_true[MVSB200_REGNET_LAYERS];
    double counts[MVSB200_REGNET_LAYERS];
    for (int i = 0; i < MVSB200_REGNET_LAYERS; ++i) {
      gam[i] = params->gamma[i]; bet[i] = params->beta[i];
      chans[i] = i == MVSB200_L_3DCONV6_2 ? 0 : p.layer[i].cout;
      chans_true[i] = i == MVSB200_L_3DCONV6_2 ? 0 : p.layer[i].cout_true;
      counts[i] = (double)p.vox[p.layer[i].out_level];
    }
    rc = launch_bn_finalize_all(stats, gam, bet, chans, chans_true, counts, MVSB200_REGNET_LAYERS, cpad, kStatsReps, eps, scale,
                                shift, s);
    if (rc) return rc;
  }
  if (profile) {
    static const char* names[MVSB200_REGNET_LAYERS] = {"3dconv1_0", "3dconv2_0", "3dconv3_0", "3dconv0_1", "3dconv1_1",
                                                       "3dconv2_1", "3dconv3_1", "3dconv4_0", "3dconv5_0", "3dconv6_0",
                                                       "3dconv6_2"};
    cudaStreamSynchronize(s);
    float total = 0.f;
    for (int i = 0; i < MVSB200_REGNET_LAYERS; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, pev[i], pev[i + 1]);
      total += ms;
      fprintf(stderr, "[regnet] %s %.3f ms\n", names[i], ms);
    }
    fprintf(stderr, "[regnet] total %.3f ms\n", total);
    for (int i = 0; i <= MVSB200_REGNET_LAYERS; ++i) cudaEventDestroy(pev[i]);
  }
  return MVSB200_OK;
}

// bf16 mode: where the fused path writes the cost volume (chunk-planar and parity-split copies)
void regnet_cost_planar(void* workspace, int D, int H, int W, int cin, int b, void** cp8, void** ps8) {
  RegnetPlan p;
  make_plan(D, H, W, cin, b, MVSB200_PRECISION_BF16, &p);
  *cp8 = (char*)workspace + p.cost_cp8_off;
  *ps8 = (char*)workspace + p.cost_ps8_off;
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_conv3d_layer(const void* x, int x_dtype, const float* x_scale, const float* x_shift,
                                    const void* skip, const float* skip_scale, const float* skip_shift,
                                    const float* kernel_tf, int depth, int height, int width, int cin, int cout,
                                    int stride, int transposed, int precision, void* y_raw, int y_dtype,
                                    double* stats, void* stream) {
  return launch_conv3d_layer(x, x_dtype, x_scale, x_shift, skip, skip_scale, skip_shift, kernel_tf, depth, height,
                             width, cin, cout, stride, transposed, precision, y_raw, y_dtype, stats,
                             (cudaStream_t)stream);
}

extern "C" int mvsb200_conv3d_plan(int depth, int height, int width, int cin, int cout, int stride, int transposed,
                                   int has_skip, int transform, int sm_count, int* numbers, char* text, int text_len) {
  return conv3d_tc_describe(depth, height, width, cin, cout, stride, transposed, has_skip, transform, sm_count, numbers,
                            text, text_len);
}

extern "C" int mvsb200_bn_finalize(const double* stats, const float* gamma, const float* beta, int channels,
                                   double count, float eps, float* scale, float* shift, void* stream) {
  return launch_bn_finalize(stats, gamma, beta, channels, count, eps, scale, shift, (cudaStream_t)stream);
}

extern "C" size_t mvsb200_regnet_workspace_bytes(int depth, int hf, int wf, int in_channels, int base_filter,
                                                 int precision) {
  if (depth <= 0 || hf <= 0 || wf <= 0 || in_channels <= 0 || base_filter <= 0) return 0;
  RegnetPlan p;
  make_plan(depth, hf, wf, in_channels, base_filter, precision, &p);
  return p.total;
}

extern "C" int mvsb200_regnet_forward(const void* cost, int cost_dtype, const mvsb200_regnet_params* params,
                                      int depth, int hf, int wf, int in_channels, int base_filter, float bn_eps,
                                      int precision, float* filtered, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  return regnet_forward_impl(cost, cost_dtype, 0, params, depth, hf, wf, in_channels, base_filter, bn_eps, precision,
                             filtered, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" const void* mvsb200_regnet_layer_raw(const void* workspace, int depth, int hf, int wf, int in_channels,
                                                int base_filter, int precision, int layer, const float** scale,
                                                const float** shift) {
  if (!workspace || layer < 0 || layer >= MVSB200_REGNET_LAYERS) return nullptr;
  RegnetPlan p;
  make_plan(depth, hf, wf, in_channels, base_filter, precision, &p);
  const char* ws = (const char*)workspace;
  const int cpad = plan_cpad(p);
  if (scale) *scale = (const float*)(ws + p.scale_off) + (size_t)layer * cpad;
  if (shift) *shift = (const float*)(ws + p.shift_off) + (size_t)layer * cpad;
  if (layer == MVSB200_L_3DCONV6_2) return nullptr;
  return ws + p.raw_off[layer];
}

// ---------------------------------------------------------------------------------------------
// D-slab mode: ONE volume split along depth over `slabs` GPUs (BASELINE config 5).  Every rank runs the
// bf16 path on D/slabs consecutive planes; between layers the host exchanges (NCCL over NVLink, see
// mvsnet_b200/dslab.py) one boundary plane per tensor with each neighbour and all-reduces the batch statistics.
// Tensors carry one halo plane before and after the local planes; the library only exposes where they are.
// ---------------------------------------------------------------------------------------------
namespace mvsb200 {
int launch_cost_volume_slab(const float* feats, const float* homographies, int n_views, int depth_num, int d0g,
                            int dloc, int hf, int wf, int channels, int order, int sampler, void* cp8, void* ps8,
                            void* feats16, const float* coef_table, cudaStream_t s);
size_t cost_volume_pair_bytes(int n_views, int hf, int wf);
bool cost_volume_planar_ok(int n_views, int hf, int wf, int channels, int sampler);

struct SlabPlan {
  RegnetPlan net;                       // layer table; dims = LOCAL extents (depth / slabs)
  size_t raw_off[MVSB200_REGNET_LAYERS], ps8_off[MVSB200_REGNET_LAYERS];      // tensors with halo planes
  size_t raw_plane[MVSB200_REGNET_LAYERS], ps8_plane[MVSB200_REGNET_LAYERS];  // bytes per plane
  size_t cost_cp8_off, cost_ps8_off, cost_cp8_plane, cost_ps8_plane;
  size_t hom_off, coef_off, pair_off, filtered_off, stats_off, stats_bytes, scratch_off, total;
  size_t gstats_off, gstats_bytes;      // peer mode: statistics of every layer per source rank [layer][slabs][2*cpad]
  size_t flags_off, flags_bytes;        // peer mode: publication flags [layer][slabs] + an error word + an abort word
  int cpad, slabs;
};

static int make_slab_plan(int n_views, int D, int slabs, int H, int W, int cin, int b, SlabPlan* sp) {
  MVS_CHECK_ARG(slabs >= 1 && D % slabs == 0 && (D / slabs) % 8 == 0,
                "slab: depth %d must split into %d slabs of a multiple of 8 planes", D, slabs);
  const int Dl = D / slabs;
  int rc = check_regnet_shape(Dl, H, W, cin, b);
  if (rc) return rc;
  MVS_CHECK_ARG(cin % 8 == 0 && b % 8 == 0, "slab: channel counts must be multiples of 8");
  make_plan(Dl, H, W, cin, b, MVSB200_PRECISION_BF16, &sp->net);
  const RegnetPlan& p = sp->net;
  size_t off = 0;
  sp->hom_off = off;  off += align_up((size_t)(n_views - 1) * D * 9 * sizeof(float), 256);
  sp->coef_off = off; off += align_up((size_t)(n_views - 1) * D * 8 * sizeof(float), 256);
  sp->pair_off = off; off += align_up(cost_volume_pair_bytes(n_views, H, W), 256);
  sp->cost_cp8_plane = planar_bytes(1, H, W, cin, 0);
  sp->cost_ps8_plane = planar_bytes(1, H, W, cin, 1);
  sp->cost_cp8_off = off; off += align_up(sp->cost_cp8_plane * (Dl + 2), 256);
  sp->cost_ps8_off = off; off += align_up(sp->cost_ps8_plane * (Dl + 2), 256);
  for (int i = 0; i < MVSB200_REGNET_LAYERS; ++i) {
    const LayerDesc& L = p.layer[i];
    const int* d = p.dims[L.out_level];
    sp->raw_plane[i] = sp->ps8_plane[i] = 0;
    sp->raw_off[i] = sp->ps8_off[i] = 0;
    if (i == MVSB200_L_3DCONV6_2) continue;
    sp->raw_plane[i] = planar_bytes(1, d[1], d[2], L.cout, 0);
    sp->raw_off[i] = off; off += align_up(sp->raw_plane[i] * (d[0] + 2), 256);
    if (p.has_ps8[i]) {
      sp->ps8_plane[i] = planar_bytes(1, d[1], d[2], L.cout, 1);
      sp->ps8_off[i] = off; off += align_up(sp->ps8_plane[i] * (d[0] + 2), 256);
    }
  }
  sp->filtered_off = off; off += align_up((size_t)Dl * H * W * sizeof(float), 256);
  sp->cpad = 64;
  // statistics [layer][copy][2*cpad]: one contiguous region per layer for the all-reduce
  sp->stats_bytes = (size_t)MVSB200_REGNET_LAYERS * kStatsReps * 2 * sp->cpad * sizeof(double);
  sp->stats_off = off; off += align_up(sp->stats_bytes, 256);
  sp->scratch_off = off; off += align_up(conv3d_tc_pack_slot_bytes() * 2 * MVSB200_REGNET_LAYERS, 256);
  sp->slabs = slabs;
  sp->gstats_bytes = (size_t)MVSB200_REGNET_LAYERS * slabs * 2 * sp->cpad * sizeof(double);
  sp->gstats_off = off; off += align_up(sp->gstats_bytes, 256);
  sp->flags_bytes = (size_t)(MVSB200_REGNET_LAYERS * slabs + 2) * sizeof(unsigned);
  sp->flags_off = off; off += align_up(sp->flags_bytes, 256);
  sp->total = off;
  return MVSB200_OK;
}
}  // namespace mvsb200

extern "C" size_t mvsb200_slab_workspace_bytes(int n_views, int depth_num, int slabs, int hf, int wf, int channels,
                                               int base_filter) {
  SlabPlan sp;
  if (n_views < 2 || make_slab_plan(n_views, depth_num, slabs, hf, wf, channels, base_filter, &sp)) return 0;
  return sp.total;
}

// Stage 0 of a slab: homographies of all planes, the slab's cost volume planes (plus its two halo planes, which
// need no exchange: every rank holds all feature maps), weight packing, cleared statistics.
extern "C" int mvsb200_slab_begin(const float* feats, const float* cams, int n_views, int depth_num, int slab, int slabs,
                                  int hf, int wf, int channels, float depth_start, float depth_interval,
                                  int inverse_depth, int order, const mvsb200_regnet_params* params, int base_filter,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  MVS_CHECK_ARG(feats && cams && params && workspace, "slab_begin: NULL pointer");
  MVS_CHECK_ARG(slab >= 0 && slab < slabs, "slab_begin: slab %d of %d", slab, slabs);
  SlabPlan sp;
  int rc = make_slab_plan(n_views, depth_num, slabs, hf, wf, channels, base_filter, &sp);
  if (rc) return rc;
  if (workspace_bytes < sp.total) {
    set_error("slab_begin: workspace %zu < required %zu bytes", workspace_bytes, sp.total);
    return MVSB200_ERR_WORKSPACE;
  }
  MVS_CHECK_ARG(cost_volume_planar_ok(n_views, hf, wf, channels, MVSB200_SAMPLER_TRANSFORM),
                "slab_begin: needs 32 feature channels and at most 8 views");
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const RegnetPlan& p = sp.net;
  const int Dl = depth_num / slabs;
  volatile float dm1 = (float)depth_num - 1.0f;
  volatile float prod = dm1 * depth_interval;
  volatile float depth_end = depth_start + prod;
  float* homs = (float*)(ws + sp.hom_off);
  float* coefs = (float*)(ws + sp.coef_off);
  rc = launch_homographies(cams, n_views, depth_num, depth_start, inverse_depth ? (float)depth_end : depth_interval,
                           inverse_depth, homs, coefs, s);
  if (rc) return rc;
  // SAME padding at the ends of the volume: the halo planes nobody writes
  auto clear_ends = [&](size_t off, size_t plane, int planes) -> int {
    if (slab == 0) MVS_CUDA(cudaMemsetAsync(ws + off, 0, plane, s));
    if (slab == slabs - 1) MVS_CUDA(cudaMemsetAsync(ws + off + plane * (size_t)(planes + 1), 0, plane, s));
    return MVSB200_OK;
  };
  if ((rc = clear_ends(sp.cost_cp8_off, sp.cost_cp8_plane, Dl))) return rc;
  if ((rc = clear_ends(sp.cost_ps8_off, sp.cost_ps8_plane, Dl))) return rc;
  for (int i = 0; i < MVSB200_REGNET_LAYERS; ++i) {
    const int dl = p.dims[p.layer[i].out_level][0];
    if (sp.raw_plane[i] && (rc = clear_ends(sp.raw_off[i], sp.raw_plane[i], dl))) return rc;
    if (sp.ps8_plane[i] && (rc = clear_ends(sp.ps8_off[i], sp.ps8_plane[i], dl))) return rc;
  }
  rc = launch_cost_volume_slab(feats, homs, n_views, depth_num, slab * Dl - 1, Dl + 2, hf, wf, channels, order,
                               MVSB200_SAMPLER_TRANSFORM, ws + sp.cost_cp8_off, ws + sp.cost_ps8_off, ws + sp.pair_off,
                               coefs, s);
  if (rc) return rc;
  MVS_CUDA(cudaMemsetAsync(ws + sp.stats_off, 0, sp.stats_bytes, s));
  TcPackJob jobs[MVSB200_REGNET_LAYERS];
  for (int i = 0; i < MVSB200_REGNET_LAYERS; ++i) {
    const LayerDesc& L = p.layer[i];
    MVS_CHECK_ARG(params->kernel[i] != nullptr, "slab_begin: kernel[%d] is NULL", i);
    const int* d = p.dims[L.in_level];
    jobs[i] = {params->kernel[i], d[0], d[1], d[2], L.cin, L.cout, L.stride, L.transposed, L.skip >= 0 ? 1 : 0,
               (L.src >= 0 || L.skip >= 0) ? 1 : 0, 2 * i};
  }
  return conv3d_tc_pack_all(jobs, MVSB200_REGNET_LAYERS, ws + sp.scratch_off, s);
}

// Layer `layer` on the local slab.  Its inputs' halo planes and its producers' statistics must have been
// exchanged (mvsb200_slab_regions says where they live).
extern "C" int mvsb200_slab_layer(int layer, int n_views, int depth_num, int slab, int slabs, int hf, int wf,
                                  int channels, const mvsb200_regnet_params* params, int base_filter, float bn_eps,
                                  void* workspace, void* stream) {
  MVS_CHECK_ARG(layer >= 0 && layer < MVSB200_REGNET_LAYERS && params && workspace, "slab_layer: bad arguments");
  SlabPlan sp;
  int rc = make_slab_plan(n_views, depth_num, slabs, hf, wf, channels, base_filter, &sp);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const RegnetPlan& p = sp.net;
  const LayerDesc& L = p.layer[layer];
  const bool last = layer == MVSB200_L_3DCONV6_2;
  const bool s2 = L.stride == 2 && !L.transposed;
  const int* d = p.dims[L.in_level];
  double* stats = (double*)(ws + sp.stats_off);
  const int lstride = kStatsReps * 2 * sp.cpad;          // doubles per layer
  const void* x = L.src < 0 ? (const void*)(ws + (s2 ? sp.cost_ps8_off : sp.cost_cp8_off))
                            : (const void*)(ws + (s2 ? sp.ps8_off[L.src] : sp.raw_off[L.src]));
  const void* sk = L.skip < 0 ? nullptr : (const void*)(ws + sp.raw_off[L.skip]);
  TcBnSrc xbn = {nullptr, nullptr, nullptr, 1.0, bn_eps, 0, 1, 0}, sbn = xbn;
  // batch statistics span the whole volume: counts are global, the sums have been all-reduced by the host
  if (L.src >= 0)
    xbn = {stats + (size_t)L.src * lstride, params->gamma[L.src], params->beta[L.src],
           (double)p.vox[p.layer[L.src].out_level] * slabs, bn_eps, p.layer[L.src].cout, kStatsReps, 2 * sp.cpad};
  if (L.skip >= 0)
    sbn = {stats + (size_t)L.skip * lstride, params->gamma[L.skip], params->beta[L.skip],
           (double)p.vox[p.layer[L.skip].out_level] * slabs, bn_eps, p.layer[L.skip].cout, kStatsReps, 2 * sp.cpad};
  const TcSlab win = {1, slab == 0 ? 1 : 0, slab == slabs - 1 ? d[0] + 1 : d[0] + 2};
  // outputs start at extended plane 1
  void* y_cp8 = last ? nullptr : ws + sp.raw_off[layer] + sp.raw_plane[layer];
  void* y_ps8 = (!last && sp.ps8_plane[layer]) ? ws + sp.ps8_off[layer] + sp.ps8_plane[layer] : nullptr;
  return launch_conv3d_tc(x, nullptr, nullptr, sk, nullptr, nullptr, params->kernel[layer], d[0], d[1], d[2], L.cin, L.cout,
                          L.stride, L.transposed, y_cp8, y_ps8, last ? (float*)(ws + sp.filtered_off) : nullptr,
                          last ? nullptr : stats + (size_t)layer * lstride, nullptr, L.src >= 0 ? &xbn : nullptr,
                          L.skip >= 0 ? &sbn : nullptr, ws + sp.scratch_off + (size_t)2 * layer * conv3d_tc_pack_slot_bytes(),
                          kStatsReps, 2 * sp.cpad, &win, nullptr, nullptr, s);
}

// ---- D-slab mode over peer memory (NVLink) -----------------------------------------------------------------------
// The exchange between layers happens inside the kernels: the producing epilogue stores its boundary planes into the
// neighbours' halo planes through peer pointers, a one-block publish kernel copies the layer's statistics into every
// rank's per-source table and then raises this rank's flag on every rank, and the consuming kernel spins (bounded) on
// its local flags before it reads anything.  No NCCL call between layers.
namespace mvsb200 {
__global__ void slab_publish_kernel(const double* __restrict__ local_stats, int reps, int rep_stride, int n,
                                    char* const* __restrict__ peers, int slabs, int slab, size_t gstats_off,
                                    size_t flags_off, int layer, int cpad, unsigned seq) {
  // totals of this rank's partial copies -> slot `slab` of the layer's table on every rank
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    double t = 0.0;
    for (int r = 0; r < reps; ++r) t += local_stats[(size_t)r * rep_stride + c];
    for (int q = 0; q < slabs; ++q) {
      double* g = reinterpret_cast<double*>(peers[q] + gstats_off) + ((size_t)layer * slabs + slab) * 2 * cpad;
      g[c] = t;
    }
  }
  __threadfence_system();
  __syncthreads();
  // (the layer kernel before us in the stream has completed, so its peer stores are done as well)
  if (threadIdx.x < slabs) {
    unsigned* f = reinterpret_cast<unsigned*>(peers[threadIdx.x] + flags_off) + layer * slabs + slab;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
  }
}
}  // namespace mvsb200

// Layer `layer` on the local slab with the exchange fused in.  peers[slabs]: device-visible base addresses of every
// rank's slab workspace (own one included, from mvsb200_ipc_open); seq: 1, 2, 3, ... per inference.
extern "C" int mvsb200_slab_layer_p2p(int layer, int n_views, int depth_num, int slab, int slabs, int hf, int wf,
                                      int channels, const mvsb200_regnet_params* params, int base_filter, float bn_eps,
                                      void* workspace, void* const* peers_dev, void* const* peers_host, unsigned seq,
                                      void* stream) {
  MVS_CHECK_ARG(layer >= 0 && layer < MVSB200_REGNET_LAYERS && params && workspace && peers_dev && peers_host,
                "slab_layer_p2p: bad arguments");
  MVS_CHECK_ARG(slabs >= 1 && slabs <= 8 && slab >= 0 && slab < slabs, "slab_layer_p2p: slab %d of %d", slab, slabs);
  SlabPlan sp;
  int rc = make_slab_plan(n_views, depth_num, slabs, hf, wf, channels, base_filter, &sp);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const RegnetPlan& p = sp.net;
  const LayerDesc& L = p.layer[layer];
  const bool last = layer == MVSB200_L_3DCONV6_2;
  const bool s2 = L.stride == 2 && !L.transposed;
  const int* d = p.dims[L.in_level];
  double* stats = (double*)(ws + sp.stats_off);
  double* gstats = (double*)(ws + sp.gstats_off);
  unsigned* flags = (unsigned*)(ws + sp.flags_off);
  const int lstride = kStatsReps * 2 * sp.cpad, gstride = slabs * 2 * sp.cpad;
  const void* x = L.src < 0 ? (const void*)(ws + (s2 ? sp.cost_ps8_off : sp.cost_cp8_off))
                            : (const void*)(ws + (s2 ? sp.ps8_off[L.src] : sp.raw_off[L.src]));
  const void* sk = L.skip < 0 ? nullptr : (const void*)(ws + sp.raw_off[L.skip]);
  TcBnSrc xbn = {nullptr, nullptr, nullptr, 1.0, bn_eps, 0, 1, 0}, sbn = xbn;
  if (L.src >= 0)
    xbn = {gstats + (size_t)L.src * gstride, params->gamma[L.src], params->beta[L.src],
           (double)p.vox[p.layer[L.src].out_level] * slabs, bn_eps, p.layer[L.src].cout, slabs, 2 * sp.cpad};
  if (L.skip >= 0)
    sbn = {gstats + (size_t)L.skip * gstride, params->gamma[L.skip], params->beta[L.skip],
           (double)p.vox[p.layer[L.skip].out_level] * slabs, bn_eps, p.layer[L.skip].cout, slabs, 2 * sp.cpad};
  const TcSlab win = {1, slab == 0 ? 1 : 0, slab == slabs - 1 ? d[0] + 1 : d[0] + 2};
  TcPeer peer = {};
  const int dlo = p.dims[L.out_level][0];
  if (!last) {
    const size_t off[2] = {sp.raw_off[layer], sp.ps8_off[layer]}, plane[2] = {sp.raw_plane[layer], sp.ps8_plane[layer]};
    for (int t = 0; t < 2; ++t) {
      if (!plane[t]) continue;
      if (slab > 0) peer.mir_prev[t] = (char*)peers_host[slab - 1] + off[t] + plane[t] * (size_t)(dlo + 1);   // its AFTER halo
      if (slab < slabs - 1) peer.mir_next[t] = (char*)peers_host[slab + 1] + off[t];                            // its BEFORE halo
    }
  }
  peer.wait_n = slabs; peer.wait_seq = seq; peer.err_flag = flags + MVSB200_REGNET_LAYERS * slabs;
  peer.wait_flags[0] = L.src >= 0 ? flags + L.src * slabs : nullptr;
  peer.wait_flags[1] = L.skip >= 0 ? flags + L.skip * slabs : nullptr;
  void* y_cp8 = last ? nullptr : ws + sp.raw_off[layer] + sp.raw_plane[layer];
  void* y_ps8 = (!last && sp.ps8_plane[layer]) ? ws + sp.ps8_off[layer] + sp.ps8_plane[layer] : nullptr;
  rc = launch_conv3d_tc(x, nullptr, nullptr, sk, nullptr, nullptr, params->kernel[layer], d[0], d[1], d[2], L.cin, L.cout,
                        L.stride, L.transposed, y_cp8, y_ps8, last ? (float*)(ws + sp.filtered_off) : nullptr,
                        last ? nullptr : stats + (size_t)layer * lstride, nullptr, L.src >= 0 ? &xbn : nullptr,
                        L.skip >= 0 ? &sbn : nullptr, ws + sp.scratch_off + (size_t)2 * layer * conv3d_tc_pack_slot_bytes(),
                        kStatsReps, 2 * sp.cpad, &win, &peer, nullptr, s);
  if (rc || last) return rc;
  slab_publish_kernel<<<1, 128, 0, s>>>(stats + (size_t)layer * lstride, kStatsReps, 2 * sp.cpad, 2 * L.cout,
                                        (char* const*)peers_dev, slabs, slab, sp.gstats_off, sp.flags_off, layer, sp.cpad,
                                        seq);
  MVS_LAUNCH_CHECK("slab_publish_kernel");
  return MVSB200_OK;
}

// 1 when a consumer gave up waiting for a publication flag (a rank died or ran out of order); clears it
extern "C" int mvsb200_slab_p2p_error(int n_views, int depth_num, int slabs, int hf, int wf, int channels, int base_filter,
                                      void* workspace, void* stream) {
  SlabPlan sp;
  if (make_slab_plan(n_views, depth_num, slabs, hf, wf, channels, base_filter, &sp)) return MVSB200_ERR_INVALID;
  unsigned v = 0;
  unsigned* f = (unsigned*)((char*)workspace + sp.flags_off) + MVSB200_REGNET_LAYERS * slabs;
  MVS_CUDA(cudaMemcpyAsync(&v, f, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  MVS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (v) MVS_CUDA(cudaMemsetAsync(f, 0, 2 * sizeof(v), (cudaStream_t)stream));       // error word + abort word
  return (int)v;
}

// Release every kernel of THIS rank that is waiting for a publication flag (it gives up at once and raises the error
// word): what a watchdog calls on the surviving ranks when a rank is lost, instead of sitting out the ~2 s bound per
// layer.  The write goes out on a stream of its own -- the compute stream is the one that is stuck.  Cleared by
// mvsb200_slab_p2p_error.
extern "C" int mvsb200_slab_p2p_abort(int n_views, int depth_num, int slabs, int hf, int wf, int channels, int base_filter,
                                      void* workspace) {
  SlabPlan sp;
  if (!workspace || make_slab_plan(n_views, depth_num, slabs, hf, wf, channels, base_filter, &sp)) return MVSB200_ERR_INVALID;
  static thread_local cudaStream_t side[64] = {};
  int dev = 0;
  MVS_CUDA(cudaGetDevice(&dev));
  if (!side[dev & 63]) MVS_CUDA(cudaStreamCreateWithFlags(&side[dev & 63], cudaStreamNonBlocking));
  static const unsigned one = 1u;
  unsigned* f = (unsigned*)((char*)workspace + sp.flags_off) + MVSB200_REGNET_LAYERS * slabs + 1;
  MVS_CUDA(cudaMemcpyAsync(f, &one, sizeof(one), cudaMemcpyHostToDevice, side[dev & 63]));
  MVS_CUDA(cudaStreamSynchronize(side[dev & 63]));
  return MVSB200_OK;
}

// Peer-visible device memory for the slab workspace (cudaMalloc + CUDA IPC): ranks are separate processes.
extern "C" int mvsb200_ipc_alloc(size_t bytes, void** ptr) {
  MVS_CHECK_ARG(ptr && bytes > 0, "ipc_alloc: bad arguments");
  MVS_CUDA(cudaMalloc(ptr, bytes));
  MVS_CUDA(cudaMemset(*ptr, 0, bytes));
  return MVSB200_OK;
}
extern "C" int mvsb200_ipc_free(void* ptr) {
  MVS_CUDA(cudaFree(ptr));
  return MVSB200_OK;
}
extern "C" int mvsb200_ipc_export(void* ptr, unsigned char* handle64) {
  MVS_CHECK_ARG(ptr && handle64, "ipc_export: NULL pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  MVS_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, 64);
  return MVSB200_OK;
}
extern "C" int mvsb200_ipc_open(const unsigned char* handle64, void** ptr) {
  MVS_CHECK_ARG(ptr && handle64, "ipc_open: NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  MVS_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return MVSB200_OK;
}
extern "C" int mvsb200_ipc_close(void* ptr) {
  MVS_CUDA(cudaIpcCloseMemHandle(ptr));
  return MVSB200_OK;
}

// Regions the host exchanges after layer `layer` (byte offsets into the workspace):
//   out[0], out[1]   statistics of the layer: offset, bytes (fp64; all-reduce SUM over the slabs)
//   per tensor t in {0: chunk-planar, 1: parity-split copy}: out[2+5t .. 6+5t] = plane bytes (0 = absent),
//   first local plane (-> previous rank's AFTER halo), last local plane (-> next rank's BEFORE halo),
//   own BEFORE halo, own AFTER halo
//   out[12], out[13] the filtered slab [D/slabs, Hf, Wf] fp32: offset, bytes
extern "C" int mvsb200_slab_regions(int layer, int n_views, int depth_num, int slabs, int hf, int wf, int channels,
                                    int base_filter, unsigned long long* out) {
  MVS_CHECK_ARG(layer >= 0 && layer < MVSB200_REGNET_LAYERS && out, "slab_regions: bad arguments");
  SlabPlan sp;
  int rc = make_slab_plan(n_views, depth_num, slabs, hf, wf, channels, base_filter, &sp);
  if (rc) return rc;
  const int dl = sp.net.dims[sp.net.layer[layer].out_level][0];
  out[0] = sp.stats_off + (size_t)layer * kStatsReps * 2 * sp.cpad * sizeof(double);
  out[1] = layer == MVSB200_L_3DCONV6_2 ? 0 : (size_t)kStatsReps * 2 * sp.cpad * sizeof(double);
  const size_t offs[2] = {sp.raw_off[layer], sp.ps8_off[layer]}, planes[2] = {sp.raw_plane[layer], sp.ps8_plane[layer]};
  for (int t = 0; t < 2; ++t) {
    out[2 + 5 * t] = planes[t];
    out[3 + 5 * t] = offs[t] + planes[t];                       // first local plane
    out[4 + 5 * t] = offs[t] + planes[t] * (size_t)dl;          // last local plane
    out[5 + 5 * t] = offs[t];                                   // before halo
    out[6 + 5 * t] = offs[t] + planes[t] * (size_t)(dl + 1);    // after halo
  }
  out[12] = sp.filtered_off;
  out[13] = (size_t)(depth_num / slabs) * hf * wf * sizeof(float);
  return MVSB200_OK;
}

// ---------------------------------------------------------------------------------------------
// whole path
// ---------------------------------------------------------------------------------------------
namespace mvsb200 {
bool cost_volume_planar_ok(int n_views, int hf, int wf, int channels, int sampler);
int launch_cost_volume_planar(const float* feats, const float* homographies, int n_views, int depth_num, int hf,
                              int wf, int channels, int order, int sampler, void* cp8, void* ps8, void* feats16,
                              const float* coef_table, cudaStream_t s, bool feats16_ready = false);
bool cost_volume_uses_window(int n_views, int dloc, int hf, int wf, int channels, int sampler);
int launch_planar_half_features(const float* feats, int n_views, int hf, int wf, void* feats16, cudaStream_t s);
int launch_cost_volume_coef(const float* feats, const float* homographies, const float* coef_table, int n_views,
                            int depth_num, int hf, int wf, int channels, int order, int sampler, int out_dtype,
                            void* out, cudaStream_t s);
size_t cost_volume_pair_bytes(int n_views, int hf, int wf);
}
namespace mvsb200 {
int launch_regress_combine(const float* partials, int slabs, const float* filtered, int dl, int d0, int D, int npix,
                           float depth_start, float depth_interval, int inverse_depth, int num_buckets, float* depth_map,
                           float* prob_partial, cudaStream_t s);
}
namespace {
inline float sub_host(float a, float b) { volatile float r = a - b; return r; }
struct InferPlan {
  size_t hom_off, coef_off, cost_off, filtered_off, pair_off, partial_off, regnet_off, total;
  size_t regnet_bytes;
};
void make_infer_plan(int n_views, int D, int hf, int wf, int C, int b, int precision, InferPlan* ip) {
  size_t off = 0;
  ip->hom_off = off;      off += align_up((size_t)(n_views - 1) * D * 9 * sizeof(float), 256);
  ip->coef_off = off;     off += align_up((size_t)(n_views - 1) * D * 8 * sizeof(float), 256);
  // (bf16 mode with the fast cost-volume kernel writes the volume straight into the regularizer's planar buffers
  // inside its workspace; the NDHWC buffer is then unused but stays reserved: the sampler is a per-call choice)
  ip->cost_off = off;     off += align_up((size_t)D * hf * wf * C * (precision == MVSB200_PRECISION_BF16 ? 2 : 4), 256);
  ip->filtered_off = off; off += align_up((size_t)D * hf * wf * sizeof(float), 256);
  ip->pair_off = off;     off += align_up(cost_volume_pair_bytes(n_views, hf, wf), 256);
  ip->partial_off = off;  off += align_up((size_t)3 * hf * wf * sizeof(float), 256);   // fused soft-argmin partials
  ip->regnet_off = off;
  ip->regnet_bytes = mvsb200_regnet_workspace_bytes(D, hf, wf, C, b, precision);
  off += ip->regnet_bytes;
  ip->total = off;
}
}  // namespace

// optional stage-boundary events (bench.py measures the dominant kernel live with these)
static thread_local cudaEvent_t t_stage_events[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
static thread_local bool t_stage_events_on = false;

extern "C" int mvsb200_infer_set_stage_events(void* const* events) {
  t_stage_events_on = events != nullptr;
  for (int i = 0; i < 5; ++i) t_stage_events[i] = events ? (cudaEvent_t)events[i] : nullptr;
  return MVSB200_OK;
}

#define MVS_STAGE_EVENT(i)                                                        \
  do {                                                                            \
    if (t_stage_events_on && t_stage_events[i]) MVS_CUDA(cudaEventRecord(t_stage_events[i], s)); \
  } while (0)

extern "C" size_t mvsb200_infer_workspace_bytes(int n_views, int depth_num, int hf, int wf, int channels,
                                                int base_filter, int precision) {
  if (n_views < 2 || depth_num <= 0 || hf <= 0 || wf <= 0 || channels <= 0 || base_filter <= 0) return 0;
  InferPlan ip;
  make_infer_plan(n_views, depth_num, hf, wf, channels, base_filter, precision, &ip);
  return ip.total;
}

// Byte offsets, inside the whole-path workspace, of the cost volume in the regularizer's planar layouts (bf16 mode;
// tests and inspection)
extern "C" int mvsb200_infer_cost_offsets(int n_views, int depth_num, int hf, int wf, int channels, int base_filter,
                                          int precision, size_t* cp8_offset, size_t* ps8_offset) {
  MVS_CHECK_ARG(cp8_offset && ps8_offset && precision == MVSB200_PRECISION_BF16, "infer_cost_offsets: bf16 mode only");
  MVS_CHECK_ARG(n_views >= 2, "infer_cost_offsets: n_views must be >= 2");
  int rc = check_regnet_shape(depth_num, hf, wf, channels, base_filter);
  if (rc) return rc;
  InferPlan ip;
  make_infer_plan(n_views, depth_num, hf, wf, channels, base_filter, precision, &ip);
  RegnetPlan p;
  make_plan(depth_num, hf, wf, channels, base_filter, precision, &p);
  *cp8_offset = ip.regnet_off + p.cost_cp8_off;
  *ps8_offset = ip.regnet_off + p.cost_ps8_off;
  return MVSB200_OK;
}

// Byte offset, inside the whole-path workspace, of the filtered cost volume [D,Hf,Wf] fp32 (output of 3dconv6_2) that
// the last mvsb200_infer call left there (tests: the regression stage against the oracle on the same volume)
extern "C" int mvsb200_infer_filtered_offset(int n_views, int depth_num, int hf, int wf, int channels, int base_filter,
                                             int precision, size_t* offset) {
  MVS_CHECK_ARG(offset != nullptr && n_views >= 2, "infer_filtered_offset: bad arguments");
  int rc = check_regnet_shape(depth_num, hf, wf, channels, base_filter);
  if (rc) return rc;
  InferPlan ip;
  make_infer_plan(n_views, depth_num, hf, wf, channels, base_filter, precision, &ip);
  *offset = ip.filtered_off;
  return MVSB200_OK;
}

// A side stream per (host thread, device) with its fork / join events, created on first use and kept for the life of
// the thread: work that does not depend on the cost volume (cleared statistics, packed weights: ~15 us) runs there,
// beside the cost-volume kernel, instead of between it and the first layer of the regularizer.
struct SideLane { cudaStream_t stream; cudaEvent_t fork, feats, join; };
static int side_lane(SideLane** out) {
  static thread_local SideLane pool[64] = {};
  int dev = 0;
  MVS_CUDA(cudaGetDevice(&dev));
  SideLane& l = pool[dev & 63];
  if (!l.stream) {
    MVS_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
    MVS_CUDA(cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming));
    MVS_CUDA(cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming));
    MVS_CUDA(cudaEventCreateWithFlags(&l.feats, cudaEventDisableTiming));
  }
  *out = &l;
  return MVSB200_OK;
}

extern "C" int mvsb200_infer(const float* feats, const float* cams, int n_views, int depth_num, int hf, int wf,
                             int channels, float depth_start, float depth_interval, int inverse_depth, int order,
                             int sampler, const mvsb200_regnet_params* params, int base_filter, float bn_eps,
                             int precision, float* depth_map, float* prob_map, void* workspace,
                             size_t workspace_bytes, void* stream) {
  MVS_CHECK_ARG(feats && cams && params && depth_map && prob_map && workspace, "infer: NULL pointer");
  MVS_CHECK_ARG(n_views >= 2, "infer: n_views must be >= 2 (got %d)", n_views);
  int rc = check_regnet_shape(depth_num, hf, wf, channels, base_filter);
  if (rc) return rc;
  InferPlan ip;
  make_infer_plan(n_views, depth_num, hf, wf, channels, base_filter, precision, &ip);
  if (workspace_bytes < ip.total) {
    set_error("infer: workspace %zu < required %zu bytes", workspace_bytes, ip.total);
    return MVSB200_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float* homs = (float*)(ws + ip.hom_off);
  void* cost = ws + ip.cost_off;
  float* filtered = (float*)(ws + ip.filtered_off);
  // model.py:378-379: depth_end = depth_start + (float(depth_num) - 1) * depth_interval (fp32)
  volatile float dm1 = (float)depth_num - 1.0f;
  volatile float prod = dm1 * depth_interval;
  volatile float depth_end = depth_start + prod;
  MVS_STAGE_EVENT(0);
  // fork: everything enqueued on `s` so far (the previous inference on this workspace) precedes the side lane's work
  SideLane* lane = nullptr;
  const bool side = tuning().infer_side != 0;
  bool feats_early = false;
  if (side) {
    rc = side_lane(&lane);
    if (rc) return rc;
    RegnetPlan rp;
    make_plan(depth_num, hf, wf, channels, base_filter, precision, &rp);
    MVS_CUDA(cudaEventRecord(lane->fork, s));
    MVS_CUDA(cudaStreamWaitEvent(lane->stream, lane->fork, 0));
    // first what the cost-volume kernel itself waits for: the fp16 copy of the feature maps (beside the homographies)
    feats_early = precision == MVSB200_PRECISION_BF16 && sampler == MVSB200_SAMPLER_TRANSFORM &&
                  cost_volume_planar_ok(n_views, hf, wf, channels, sampler) && tuning().cv_fp32_taps == 0 &&
                  cost_volume_uses_window(n_views, depth_num, hf, wf, channels, sampler);
    if (feats_early) {
      rc = launch_planar_half_features(feats, n_views, hf, wf, ws + ip.pair_off, lane->stream);
      if (rc) feats_early = false;
      cudaEventRecord(lane->feats, lane->stream);
    }
    rc = regnet_prepare(rp, params, depth_num, hf, wf, channels, base_filter, precision == MVSB200_PRECISION_BF16,
                        ws + ip.regnet_off, lane->stream);
    // (join even after an error: the side stream must not be left waiting inside a capture)
    cudaEventRecord(lane->join, lane->stream);
    if (rc) { cudaStreamWaitEvent(s, lane->join, 0); return rc; }
  }
  float* coefs = (float*)(ws + ip.coef_off);      // pixel-coordinate transform rows, private to this call's workspace
  rc = launch_homographies(cams, n_views, depth_num, depth_start, inverse_depth ? (float)depth_end : depth_interval,
                           inverse_depth, homs, coefs, s);
  if (rc) return rc;
  MVS_STAGE_EVENT(1);
  const int cost_dtype = precision == MVSB200_PRECISION_BF16 ? MVSB200_BF16 : MVSB200_F32;
  const bool planar = precision == MVSB200_PRECISION_BF16 && sampler == MVSB200_SAMPLER_TRANSFORM &&
                      cost_volume_planar_ok(n_views, hf, wf, channels, sampler);
  if (planar) {
    void *cp8 = nullptr, *ps8 = nullptr;
    regnet_cost_planar(ws + ip.regnet_off, depth_num, hf, wf, channels, base_filter, &cp8, &ps8);
    if (regnet_fuse01(depth_num, hf, wf, channels, base_filter)) ps8 = nullptr;      // nobody reads the parity-split copy
    const bool fp32_taps = tuning().cv_fp32_taps != 0;
    if (feats_early) MVS_CUDA(cudaStreamWaitEvent(s, lane->feats, 0));
    rc = launch_cost_volume_planar(feats, homs, n_views, depth_num, hf, wf, channels, order, sampler, cp8, ps8,
                                   fp32_taps ? nullptr : ws + ip.pair_off, coefs, s, feats_early);
  } else {
    rc = launch_cost_volume_coef(feats, homs, coefs, n_views, depth_num, hf, wf, channels, order, sampler, cost_dtype,
                                 cost, s);
  }
  if (rc) return rc;
  MVS_STAGE_EVENT(2);
  // bf16 mode, linear depth: the soft-argmin runs inside the last layer's epilogue when its plan allows it
  // (tf.linspace step in fp32, model.py:487-490)
  volatile float lin_num = sub_host((float)depth_num, 1.0f);
  volatile float lin_span = sub_host((float)depth_end, depth_start);
  volatile float lin_step = depth_num > 1 ? lin_span / lin_num : 0.0f;
  const bool no_fused_regress = tuning().no_fused_regress != 0;      // tests compare the two paths
  TcRegress rg = {(float*)(ws + ip.partial_off), depth_start, (float)lin_step, 0};
  const bool try_fuse = precision == MVSB200_PRECISION_BF16 && !inverse_depth && !no_fused_regress;
  if (side) MVS_CUDA(cudaStreamWaitEvent(s, lane->join, 0));
  rc = regnet_forward_impl(cost, cost_dtype, planar ? 1 : 0, params, depth_num, hf, wf, channels, base_filter, bn_eps,
                           precision, filtered, ws + ip.regnet_off, ip.regnet_bytes, s, try_fuse ? &rg : nullptr, false,
                           side);
  if (rc) return rc;
  MVS_STAGE_EVENT(3);
  if (try_fuse && rg.fused)
    rc = launch_regress_combine(rg.partial, 1, filtered, depth_num, 0, depth_num, hf * wf, depth_start, depth_interval,
                                inverse_depth, 4, depth_map, prob_map, s);
  else
    rc = launch_depth_regress(filtered, depth_num, hf, wf, depth_start, depth_interval, inverse_depth, 4, depth_map,
                              prob_map, nullptr, s);
  if (rc) return rc;
  MVS_STAGE_EVENT(4);
  return MVSB200_OK;
}

extern "C" size_t mvsb200_infer_host_staging_bytes(int n_views, int hf, int wf, int channels) {
  if (n_views < 2 || hf <= 0 || wf <= 0 || channels <= 0) return 0;
  return align_up((size_t)n_views * hf * wf * channels * sizeof(float), 256) +
         align_up((size_t)n_views * 32 * sizeof(float), 256) + 2 * align_up((size_t)hf * wf * sizeof(float), 256);
}

// one disable-timing event per (host thread, device), created on first use and kept for the life of the thread
static int chain_event(cudaEvent_t* out) {
  static thread_local cudaEvent_t pool[64] = {};
  int dev = 0;
  MVS_CUDA(cudaGetDevice(&dev));
  cudaEvent_t& e = pool[dev & 63];
  if (!e) MVS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *out = e;
  return MVSB200_OK;
}

// copies on `copy_stream`, kernels on `stream`; when they differ the two are chained with events (feed -> kernels ->
// fetch), so that a caller with one compute stream and one copy stream keeps the kernels of consecutive reference
// views back to back while the feed of the next view and the fetch of the previous one run beside them
static int infer_host_enqueue(const float* feats_host, const float* cams_host, int n_views, int depth_num, int hf,
                              int wf, int channels, float depth_start, float depth_interval, int inverse_depth, int order,
                              int sampler, const mvsb200_regnet_params* params, int base_filter, float bn_eps,
                              int precision, float* depth_map_host, float* prob_map_host, void* staging_dev,
                              void* workspace, size_t workspace_bytes, void* stream, void* copy_stream) {
  MVS_CHECK_ARG(feats_host && cams_host && depth_map_host && prob_map_host && staging_dev, "infer_host: NULL pointer");
  MVS_CHECK_ARG(n_views >= 2 && hf > 0 && wf > 0 && channels > 0, "infer_host: bad shape");
  cudaStream_t s = (cudaStream_t)copy_stream, sk = (cudaStream_t)stream;
  const bool split = s != sk;
  char* st = (char*)staging_dev;
  const size_t feat_bytes = (size_t)n_views * hf * wf * channels * sizeof(float);
  const size_t cam_bytes = (size_t)n_views * 32 * sizeof(float);
  const size_t map_bytes = (size_t)hf * wf * sizeof(float);
  float* d_feats = (float*)st;
  float* d_cams = (float*)(st + align_up(feat_bytes, 256));
  float* d_depth = (float*)((char*)d_cams + align_up(cam_bytes, 256));
  float* d_prob = (float*)((char*)d_depth + align_up(map_bytes, 256));
  MVS_CUDA(cudaMemcpyAsync(d_feats, feats_host, feat_bytes, cudaMemcpyHostToDevice, s));
  MVS_CUDA(cudaMemcpyAsync(d_cams, cams_host, cam_bytes, cudaMemcpyHostToDevice, s));
  // the two streams are chained feed -> kernels -> fetch through ONE pooled event per (thread, device): a wait
  // captures the record that precedes it, so the event can be re-recorded at once and is never destroyed per call
  cudaEvent_t ev = nullptr;
  if (split) {
    int rc_ev = chain_event(&ev);
    if (rc_ev) return rc_ev;
    MVS_CUDA(cudaEventRecord(ev, s));
    MVS_CUDA(cudaStreamWaitEvent(sk, ev, 0));
  }
  int rc = mvsb200_infer(d_feats, d_cams, n_views, depth_num, hf, wf, channels, depth_start, depth_interval,
                         inverse_depth, order, sampler, params, base_filter, bn_eps, precision, d_depth, d_prob,
                         workspace, workspace_bytes, stream);
  if (rc) return rc;
  if (split) {
    MVS_CUDA(cudaEventRecord(ev, sk));
    MVS_CUDA(cudaStreamWaitEvent(s, ev, 0));
  }
  MVS_CUDA(cudaMemcpyAsync(depth_map_host, d_depth, map_bytes, cudaMemcpyDeviceToHost, s));
  MVS_CUDA(cudaMemcpyAsync(prob_map_host, d_prob, map_bytes, cudaMemcpyDeviceToHost, s));
  return MVSB200_OK;
}

extern "C" int mvsb200_infer_host(const float* feats_host, const float* cams_host, int n_views, int depth_num,
                                  int hf, int wf, int channels, float depth_start, float depth_interval,
                                  int inverse_depth, int order, int sampler, const mvsb200_regnet_params* params,
                                  int base_filter, float bn_eps, int precision, float* depth_map_host,
                                  float* prob_map_host, void* staging_dev, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  int rc = infer_host_enqueue(feats_host, cams_host, n_views, depth_num, hf, wf, channels, depth_start, depth_interval,
                              inverse_depth, order, sampler, params, base_filter, bn_eps, precision, depth_map_host,
                              prob_map_host, staging_dev, workspace, workspace_bytes, stream, stream);
  if (rc) return rc;
  MVS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return MVSB200_OK;
}

extern "C" int mvsb200_infer_host_async(const float* feats_host, const float* cams_host, int n_views, int depth_num,
                                        int hf, int wf, int channels, float depth_start, float depth_interval,
                                        int inverse_depth, int order, int sampler,
                                        const mvsb200_regnet_params* params, int base_filter, float bn_eps,
                                        int precision, float* depth_map_host, float* prob_map_host, void* staging_dev,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  return infer_host_enqueue(feats_host, cams_host, n_views, depth_num, hf, wf, channels, depth_start, depth_interval,
                            inverse_depth, order, sampler, params, base_filter, bn_eps, precision, depth_map_host,
                            prob_map_host, staging_dev, workspace, workspace_bytes, stream, stream);
}

extern "C" int mvsb200_infer_host_pipelined(const float* feats_host, const float* cams_host, int n_views, int depth_num,
                                            int hf, int wf, int channels, float depth_start, float depth_interval,
                                            int inverse_depth, int order, int sampler,
                                            const mvsb200_regnet_params* params, int base_filter, float bn_eps,
                                            int precision, float* depth_map_host, float* prob_map_host,
                                            void* staging_dev, void* workspace, size_t workspace_bytes,
                                            void* compute_stream, void* copy_stream) {
  return infer_host_enqueue(feats_host, cams_host, n_views, depth_num, hf, wf, channels, depth_start, depth_interval,
                            inverse_depth, order, sampler, params, base_filter, bn_eps, precision, depth_map_host,
                            prob_map_host, staging_dev, workspace, workspace_bytes, compute_stream, copy_stream);
}
