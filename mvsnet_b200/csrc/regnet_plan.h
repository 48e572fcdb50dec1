// Workspace layout of RegNetUS0 (mvsnetworks.py:122-158) shared by the forward (regnet.cu) and the backward (backward.cu).
#pragma once
#include "common.cuh"
#include "conv3d_tc.h"

namespace mvsb200 {

constexpr int kStatsReps = 16;      // partial copies of every layer's statistics (bf16 mode), summed by the consumers

// ---------------------------------------------------------------------------------------------
// workspace layout
// ---------------------------------------------------------------------------------------------
struct LayerDesc {
  int cin, cout, stride, transposed;
  int in_level, out_level;     // U-Net level of input / output volume (0 = full resolution)
  int src;                     // producing layer of the input, -1 = cost volume
  int skip;                    // layer added to the input (skip connection), -1 = none
  int cin_true, cout_true;     // channel counts of the layer's variables; cin / cout may be padded to whole 8-channel cells
};

struct RegnetPlan {
  LayerDesc layer[MVSB200_REGNET_LAYERS];
  size_t raw_off[MVSB200_REGNET_LAYERS];     // raw output of each layer (layer 10 writes `filtered` instead);
                                             // fp32 mode: NDHWC, bf16 mode: CP8 chunk-planar (conv3d_tc.cu)
  size_t ps8_off[MVSB200_REGNET_LAYERS];     // bf16 mode: parity-split copy for layers feeding a stride-2 conv
  bool has_ps8[MVSB200_REGNET_LAYERS];
  size_t cost_cp8_off, cost_ps8_off;         // bf16 mode: the cost volume in both planar layouts
  size_t stats_off, scale_off, shift_off, scratch_off;
  size_t stats_bytes, total;
  int elem;                                  // bytes per activation element
  size_t vox[4];
  int dims[4][3];
};

inline void make_plan(int D, int H, int W, int cin, int b, int precision, RegnetPlan* p) {
  // mvsnetworks.py:131-158
  const LayerDesc L[MVSB200_REGNET_LAYERS] = {
      {cin, 2 * b, 2, 0, 0, 1, -1, -1},                                     // 3dconv1_0
      {2 * b, 4 * b, 2, 0, 1, 2, MVSB200_L_3DCONV1_0, -1},                  // 3dconv2_0
      {4 * b, 8 * b, 2, 0, 2, 3, MVSB200_L_3DCONV2_0, -1},                  // 3dconv3_0
      {cin, b, 1, 0, 0, 0, -1, -1},                                         // 3dconv0_1
      {2 * b, 2 * b, 1, 0, 1, 1, MVSB200_L_3DCONV1_0, -1},                  // 3dconv1_1
      {4 * b, 4 * b, 1, 0, 2, 2, MVSB200_L_3DCONV2_0, -1},                  // 3dconv2_1
      {8 * b, 8 * b, 1, 0, 3, 3, MVSB200_L_3DCONV3_0, -1},                  // 3dconv3_1
      {8 * b, 4 * b, 2, 1, 3, 2, MVSB200_L_3DCONV3_1, -1},                  // 3dconv4_0
      {4 * b, 2 * b, 2, 1, 2, 1, MVSB200_L_3DCONV4_0, MVSB200_L_3DCONV2_1}, // 3dconv5_0 (input 3dconv4_1 = add)
      {2 * b, b, 2, 1, 1, 0, MVSB200_L_3DCONV5_0, MVSB200_L_3DCONV1_1},     // 3dconv6_0 (input 3dconv5_1 = add)
      {b, 1, 1, 0, 0, 0, MVSB200_L_3DCONV6_0, MVSB200_L_3DCONV0_1},         // 3dconv6_2 (input 3dconv6_1 = add)
  };
  for (int l = 0; l < 4; ++l) {
    p->dims[l][0] = D >> l; p->dims[l][1] = H >> l; p->dims[l][2] = W >> l;
    p->vox[l] = (size_t)(D >> l) * (H >> l) * (W >> l);
  }
  p->elem = precision == MVSB200_PRECISION_BF16 ? 2 : 4;
  size_t off = 0;
  int max_c = 0;
  // bf16 mode keeps activations in 16-byte cells of 8 channels: the narrow network modes (network.py:75-85: lite = 4,
  // ultralite = 2 base filters) run with every channel count padded to a whole cell; padded channels carry zero weights
  // and zero scale / shift, so they hold exact zeros everywhere
  const bool pad8 = precision == MVSB200_PRECISION_BF16 && b % 8 != 0;
  for (int i = 0; i < MVSB200_REGNET_LAYERS; ++i) {
    p->layer[i] = L[i];
    p->layer[i].cin_true = L[i].cin;
    p->layer[i].cout_true = L[i].cout;
    if (pad8) {
      p->layer[i].cin = (L[i].cin + 7) / 8 * 8;
      if (i != MVSB200_L_3DCONV6_2) p->layer[i].cout = (L[i].cout + 7) / 8 * 8;
    }
    p->raw_off[i] = off;
    if (i != MVSB200_L_3DCONV6_2) off += align_up(p->vox[L[i].out_level] * p->layer[i].cout * p->elem, 256);
    if (p->layer[i].cout > max_c) max_c = p->layer[i].cout;
    p->has_ps8[i] = false;
    p->ps8_off[i] = 0;
  }
  p->cost_cp8_off = p->cost_ps8_off = 0;
  if (precision == MVSB200_PRECISION_BF16) {
    // outputs that feed a stride-2 conv are also written parity-split
    for (int i = 0; i < MVSB200_REGNET_LAYERS; ++i)
      if (L[i].stride == 2 && !L[i].transposed && L[i].src >= 0) p->has_ps8[L[i].src] = true;
    for (int i = 0; i < MVSB200_REGNET_LAYERS; ++i)
      if (p->has_ps8[i]) {
        const int* d = p->dims[L[i].out_level];
        p->ps8_off[i] = off;
        off += align_up(planar_bytes(d[0], d[1], d[2], p->layer[i].cout, 1), 256);
      }
    p->cost_cp8_off = off; off += align_up(planar_bytes(D, H, W, cin, 0), 256);
    p->cost_ps8_off = off; off += align_up(planar_bytes(D, H, W, cin, 1), 256);
  }
  const int cpad = (max_c + 63) / 64 * 64;
  p->stats_off = off;
  p->stats_bytes = (size_t)(precision == MVSB200_PRECISION_BF16 ? kStatsReps : 1) * MVSB200_REGNET_LAYERS * 2 * cpad * sizeof(double);
  off += align_up(p->stats_bytes, 256);
  p->scale_off = off;  off += align_up((size_t)MVSB200_REGNET_LAYERS * cpad * sizeof(float), 256);
  p->shift_off = off;  off += align_up((size_t)MVSB200_REGNET_LAYERS * cpad * sizeof(float), 256);
  p->scratch_off = off;
  size_t scratch = 0;
  if (precision == MVSB200_PRECISION_BF16) scratch = conv3d_tc_pack_slot_bytes() * 2 * MVSB200_REGNET_LAYERS;   // 2 launch slots per layer
  off += align_up(scratch, 256);
  p->total = off;
}

inline int plan_cpad(const RegnetPlan& p) {
  return (int)(p.stats_bytes / ((p.elem == 2 ? kStatsReps : 1) * MVSB200_REGNET_LAYERS * 2 * sizeof(double)));
}

inline int check_regnet_shape(int D, int H, int W, int cin, int b) {
  MVS_CHECK_ARG(D > 0 && H > 0 && W > 0 && cin > 0 && b > 0, "regnet: bad shape D=%d H=%d W=%d Cin=%d base=%d", D, H, W,
                cin, b);
  // the reference graph only closes when every extent halves three times (mvsnetworks.py:148,152,156)
  MVS_CHECK_ARG(D % 8 == 0 && H % 8 == 0 && W % 8 == 0,
                "regnet: D, Hf, Wf must be multiples of 8 (got %d, %d, %d): the skip adds of RegNetUS0 do not "
                "line up otherwise", D, H, W);
  return MVSB200_OK;
}


}  // namespace mvsb200
