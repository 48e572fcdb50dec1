// Layer table of the image feature tower UNetDS2GN (mvsnetworks.py:53-115) shared by the fp32 CUDA-core implementation
// (feature2d.cu) and the bf16 tensor-core implementation (feature2d_tc.cu).
#pragma once
#include "common.cuh"

namespace mvsb200 {

struct F2Layer {
  const char* name;
  int transposed, k, stride, mult;   // filters = base_filter * mult
  int src_a, src_b;                  // -1 = the images, else a layer index; src_b = -2: single source
  int gn, relu;
};

// mvsnetworks.py:58-115, in the order the reference builds them
static const F2Layer kUnet[MVSB200_UNET_LAYERS] = {
    {"2dconv1_0", 0, 3, 2, 2, -1, -2, 1, 1},  {"2dconv2_0", 0, 3, 2, 4, 0, -2, 1, 1},
    {"2dconv3_0", 0, 3, 2, 8, 1, -2, 1, 1},   {"2dconv4_0", 0, 3, 2, 16, 2, -2, 1, 1},
    {"2dconv0_1", 0, 3, 1, 1, -1, -2, 1, 1},  {"2dconv0_2", 0, 3, 1, 1, 4, -2, 1, 1},
    {"2dconv1_1", 0, 3, 1, 2, 0, -2, 1, 1},   {"2dconv1_2", 0, 3, 1, 2, 6, -2, 1, 1},
    {"2dconv2_1", 0, 3, 1, 4, 1, -2, 1, 1},   {"2dconv2_2", 0, 3, 1, 4, 8, -2, 1, 1},
    {"2dconv3_1", 0, 3, 1, 8, 2, -2, 1, 1},   {"2dconv3_2", 0, 3, 1, 8, 10, -2, 1, 1},
    {"2dconv4_1", 0, 3, 1, 16, 3, -2, 1, 1},  {"2dconv4_2", 0, 3, 1, 16, 12, -2, 1, 1},
    {"2dconv5_0", 1, 3, 2, 8, 13, -2, 1, 0},  {"2dconv5_1", 0, 3, 1, 8, 14, 11, 1, 1},
    {"2dconv5_2", 0, 3, 1, 8, 15, -2, 1, 1},  {"2dconv6_0", 1, 3, 2, 4, 16, -2, 1, 0},
    {"2dconv6_1", 0, 3, 1, 4, 17, 9, 1, 1},   {"2dconv6_2", 0, 3, 1, 4, 18, -2, 1, 1},
    {"2dconv7_0", 1, 3, 2, 2, 19, -2, 1, 0},  {"2dconv7_1", 0, 3, 1, 2, 20, 7, 1, 1},
    {"2dconv7_2", 0, 3, 1, 2, 21, -2, 1, 1},  {"2dconv8_0", 1, 3, 2, 1, 22, -2, 1, 0},
    {"2dconv8_1", 0, 3, 1, 1, 23, 5, 1, 1},   {"2dconv8_2", 0, 3, 1, 1, 24, -2, 1, 1},
    {"conv9_0", 0, 5, 2, 2, 25, -2, 1, 1},    {"conv9_1", 0, 3, 1, 2, 26, -2, 1, 1},
    {"conv9_2", 0, 3, 1, 2, 27, -2, 1, 1},    {"conv10_0", 0, 5, 2, 4, 28, -2, 1, 1},
    {"conv10_1", 0, 3, 1, 4, 29, -2, 1, 1},   {"conv10_2", 0, 3, 1, 4, 30, -2, 0, 0},
};

inline int unet_out_extent(int in, int stride, int transposed) { return transposed ? in * 2 : (in + stride - 1) / stride; }

}  // namespace mvsb200
