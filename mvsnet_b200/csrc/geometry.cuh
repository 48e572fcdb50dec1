// Plane-sweep geometry in strict fp32 (one IEEE rounding per operation, no FMA contraction):
// every function here must agree bit-for-bit with oracle/mvs_oracle.py.  The round-to-nearest
// intrinsics (__fmul_rn ...) are never fused by nvcc, which is what makes that possible.
#pragma once
#include "common.cuh"

namespace mvsb200 {

__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }

// C = A(3x3) * B(3xn), C_ij = (a_i0 b_0j + a_i1 b_1j) + a_i2 b_2j   (oracle matmul3)
template <int NCOL>
__device__ __forceinline__ void mm3(const float* A, const float* B, float* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < NCOL; ++j)
      C[i * NCOL + j] = add_(add_(mul_(A[i * 3 + 0], B[0 * NCOL + j]), mul_(A[i * 3 + 1], B[1 * NCOL + j])),
                             mul_(A[i * 3 + 2], B[2 * NCOL + j]));
}

// tf.matrix_inverse restated as partial-pivot LU + three triangular solves (oracle inv3x3_lu).
__device__ inline void inv3x3_lu(const float* A, float* inv) {
  float lu[3][3];
  int perm[3] = {0, 1, 2};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) lu[i][j] = A[i * 3 + j];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int p = k;
    float best = fabsf(lu[k][k]);
#pragma unroll
    for (int i = k + 1; i < 3; ++i) {
      float a = fabsf(lu[i][k]);
      if (a > best) { best = a; p = i; }
    }
    if (p != k) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { float t = lu[k][j]; lu[k][j] = lu[p][j]; lu[p][j] = t; }
      int tp = perm[k]; perm[k] = perm[p]; perm[p] = tp;
    }
#pragma unroll
    for (int i = k + 1; i < 3; ++i) {
      lu[i][k] = div_(lu[i][k], lu[k][k]);
#pragma unroll
      for (int j = k + 1; j < 3; ++j) lu[i][j] = sub_(lu[i][j], mul_(lu[i][k], lu[k][j]));
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float b0 = perm[0] == c ? 1.0f : 0.0f, b1 = perm[1] == c ? 1.0f : 0.0f, b2 = perm[2] == c ? 1.0f : 0.0f;
    float y0 = b0;
    float y1 = sub_(b1, mul_(lu[1][0], y0));
    float y2 = sub_(sub_(b2, mul_(lu[2][0], y0)), mul_(lu[2][1], y1));
    float x2 = div_(y2, lu[2][2]);
    float x1 = div_(sub_(y1, mul_(lu[1][2], x2)), lu[1][1]);
    float x0 = div_(sub_(sub_(y0, mul_(lu[0][1], x1)), mul_(lu[0][2], x2)), lu[0][0]);
    inv[0 * 3 + c] = x0; inv[1 * 3 + c] = x1; inv[2 * 3 + c] = x2;
  }
}

// Depth of plane i.  linear: float(i)*interval + start (homography_warping.py:28-30);
// inverse: 1/linspace(1/start, 1/end, D)[i] with TF LinSpace step=(stop-start)/(num-1) (:74-77).
__device__ __forceinline__ float plane_depth(int i, int depth_num, float depth_start, float depth_step,
                                             int inverse_depth) {
  if (!inverse_depth) return add_(mul_((float)i, depth_step), depth_start);
  float inv_start = div_(1.0f, depth_start);
  float inv_end = div_(1.0f, depth_step);
  if (depth_num == 1) return div_(1.0f, inv_start);
  float step = div_(sub_(inv_end, inv_start), (float)(depth_num - 1));
  return div_(1.0f, add_(inv_start, mul_(step, (float)i)));
}

// One homography H = K_r R_r (I - (c_r - c_l) n^T / d) R_l^T K_l^-1  (homography_warping.py:33-56).
// cam layout [2][4][4]: cam[0] extrinsic, cam[1][:3][:3] intrinsic (mvs_cluster.py:103-111).
__device__ inline void plane_homography(const float* left_cam, const float* right_cam, float depth, float* H) {
  float Rl[9], Rr[9], Kl[9], Kr[9], tl[3], tr[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      Rl[i * 3 + j] = left_cam[i * 4 + j];
      Rr[i * 3 + j] = right_cam[i * 4 + j];
      Kl[i * 3 + j] = left_cam[16 + i * 4 + j];
      Kr[i * 3 + j] = right_cam[16 + i * 4 + j];
    }
    tl[i] = left_cam[i * 4 + 3];
    tr[i] = right_cam[i * 4 + 3];
  }
  float Kl_inv[9], RlT[9], RrT[9];
  inv3x3_lu(Kl, Kl_inv);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { RlT[i * 3 + j] = Rl[j * 3 + i]; RrT[i * 3 + j] = Rr[j * 3 + i]; }
  float cl[3], cr[3], crel[3];
  mm3<1>(RlT, tl, cl);
  mm3<1>(RrT, tr, cr);
#pragma unroll
  for (int i = 0; i < 3; ++i) crel[i] = sub_(-cr[i], -cl[i]);
  float M0[9], M1[9], M2[9], M3[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      M0[i * 3 + j] = sub_(i == j ? 1.0f : 0.0f, div_(mul_(crel[i], Rl[2 * 3 + j]), depth));
  mm3<3>(RlT, Kl_inv, M1);
  mm3<3>(M0, M1, M2);
  mm3<3>(Rr, M2, M3);
  mm3<3>(Kr, M3, H);
}

// tf_transform_homography coefficient conversion (homography_warping.py:216-250).
__device__ __forceinline__ void transform_coefs(const float* h, float* t) {
  const float a0 = h[0], a1 = h[1], a2 = h[2], b0 = h[3], b1 = h[4], b2 = h[5], c0 = h[6], c1 = h[7], c2 = h[8];
  float a_0 = sub_(a0, div_(c0, 2.0f));
  float a_1 = sub_(a1, div_(c1, 2.0f));
  float a_2 = sub_(sub_(add_(div_(add_(a0, a1), 2.0f), a2), div_(add_(c0, c1), 4.0f)), div_(c2, 2.0f));
  float b_0 = sub_(b0, div_(c0, 2.0f));
  float b_1 = sub_(b1, div_(c1, 2.0f));
  float b_2 = sub_(sub_(add_(div_(add_(b0, b1), 2.0f), b2), div_(add_(c0, c1), 4.0f)), div_(c2, 2.0f));
  float c_2 = add_(c2, div_(add_(c0, c1), 2.0f));
  t[0] = div_(a_0, c_2); t[1] = div_(a_1, c_2); t[2] = div_(a_2, c_2);
  t[3] = div_(b_0, c_2); t[4] = div_(b_1, c_2); t[5] = div_(b_2, c_2);
  t[6] = div_(c0, c_2);  t[7] = div_(c1, c_2);
}

// TF ImageProjectiveTransform sample position of output pixel (x,y).
__device__ __forceinline__ void transform_coords(const float* t, float x, float y, float& ix, float& iy) {
  float proj = add_(add_(mul_(t[6], x), mul_(t[7], y)), 1.0f);
  ix = div_(add_(add_(mul_(t[0], x), mul_(t[1], y)), t[2]), proj);
  iy = div_(add_(add_(mul_(t[3], x), mul_(t[4], y)), t[5]), proj);
}

// Legacy homography_warping image coordinates (homography_warping.py:186-203): pixel-centre grid
// from TF linspace, affine and divide with the +1e-7 guard.
__device__ __forceinline__ float tf_linspace_at(float start, float stop, int num, int i) {
  if (num == 1) return start;
  float step = div_(sub_(stop, start), (float)(num - 1));
  return add_(start, mul_(step, (float)i));
}
__device__ __forceinline__ void legacy_coords(const float* h, int x, int y, int width, int height, float& xw,
                                              float& yw) {
  float gx = tf_linspace_at(0.5f, sub_((float)width, 0.5f), width, x);
  float gy = tf_linspace_at(0.5f, sub_((float)height, 0.5f), height, y);
  float ax = add_(add_(mul_(h[0], gx), mul_(h[1], gy)), mul_(h[2], 1.0f));
  float ay = add_(add_(mul_(h[3], gx), mul_(h[4], gy)), mul_(h[5], 1.0f));
  float dv = add_(add_(mul_(h[6], gx), mul_(h[7], gy)), mul_(h[8], 1.0f));
  dv = add_(dv, mul_(dv == 0.0f ? 1.0f : 0.0f, 1e-7f));
  xw = div_(ax, dv);
  yw = div_(ay, dv);
}

// floor -> int with NaN -> 0 and clamp to +-2^30 (oracle _floor_to_int).
__device__ __forceinline__ int floor_to_int(float v) {
  float f = floorf(v);
  if (f != f) f = 0.0f;
  f = fminf(fmaxf(f, -1073741824.0f), 1073741824.0f);
  return (int)f;
}

// Bilinear footprint of the contrib kernel for one sample position: the 2x2 corner (x0,y0),
// validity of the four taps and the weights.  Non-finite positions read as outside (all invalid).
struct Footprint {
  int x0, y0;          // top-left tap, valid range [-1, W-1] / [-1, H-1] when any tap is valid
  float wxl, wxr, wyl, wyr;
  bool vx0, vx1, vy0, vy1;
};
__device__ __forceinline__ Footprint make_footprint(float ix, float iy, int width, int height) {
  Footprint f;
  bool finite = (fabsf(ix) <= 3.0e38f) && (fabsf(iy) <= 3.0e38f);  // false for NaN and inf
  float xf = floorf(ix), yf = floorf(iy);
  float xc = xf + 1.0f, yc = yf + 1.0f;
  f.wxl = xc - ix; f.wxr = ix - xf;
  f.wyl = yc - iy; f.wyr = iy - yf;
  f.vx0 = finite && xf >= 0.0f && xf < (float)width;
  f.vx1 = finite && xc >= 0.0f && xc < (float)width;
  f.vy0 = finite && yf >= 0.0f && yf < (float)height;
  f.vy1 = finite && yc >= 0.0f && yc < (float)height;
  // clamp before the int conversion so wild coordinates cannot overflow
  f.x0 = (int)fminf(fmaxf(xf, -2.0f), (float)width);
  f.y0 = (int)fminf(fmaxf(yf, -2.0f), (float)height);
  if (!finite) { f.wxl = f.wxr = f.wyl = f.wyr = 0.0f; f.x0 = -2; f.y0 = -2; }
  return f;
}

}  // namespace mvsb200
