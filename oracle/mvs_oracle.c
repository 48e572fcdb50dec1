/*
 * mvs_oracle.c -- plain-C restatement of the geometry / warp / variance / regression part of the
 * MVSNet cost-volume hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ to
 * cross-check the NumPy oracle with an independent implementation and by bench.py's CPU-baseline leg.
 * PARITY UNPINNED: the reference ships no golden vectors and TensorFlow 1.12 cannot run here.
 *
 * Every function cites the reference lines it follows (paths under /root/reference).  All geometry is
 * IEEE fp32 with one rounding per operation: compile with -ffp-contract=off (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* C = A(3x3) * B(3xn), C_ij = (a_i0 b_0j + a_i1 b_1j) + a_i2 b_2j  -- tf.matmul, homography_warping.py:39-56 */
static void mm3(const float* A, const float* B, float* C, int n) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < n; ++j) {
      float p0 = A[i * 3 + 0] * B[0 * n + j], p1 = A[i * 3 + 1] * B[1 * n + j], p2 = A[i * 3 + 2] * B[2 * n + j];
      float s = p0 + p1;
      C[i * n + j] = s + p2;
    }
}

/* tf.matrix_inverse (homography_warping.py:33) as partial-pivot LU + triangular solves */
static void inv3x3_lu(const float* A, float* inv) {
  float lu[3][3];
  int perm[3] = {0, 1, 2};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) lu[i][j] = A[i * 3 + j];
  for (int k = 0; k < 3; ++k) {
    int p = k;
    float best = fabsf(lu[k][k]);
    for (int i = k + 1; i < 3; ++i) if (fabsf(lu[i][k]) > best) { best = fabsf(lu[i][k]); p = i; }
    if (p != k) {
      for (int j = 0; j < 3; ++j) { float t = lu[k][j]; lu[k][j] = lu[p][j]; lu[p][j] = t; }
      int tp = perm[k]; perm[k] = perm[p]; perm[p] = tp;
    }
    for (int i = k + 1; i < 3; ++i) {
      lu[i][k] = lu[i][k] / lu[k][k];
      for (int j = k + 1; j < 3; ++j) { float m = lu[i][k] * lu[k][j]; lu[i][j] = lu[i][j] - m; }
    }
  }
  for (int c = 0; c < 3; ++c) {
    float b0 = perm[0] == c, b1 = perm[1] == c, b2 = perm[2] == c;
    float y0 = b0;
    float m10 = lu[1][0] * y0; float y1 = b1 - m10;
    float m20 = lu[2][0] * y0; float t2 = b2 - m20; float m21 = lu[2][1] * y1; float y2 = t2 - m21;
    float x2 = y2 / lu[2][2];
    float m12 = lu[1][2] * x2; float x1 = (y1 - m12) / lu[1][1];
    float m01 = lu[0][1] * x1; float t0 = y0 - m01; float m02 = lu[0][2] * x2; float x0 = (t0 - m02) / lu[0][0];
    inv[0 * 3 + c] = x0; inv[1 * 3 + c] = x1; inv[2 * 3 + c] = x2;
  }
}

static float plane_depth(int i, int depth_num, float start, float step, int inverse) {
  if (!inverse) { float m = (float)i * step; return m + start; }                  /* :28-30 */
  float inv_start = 1.0f / start, inv_end = 1.0f / step;                            /* :74-77 */
  if (depth_num == 1) return 1.0f / inv_start;
  float st = (inv_end - inv_start) / (float)(depth_num - 1);
  float m = st * (float)i; float v = inv_start + m;
  return 1.0f / v;
}

/* get_homographies / get_homographies_inv_depth (homography_warping.py:10-106).
 * cams [n_views][2][4][4]; out H [(n_views-1)][depth_num][9]. */
void oracle_homographies(const float* cams, int n_views, int depth_num, float depth_start, float depth_step,
                         int inverse, float* H) {
  const float* L = cams;
  float Rl[9], Kl[9], tl[3], RlT[9], Kl_inv[9], cl[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) { Rl[i * 3 + j] = L[i * 4 + j]; Kl[i * 3 + j] = L[16 + i * 4 + j]; }
    tl[i] = L[i * 4 + 3];
  }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) RlT[i * 3 + j] = Rl[j * 3 + i];
  inv3x3_lu(Kl, Kl_inv);
  mm3(RlT, tl, cl, 1);
  float M1[9];
  mm3(RlT, Kl_inv, M1, 3);
  for (int v = 1; v < n_views; ++v) {
    const float* R = cams + (size_t)v * 32;
    float Rr[9], Kr[9], tr[3], RrT[9], cr[3], crel[3];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) { Rr[i * 3 + j] = R[i * 4 + j]; Kr[i * 3 + j] = R[16 + i * 4 + j]; }
      tr[i] = R[i * 4 + 3];
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) RrT[i * 3 + j] = Rr[j * 3 + i];
    mm3(RrT, tr, cr, 1);
    for (int i = 0; i < 3; ++i) crel[i] = (-cr[i]) - (-cl[i]);
    for (int d = 0; d < depth_num; ++d) {
      float depth = plane_depth(d, depth_num, depth_start, depth_step, inverse);
      float M0[9], M2[9], M3[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          float t = crel[i] * Rl[2 * 3 + j];
          float q = t / depth;
          M0[i * 3 + j] = (i == j ? 1.0f : 0.0f) - q;
        }
      mm3(M0, M1, M2, 3);
      mm3(Rr, M2, M3, 3);
      mm3(Kr, M3, H + ((size_t)(v - 1) * depth_num + d) * 9, 3);
    }
  }
}

/* tf_transform_homography coefficient conversion (homography_warping.py:216-250) */
void oracle_transform_coefs(const float* h, float* t) {
  float a0 = h[0], a1 = h[1], a2 = h[2], b0 = h[3], b1 = h[4], b2 = h[5], c0 = h[6], c1 = h[7], c2 = h[8];
  float hc0 = c0 / 2.0f, hc1 = c1 / 2.0f, hc2 = c2 / 2.0f;
  float sc = c0 + c1; float qc = sc / 4.0f; float hsc = sc / 2.0f;
  float a_0 = a0 - hc0, a_1 = a1 - hc1;
  float sa = a0 + a1; float ha = sa / 2.0f; float a_2 = ha + a2; a_2 = a_2 - qc; a_2 = a_2 - hc2;
  float b_0 = b0 - hc0, b_1 = b1 - hc1;
  float sb = b0 + b1; float hb = sb / 2.0f; float b_2 = hb + b2; b_2 = b_2 - qc; b_2 = b_2 - hc2;
  float c_2 = c2 + hsc;
  t[0] = a_0 / c_2; t[1] = a_1 / c_2; t[2] = a_2 / c_2; t[3] = b_0 / c_2; t[4] = b_1 / c_2; t[5] = b_2 / c_2;
  t[6] = c0 / c_2; t[7] = c1 / c_2;
}

static inline float read_fill0(const float* img, int H, int W, int C, float yy, float xx, int c) {
  if (!(yy >= 0.0f && yy < (float)H && xx >= 0.0f && xx < (float)W)) return 0.0f;
  return img[((size_t)(int)yy * W + (int)xx) * C + c];
}

/* tf.contrib.image.transform BILINEAR, zero fill (homography_warping.py:251).
 * image [H][W][C], t[8] -> out [H][W][C]; coords (optional) [H][W][2]. */
void oracle_transform_warp(const float* img, int H, int W, int C, const float* t, float* out, float* coords) {
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float fx = (float)x, fy = (float)y;
      float p6 = t[6] * fx, p7 = t[7] * fy; float pr = p6 + p7; pr = pr + 1.0f;
      float ax = t[0] * fx, bx = t[1] * fy; float sx = ax + bx; sx = sx + t[2];
      float ay = t[3] * fx, by = t[4] * fy; float sy = ay + by; sy = sy + t[5];
      float ix = sx / pr, iy = sy / pr;
      if (coords) { coords[((size_t)y * W + x) * 2] = ix; coords[((size_t)y * W + x) * 2 + 1] = iy; }
      if (!out) continue;
      float* o = out + ((size_t)y * W + x) * C;
      if (!(isfinite(ix) && isfinite(iy))) { memset(o, 0, sizeof(float) * C); continue; }
      float xf = floorf(ix), yf = floorf(iy), xc = xf + 1.0f, yc = yf + 1.0f;
      float wxl = xc - ix, wxr = ix - xf, wyl = yc - iy, wyr = iy - yf;
      for (int c = 0; c < C; ++c) {
        float m0 = wxl * read_fill0(img, H, W, C, yf, xf, c), m1 = wxr * read_fill0(img, H, W, C, yf, xc, c);
        float v0 = m0 + m1;
        float m2 = wxl * read_fill0(img, H, W, C, yc, xf, c), m3 = wxr * read_fill0(img, H, W, C, yc, xc, c);
        float v1 = m2 + m3;
        float u0 = wyl * v0, u1 = wyr * v1;
        o[c] = u0 + u1;
      }
    }
}

/* N-view variance cost volume (model.py:423-463 order 0 / :315-334 order 1), transform sampler.
 * feats [N][Hf][Wf][C], H [(N-1)][D][9] -> out [D][Hf][Wf][C].  OpenMP over depth planes. */
void oracle_cost_volume(const float* feats, const float* Hm, int N, int D, int Hf, int Wf, int C, int order, float* out) {
  const size_t plane = (size_t)Hf * Wf * C;
  const float n_f = (float)N, nn_f = (float)(N * N);
#pragma omp parallel
  {
    float* w = (float*)malloc(plane * sizeof(float));
    float* S = (float*)malloc(plane * sizeof(float));
    float* Q = (float*)malloc(plane * sizeof(float));
#pragma omp for schedule(dynamic)
    for (int d = 0; d < D; ++d) {
      for (size_t i = 0; i < plane; ++i) { S[i] = feats[i]; Q[i] = feats[i] * feats[i]; }
      for (int v = 0; v < N - 1; ++v) {
        float t[8];
        oracle_transform_coefs(Hm + ((size_t)v * D + d) * 9, t);
        oracle_transform_warp(feats + (size_t)(v + 1) * plane, Hf, Wf, C, t, w, NULL);
        for (size_t i = 0; i < plane; ++i) { S[i] = S[i] + w[i]; float sq = w[i] * w[i]; Q[i] = Q[i] + sq; }
      }
      float* o = out + (size_t)d * plane;
      for (size_t i = 0; i < plane; ++i) {
        if (order == 0) { float ss = S[i] * S[i]; float A = ss / nn_f; float q = Q[i] / n_f; o[i] = q - A; }
        else { float m = S[i] / n_f, m2 = Q[i] / n_f; float mm = m * m; o[i] = m2 - mm; }
      }
    }
    free(w); free(S); free(Q);
  }
}

static float linspace_at(float start, float stop, int num, int i) {
  if (num == 1) return start;
  float st = (stop - start) / (float)(num - 1);
  float m = st * (float)i;
  return start + m;
}

static int floor_to_int(float v) {
  float f = floorf(v);
  if (f != f) f = 0.0f;
  if (f < -1073741824.0f) f = -1073741824.0f;
  if (f > 1073741824.0f) f = 1073741824.0f;
  return (int)f;
}
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* softmax(-F) over depth, soft-argmin, 4-neighbour probability map (model.py:472-498, 45-144), linear depth.
 * F [D][npix] -> depth [npix], prob [npix], P [D][npix] (optional). */
void oracle_depth_regress(const float* F, int D, int npix, float start, float interval, int num_buckets, float* depth,
                          float* prob, float* P) {
  float dm1 = (float)D - 1.0f; float pr = dm1 * interval; float end = start + pr;   /* model.py:378-379 */
#pragma omp parallel for schedule(static)
  for (int p = 0; p < npix; ++p) {
    float m = -F[p];
    for (int d = 1; d < D; ++d) { float x = -F[(size_t)d * npix + p]; if (x > m) m = x; }
    float sum = 0.0f;
    for (int d = 0; d < D; ++d) { float e = expf(-F[(size_t)d * npix + p] - m); sum = sum + e; }
    float dep = 0.0f;
    for (int d = 0; d < D; ++d) {
      float e = expf(-F[(size_t)d * npix + p] - m);
      float pv = e / sum;
      if (P) P[(size_t)d * npix + p] = pv;
      float t = linspace_at(start, end, D, d) * pv;
      dep = dep + t;
    }
    depth[p] = dep;
    float idx = (dep - start) / interval;
    int l0 = clampi(floor_to_int(idx), 0, D - 1), r0 = clampi(floor_to_int(ceilf(idx)), 0, D - 1);
    int l1 = clampi(l0 - 1, 0, D - 1), r1 = clampi(r0 + 1, 0, D - 1);
#define PV(dd) (expf(-F[(size_t)(dd) * npix + p] - m) / sum)
    float pb = PV(l0) + PV(r0);
    if (num_buckets == 4) { float q = PV(l1) + PV(r1); pb = pb + q; }
#undef PV
    prob[p] = pb;
  }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
