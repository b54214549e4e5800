// Rigid-motion device math (fp64): SE(3) frames, adjoint actions, twist
// adjacency and exponential.  Device-side counterpart of the reference's
// arboris/homogeneousmatrix.py (inv :254-275, adjoint :277-319, iadjoint :321-325,
// zaligned :201-232), arboris/twistvector.py (adjacency :10-33, exp :35-70) and
// arboris/rigidmotion.py (:36-75).  Nothing here is copied from the reference:
// frames are kept as (R, p) instead of 4x4 arrays and adjoints are never
// materialised as 6x6 matrices -- they are applied as  [R w ; p x (R w) + R v].
#pragma once
#include <math.h>

// The per-world routines are plain scalar code; they compile for the device
// (product) and, for CPU-only unit tests of the arithmetic, for the host
// (tests/hosttest builds them with g++; the package never loads that build).
#ifdef __CUDACC__
#define ARB_HD __host__ __device__ __forceinline__
#define ARB_NOINLINE static __host__ __device__ __noinline__
#else
#define ARB_HD inline
#define ARB_NOINLINE inline
#endif
#define ARB_D ARB_HD

struct Se3 {
  double R[9];  // row-major
  double p[3];
};

// 6x6 matrices of the form [[A,0],[B,A]] (adjoints, adjacency matrices and
// their products) are stored as the pair of 3x3 blocks.
struct Blk6 {
  double A[9];
  double B[9];
};

ARB_HD void m3_mul(const double* a, const double* b, double* c) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
ARB_HD void m3_mulv(const double* a, const double* x, double* y) {
#pragma unroll
  for (int i = 0; i < 3; ++i) y[i] = a[3 * i] * x[0] + a[3 * i + 1] * x[1] + a[3 * i + 2] * x[2];
}
ARB_HD void m3t_mulv(const double* a, const double* x, double* y) {
#pragma unroll
  for (int i = 0; i < 3; ++i) y[i] = a[i] * x[0] + a[3 + i] * x[1] + a[6 + i] * x[2];
}
ARB_HD void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
ARB_HD void skew3(const double* p, double* s) {
  s[0] = 0.;    s[1] = -p[2]; s[2] = p[1];
  s[3] = p[2];  s[4] = 0.;    s[5] = -p[0];
  s[6] = -p[1]; s[7] = p[0];  s[8] = 0.;
}

ARB_HD void se3_identity(Se3& h) {
#pragma unroll
  for (int i = 0; i < 9; ++i) h.R[i] = (i % 4 == 0) ? 1. : 0.;
  h.p[0] = h.p[1] = h.p[2] = 0.;
}
// c = a * b
ARB_HD void se3_mul(const Se3& a, const Se3& b, Se3& c) {
  m3_mul(a.R, b.R, c.R);
  double t[3];
  m3_mulv(a.R, b.p, t);
  c.p[0] = t[0] + a.p[0]; c.p[1] = t[1] + a.p[1]; c.p[2] = t[2] + a.p[2];
}
// b = a^-1 = [R^T, -R^T p]   (homogeneousmatrix.py:272-275)
ARB_HD void se3_inv(const Se3& a, Se3& b) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) b.R[3 * i + j] = a.R[3 * j + i];
  double t[3];
  m3t_mulv(a.R, a.p, t);
  b.p[0] = -t[0]; b.p[1] = -t[1]; b.p[2] = -t[2];
}
// load/store from a row-major 4x4 (16 doubles)
ARB_HD void se3_from16(const double* m, Se3& h) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) h.R[3 * i + j] = m[4 * i + j];
    h.p[i] = m[4 * i + 3];
  }
}

// frame j of a table of constant frames stored as 12 doubles (R row-major, p)
ARB_HD void load_se3_const(const double* tab, int j, Se3& h) {
#pragma unroll
  for (int i = 0; i < 9; ++i) h.R[i] = tab[12 * j + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) h.p[i] = tab[12 * j + 9 + i];
}

// y = Ad(H) x,  Ad = [[R,0],[p^ R, R]]   (homogeneousmatrix.py:308-319)
ARB_HD void ad_apply(const Se3& h, const double* x, double* y) {
  double rw[3], rv[3], c[3];
  m3_mulv(h.R, x, rw);
  m3_mulv(h.R, x + 3, rv);
  cross3(h.p, rw, c);
  y[0] = rw[0]; y[1] = rw[1]; y[2] = rw[2];
  y[3] = c[0] + rv[0]; y[4] = c[1] + rv[1]; y[5] = c[2] + rv[2];
}
// y = Ad(H^-1) x = [R^T w ; R^T (v - p x w)]   (homogeneousmatrix.py:321-325)
ARB_HD void iad_apply(const Se3& h, const double* x, double* y) {
  double c[3], t[3];
  cross3(h.p, x, c);
  t[0] = x[3] - c[0]; t[1] = x[4] - c[1]; t[2] = x[5] - c[2];
  m3t_mulv(h.R, x, y);
  m3t_mulv(h.R, t, y + 3);
}
// Adjoint of H as block pair: A = R, B = p^ R
ARB_HD void blk_from_se3(const Se3& h, Blk6& m) {
  double s[9];
  skew3(h.p, s);
#pragma unroll
  for (int i = 0; i < 9; ++i) m.A[i] = h.R[i];
  m3_mul(s, h.R, m.B);
}
// adjacency(T) = [[w^,0],[v^,w^]]   (twistvector.py:26-33)
ARB_HD void blk_adjacency(const double* tw, Blk6& m) {
  skew3(tw, m.A);
  skew3(tw + 3, m.B);
}
// c = a * b for [[A,0],[B,A]] matrices: (A1 A2, B1 A2 + A1 B2)
ARB_HD void blk_mul(const Blk6& a, const Blk6& b, Blk6& c) {
  double t1[9], t2[9];
  m3_mul(a.A, b.A, c.A);
  m3_mul(a.B, b.A, t1);
  m3_mul(a.A, b.B, t2);
#pragma unroll
  for (int i = 0; i < 9; ++i) c.B[i] = t1[i] + t2[i];
}
// y = [[A,0],[B,A]] x
ARB_HD void blk_apply(const Blk6& m, const double* x, double* y) {
  double t[3];
  m3_mulv(m.A, x, y);
  m3_mulv(m.B, x, y + 3);
  m3_mulv(m.A, x + 3, t);
  y[3] += t[0]; y[4] += t[1]; y[5] += t[2];
}

// SE(3) exponential of a twist (twistvector.py:49-70): Rodrigues with the
// series switch at |w| < 1e-3.
ARB_HD void se3_exp(const double* tw, Se3& h) {
  const double* w = tw;
  const double* v = tw + 3;
  double t = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double cc, sc, dsc;
  if (t >= 0.001) {
    double s, c;
    sincos(t, &s, &c);
    cc = (1. - c) / (t * t);
    sc = s / t;
    dsc = (t - s) / (t * t * t);
  } else {
    cc = 0.5;
    sc = 1. - t * t / 6.;
    dsc = 1. / 6.;
  }
  double wx[9], wx2[9];
  skew3(w, wx);
  m3_mul(wx, wx, wx2);
#pragma unroll
  for (int i = 0; i < 9; ++i) h.R[i] = ((i % 4 == 0) ? 1. : 0.) + sc * wx[i] + cc * wx2[i];
  // p = (sc I + cc w^ + dsc w w^T) v
  double wv = w[0] * v[0] + w[1] * v[1] + w[2] * v[2];
  double c3[3];
  cross3(w, v, c3);
#pragma unroll
  for (int i = 0; i < 3; ++i) h.p[i] = sc * v[i] + cc * c3[i] + dsc * w[i] * wv;
}
