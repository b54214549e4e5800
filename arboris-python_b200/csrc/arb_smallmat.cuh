// Small dense fp64 routines run by ONE thread on private arrays: the LAPACK
// calls the reference makes on 1x1..6x6 operands inside the Gauss-Seidel solve:
//   numpy.linalg.pinv   (constraints.py:79,83,235,795)  -> pinv_small  (one-sided Jacobi SVD,
//                                                          cutoff 1e-15*sigma_max like numpy)
//   numpy.linalg.eigvals (constraints.py:825)           -> eig_real_hess (balance, Hessenberg,
//                                                          shifted double-step QR, EISPACK hqr scheme)
//   numpy.linalg.solve  (constraints.py:834)            -> solve_small (LU, partial pivoting)
#pragma once
#include <math.h>
#include "arb_math.cuh"

// ---- Moore-Penrose pseudo-inverse of an n x n matrix, n <= 4 (row-major) -----
template <int N>
ARB_D void pinv_small(const double* a, double* out) {
  if (N == 1) {
    out[0] = (a[0] != 0.) ? 1. / a[0] : 0.;
    return;
  }
  double U[N * N], V[N * N];
#pragma unroll
  for (int i = 0; i < N * N; ++i) { U[i] = a[i]; V[i] = ((i % (N + 1)) == 0) ? 1. : 0.; }
  for (int sweep = 0; sweep < 40; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int p = 0; p < N - 1; ++p) {
#pragma unroll
      for (int q = p + 1; q < N; ++q) {
        double al = 0., be = 0., ga = 0.;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          al += U[i * N + p] * U[i * N + p];
          be += U[i * N + q] * U[i * N + q];
          ga += U[i * N + p] * U[i * N + q];
        }
        if (ga != 0. && fabs(ga) > 1e-17 * sqrt(al * be)) {
          rotated = true;
          double zeta = (be - al) / (2. * ga);
          double t = copysign(1., zeta) / (fabs(zeta) + sqrt(1. + zeta * zeta));
          double c = 1. / sqrt(1. + t * t), s = c * t;
#pragma unroll
          for (int i = 0; i < N; ++i) {
            double up = U[i * N + p], uq = U[i * N + q];
            U[i * N + p] = c * up - s * uq;
            U[i * N + q] = s * up + c * uq;
            double vp = V[i * N + p], vq = V[i * N + q];
            V[i * N + p] = c * vp - s * vq;
            V[i * N + q] = s * vp + c * vq;
          }
        }
      }
    }
    if (!rotated) break;
  }
  double sig2[N], smax2 = 0.;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double t = 0.;
#pragma unroll
    for (int i = 0; i < N; ++i) t += U[i * N + j] * U[i * N + j];
    sig2[j] = t;
    smax2 = fmax(smax2, t);
  }
  // keep sigma_j > 1e-15 * sigma_max  (numpy.linalg.pinv default rcond)
  double cut2 = 1e-30 * smax2;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int k = 0; k < N; ++k) {
      double t = 0.;
#pragma unroll
      for (int j = 0; j < N; ++j)
        if (sig2[j] > cut2) t += V[i * N + j] * U[k * N + j] / sig2[j];
      out[i * N + k] = t;
    }
}

// ---- x = A^-1 b, N <= 4, partial pivoting; returns false if singular ------------
template <int N>
ARB_D bool solve_small(const double* a_in, const double* b_in, double* x) {
  double a[N * N], b[N];
#pragma unroll
  for (int i = 0; i < N * N; ++i) a[i] = a_in[i];
#pragma unroll
  for (int i = 0; i < N; ++i) b[i] = b_in[i];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int piv = k;
    double best = fabs(a[k * N + k]);
#pragma unroll
    for (int i = k + 1; i < N; ++i)
      if (fabs(a[i * N + k]) > best) { best = fabs(a[i * N + k]); piv = i; }
    if (best == 0.) ok = false;
#pragma unroll
    for (int i = k + 1; i < N; ++i)
      if (i == piv) {
#pragma unroll
        for (int j = 0; j < N; ++j) { double t = a[k * N + j]; a[k * N + j] = a[i * N + j]; a[i * N + j] = t; }
        double t = b[k]; b[k] = b[i]; b[i] = t;
      }
    double inv = 1. / a[k * N + k];
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      double f = a[i * N + k] * inv;
#pragma unroll
      for (int j = k + 1; j < N; ++j) a[i * N + j] -= f * a[k * N + j];
      b[i] -= f * b[k];
    }
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double t = b[i];
#pragma unroll
    for (int j = i + 1; j < N; ++j) t -= a[i * N + j] * x[j];
    x[i] = t / a[i * N + i];
  }
  return ok;
}

// ---- eigenvalues of a real general 6x6 matrix -------------------------------------
// a is destroyed.  wr/wi receive real and imaginary parts.  Returns false if the
// QR iteration did not converge (30 iterations per eigenvalue, as EISPACK).
#define EIGN 6
ARB_NOINLINE bool eig_real_general6(double* a, double* wr, double* wi) {
  const int n = EIGN;
#define A_(i, j) a[(i) * EIGN + (j)]
  // balance (radix 2 scaling, no permutation)
  {
    const double RADIX = 2.0, sqrdx = RADIX * RADIX;
    int last = 0;
    int guard = 0;
    while (last == 0 && guard++ < 64) {
      last = 1;
      for (int i = 0; i < n; ++i) {
        double r = 0., c = 0.;
        for (int j = 0; j < n; ++j)
          if (j != i) { c += fabs(A_(j, i)); r += fabs(A_(i, j)); }
        if (c != 0. && r != 0.) {
          double g = r / RADIX, f = 1., s = c + r;
          while (c < g) { f *= RADIX; c *= sqrdx; }
          g = r * RADIX;
          while (c > g) { f /= RADIX; c /= sqrdx; }
          if ((c + r) / f < 0.95 * s) {
            last = 0;
            g = 1. / f;
            for (int j = 0; j < n; ++j) A_(i, j) *= g;
            for (int j = 0; j < n; ++j) A_(j, i) *= f;
          }
        }
      }
    }
  }
  // reduction to upper Hessenberg form by stabilised elementary transformations
  for (int m = 1; m < n - 1; ++m) {
    double x = 0.;
    int i = m;
    for (int j = m; j < n; ++j)
      if (fabs(A_(j, m - 1)) > fabs(x)) { x = A_(j, m - 1); i = j; }
    if (i != m) {
      for (int j = m - 1; j < n; ++j) { double t = A_(i, j); A_(i, j) = A_(m, j); A_(m, j) = t; }
      for (int j = 0; j < n; ++j) { double t = A_(j, i); A_(j, i) = A_(j, m); A_(j, m) = t; }
    }
    if (x != 0.) {
      for (int ii = m + 1; ii < n; ++ii) {
        double y = A_(ii, m - 1);
        if (y != 0.) {
          y /= x;
          A_(ii, m - 1) = y;
          for (int j = m; j < n; ++j) A_(ii, j) -= y * A_(m, j);
          for (int j = 0; j < n; ++j) A_(j, m) += y * A_(j, ii);
        }
      }
    }
  }
  for (int i = 2; i < n; ++i)
    for (int j = 0; j < i - 1; ++j) A_(i, j) = 0.;
  // shifted double-step QR on the Hessenberg matrix
  double anorm = 0.;
  for (int i = 0; i < n; ++i)
    for (int j = (i > 0 ? i - 1 : 0); j < n; ++j) anorm += fabs(A_(i, j));
  int nn = n - 1;
  double t = 0.;
  bool ok = true;
  double p = 0., q = 0., r = 0., s, x, y, z, w;
  while (nn >= 0) {
    int its = 0, l;
    do {
      for (l = nn; l >= 1; --l) {
        s = fabs(A_(l - 1, l - 1)) + fabs(A_(l, l));
        if (s == 0.) s = anorm;
        if (fabs(A_(l, l - 1)) + s == s) { A_(l, l - 1) = 0.; break; }
      }
      x = A_(nn, nn);
      if (l == nn) {  // one real root
        wr[nn] = x + t;
        wi[nn--] = 0.;
      } else {
        y = A_(nn - 1, nn - 1);
        w = A_(nn, nn - 1) * A_(nn - 1, nn);
        if (l == nn - 1) {  // a pair
          p = 0.5 * (y - x);
          q = p * p + w;
          z = sqrt(fabs(q));
          x += t;
          if (q >= 0.) {
            z = p + copysign(z, p);
            wr[nn - 1] = wr[nn] = x + z;
            if (z != 0.) wr[nn] = x - w / z;
            wi[nn - 1] = wi[nn] = 0.;
          } else {
            wr[nn - 1] = wr[nn] = x + p;
            wi[nn - 1] = -(wi[nn] = z);
          }
          nn -= 2;
        } else {
          if (its == 30) {  // give up on the remaining block
            ok = false;
            for (int i = 0; i <= nn; ++i) { wr[i] = 0.; wi[i] = 1.; }
            nn = -1;
            break;
          }
          if (its == 10 || its == 20) {  // exceptional shift
            t += x;
            for (int i = 0; i <= nn; ++i) A_(i, i) -= x;
            s = fabs(A_(nn, nn - 1)) + fabs(A_(nn - 1, nn - 2));
            y = x = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          int m;
          for (m = nn - 2; m >= l; --m) {
            z = A_(m, m);
            r = x - z;
            s = y - z;
            p = (r * s - w) / A_(m + 1, m) + A_(m, m + 1);
            q = A_(m + 1, m + 1) - z - r - s;
            r = A_(m + 2, m + 1);
            s = fabs(p) + fabs(q) + fabs(r);
            p /= s; q /= s; r /= s;
            if (m == l) break;
            double u = fabs(A_(m, m - 1)) * (fabs(q) + fabs(r));
            double v = fabs(p) * (fabs(A_(m - 1, m - 1)) + fabs(z) + fabs(A_(m + 1, m + 1)));
            if (u + v == v) break;
          }
          for (int i = m + 2; i <= nn; ++i) {
            A_(i, i - 2) = 0.;
            if (i != m + 2) A_(i, i - 3) = 0.;
          }
          for (int k = m; k <= nn - 1; ++k) {
            if (k != m) {
              p = A_(k, k - 1);
              q = A_(k + 1, k - 1);
              r = 0.;
              if (k != nn - 1) r = A_(k + 2, k - 1);
              if ((x = fabs(p) + fabs(q) + fabs(r)) != 0.) { p /= x; q /= x; r /= x; }
            }
            s = copysign(sqrt(p * p + q * q + r * r), p);
            if (s != 0.) {
              if (k == m) {
                if (l != m) A_(k, k - 1) = -A_(k, k - 1);
              } else {
                A_(k, k - 1) = -s * x;
              }
              p += s;
              x = p / s; y = q / s; z = r / s;
              q /= p; r /= p;
              for (int j = k; j <= nn; ++j) {
                p = A_(k, j) + q * A_(k + 1, j);
                if (k != nn - 1) { p += r * A_(k + 2, j); A_(k + 2, j) -= p * z; }
                A_(k + 1, j) -= p * y;
                A_(k, j) -= p * x;
              }
              int mmin = nn < k + 3 ? nn : k + 3;
              for (int i = l; i <= mmin; ++i) {
                p = x * A_(i, k) + y * A_(i, k + 1);
                if (k != nn - 1) { p += z * A_(i, k + 2); A_(i, k + 2) -= p * r; }
                A_(i, k + 1) -= p * q;
                A_(i, k) -= p;
              }
            }
          }
        }
      }
    } while (nn >= 0 && l < nn - 1);
  }
#undef A_
  return ok;
}
