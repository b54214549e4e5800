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
// Inverse by Gauss-Jordan elimination with partial pivoting; returns the 1-norm condition
// number ||A||_1 ||A^-1||_1 (inf when a pivot vanishes).
template <int N>
ARB_D double inv_small_cond(const double* a_in, double* out) {
  double a[N * N], x[N * N];
  double an = 0.;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double c = 0.;
#pragma unroll
    for (int i = 0; i < N; ++i) { a[i * N + j] = a_in[i * N + j]; x[i * N + j] = (i == j) ? 1. : 0.; c += fabs(a_in[i * N + j]); }
    an = fmax(an, c);
  }
  bool ok = true;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int piv = k;
    double best = fabs(a[k * N + k]);
#pragma unroll
    for (int i = k + 1; i < N; ++i)
      if (fabs(a[i * N + k]) > best) { best = fabs(a[i * N + k]); piv = i; }
    if (!(best > 0.)) ok = false;
#pragma unroll
    for (int i = k + 1; i < N; ++i)
      if (i == piv) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          double t = a[k * N + j]; a[k * N + j] = a[i * N + j]; a[i * N + j] = t;
          t = x[k * N + j]; x[k * N + j] = x[i * N + j]; x[i * N + j] = t;
        }
      }
    const double inv = 1. / a[k * N + k];
#pragma unroll
    for (int j = 0; j < N; ++j) { a[k * N + j] *= inv; x[k * N + j] *= inv; }
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (i != k) {
        const double f = a[i * N + k];
#pragma unroll
        for (int j = 0; j < N; ++j) { a[i * N + j] -= f * a[k * N + j]; x[i * N + j] -= f * x[k * N + j]; }
      }
  }
  double xn = 0.;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double c = 0.;
#pragma unroll
    for (int i = 0; i < N; ++i) { out[i * N + j] = x[i * N + j]; c += fabs(x[i * N + j]); }
    xn = fmax(xn, c);
  }
  const double cond = an * xn;
  return (ok && cond == cond) ? cond : 1e300;
}

template <int N>
ARB_NOINLINE void pinv_jacobi(const double* a, double* out);

template <int N>
ARB_D void pinv_small(const double* a, double* out) {
  if (N == 1) {
    out[0] = (a[0] != 0.) ? 1. / a[0] : 0.;
    return;
  }
  // numpy.linalg.pinv zeroes singular values below 1e-15 sigma_max.  cond_2 <= N cond_1, so a
  // block with cond_1 < 1e12 loses no singular value and its pseudo-inverse IS its inverse:
  // take the elimination result and leave the SVD to the (near-)singular blocks.
  if (inv_small_cond<N>(a, out) < 1e12) return;
  pinv_jacobi<N>(a, out);
}

// one-sided Jacobi SVD (out of line: only (near-)singular blocks get here)
template <int N>
ARB_NOINLINE void pinv_jacobi(const double* a, double* out) {
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
        if (ga != 0. && fabs(ga) > 4e-16 * sqrt(al * be)) {   // columns orthogonal to working precision
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

// ---- smallest real eigenvalue <= 0 of the sliding-friction matrix, structured ---------------
// The matrix of constraints.py:815-821 (with eps = (1,1,1), the value the reference fixes at
// constraints.py:423) is  B = [[Q + c J, al I], [ga I, Q]]  with J the all-ones 3x3 matrix.
// Its lower-left block commutes with everything, so
//     det(B - s I) = det( (Q + c J - s I)(Q - s I) - al ga I )  = det( s^2 I - s C1 + C0 ),
// a 3x3 quadratic matrix polynomial.  With s = -sigma t the wanted root (min real s <= 0,
// constraints.py:825-830) is the LARGEST real root t >= 0 of a monic sextic.  Real roots are
// isolated by derivative interlacing (between two consecutive real roots of p' the polynomial
// p is monotone), each bracket refined by safeguarded Newton, and the result polished by
// Newton on the 3x3 determinant itself (not on the expanded coefficients).  Control flow is
// data independent apart from iteration counts -- unlike a shifted QR iteration, lanes of a
// warp do not diverge.  eig_real_general6 above stays as the general path.
template <int D>
ARB_HD double poly_val(const double* c, double x) {
  double f = c[D];
#pragma unroll
  for (int k = D - 1; k >= 0; --k) f = f * x + c[k];
  return f;
}
// root of c in (lo, hi) given f(lo) = flo != 0 and a sign change on the bracket.
// Safeguarded Newton; stops when |f| is below its own rounding noise (running Horner bound)
// or the step is below one ulp -- the older "|dx| <= 4e-16 |x|" test alone let Newton
// wander inside the noise floor for the whole iteration budget.
template <int D>
ARB_HD double poly_refine(const double* c, double lo, double hi, double flo) {
  double x = 0.5 * (lo + hi);
  const bool lo_neg = flo < 0.;
  for (int it = 0; it < 64; ++it) {
    double f = c[D], df = 0., e = fabs(c[D]);
    const double ax = fabs(x);
#pragma unroll
    for (int k = D - 1; k >= 0; --k) { df = df * x + f; f = f * x + c[k]; e = e * ax + fabs(f); }
    if (fabs(f) <= 2.5e-16 * e) break;
    if ((f < 0.) == lo_neg) lo = x; else hi = x;
    double xn = x - f / df;
    if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
    if (xn == x || !(hi - lo > 0.)) break;
    const bool done = fabs(xn - x) <= 2.5e-16 * fabs(xn);
    x = xn;
    if (done) break;
  }
  return x;
}
// real roots in [0, T] of the degree-D polynomial c[0..D] (c[D] != 0), ascending; returns count
template <int D>
struct PolyRoots {
  static ARB_HD int run(const double* c, double T, double* roots) {
    double dc[D], crit[D];
#pragma unroll
    for (int k = 0; k < D; ++k) dc[k] = (k + 1) * c[k + 1];
    const int nc = PolyRoots<D - 1>::run(dc, T, crit);
    int n = 0;
    double lo = 0., flo = c[0];
    if (flo == 0.) roots[n++] = 0.;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if (i <= nc) {
        const double hi = (i < nc) ? crit[i] : T;
        const double fhi = poly_val<D>(c, hi);
        if ((flo < 0. && fhi > 0.) || (flo > 0. && fhi < 0.)) roots[n++] = poly_refine<D>(c, lo, hi, flo);
        else if (fhi == 0. && hi > lo) roots[n++] = hi;
        lo = hi;
        flo = fhi;
      }
    }
    return n;
  }
};
template <>
struct PolyRoots<1> {
  static ARB_HD int run(const double* c, double T, double* roots) {
    const double x = -c[0] / c[1];
    if (x >= 0. && x <= T) { roots[0] = x; return 1; }
    return 0;
  }
};

// Largest real root of the MONIC sextic p when it can be certified cheaply.
//  - start at the Laguerre-Samuelson bound  mean + sqrt(5) * std  of the roots (an upper bound
//    of every root when all six are real -- the case met in practice: the sliding matrices of
//    constraints.py:815-821 had all-real spectra in every sampled contact problem);
//  - Laguerre's iteration from the right of the largest root of a real-rooted polynomial
//    decreases monotonically onto it with cubic convergence (4-6 iterations, no bracketing,
//    the same trip count on every lane);
//  - certificate, valid whatever the roots are: if all Taylor coefficients of p at x* are
//    positive, p > 0 on (x*, inf), so x* is the largest real root.
// Returns 1 (root in *root), or 0 when the fast path does not apply (complex roots near the
// path, a multiple root, slow convergence): the caller then isolates the roots rigorously.
// out-of-line: the rigorous root isolation is the rare path, keep its code out of the hot loop
ARB_NOINLINE int poly6_roots_slow(const double* p, double T, double* roots) { return PolyRoots<6>::run(p, T, roots); }

#ifdef ARB_HOSTTEST_COUNTERS
static long arb_fastroot_hits = 0;   // host unit tests only: how often the fast path certified its root
static long arb_fastroot_fail[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // ... and why it did not (by exit), [7] = Laguerre iterations
#endif
#ifndef ARB_SIGMA_SUM
#define ARB_SIGMA_SUM 1
#endif
#ifndef ARB_LAGUERRE_FAST
#define ARB_LAGUERRE_FAST 1
#endif
#ifdef ARB_HOSTTEST_COUNTERS
#define ARB_FAIL(i) (++arb_fastroot_fail[i], 0)
#else
#define ARB_FAIL(i) 0
#endif
// Certificate, valid whatever the roots are: if the Taylor coefficients 1..5 of the monic sextic p at x
// are positive (repeated synthetic division), p' > 0 on (x, inf): a root x is then the largest real root.
ARB_HD bool poly6_certify(const double* p, double x) {
  double b[7];
#pragma unroll
  for (int k = 0; k < 6; ++k) b[k] = p[k];
  b[6] = 1.;
#pragma unroll
  for (int j = 0; j < 6; ++j)
#pragma unroll
    for (int k = 5; k >= j; --k) b[k] += x * b[k + 1];
  bool pos = true;
#pragma unroll
  for (int k = 1; k < 6; ++k) pos = pos && (b[k] > 0.);
  return pos;
}
// Sampling search of the largest root in [0, T] of a sextic that left the fast path (not real-rooted:
// the start bound sits where p' <= 0, or the variance of the roots is negative; 1 % of the sliding
// solves).  Each round evaluates p at 32 equidistant points of the bracket; the rightmost sample with
// p <= 0 gives the next bracket (p > 0 at its right end), 33 times narrower: four rounds to 1e-6 T.  A pair
// of roots between two samples is missed, so the result is only a candidate: the caller refines it
// (poly_refine) and runs poly6_certify; what fails the certificate goes to the rigorous isolation as
// before.  coop != 0 (device): the lanes of coop -- all of them here, with the same q: the sliding lanes
// of one contact visit helping ONE of them -- share the 32 samples of a round; coop == 0: the calling
// lane evaluates them itself.  Same sample points, same Horner evaluations, same decisions either way:
// the result does not depend on who computes it (bit-identical kernels and batch compositions).
// ~10 instructions per round and lane with a full warp where the isolation costs a few thousand.
// MEASURED SLOWER (ARB_SAMPLE_ROOT = 0 is the default): only a third of the polynomials that reach this
// point pass the certificate afterwards (host counters: 2 087 of 6 267) -- where p' <= 0 somewhere right
// of the largest root the Taylor coefficients there cannot all be positive -- so two thirds pay the search
// AND the isolation: gs 7.59 ms against 6.93 at 262 144 worlds, 1.92 against 1.82 at 32 768
// (profiles/ab_r04/r04n_*).  A certificate that admits those roots would have to bound p from below
// between the root and the Cauchy bound.
// Returns 1: bracket (lo, hi) with p(lo) <= 0 < p(hi); 0: p > 0 at every sample (no candidate).
ARB_HD int poly6_sample_bracket(const double* q, unsigned coop, double* lo_out, double* hi_out) {
  double T = 0.;
#pragma unroll
  for (int k = 0; k < 6; ++k) T = fmax(T, fabs(q[k]));
  T += 1.;
  if (!(T < 1e300)) return 0;
#ifdef __CUDA_ARCH__
  const unsigned lane = threadIdx.x & 31u;
  const int n = coop ? __popc(coop) : 1, rank = coop ? __popc(coop & ((1u << lane) - 1u)) : 0;
#endif
  double lo = 0., hi = T;
  bool have = !(q[0] > 0.);          // p(0) <= 0: (0, T) holds a root
  for (int round = 0; round < 16; ++round) {
    const double h = (hi - lo) * (1. / 33.);
    int jbest = -1;                   // the rightmost sample x_j = lo + h (j + 1), j = 0..31, with p <= 0
#ifdef __CUDA_ARCH__
    if (coop != 0u) {
      for (int j = rank; j < 32; j += n) {
        const double x = lo + h * (double)(j + 1);
        double f = 1.;
#pragma unroll
        for (int k = 5; k >= 0; --k) f = f * x + q[k];
        if (!(f > 0.)) jbest = j;
      }
      jbest = __reduce_max_sync(coop, jbest);
    } else
#endif
    {
      for (int j = 31; j >= 0; --j) {
        const double x = lo + h * (double)(j + 1);
        double f = 1.;
#pragma unroll
        for (int k = 5; k >= 0; --k) f = f * x + q[k];
        if (!(f > 0.)) { jbest = j; break; }
      }
    }
    if (jbest >= 0) {
      const double nlo = lo + h * (double)(jbest + 1);
      hi = (jbest + 1 < 32) ? lo + h * (double)(jbest + 2) : hi;
      lo = nlo;
      have = true;
    } else if (have) {
      hi = lo + h;                    // p(lo) <= 0 < p(lo + h)
    } else {
      return 0;
    }
    if (!(hi - lo > 1e-6 * T)) break;
  }
  *lo_out = lo;
  *hi_out = hi;
  return 1;
}
// root of the sextic in a bracket the failed fast iteration already holds (out of line: rare lanes)
ARB_NOINLINE double poly6_refine_bracket(const double* p, double lo, double hi, double flo) {
  return poly_refine<6>(p, lo, hi, flo);
}
// candidate of poly6_sample_bracket -> certified largest root (1), "no root t >= 0" as t = -1 (1), or 0
ARB_HD int poly6_from_samples(const double* p, int got, double lo, double hi, double* t) {
  if (got) {
    const double flo = poly_val<6>(p, lo);
    const double tc = (flo == 0.) ? lo : poly6_refine_bracket(p, lo, hi, flo);
    if (poly6_certify(p, tc)) { *t = tc; return 1; }
    return 0;
  }
  if (p[0] > 0. && p[1] > 0. && p[2] > 0. && p[3] > 0. && p[4] > 0. && p[5] > 0.) {
    *t = -1.;       // all coefficients positive: no root t >= 0
    return 1;
  }
  return 0;
}
#ifndef ARB_LAGUERRE_EARLY
#define ARB_LAGUERRE_EARLY 0     /* 1: stop on a step in the cubic regime without the confirming evaluation -- measured slower, see there */
#endif
#ifndef ARB_FASTROOT_RECOVER
#define ARB_FASTROOT_RECOVER 1   /* cheap recoveries of the fast path before the rigorous isolation (0: A/B builds) */
#endif
// Recovery (ARB_FASTROOT_RECOVER), validated by the same certificate, so a wrong guess only costs the
// rigorous path it would have taken anyway: when an iterate lands on the left of a root (p < 0; in
// practice the start bound itself, which only bounds the roots when all six are real), that iterate and
// the one before it -- or the Cauchy bound 1 + max |p_k| -- bracket a root: refine it there by safeguarded
// Newton (poly_refine) instead of isolating all the roots.  Host counters (profiles/fastroot_counters.py,
// 128 worlds x 200 steps, 638 724 sliding solves): 2.01 % of the solves left the fast path before (left of
// a root 1.03 %, p' <= 0 at an iterate 0.85 %, negative variance 0.13 %), 0.98 % now.  Also tried, no use:
// a Newton step where the discriminant is negative, a start (or restart) from the Cauchy bound where the
// variance is negative or p' <= 0 (those polynomials do need the isolation).  A warp runs the rigorous
// isolation (a few thousand instructions) whenever ONE of its lanes does -- every second visit of a warp
// whose 32 worlds all slide, the warps that decide the duration of a single wave (strong scaling).
ARB_HD int poly6_largest_root_fast(const double* p, double* root) {
  const double mean = -p[5] * (1. / 6.);
  const double var = (p[5] * p[5] - 2. * p[4]) * (1. / 6.) - mean * mean;
  if (!(var >= 0.)) return ARB_FAIL(0);
  double x = mean + 2.2360679774997898 * sqrt(var);
  x += 1e-12 * fabs(x) + 1e-300;
  bool conv = false;
#if ARB_FASTROOT_RECOVER
  double xprev = x;
  bool have_prev = false;
#endif
#if ARB_LAGUERRE_EARLY
  double aprev = 1e300;
#endif
  for (int it = 0; it < 12; ++it) {
    // p, p', p''/2 by one Horner pass, with the running rounding bound of p
    double f = 1., d1 = 0., d2 = 0., e = 1.;
    const double ax = fabs(x);
#pragma unroll
    for (int k = 5; k >= 0; --k) {
      d2 = d2 * x + d1;
      d1 = d1 * x + f;
      f = f * x + p[k];
      e = e * ax + fabs(f);
    }
#ifdef ARB_HOSTTEST_COUNTERS
    ++arb_fastroot_fail[7];
#endif
    if (fabs(f) <= 2.5e-16 * e) { conv = true; break; }
    if (!(f > 0.)) {                                      // fell to the left of a root: not real-rooted
#if ARB_FASTROOT_RECOVER
      if (f < 0.) {
        if (!have_prev) {      // the start bound itself is left of a root: the Cauchy bound is right of all
          double T = 0.;
#pragma unroll
          for (int k = 0; k < 6; ++k) T = fmax(T, fabs(p[k]));
          xprev = T + 1.;
          if (!(xprev < 1e300)) return ARB_FAIL(1);
        }
        x = poly6_refine_bracket(p, x, xprev, f);
#ifdef ARB_HOSTTEST_COUNTERS
        ++arb_fastroot_fail[5];
#endif
        conv = true;
        break;
      }
#endif
      return ARB_FAIL(1);
    }
    // Laguerre step n / (G + sqrt((n-1)(n H - G^2))), G = p'/p, H = G^2 - p''/p, n = 6, with the
    // common factor 1/p taken out (d2 holds p''/2): one square root and one division
    const double disc = 5. * (5. * d1 * d1 - 12. * f * d2);
    if (!(disc >= 0.) || !(d1 > 0.)) return ARB_FAIL(2);
#if defined(__CUDA_ARCH__) && ARB_LAGUERRE_FAST
    // sqrt through the reciprocal square root and a reciprocal instead of a division: the step is
    // self-correcting (the iteration stops on |p| against its own rounding bound), a few ulps in it
    // cost nothing, and the two IEEE-exact routines were a third of the iteration's instructions.
    // Both from the hardware's 20-bit seeds (MUFU.RSQ64H / RCP64H) and two Newton steps: the operands
    // are positive and normal here (disc > 0, d1 > 0), no special cases to handle.
    double sq = 0.;
    if (disc > 0.) {
      double r;
      asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(disc));
      const double h = -0.5 * disc;
      r = r * fma(h, r * r, 1.5);
      r = r * fma(h, r * r, 1.5);
      sq = disc * r;
    }
    const double den = d1 + sq;
    double rc;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(den));
    rc = fma(rc, fma(-den, rc, 1.), rc);
    rc = fma(rc, fma(-den, rc, 1.), rc);
    const double a = 6. * f * rc;
#else
    const double a = 6. * f / (d1 + sqrt(disc));
#endif
    const double xn = x - a;
    if (!(a > 2.5e-16 * fabs(x))) { conv = true; break; }
#if ARB_FASTROOT_RECOVER
    xprev = x; have_prev = true;
#endif
    x = xn;
#if ARB_LAGUERRE_EARLY
    // Cubic regime: a step below 1e-6 |x| that is a thousand times shorter than the one before leaves an
    // error of the order of its cube -- below rounding; take it without the evaluation that would only
    // confirm it (one Horner pass of four chains per solve; the warp runs as many iterations as its
    // slowest lane: 8 instead of the mean 5.3 in a warp of 32 sliding lanes).  A multiple root converges
    // linearly, fails the ratio test and iterates to the rounding floor as before.  The root is polished
    // on det M(t) afterwards in any case.
    // MEASURED SLOWER (default off): 4.55 instead of 5.34 iterations per solve on the host, but gs 6.72 ms
    // against 6.65 at 262 144 worlds and 1.549 = 1.549 at 32 768 (profiles/ab_r04/r04r_*): the warp still runs
    // to its slowest lane, and the two extra compares sit in every iteration.
    if (a < 1e-6 * fabs(xn) && a < 1e-3 * aprev) { conv = true; break; }
    aprev = a;
#endif
  }
  if (!conv) return ARB_FAIL(3);
  if (!poly6_certify(p, x)) return ARB_FAIL(4);
  *root = x;
#ifdef ARB_HOSTTEST_COUNTERS
  ++arb_fastroot_hits;
#endif
  return 1;
}

// p(t) > 0 for every t >= 0, proven by a march of Taylor expansions (true = proven; false = not proven).
// At the expansion point a (p(a) = c_0 > 0):  p(a + y) >= c_0 + sum_{k >= 1} min(c_k, 0) y^k  for y >= 0, and
// the right-hand side decreases in y: if it is positive at h, p has no root in [a, a + h].  Move there
// (Taylor shift, 15 multiply-adds), double h, repeat; once c_1..c_5 >= 0 the polynomial only grows.
// This is what the polynomials that leave the fast path need: on 6 267 of them (host run of 128 falling
// humanoids x 200 steps, 638 724 sliding solves) 6 176 have NO root t >= 0 -- the contact's friction cone
// admits no sliding solution and the reference clamps s to -1e10 (constraints.py:827-830) -- which the
// fast path cannot certify and the rigorous isolation establishes by finding the real roots of all five
// derivatives first (~2 000 instructions, for the whole warp whenever one lane needs it).  The march
// proves all 6 176 in 1.4 steps on average (19 at most) and none of the 91 others.  The outcome "no root"
// is discrete: results are bit-identical to the isolation's.
#ifndef ARB_POSITIVE_MARCH
#define ARB_POSITIVE_MARCH 1      /* 0: A/B builds */
#endif
ARB_NOINLINE bool poly6_positive_on_halfline(const double* p) {
  double c[7];
#pragma unroll
  for (int k = 0; k < 6; ++k) c[k] = p[k];
  c[6] = 1.;
  if (!(c[0] > 0.)) return false;
  double h = 1.;                                 // (the coefficients are scaled to O(1))
  for (int step = 0; step < 24; ++step) {
    bool grows = true;
#pragma unroll
    for (int k = 1; k < 6; ++k) grows = grows && (c[k] >= 0.);
    if (grows) return true;
    double n[6];
#pragma unroll
    for (int k = 1; k < 6; ++k) n[k] = fmin(c[k], 0.);
    bool ok = false;
    for (int tr = 0; tr < 30; ++tr) {
      const double lb = c[0] + h * (n[1] + h * (n[2] + h * (n[3] + h * (n[4] + h * n[5]))));
      if (lb > 0.25 * c[0]) { ok = true; break; }
      h *= 0.5;
    }
    if (!ok) return false;
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
      for (int k = 5; k >= j; --k) c[k] += h * c[k + 1];
    if (!(c[0] > 0.)) return false;
    h *= 2.;
  }
  return false;
}
#ifndef ARB_SAMPLE_ROOT
#define ARB_SAMPLE_ROOT 0     /* 1: sampling search (poly6_sample_bracket) before the rigorous isolation -- measured slower, see there */
#endif
#ifndef ARB_POLISH_ITERS
#define ARB_POLISH_ITERS 1   /* one Newton step on det M(t) brings the root of the expanded sextic to ~1e-14 of LAPACK's */
#endif
// A: 4x4 contact admittance block, alpha: the vector of constraints.py:807-808, mu: friction.
// Returns false if the structured path does not apply (caller falls back to the general
// eigenvalue routine); else *found tells whether a real eigenvalue <= 0 exists and *s_out is it.
// coop (device): the lanes of the warp that are in this call together (0: none known) -- they help the
// lanes whose polynomial leaves the fast path (poly6_coop_bracket); every lane of coop must get here,
// so nothing returns before that point.
ARB_HD bool sliding_root_structured(const double* A, const double* alpha, double mu, double* s_out, bool* found,
                                    unsigned coop = 0u) {
  const double Yc[3] = {A[3], A[7], A[11]};
  const double yn = A[15];
  // one reciprocal of y_n and one of a instead of five divisions (the reference divides: mu/y_n*alpha_n
  // ...; the operands of B move by an ulp, its admissible root by that much times its conditioning --
  // far inside the 1e-14 at which the root agrees with LAPACK's anyway)
  const double iyn = 1. / yn;
  const double mu_yn = mu * iyn, an_yn = alpha[3] * iyn;
  const double a = mu_yn * alpha[3];
  double beta[3], bb[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    beta[i] = alpha[i] - an_yn * Yc[i];
    bb[i] = mu_yn * Yc[i];
  }
  const double kappa = (Yc[0] * Yc[0] + Yc[1] * Yc[1] + Yc[2] * Yc[2]) * iyn;
  const double ia = 1. / a;
  const double cc = 2. * ia * (beta[0] * bb[0] + beta[1] * bb[1] + beta[2] * bb[2]);
  const double al = -((beta[0] * beta[0] + beta[1] * beta[1] + beta[2] * beta[2]) * (ia * ia));
  const double ga = (bb[0] * bb[0] + bb[1] * bb[1] + bb[2] * bb[2]) - 1.;
  const double delta = al * ga;
  double Q[9];
#if ARB_SIGMA_SUM
  // scale of the matrix polynomial: any sigma within a small factor of max(|Q_ij| + |c|, sqrt|delta|)
  // serves (it only keeps the sextic's coefficients O(1)); the sum of the magnitudes costs ten additions
  // where nine fp64 maxima cost sixty instructions
  double sigma = sqrt(fabs(delta)) + fabs(cc);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      Q[3 * i + j] = A[4 * i + j] - kappa;
      sigma += fabs(Q[3 * i + j]);
    }
#else
  double sigma = sqrt(fabs(delta));
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      Q[3 * i + j] = A[4 * i + j] - kappa;
      sigma = fmax(sigma, fabs(Q[3 * i + j]) + fabs(cc));
    }
#endif
  bool bad = !(sigma > 0.) || !(sigma < 1e300);
  if (bad && coop == 0u) return false;
  // scaled coefficients of  M(t) = t^2 I + t C1 + C0,  s = -sigma t
  const double is = 1. / sigma;
  double C1[9], C0[9];
  const double cs[3] = {Q[0] + Q[3] + Q[6], Q[1] + Q[4] + Q[7], Q[2] + Q[5] + Q[8]};  // J Q rows
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      C1[3 * i + j] = (2. * Q[3 * i + j] + cc) * is;
      double t = Q[3 * i] * Q[j] + Q[3 * i + 1] * Q[3 + j] + Q[3 * i + 2] * Q[6 + j] + cc * cs[j];
      if (i == j) t -= delta;
      C0[3 * i + j] = t * is * is;
    }
  // sextic p(t) = det M(t): entries m_ij(t) = d_ij t^2 + C1_ij t + C0_ij
  double p[7] = {0., 0., 0., 0., 0., 0., 0.};
  {
    // 2x2 minors of rows 1,2 as quartics, then expand along row 0
#define ARB_QUAD(i, j, q) { (q)[0] = C0[3 * (i) + (j)]; (q)[1] = C1[3 * (i) + (j)]; (q)[2] = ((i) == (j)) ? 1. : 0.; }
    const int cols[3][2] = {{1, 2}, {0, 2}, {0, 1}};
#pragma unroll
    for (int e = 0; e < 3; ++e) {
      double u[3], v[3], x[3], y[3], mnr[5] = {0., 0., 0., 0., 0.};
      ARB_QUAD(1, cols[e][0], u); ARB_QUAD(2, cols[e][1], v); ARB_QUAD(1, cols[e][1], x); ARB_QUAD(2, cols[e][0], y);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) mnr[i + j] += u[i] * v[j] - x[i] * y[j];
      double r0[3];
      ARB_QUAD(0, e, r0);
      const double sgn = (e == 1) ? -1. : 1.;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) p[i + j] += sgn * r0[i] * mnr[j];
    }
#undef ARB_QUAD
  }
  bad = bad || !(fabs(p[6] - 1.) < 1e-9);
  if (bad && coop == 0u) return false;
  p[6] = 1.;
  double t = 0.;
  int ok = bad ? 0 : poly6_largest_root_fast(p, &t);
#if ARB_SAMPLE_ROOT
#ifdef __CUDA_ARCH__
  if (coop != 0u) {
    unsigned fm = __ballot_sync(coop, !bad && !ok);
    while (fm != 0u) {
      const int src = __ffs(fm) - 1;
      fm &= fm - 1u;
      double q[6], lo = 0., hi = 0.;
#pragma unroll
      for (int k = 0; k < 6; ++k) q[k] = __shfl_sync(coop, p[k], src);
      const int got = poly6_sample_bracket(q, coop, &lo, &hi);
      if ((int)(threadIdx.x & 31u) == src) ok = poly6_from_samples(p, got, lo, hi, &t);
    }
    if (bad) return false;
  } else
#endif
  if (!ok) {
    double lo = 0., hi = 0.;
    const int got = poly6_sample_bracket(p, 0u, &lo, &hi);
    ok = poly6_from_samples(p, got, lo, hi, &t);
#ifdef ARB_HOSTTEST_COUNTERS
    if (ok) { ++arb_fastroot_fail[6]; }
#endif
  }
#else
  if (bad) return false;
#endif
#if ARB_POSITIVE_MARCH
  if (!ok && poly6_positive_on_halfline(p)) {
    ok = 1;
    t = -1.;        // no root t >= 0
#ifdef ARB_HOSTTEST_COUNTERS
    ++arb_fastroot_fail[6];
#endif
  }
#endif
  if (ok) {
    if (t < 0.) { *found = false; *s_out = 0.; return true; }   // every real eigenvalue is > 0
  } else {
    // (Cauchy bound of the roots, needed by the rigorous isolation only: the fast path rejects
    // non-finite coefficients by itself -- its variance test fails on a NaN)
    double T = 0.;
#pragma unroll
    for (int k = 0; k < 6; ++k) T = fmax(T, fabs(p[k]));
    T += 1.;
    if (!(T < 1e300)) return false;
    double roots[6];
    const int nr = poly6_roots_slow(p, T, roots);
    if (nr == 0) { *found = false; *s_out = 0.; return true; }
    t = roots[nr - 1];
  }
  // polish on det M(t):  f' = tr(adj(M) (2 t I + C1))
  for (int it = 0; it < ARB_POLISH_ITERS; ++it) {
    double M[9], dM[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double id = (i % 4 == 0) ? 1. : 0.;
      M[i] = (id * t + C1[i]) * t + C0[i];
      dM[i] = 2. * id * t + C1[i];
    }
    const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
    const double c10 = M[2] * M[7] - M[1] * M[8], c11 = M[0] * M[8] - M[2] * M[6], c12 = M[1] * M[6] - M[0] * M[7];
    const double c20 = M[1] * M[5] - M[2] * M[4], c21 = M[2] * M[3] - M[0] * M[5], c22 = M[0] * M[4] - M[1] * M[3];
    const double f = M[0] * c00 + M[1] * c01 + M[2] * c02;
    // adj(M)_ji = cofactor_ij ;  tr(adj(M) dM) = sum_ij cof_ij dM_ij
    const double df = c00 * dM[0] + c01 * dM[1] + c02 * dM[2] + c10 * dM[3] + c11 * dM[4] + c12 * dM[5] +
                      c20 * dM[6] + c21 * dM[7] + c22 * dM[8];
    const double tn = t - f / df;
    if (!(fabs(tn - t) <= 1e-6 * (fabs(t) + 1e-6)) || tn < 0.) break;   // not a simple, well-isolated root
    t = tn;
  }
  *found = true;
  *s_out = -sigma * t;
  return true;
}
