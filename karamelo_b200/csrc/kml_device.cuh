// kml_device.cuh - device-side building blocks of the MPM step: shape functions,
// per-axis stencil weights, 3x3 helpers and the constitutive functors.
// Reference semantics are cited per function (paths relative to the reference tree).
#pragma once
#include "../../include/kml.h"
#include <cuda_runtime.h>
#include <math.h>

#define KML_SQRT_3_OVER_2 1.224744871                                  /* src/solid.cpp:39 (truncated in the reference) */
#define KML_FOUR_THIRD 1.333333333333333333333333333333333333333      /* src/solid.cpp:40 */

namespace kml {

// ------------------------------------------------------------------------------------------
// Shape functions, src/basis_functions.h:21-241.  Value and derivative are evaluated
// together; the derivative is forced to 0 where the value is 0 because the reference
// drops a (particle,node) pair whenever its weight is 0 (src/ulmpm.cpp:265, src/tlmpm.cpp:282).
// ------------------------------------------------------------------------------------------
template <int SHAPE> struct Basis;

// Neighbour membership is decided by the reference's arithmetic.  The reference drops a (particle, node) pair whose weight is exactly 0
// (src/ulmpm.cpp:252-263, src/tlmpm.cpp:282) and its weights are Horner forms with separately rounded products and sums (a stock x86-64
// build has no FMA contraction).  Near the end of a spline's support the true weight is far below one ulp of the constant term, so
// whether the computed weight IS zero depends on that operation sequence, while the gradient there is not small at all.  The kernels
// evaluate the polynomials with FMAs; for a tiny result they replay the reference's sequence, which makes the weight the reference's bit
// for bit in exactly the regime where its zero-ness matters (tests/test_weight_zero_skip.py).
__device__ __forceinline__ bool weight_tiny(double w) { return (__double2hiint(w) & 0x7fffffff) < 0x3cd00000; } // |w| < 2^-50
__device__ __forceinline__ double horner3_unfused(double c3, double c2, double c1, double c0, double r) { // ((c3 r + c2) r + c1) r + c0
  return __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c3, r), c2), r), c1), r), c0);
}
__device__ __forceinline__ double horner2_unfused(double c2, double c1, double c0, double r) { // (c2 r + c1) r + c0
  return __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c2, r), c1), r), c0);
}

template <> struct Basis<KML_SHAPE_LINEAR> {
  static constexpr int SPAN = 2;
  __device__ __forceinline__ static void eval(double r, int, double ih, double &s, double &sd) {
    double ar = fabs(r);
    if (ar >= 1.0) { s = 0.0; sd = 0.0; return; }
    s = 1.0 - ar;
    sd = (r == 0.0) ? 0.0 : (r > 0.0 ? -ih : ih);
  }
};

template <> struct Basis<KML_SHAPE_CUBIC_SPLINE> {
  static constexpr int SPAN = 4;
  __device__ __forceinline__ static void eval(double r, int nt, double ih, double &s, double &sd) {
    if (r >= 1 && r < 2) {
      if (nt == 1) { s = 0; sd = -ih; }
      else { s = ((-1.0 / 6.0 * r + 1) * r - 2) * r + 4.0 / 3.0; sd = ih * ((-0.5 * r + 2) * r - 2); if (weight_tiny(s)) s = horner3_unfused(-1.0 / 6.0, 1.0, -2.0, 4.0 / 3.0, r); }
    } else if (r >= 0 && r < 1) {
      if (nt == -2) { s = (1.0 / 6.0 * r * r - 1) * r + 1; sd = ih * (0.5 * r * r - 1); }
      else if (nt == 2) { s = 1; sd = ih; }
      else if (nt == 1) { s = (1.0 / 3.0 * r - 1) * r * r + 2.0 / 3.0; sd = ih * r * (r - 2); }
      else { s = (0.5 * r - 1) * r * r + 2.0 / 3.0; sd = ih * (3.0 / 2.0 * r - 2) * r; }
    } else if (r >= -1 && r < 0) {
      if (nt == 2) { s = (-1.0 / 6.0 * r * r + 1) * r + 1; sd = ih * (-0.5 * r * r + 1); }
      else if (nt == -1) { s = (-1.0 / 3.0 * r - 1) * r * r + 2.0 / 3.0; sd = ih * (-r - 2) * r; }
      else { s = (-0.5 * r - 1) * r * r + 2.0 / 3.0; sd = ih * (-3.0 / 2.0 * r - 2) * r; }
    } else if (r >= -2 && r < -1) {
      s = ((1.0 / 6.0 * r + 1) * r + 2) * r + 4.0 / 3.0; sd = ih * ((0.5 * r + 2) * r + 2); if (weight_tiny(s)) s = horner3_unfused(1.0 / 6.0, 1.0, 2.0, 4.0 / 3.0, r);
    } else { s = 0; sd = 0; }
    if (s == 0) sd = 0;
  }
};

template <> struct Basis<KML_SHAPE_QUADRATIC_SPLINE> { // incl. the interval quirk of src/basis_functions.h:146,198
  static constexpr int SPAN = 4;
  __device__ __forceinline__ static void eval(double r, int nt, double ih, double &s, double &sd) {
    s = 0; sd = 0;
    if (nt == 0) {
      if (r >= 0.5 && r < 1.5) { s = (0.5 * r - 1.5) * r + 1.125; sd = ih * (r - 1.5); }
      else if (r >= -0.5 && r < 0.5) { s = -r * r + 0.75; sd = -2 * ih * r; }
      else if (r >= -1.5 && r < 0.5) { s = (0.5 * r + 1.5) * r + 1.125; sd = ih * (r + 1.5); }
    } else if (nt == -2) {
      if (r >= 0. && r < 0.5) { s = 1 - r; sd = -ih; }
      else if (r >= 0.5 && r < 1.5) { s = (0.5 * r - 1.5) * r + 1.125; sd = ih * (r - 1.5); }
    } else if (nt == -1) {
      if (r >= -1. && r < -0.5) { s = 1 + r; sd = ih; }
      else if (r >= -0.5 && r < 0.5) { s = -r * r + 0.75; sd = -2 * ih * r; }
      else if (r >= 0.5 && r < 1.5) { s = (0.5 * r - 1.5) * r + 1.125; sd = ih * (r - 1.5); }
    } else if (nt == 1) {
      if (r >= -1.5 && r < -0.5) { s = (0.5 * r + 1.5) * r + 1.125; sd = ih * (r + 1.5); }
      else if (r >= -0.5 && r < 0.5) { s = -r * r + 0.75; sd = -2 * ih * r; }
      else if (r >= 0.5 && r < 1.) { s = 1 - r; sd = -ih; }
    } else {
      if (r >= -1.5 && r < -0.5) { s = (0.5 * r + 1.5) * r + 1.125; sd = ih * (r + 1.5); }
      else if (r >= -0.5 && r <= 0.) { s = 1 + r; sd = ih; }
    }
    if (weight_tiny(s)) { // the outer polynomial pieces near +-1.5 (a tiny value anywhere else is exact)
      if (r >= 0.5 && r < 1.5 && !(nt == 1)) s = horner2_unfused(0.5, -1.5, 1.125, r);
      else if (r < -0.5 && r >= -1.5 && nt != -2 && nt != -1) s = horner2_unfused(0.5, 1.5, 1.125, r);
    }
    if (s == 0) sd = 0;
  }
};

template <> struct Basis<KML_SHAPE_BERNSTEIN> {
  static constexpr int SPAN = 3; // TLMPM stencil (src/tlmpm.cpp:189-226); ULMPM walks 4 nodes per axis like the splines
  __device__ __forceinline__ static void eval(double rs, int nt, double ih, double &s, double &sd) {
    double r = fabs(rs);
    if (r >= 1.0) { s = 0; sd = 0; return; }
    if (nt == 1) {
      if (r >= 0.5) s = 0; else s = 0.5 - 2 * r * r;
      sd = (r > 0.5) ? 0.0 : -4 * rs * ih;
    } else {
      s = (1 - r) * (1 - r);
      sd = rs > 0 ? -2 * (1 - rs) * ih : 2 * (1 + rs) * ih;
    }
    if (s == 0) sd = 0;
  }
};

// Grid description as the kernels see it.
struct GridDev {
  double lo[3]; double h; double inv_cellsize; double cellsize;
  int n[3]; long long nn;
  // slab decomposition along x: local node plane i is global plane i + goff0 of gn0; lo[] is always the GLOBAL
  // origin so node positions and node types are bit-identical to the undecomposed grid
  int goff0, gn0;
  int own_lo, own_hi; // local planes [own_lo, own_hi) are owned by this rank (shared planes belong to the right neighbour)
  // Node records read by the gather kernels are 32-byte AoS so that one LDG.128 pair fetches a node:
  //   nv  = {vx, vy, vz, mass}      (momentum until the grid kernel divides by the mass)
  //   nvu = {vux, vuy, vuz, T_update}
  // The scatter-only quantities stay SoA.
  double4 *nv, *nvu;
  double *f[3]; double *mb[3];
  double *T, *Qext, *Qint; double *x[3];
  int *mask; int *rigid;
};
__device__ __forceinline__ double4 ldg4(const double4 *p) { // read-only path, two LDG.128
  const double2 lo = __ldg((const double2 *)p), hi = __ldg((const double2 *)p + 1);
  return make_double4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ double &comp(double4 &r, int d) { return d == 0 ? r.x : (d == 1 ? r.y : r.z); }
__device__ __forceinline__ double *comp_ptr(double4 *r, int d) { return d == 0 ? &r->x : (d == 1 ? &r->y : &r->z); }

// ntype per axis, src/grid.cpp:229-240
template <int SHAPE> __device__ __forceinline__ int node_type(int i, int n) {
  if (SHAPE == KML_SHAPE_LINEAR) return 0;
  if (SHAPE == KML_SHAPE_BERNSTEIN) return i & 1;
  return min(2, i) - min(n - 1 - i, 2);
}

// Per-axis stencil of one particle: base node index and SPAN (value, derivative) pairs.
// i0 follows the C truncation of src/ulmpm.cpp:163-165,206-208 / src/tlmpm.cpp:154-156,189-195,229-231,
// computed with explicitly rounded operations so it is bit-identical to the CPU expression.
template <int SHAPE, bool TL> struct StencilSpan {
  static constexpr int value = (SHAPE == KML_SHAPE_LINEAR) ? 2 : ((SHAPE == KML_SHAPE_BERNSTEIN && TL) ? 3 : 4);
};

template <int SHAPE, bool TL, int SPAN>
__device__ __forceinline__ int stencil_base(double xp, double lo, double ih) { // GLOBAL base node index
  double t = __dmul_rn(__dsub_rn(xp, lo), ih);
  if (SHAPE == KML_SHAPE_LINEAR) return (int)t;
  if (SHAPE == KML_SHAPE_BERNSTEIN && TL) { int i0 = 2 * (int)t; if (i0 >= 1 && (i0 & 1)) i0--; return i0; }
  return (int)__dsub_rn(t, 1.0);
}
// n = local node count, goff = global index of local node 0, gn = global node count; i0 is returned LOCAL
template <int SHAPE, bool TL, int SPAN>
__device__ __forceinline__ void axis_weights(double xp, double lo, double h, double ih, int n, int goff, int gn, int &i0, double (&w)[SPAN], double (&dw)[SPAN]) {
  const int ig0 = stencil_base<SHAPE, TL, SPAN>(xp, lo, ih);
  i0 = ig0 - goff;
#pragma unroll
  for (int a = 0; a < SPAN; a++) {
    const int i = i0 + a, ig = ig0 + a;
    if (i < 0 || i >= n) { w[a] = 0; dw[a] = 0; continue; }
    double xn = __dadd_rn(lo, __dmul_rn((double)ig, h));         // node position, src/grid.cpp:222-227
    double r = __dmul_rn(__dsub_rn(xp, xn), ih);                  // src/ulmpm.cpp:246
    Basis<SHAPE>::eval(r, node_type<SHAPE>(ig, gn), ih, w[a], dw[a]);
  }
}

// ------------------------------------------------------------------------------------------
// 3x3 helpers (row-major double[9]) and symmetric storage (xx,yy,zz,xy,xz,yz)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double det3(const double *m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
__device__ __forceinline__ double inv3(const double *m, double *r) { // returns 1 / det(m)
  double c00 = m[4] * m[8] - m[5] * m[7], c10 = m[5] * m[6] - m[3] * m[8], c20 = m[3] * m[7] - m[4] * m[6];
  double id = 1.0 / (m[0] * c00 + m[1] * c10 + m[2] * c20);
  r[0] = c00 * id; r[3] = c10 * id; r[6] = c20 * id;
  r[1] = (m[2] * m[7] - m[1] * m[8]) * id; r[4] = (m[0] * m[8] - m[2] * m[6]) * id; r[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  r[2] = (m[1] * m[5] - m[2] * m[4]) * id; r[5] = (m[2] * m[3] - m[0] * m[5]) * id; r[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return id;
}
__device__ __forceinline__ void mul3(const double *a, const double *b, double *c) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
__device__ __forceinline__ void mul3_bt(const double *a, const double *b, double *c) { // c = a * b^T
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) c[3 * i + j] = a[3 * i] * b[3 * j] + a[3 * i + 1] * b[3 * j + 1] + a[3 * i + 2] * b[3 * j + 2];
}
__device__ __forceinline__ void mul3_at(const double *a, const double *b, double *c) { // c = a^T * b
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) c[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
__device__ __forceinline__ double frob3(const double *m) {
  double s = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) s += m[i] * m[i];
  return sqrt(s);
}
__device__ __forceinline__ void deviator3(const double *m, double *d) { // MPM_Math::Deviator, src/mpm_math.h:28-33
  double t = (m[0] + m[4] + m[8]) * (1.0 / 3.0); // a multiplication, not an FP64 division (<= 1 ulp from the reference's / 3.0)
#pragma unroll
  for (int i = 0; i < 9; i++) d[i] = m[i];
  d[0] -= t; d[4] -= t; d[8] -= t;
}

// Polar decomposition F = R U: R from the SVD like MPM_Math::PolDec (src/mpm_math.h:104-133);
// one-sided Jacobi on a 3x3 (the reference uses Eigen::JacobiSVD, an un-vendored dependency).
__device__ inline bool poldec3(const double *M, double *R) {
  double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; i++) A[i] = M[i];
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < 3; i++) { alpha += A[3 * i + p] * A[3 * i + p]; beta += A[3 * i + q] * A[3 * i + q]; gamma += A[3 * i + p] * A[3 * i + q]; }
        if (gamma == 0.0) continue;
        double lim = sqrt(alpha * beta);
        if (fabs(gamma) <= 1e-300 || fabs(gamma) <= 2.2e-16 * lim * 0.25) continue;
        off = fmax(off, fabs(gamma) / (lim > 0 ? lim : 1));
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < 3; i++) {
          double ap = A[3 * i + p], aq = A[3 * i + q]; A[3 * i + p] = c * ap - s * aq; A[3 * i + q] = s * ap + c * aq;
          double vp = V[3 * i + p], vq = V[3 * i + q]; V[3 * i + p] = c * vp - s * vq; V[3 * i + q] = s * vp + c * vq;
        }
      }
    if (off == 0) break;
  }
  double sv[3], U[9];
  for (int j = 0; j < 3; j++) {
    double s = 0; for (int i = 0; i < 3; i++) s += A[3 * i + j] * A[3 * i + j];
    sv[j] = sqrt(s);
    for (int i = 0; i < 3; i++) U[3 * i + j] = sv[j] > 0 ? A[3 * i + j] / sv[j] : 0.0;
  }
  mul3_bt(U, V, R); // R = U V^T (independent of the ordering of the singular triplets)
  if (det3(R) < 0.0) { // improper rotation: flip the smallest singular value, R = M V S^-1 V^T
    int imin = 0; for (int i = 1; i < 3; i++) if (sv[i] < sv[imin]) imin = i;
    double Sinv[3]; for (int i = 0; i < 3; i++) Sinv[i] = 1.0 / (i == imin ? -sv[i] : sv[i]);
    double T1[9], T2[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T1[3 * i + j] = V[3 * i + j] * Sinv[j];
    mul3_bt(T1, V, T2); mul3(M, T2, R);
  }
  return det3(R) > 0.0;
}

// min |Re(lambda_i)| of a general real 3x3 matrix (TL time-step limiter, src/solid.cpp:1400-1408;
// the reference uses Eigen::EigenSolver).  Reduction to Hessenberg form followed by the implicit double-shift QR iteration for the
// eigenvalues only, i.e. the algorithm of EISPACK's `hqr` (Martin, Peters & Wilkinson, "The QR algorithm for real Hessenberg matrices",
// Numer. Math. 14, 219-231 (1970); Handbook for Automatic Computation II, contribution II/14; the same routine appears as `hqr` in
// Numerical Recipes in C, section 11.6), restated here for a fixed 3 x 3 size - third-party algorithm, not reference code.
__device__ inline bool eig3_min_abs_real(const double *M, double &out) {
  const int n = 3;
  double a[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i][j] = M[3 * i + j];
  {
    double alpha = sqrt(a[1][0] * a[1][0] + a[2][0] * a[2][0]);
    if (a[2][0] != 0.0 && alpha > 0) {
      if (a[1][0] > 0) alpha = -alpha;
      double v1 = a[1][0] - alpha, v2 = a[2][0], vn = v1 * v1 + v2 * v2;
      if (vn > 0) {
        for (int j = 0; j < n; j++) { double d = 2.0 * (v1 * a[1][j] + v2 * a[2][j]) / vn; a[1][j] -= d * v1; a[2][j] -= d * v2; }
        for (int i = 0; i < n; i++) { double d = 2.0 * (a[i][1] * v1 + a[i][2] * v2) / vn; a[i][1] -= d * v1; a[i][2] -= d * v2; }
        a[2][0] = 0.0;
      }
    }
  }
  double wr[3];
  int nn, m, l, k, j, its, i, mmin;
  double z, y, x, w, v, u, t, s, r = 0, q = 0, p = 0, anorm = 0;
  for (i = 0; i < n; i++) for (j = (i - 1 > 0 ? i - 1 : 0); j < n; j++) anorm += fabs(a[i][j]);
  nn = n - 1; t = 0.0;
  while (nn >= 0) {
    its = 0;
    do {
      for (l = nn; l >= 1; l--) {
        s = fabs(a[l - 1][l - 1]) + fabs(a[l][l]);
        if (s == 0.0) s = anorm;
        if (fabs(a[l][l - 1]) + s == s) { a[l][l - 1] = 0.0; break; }
      }
      x = a[nn][nn];
      if (l == nn) { wr[nn] = x + t; nn--; }
      else {
        y = a[nn - 1][nn - 1]; w = a[nn][nn - 1] * a[nn - 1][nn];
        if (l == nn - 1) {
          p = 0.5 * (y - x); q = p * p + w; z = sqrt(fabs(q)); x += t;
          if (q >= 0.0) { z = p + (p >= 0 ? fabs(z) : -fabs(z)); wr[nn - 1] = wr[nn] = x + z; if (z != 0.0) wr[nn] = x - w / z; }
          else { wr[nn - 1] = wr[nn] = x + p; }
          nn -= 2;
        } else {
          if (its == 60) return false;
          if (its == 10 || its == 20) {
            t += x; for (i = 0; i <= nn; i++) a[i][i] -= x;
            s = fabs(a[nn][nn - 1]) + fabs(a[nn - 1][nn - 2]); y = x = 0.75 * s; w = -0.4375 * s * s;
          }
          ++its;
          for (m = nn - 2; m >= l; m--) {
            z = a[m][m]; r = x - z; s = y - z;
            p = (r * s - w) / a[m + 1][m] + a[m][m + 1]; q = a[m + 1][m + 1] - z - r - s; r = a[m + 2][m + 1];
            s = fabs(p) + fabs(q) + fabs(r); p /= s; q /= s; r /= s;
            if (m == l) break;
            u = fabs(a[m][m - 1]) * (fabs(q) + fabs(r));
            v = fabs(p) * (fabs(a[m - 1][m - 1]) + fabs(z) + fabs(a[m + 1][m + 1]));
            if (u + v == v) break;
          }
          for (i = m + 2; i <= nn; i++) { a[i][i - 2] = 0.0; if (i != m + 2) a[i][i - 3] = 0.0; }
          for (k = m; k <= nn - 1; k++) {
            if (k != m) {
              p = a[k][k - 1]; q = a[k + 1][k - 1]; r = 0.0; if (k != nn - 1) r = a[k + 2][k - 1];
              if ((x = fabs(p) + fabs(q) + fabs(r)) != 0.0) { p /= x; q /= x; r /= x; }
            }
            double sq = sqrt(p * p + q * q + r * r);
            s = p >= 0 ? sq : -sq;
            if (s != 0.0) {
              if (k == m) { if (l != m) a[k][k - 1] = -a[k][k - 1]; }
              else a[k][k - 1] = -s * x;
              p += s; x = p / s; y = q / s; z = r / s; q /= p; r /= p;
              for (j = k; j <= nn; j++) {
                p = a[k][j] + q * a[k + 1][j];
                if (k != nn - 1) { p += r * a[k + 2][j]; a[k + 2][j] -= p * z; }
                a[k + 1][j] -= p * y; a[k][j] -= p * x;
              }
              mmin = nn < k + 3 ? nn : k + 3;
              for (i = l; i <= mmin; i++) {
                p = x * a[i][k] + y * a[i][k + 1];
                if (k != nn - 1) { p += z * a[i][k + 2]; a[i][k + 2] -= p * r; }
                a[i][k + 1] -= p * q; a[i][k] -= p;
              }
            }
          }
        }
      }
    } while (l < nn - 1);
  }
  out = fmin(fmin(fabs(wr[0]), fabs(wr[1])), fabs(wr[2]));
  return true;
}

// ------------------------------------------------------------------------------------------
// Constitutive functors.  The material is uniform per launch, so every branch on its type is
// warp-uniform (no divergence); the data-dependent branches (yield, damage) are the
// reference's own.
// ------------------------------------------------------------------------------------------
// EOS*::compute_pressure, src/eos_linear.cpp:71-74, src/eos_shock.cpp:99-133, src/eos_fluid.cpp:68-73
__device__ __forceinline__ double eos_pressure(const kml_material &m, double &e, double J, double rho, double damage, double trD, double cellsize, double T) {
  double p;
  if (m.eos_type == KML_EOS_LINEAR) { e = 0; p = m.eos_K * (1 - J) * (1 - damage); }
  else if (m.eos_type == KML_EOS_FLUID) { double mu = rho / m.rho0; p = m.eos_K * (pow(mu, m.eos_Gamma) - 1.0); e = 0; }
  else {
    double mu = rho / m.rho0 - 1.0;
    double sq = 1.0 - (m.eos_S - 1.0) * mu;
    double pH = m.rho0 * (m.eos_c0 * m.eos_c0) * mu * (1.0 + mu) / (sq * sq);
    if (T > m.eos_Tr) e = (m.eos_cv * m.rho0) * (T - m.eos_Tr); else e = 0;
    p = pH + m.eos_Gamma * (e - 0.0);
    if (damage > 0.0 && p < 0.0) { if (damage >= 1.0) p = 0; else p *= 1.0 - damage; }
    if (!(m.eos_Q1 == 0 && m.eos_Q2 == 0) && trD < 0)
      p += rho * cellsize * (m.eos_Q1 * cellsize * trD * trD - m.eos_Q2 * m.eos_c0 * sqrt(J) * trD);
  }
  return p;
}

// Strength*::update_deviatoric_stress: src/strength_linear.cpp:59-71, src/strength_plastic.cpp:62-115,
// src/strength_jc.cpp:100-184, src/strength_swift.cpp:77-147, src/strength_fluid.cpp:46-58.
// sigma, D full 3x3 row-major; returns the deviatoric stress in sdev and the plastic strain increment.
__device__ __forceinline__ void strength_dev(const kml_material &m, double dt, const double *sigma, const double *D, double *sdev, double &dep,
                                              double eps, double epsdot, double damage, double T) {
  const double G_ = m.str_G;
  dep = 0;
  if (m.strength_type == KML_STRENGTH_LINEAR) {
    double dD[9], ds[9]; deviator3(D, dD); deviator3(sigma, ds);
    double c = 2.0 * G_ * (1 - damage);
#pragma unroll
    for (int i = 0; i < 9; i++) sdev[i] = ds[i] + dt * (c * dD[i]);
    return;
  }
  if (m.strength_type == KML_STRENGTH_FLUID) {
    double dD[9]; deviator3(D, dD);
#pragma unroll
    for (int i = 0; i < 9; i++) sdev[i] = 2.0 * G_ * dD[i];
    return;
  }
  if (m.strength_type == KML_STRENGTH_PLASTIC) {
    double Gd = G_ * (1 - damage), yd = m.str_A * (1 - damage);
    double dD[9], ds[9]; deviator3(D, dD); deviator3(sigma, ds);
#pragma unroll
    for (int i = 0; i < 9; i++) sdev[i] = ds[i] + dt * (2.0 * Gd * dD[i]);
    double J2 = sqrt(3. / 2.) * frob3(sdev);
    if (!(J2 < yd)) {
      dep = (J2 - yd) / (3.0 * Gd);
      double sc = yd / J2;
#pragma unroll
      for (int i = 0; i < 9; i++) sdev[i] *= sc;
    }
    return;
  }
  // Johnson-Cook / Swift
  if (damage >= 1.0) {
#pragma unroll
    for (int i = 0; i < 9; i++) sdev[i] = 0;
    return;
  }
  double ys;
  if (m.strength_type == KML_STRENGTH_JOHNSON_COOK) {
    double ratio = epsdot / m.str_epsdot0;
    ratio = (ratio > 1.0) ? ratio : 1.0; // MAX macro semantics incl. NaN -> 1.0 (src/pointers.h:23-24)
    ys = (eps < 1.0e-10) ? m.str_A : m.str_A + m.str_B * pow(eps, m.str_n);
    if (m.str_C != 0) ys *= pow(1.0 + ratio, m.str_C);
    if (T < m.str_Tm) { if (m.str_m != 0 && T >= m.str_Tr) ys *= 1.0 - pow((T - m.str_Tr) / (m.str_Tm - m.str_Tr), m.str_m); }
    else ys = 0;
  } else {
    ys = (eps > 1.0e-10 && eps > m.str_C) ? m.str_A + m.str_B * pow(eps - m.str_C, m.str_n) : m.str_A;
  }
  double Gd = G_;
  if (damage > 0) { Gd *= (1 - damage); ys *= (1 - damage); }
  double tr[9]; const double c = dt * 2.0 * Gd;
#pragma unroll
  for (int i = 0; i < 9; i++) tr[i] = sigma[i] + c * D[i];
  deviator3(tr, sdev);
  double J2 = KML_SQRT_3_OVER_2 * frob3(sdev);
  if (!(J2 < ys)) {
    dep = (J2 - ys) / (3.0 * Gd);
    double sc = ys / J2;
#pragma unroll
    for (int i = 0; i < 9; i++) sdev[i] *= sc;
  }
}

// DamageJohnsonCook::compute_damage, src/damage_jc.cpp:91-143
__device__ __forceinline__ void damage_jc(const kml_material &m, double &damage_init, double &damage, double pH, const double *sdev, double epsdot, double dep, double T) {
  if (dep == 0 && damage >= 1.0) return;
  double vm = KML_SQRT_3_OVER_2 * frob3(sdev);
  double triax = 0.0;
  if (pH != 0.0 && vm != 0.0) triax = -pH / (vm + 0.001 * fabs(pH));
  if (triax <= -3) { damage_init = 0; return; }
  double fs = m.dmg_d1 + m.dmg_d2 * exp(m.dmg_d3 * triax);
  if (m.dmg_d4 > 0.0 && epsdot > m.dmg_epsdot0) fs *= (1.0 + m.dmg_d4 * log(epsdot / m.dmg_epsdot0));
  if (m.dmg_d5 > 0.0 && T >= m.dmg_Tr) fs *= 1 + m.dmg_d5 * (T - m.dmg_Tr) / (m.dmg_Tm - m.dmg_Tr);
  damage_init += dep / fs;
  if (damage_init >= 1.0) damage = fmin((damage_init - 1.0) * 10, 1.0);
}

// positive doubles order like their bit patterns: atomic max / min on the 64-bit image
__device__ __forceinline__ void atomic_max_pos(double *addr, double v) { atomicMax((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v)); }
__device__ __forceinline__ void atomic_min_pos(double *addr, double v) { atomicMin((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v)); }

} // namespace kml
