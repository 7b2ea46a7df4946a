// oracle_math.h - small fixed-size linear algebra for the CPU oracle.
//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Restates the Eigen 3 fixed-size
// operations the reference relies on (Eigen is a third-party dependency that is
// NOT vendored under /root/reference; version unpinned, Install.txt:5,23 mention
// libeigen3-dev / eigen 3.3.7): Vector3d / Matrix3d arithmetic, inverse,
// determinant, JacobiSVD (used by MPM_Math::PolDec, reference src/mpm_math.h:104-133)
// and EigenSolver (reference src/solid.cpp:1401-1408).
#ifndef KML_ORACLE_MATH_H
#define KML_ORACLE_MATH_H
#include <cmath>
#include <algorithm>

namespace okml {

struct Vec3 {
  double a[3];
  double &operator[](int i) { return a[i]; }
  const double &operator[](int i) const { return a[i]; }
  void setZero() { a[0] = a[1] = a[2] = 0; }
  double norm() const { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
  double dot(const Vec3 &o) const { return a[0] * o.a[0] + a[1] * o.a[1] + a[2] * o.a[2]; }
};
inline Vec3 operator+(const Vec3 &x, const Vec3 &y) { return {{x[0] + y[0], x[1] + y[1], x[2] + y[2]}}; }
inline Vec3 operator-(const Vec3 &x, const Vec3 &y) { return {{x[0] - y[0], x[1] - y[1], x[2] - y[2]}}; }
inline Vec3 operator*(double s, const Vec3 &x) { return {{s * x[0], s * x[1], s * x[2]}}; }
inline Vec3 operator*(const Vec3 &x, double s) { return {{x[0] * s, x[1] * s, x[2] * s}}; }
inline Vec3 operator/(const Vec3 &x, double s) { return {{x[0] / s, x[1] / s, x[2] / s}}; }
inline Vec3 &operator+=(Vec3 &x, const Vec3 &y) { x[0] += y[0]; x[1] += y[1]; x[2] += y[2]; return x; }
inline Vec3 &operator-=(Vec3 &x, const Vec3 &y) { x[0] -= y[0]; x[1] -= y[1]; x[2] -= y[2]; return x; }
inline Vec3 &operator*=(Vec3 &x, double s) { x[0] *= s; x[1] *= s; x[2] *= s; return x; }
inline Vec3 &operator/=(Vec3 &x, double s) { x[0] /= s; x[1] /= s; x[2] /= s; return x; }

struct Mat3 {
  double m[3][3]; // row-major: m[i][j]
  double &operator()(int i, int j) { return m[i][j]; }
  const double &operator()(int i, int j) const { return m[i][j]; }
  void setZero() { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = 0; }
  void setIdentity() { setZero(); m[0][0] = m[1][1] = m[2][2] = 1; }
  double trace() const { return m[0][0] + m[1][1] + m[2][2]; }
  double norm() const { // Frobenius; Eigen sums in storage (column-major) order
    double s = 0;
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) s += m[i][j] * m[i][j];
    return std::sqrt(s);
  }
  Mat3 transpose() const { Mat3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = m[j][i]; return r; }
  double determinant() const {
    return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
           m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
  }
  Mat3 inverse() const {
    Mat3 r;
    double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
    double c10 = m[1][2] * m[2][0] - m[1][0] * m[2][2];
    double c20 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
    double id = 1.0 / (m[0][0] * c00 + m[0][1] * c10 + m[0][2] * c20);
    r.m[0][0] = c00 * id; r.m[1][0] = c10 * id; r.m[2][0] = c20 * id;
    r.m[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id;
    r.m[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id;
    r.m[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id;
    r.m[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
    r.m[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
    r.m[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
    return r;
  }
};
inline Mat3 operator+(const Mat3 &x, const Mat3 &y) { Mat3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = x.m[i][j] + y.m[i][j]; return r; }
inline Mat3 operator-(const Mat3 &x, const Mat3 &y) { Mat3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = x.m[i][j] - y.m[i][j]; return r; }
inline Mat3 operator*(double s, const Mat3 &x) { Mat3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = s * x.m[i][j]; return r; }
inline Mat3 operator*(const Mat3 &x, double s) { return s * x; }
inline Mat3 operator/(const Mat3 &x, double s) { Mat3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = x.m[i][j] / s; return r; }
inline Mat3 operator*(const Mat3 &x, const Mat3 &y) {
  Mat3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[i][j] = x.m[i][0] * y.m[0][j] + x.m[i][1] * y.m[1][j] + x.m[i][2] * y.m[2][j];
  return r;
}
inline Vec3 operator*(const Mat3 &x, const Vec3 &v) {
  Vec3 r;
  for (int i = 0; i < 3; i++) r[i] = x.m[i][0] * v[0] + x.m[i][1] * v[1] + x.m[i][2] * v[2];
  return r;
}
inline Mat3 &operator+=(Mat3 &x, const Mat3 &y) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) x.m[i][j] += y.m[i][j]; return x; }
inline Mat3 &operator*=(Mat3 &x, double s) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) x.m[i][j] *= s; return x; }

// MPM_Math::Deviator, reference src/mpm_math.h:28-33
inline Mat3 Deviator(const Mat3 &M) {
  Mat3 eye; eye.setIdentity(); eye *= M.trace() / 3.0;
  return M - eye;
}

// One-sided (Hestenes) Jacobi SVD: M = U diag(S) V^T, S descending.  Stands in for
// Eigen::JacobiSVD<Matrix3d>(M, ComputeFullU | ComputeFullV).
inline void svd3(const Mat3 &M, Mat3 &U, double S[3], Mat3 &V) {
  Mat3 A = M; Mat3 W; W.setIdentity();
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < 3; i++) { alpha += A(i, p) * A(i, p); beta += A(i, q) * A(i, q); gamma += A(i, p) * A(i, q); }
        if (gamma == 0.0) continue;
        double lim = std::sqrt(alpha * beta);
        if (std::fabs(gamma) <= 1e-300 || std::fabs(gamma) <= 2.2e-16 * lim * 0.25) continue;
        off = std::max(off, std::fabs(gamma) / (lim > 0 ? lim : 1));
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < 3; i++) {
          double ap = A(i, p), aq = A(i, q); A(i, p) = c * ap - s * aq; A(i, q) = s * ap + c * aq;
          double vp = W(i, p), vq = W(i, q); W(i, p) = c * vp - s * vq; W(i, q) = s * vp + c * vq;
        }
      }
    if (off == 0) break;
  }
  double sv[3]; int order[3] = {0, 1, 2};
  for (int j = 0; j < 3; j++) { double s = 0; for (int i = 0; i < 3; i++) s += A(i, j) * A(i, j); sv[j] = std::sqrt(s); }
  std::stable_sort(order, order + 3, [&](int x, int y) { return sv[x] > sv[y]; });
  for (int jj = 0; jj < 3; jj++) {
    int j = order[jj]; S[jj] = sv[j];
    for (int i = 0; i < 3; i++) { V(i, jj) = W(i, j); U(i, jj) = sv[j] > 0 ? A(i, j) / sv[j] : 0.0; }
  }
}

// MPM_Math::PolDec(M, R), reference src/mpm_math.h:104-133
inline bool PolDec(const Mat3 &M, Mat3 &R) {
  Mat3 U, V; double S[3];
  svd3(M, U, S, V);
  R = U * V.transpose();
  if (R.determinant() < 0.0) {
    int imin = 0; for (int i = 1; i < 3; i++) if (S[i] < S[imin]) imin = i;
    Mat3 Sm; Sm.setZero(); for (int i = 0; i < 3; i++) Sm(i, i) = S[i];
    Sm(imin, imin) *= -1.0;
    R = M * V * Sm.inverse() * V.transpose();
  }
  return R.determinant() > 0.0;
}

// Real parts of the eigenvalues of a general real 3x3 matrix: Householder Hessenberg
// reduction + shifted QR (EISPACK hqr scheme); stands in for Eigen::EigenSolver<Matrix3d>.
inline bool eig3_real_parts(const Mat3 &M, double wr[3], double wi[3]) {
  const int n = 3;
  double a[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i][j] = M(i, j);
  {
    double alpha = std::sqrt(a[1][0] * a[1][0] + a[2][0] * a[2][0]);
    if (a[2][0] != 0.0 && alpha > 0) {
      if (a[1][0] > 0) alpha = -alpha;
      double v1 = a[1][0] - alpha, v2 = a[2][0];
      double vn = v1 * v1 + v2 * v2;
      if (vn > 0) {
        for (int j = 0; j < n; j++) { double d = 2.0 * (v1 * a[1][j] + v2 * a[2][j]) / vn; a[1][j] -= d * v1; a[2][j] -= d * v2; }
        for (int i = 0; i < n; i++) { double d = 2.0 * (a[i][1] * v1 + a[i][2] * v2) / vn; a[i][1] -= d * v1; a[i][2] -= d * v2; }
        a[2][0] = 0.0;
      }
    }
  }
  auto sign = [](double x, double y) { return y >= 0 ? std::fabs(x) : -std::fabs(x); };
  int nn, m, l, k, j, its, i, mmin;
  double z, y, x, w, v, u, t, s, r = 0, q = 0, p = 0, anorm = 0;
  for (i = 0; i < n; i++) for (j = std::max(i - 1, 0); j < n; j++) anorm += std::fabs(a[i][j]);
  nn = n - 1; t = 0.0;
  while (nn >= 0) {
    its = 0;
    do {
      for (l = nn; l >= 1; l--) {
        s = std::fabs(a[l - 1][l - 1]) + std::fabs(a[l][l]);
        if (s == 0.0) s = anorm;
        if (std::fabs(a[l][l - 1]) + s == s) { a[l][l - 1] = 0.0; break; }
      }
      x = a[nn][nn];
      if (l == nn) { wr[nn] = x + t; wi[nn--] = 0.0; }
      else {
        y = a[nn - 1][nn - 1]; w = a[nn][nn - 1] * a[nn - 1][nn];
        if (l == nn - 1) {
          p = 0.5 * (y - x); q = p * p + w; z = std::sqrt(std::fabs(q)); x += t;
          if (q >= 0.0) {
            z = p + sign(z, p); wr[nn - 1] = wr[nn] = x + z; if (z != 0.0) wr[nn] = x - w / z; wi[nn - 1] = wi[nn] = 0.0;
          } else { wr[nn - 1] = wr[nn] = x + p; wi[nn - 1] = -(wi[nn] = z); }
          nn -= 2;
        } else {
          if (its == 60) return false;
          if (its == 10 || its == 20) {
            t += x; for (i = 0; i <= nn; i++) a[i][i] -= x;
            s = std::fabs(a[nn][nn - 1]) + std::fabs(a[nn - 1][nn - 2]); y = x = 0.75 * s; w = -0.4375 * s * s;
          }
          ++its;
          for (m = nn - 2; m >= l; m--) {
            z = a[m][m]; r = x - z; s = y - z;
            p = (r * s - w) / a[m + 1][m] + a[m][m + 1]; q = a[m + 1][m + 1] - z - r - s; r = a[m + 2][m + 1];
            s = std::fabs(p) + std::fabs(q) + std::fabs(r); p /= s; q /= s; r /= s;
            if (m == l) break;
            u = std::fabs(a[m][m - 1]) * (std::fabs(q) + std::fabs(r));
            v = std::fabs(p) * (std::fabs(a[m - 1][m - 1]) + std::fabs(z) + std::fabs(a[m + 1][m + 1]));
            if (u + v == v) break;
          }
          for (i = m + 2; i <= nn; i++) { a[i][i - 2] = 0.0; if (i != m + 2) a[i][i - 3] = 0.0; }
          for (k = m; k <= nn - 1; k++) {
            if (k != m) {
              p = a[k][k - 1]; q = a[k + 1][k - 1]; r = 0.0; if (k != nn - 1) r = a[k + 2][k - 1];
              if ((x = std::fabs(p) + std::fabs(q) + std::fabs(r)) != 0.0) { p /= x; q /= x; r /= x; }
            }
            if ((s = sign(std::sqrt(p * p + q * q + r * r), p)) != 0.0) {
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
  return true;
}

} // namespace okml
#endif
