// TEST INFRASTRUCTURE — CPU oracle for the svin_b200 hot paths.  Not product code:
// only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may use it.
//
// Small dependency-free fp64 kinematics restating okvis_kinematics (paths
// relative to okvis_ros/okvis/ of the reference):
//   okvis_kinematics/include/okvis/kinematics/operators.hpp:50-138   (crossMx, plus, oplus)
//   okvis_kinematics/include/okvis/kinematics/implementation/Transformation.hpp:51-253
//   (sinc, deltaQ, rightJacobian, Transformation::{inverse,operator*,oplus})
// Quaternions are stored [x y z w] like Eigen::Quaterniond::coeffs().
#pragma once
#include <cmath>
#include <cstring>

namespace orc {

struct Quat {
  double x, y, z, w;
};

inline Quat quat_mul(const Quat& a, const Quat& b) {  // Eigen Hamilton product a*b
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
inline double quat_sqnorm(const Quat& q) { return q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w; }
inline Quat quat_normalized(const Quat& q) {
  double n = std::sqrt(quat_sqnorm(q));
  return Quat{q.x / n, q.y / n, q.z / n, q.w / n};
}
inline Quat quat_conj(const Quat& q) { return Quat{-q.x, -q.y, -q.z, q.w}; }
inline Quat quat_inverse(const Quat& q) {  // Eigen: conjugate / squaredNorm
  double n2 = quat_sqnorm(q);
  return Quat{-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2};
}
// Eigen::QuaternionBase::toRotationMatrix (no normalisation); C row-major 3x3.
inline void quat_to_rot(const Quat& q, double C[9]) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  C[0] = 1 - (tyy + tzz);
  C[1] = txy - twz;
  C[2] = txz + twy;
  C[3] = txy + twz;
  C[4] = 1 - (txx + tzz);
  C[5] = tyz - twx;
  C[6] = txz - twy;
  C[7] = tyz + twx;
  C[8] = 1 - (txx + tyy);
}

inline void cross_mx(const double v[3], double C[9]) {  // operators.hpp:62-75
  C[0] = 0;
  C[1] = -v[2];
  C[2] = v[1];
  C[3] = v[2];
  C[4] = 0;
  C[5] = -v[0];
  C[6] = -v[1];
  C[7] = v[0];
  C[8] = 0;
}
// plus(q): q*p = plus(q) p.coeffs()   (operators.hpp:92-112), row-major 4x4
inline void quat_plus(const Quat& q_, double Q[16]) {
  const double q[4] = {q_.x, q_.y, q_.z, q_.w};
  Q[0] = q[3];  Q[1] = -q[2]; Q[2] = q[1];  Q[3] = q[0];
  Q[4] = q[2];  Q[5] = q[3];  Q[6] = -q[0]; Q[7] = q[1];
  Q[8] = -q[1]; Q[9] = q[0];  Q[10] = q[3]; Q[11] = q[2];
  Q[12] = -q[0]; Q[13] = -q[1]; Q[14] = -q[2]; Q[15] = q[3];
}
// oplus(q): p*q = oplus(q) p.coeffs()  (operators.hpp:116-136)
inline void quat_oplus(const Quat& q_, double Q[16]) {
  const double q[4] = {q_.x, q_.y, q_.z, q_.w};
  Q[0] = q[3];  Q[1] = q[2];  Q[2] = -q[1]; Q[3] = q[0];
  Q[4] = -q[2]; Q[5] = q[3];  Q[6] = q[0];  Q[7] = q[1];
  Q[8] = q[1];  Q[9] = -q[0]; Q[10] = q[3]; Q[11] = q[2];
  Q[12] = -q[0]; Q[13] = -q[1]; Q[14] = -q[2]; Q[15] = q[3];
}

inline double sinc(double x) {  // Transformation.hpp:51-63
  if (std::fabs(x) > 1e-6) return std::sin(x) / x;
  const double c_2 = 1.0 / 6.0, c_4 = 1.0 / 120.0, c_6 = 1.0 / 5040.0;
  const double x_2 = x * x, x_4 = x_2 * x_2, x_6 = x_2 * x_2 * x_2;
  return 1.0 - c_2 * x_2 + c_4 * x_4 - c_6 * x_6;
}
inline Quat delta_q(const double dAlpha[3]) {  // Transformation.hpp:65-71
  const double halfnorm = 0.5 * std::sqrt(dAlpha[0] * dAlpha[0] + dAlpha[1] * dAlpha[1] + dAlpha[2] * dAlpha[2]);
  const double s = sinc(halfnorm) * 0.5;
  return Quat{s * dAlpha[0], s * dAlpha[1], s * dAlpha[2], std::cos(halfnorm)};
}

// generic small row-major helpers ------------------------------------------------
// C(m x n) = A(m x k) * B(k x n)
inline void mm(const double* A, const double* B, double* C, int m, int k, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int l = 0; l < k; ++l) s += A[i * k + l] * B[l * n + j];
      C[i * n + j] = s;
    }
}
// C(m x n) = A(m x k) * B^T, B is (n x k)
inline void mmt(const double* A, const double* B, double* C, int m, int k, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int l = 0; l < k; ++l) s += A[i * k + l] * B[j * k + l];
      C[i * n + j] = s;
    }
}
inline void mv(const double* A, const double* x, double* y, int m, int n) {
  for (int i = 0; i < m; ++i) {
    double s = 0;
    for (int j = 0; j < n; ++j) s += A[i * n + j] * x[j];
    y[i] = s;
  }
}
inline void transpose(const double* A, double* At, int m, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) At[j * m + i] = A[i * n + j];
}
inline void mat3_mul(const double* A, const double* B, double* C) { mm(A, B, C, 3, 3, 3); }
inline void mat3_vec(const double* A, const double* x, double* y) { mv(A, x, y, 3, 3); }
inline void mat3_t(const double* A, double* At) { transpose(A, At, 3, 3); }

// rightJacobian (Transformation.hpp:74-88)
inline void right_jacobian(const double phi[3], double R[9]) {
  const double Phi = std::sqrt(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2]);
  double Px[9], Px2[9];
  cross_mx(phi, Px);
  mat3_mul(Px, Px, Px2);
  double a, b;
  if (Phi < 1.0e-4) {
    a = -0.5;
    b = 1.0 / 6.0;
  } else {
    const double Phi2 = Phi * Phi, Phi3 = Phi2 * Phi;
    a = -(1.0 - std::cos(Phi)) / Phi2;
    b = (Phi - std::sin(Phi)) / Phi3;
  }
  for (int i = 0; i < 9; ++i) R[i] = a * Px[i] + b * Px2[i];
  R[0] += 1.0;
  R[4] += 1.0;
  R[8] += 1.0;
}

// okvis::kinematics::Transformation restated: r, q (normalised on construction), C cache.
struct Transform {
  double r[3];
  Quat q;
  double C[9];
  Transform() {
    r[0] = r[1] = r[2] = 0;
    q = Quat{0, 0, 0, 1};
    quat_to_rot(q, C);
  }
  Transform(const double r_[3], const Quat& q_) {  // Transformation.hpp:104-109 (normalises)
    r[0] = r_[0];
    r[1] = r_[1];
    r[2] = r_[2];
    q = quat_normalized(q_);
    quat_to_rot(q, C);
  }
  static Transform from_params(const double* p) {  // [x y z qx qy qz qw]
    return Transform(p, Quat{p[3], p[4], p[5], p[6]});
  }
  Transform inverse() const {  // Transformation.hpp:147
    double Ct[9], t[3];
    mat3_t(C, Ct);
    mat3_vec(Ct, r, t);
    double nr[3] = {-t[0], -t[1], -t[2]};
    return Transform(nr, quat_inverse(q));
  }
  Transform operator*(const Transform& rhs) const {  // Transformation.hpp:185-187
    double t[3];
    mat3_vec(C, rhs.r, t);
    double nr[3] = {t[0] + r[0], t[1] + r[1], t[2] + r[2]};
    return Transform(nr, quat_mul(q, rhs.q));
  }
  void oplus(const double delta[6]) {  // Transformation.hpp:206-217
    r[0] += delta[0];
    r[1] += delta[1];
    r[2] += delta[2];
    Quat dq = delta_q(delta + 3);
    q = quat_normalized(quat_mul(dq, q));
    quat_to_rot(q, C);
  }
  void to_params(double* p) const {
    p[0] = r[0];
    p[1] = r[1];
    p[2] = r[2];
    p[3] = q.x;
    p[4] = q.y;
    p[5] = q.z;
    p[6] = q.w;
  }
};

// Eigen::LLT<Matrix, Lower> as executed by llt_inplace<Lower>::unblocked (size < 32):
// left-looking, and on a non-positive pivot it RETURNS EARLY leaving the remaining
// columns untouched.  The reference calls it on singular information matrices
// (Estimator.cpp:321-326 -> PoseError.cpp:70-76), so the early exit is observable.
// A (n x n row-major, symmetric) -> L (row-major, lower; strict upper zeroed).
// Returns the failing pivot index or -1.
inline int eigen_llt_lower(const double* A, double* L, int n) {
  for (int i = 0; i < n * n; ++i) L[i] = A[i];
  int fail = -1;
  for (int k = 0; k < n; ++k) {
    double x = L[k * n + k];
    for (int j = 0; j < k; ++j) x -= L[k * n + j] * L[k * n + j];
    if (x <= 0.0) {
      fail = k;
      break;
    }
    x = std::sqrt(x);
    L[k * n + k] = x;
    for (int i = k + 1; i < n; ++i) {
      double s = L[i * n + k];
      for (int j = 0; j < k; ++j) s -= L[i * n + j] * L[k * n + j];
      L[i * n + k] = s / x;
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) L[i * n + j] = 0.0;
  return fail;
}
// squareRootInformation_ = lltOfInformation.matrixL().transpose()  -> U = L^T
inline void sqrt_information(const double* info, double* U, int n) {
  double L[81 * 4];
  eigen_llt_lower(info, L, n);
  transpose(L, U, n, n);
}

// Inverse by LU with partial pivoting (what Eigen's fixed-size inverse() does for n > 4).
inline bool lu_inverse(const double* A_, double* Ainv, int n) {
  double A[15 * 15];
  int piv[15];
  for (int i = 0; i < n * n; ++i) A[i] = A_[i];
  for (int i = 0; i < n; ++i) piv[i] = i;
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = std::fabs(A[k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(A[i * n + k]) > best) {
        best = std::fabs(A[i * n + k]);
        p = i;
      }
    if (best == 0.0) return false;
    if (p != k) {
      for (int j = 0; j < n; ++j) {
        double t = A[k * n + j];
        A[k * n + j] = A[p * n + j];
        A[p * n + j] = t;
      }
      int t = piv[k];
      piv[k] = piv[p];
      piv[p] = t;
    }
    for (int i = k + 1; i < n; ++i) {
      A[i * n + k] /= A[k * n + k];
      const double f = A[i * n + k];
      for (int j = k + 1; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
    }
  }
  for (int c = 0; c < n; ++c) {  // solve A x = e_c
    double y[15];
    for (int i = 0; i < n; ++i) {
      double s = (piv[i] == c) ? 1.0 : 0.0;
      for (int j = 0; j < i; ++j) s -= A[i * n + j] * y[j];
      y[i] = s;
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = y[i];
      for (int j = i + 1; j < n; ++j) s -= A[i * n + j] * Ainv[j * n + c];
      Ainv[i * n + c] = s / A[i * n + i];
    }
  }
  return true;
}

// okvis::Duration::toSec on a nanosecond difference (Duration.hpp:102; sec floor, nsec in [0,1e9))
inline double ns_to_sec(int64_t ns) {
  int64_t sec = ns / 1000000000ll;
  int64_t nsec = ns % 1000000000ll;
  if (nsec < 0) {
    nsec += 1000000000ll;
    sec -= 1;
  }
  return static_cast<double>(sec) + 1e-9 * static_cast<double>(nsec);
}

}  // namespace orc
