// Device-side fp64 kinematics and camera model for the BA kernels (sm_100a).
// Independent CUDA statement of the arithmetic the reference performs in
//   okvis_kinematics/include/okvis/kinematics/operators.hpp:50-138,
//   okvis_kinematics/include/okvis/kinematics/implementation/Transformation.hpp:51-253,
//   okvis_cv/include/okvis/cameras/implementation/PinholeCamera.hpp:143-212,320-348,
//   okvis_cv/include/okvis/cameras/implementation/RadialTangentialDistortion.hpp:96-111
// (paths relative to okvis_ros/okvis/).  Everything lives in registers; 3x3 matrices are
// passed as plain structs so the compiler can scalarise them.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace svin {

struct V3 {
  double x, y, z;
};
struct Q4 {  // Eigen coeffs order
  double x, y, z, w;
};
struct M3 {  // row-major
  double m[9];
};

__device__ __forceinline__ Q4 qmul(const Q4& a, const Q4& b) {
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
__device__ __forceinline__ double qsqnorm(const Q4& q) { return q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w; }
__device__ __forceinline__ Q4 qnormalized(const Q4& q) {
  const double n = sqrt(qsqnorm(q));
  return Q4{q.x / n, q.y / n, q.z / n, q.w / n};
}
__device__ __forceinline__ Q4 qinverse(const Q4& q) {
  const double n2 = qsqnorm(q);
  return Q4{-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2};
}
// Eigen toRotationMatrix (no normalisation)
__device__ __forceinline__ M3 qrot(const Q4& q) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3 C;
  C.m[0] = 1 - (tyy + tzz);
  C.m[1] = txy - twz;
  C.m[2] = txz + twy;
  C.m[3] = txy + twz;
  C.m[4] = 1 - (txx + tzz);
  C.m[5] = tyz - twx;
  C.m[6] = txz - twy;
  C.m[7] = tyz + twx;
  C.m[8] = 1 - (txx + tyy);
  return C;
}
__device__ __forceinline__ M3 m3t(const M3& A) {
  M3 T;
  T.m[0] = A.m[0]; T.m[1] = A.m[3]; T.m[2] = A.m[6];
  T.m[3] = A.m[1]; T.m[4] = A.m[4]; T.m[5] = A.m[7];
  T.m[6] = A.m[2]; T.m[7] = A.m[5]; T.m[8] = A.m[8];
  return T;
}
__device__ __forceinline__ M3 m3mul(const M3& A, const M3& B) {
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C.m[i * 3 + j] = A.m[i * 3] * B.m[j] + A.m[i * 3 + 1] * B.m[3 + j] + A.m[i * 3 + 2] * B.m[6 + j];
  return C;
}
__device__ __forceinline__ V3 m3v(const M3& A, const V3& v) {
  return V3{A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z,
            A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z};
}
__device__ __forceinline__ M3 crossmx(const V3& v) {
  M3 C;
  C.m[0] = 0;    C.m[1] = -v.z; C.m[2] = v.y;
  C.m[3] = v.z;  C.m[4] = 0;    C.m[5] = -v.x;
  C.m[6] = -v.y; C.m[7] = v.x;  C.m[8] = 0;
  return C;
}
__device__ __forceinline__ double sinc_okvis(double x) {
  if (fabs(x) > 1e-6) return sin(x) / x;
  const double x2 = x * x, x4 = x2 * x2, x6 = x2 * x2 * x2;
  return 1.0 - (1.0 / 6.0) * x2 + (1.0 / 120.0) * x4 - (1.0 / 5040.0) * x6;
}
__device__ __forceinline__ Q4 delta_q(const V3& a) {
  const double half = 0.5 * sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
  const double s = sinc_okvis(half) * 0.5;
  return Q4{s * a.x, s * a.y, s * a.z, cos(half)};
}
// top-left 3x3 of plus(q) / oplus(q) (operators.hpp:92-136)
__device__ __forceinline__ M3 qplus33(const Q4& q) {
  M3 Q;
  Q.m[0] = q.w;  Q.m[1] = -q.z; Q.m[2] = q.y;
  Q.m[3] = q.z;  Q.m[4] = q.w;  Q.m[5] = -q.x;
  Q.m[6] = -q.y; Q.m[7] = q.x;  Q.m[8] = q.w;
  return Q;
}
__device__ __forceinline__ M3 qoplus33(const Q4& q) {
  M3 Q;
  Q.m[0] = q.w;  Q.m[1] = q.z;  Q.m[2] = -q.y;
  Q.m[3] = -q.z; Q.m[4] = q.w;  Q.m[5] = q.x;
  Q.m[6] = q.y;  Q.m[7] = -q.x; Q.m[8] = q.w;
  return Q;
}
// full 4x4 plus/oplus into row-major arrays
__device__ __forceinline__ void qplus44(const Q4& q, double* Q) {
  Q[0] = q.w;  Q[1] = -q.z; Q[2] = q.y;  Q[3] = q.x;
  Q[4] = q.z;  Q[5] = q.w;  Q[6] = -q.x; Q[7] = q.y;
  Q[8] = -q.y; Q[9] = q.x;  Q[10] = q.w; Q[11] = q.z;
  Q[12] = -q.x; Q[13] = -q.y; Q[14] = -q.z; Q[15] = q.w;
}
__device__ __forceinline__ void qoplus44(const Q4& q, double* Q) {
  Q[0] = q.w;  Q[1] = q.z;  Q[2] = -q.y; Q[3] = q.x;
  Q[4] = -q.z; Q[5] = q.w;  Q[6] = q.x;  Q[7] = q.y;
  Q[8] = q.y;  Q[9] = -q.x; Q[10] = q.w; Q[11] = q.z;
  Q[12] = -q.x; Q[13] = -q.y; Q[14] = -q.z; Q[15] = q.w;
}
__device__ __forceinline__ void mat44mul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      C[i * 4 + j] = A[i * 4] * B[j] + A[i * 4 + 1] * B[4 + j] + A[i * 4 + 2] * B[8 + j] + A[i * 4 + 3] * B[12 + j];
}
__device__ __forceinline__ M3 right_jacobian(const V3& phi) {
  const double Phi = sqrt(phi.x * phi.x + phi.y * phi.y + phi.z * phi.z);
  const M3 Px = crossmx(phi);
  const M3 Px2 = m3mul(Px, Px);
  double a, b;
  if (Phi < 1.0e-4) {
    a = -0.5;
    b = 1.0 / 6.0;
  } else {
    const double Phi2 = Phi * Phi, Phi3 = Phi2 * Phi;
    a = -(1.0 - cos(Phi)) / Phi2;
    b = (Phi - sin(Phi)) / Phi3;
  }
  M3 R;
#pragma unroll
  for (int i = 0; i < 9; ++i) R.m[i] = a * Px.m[i] + b * Px2.m[i];
  R.m[0] += 1.0;
  R.m[4] += 1.0;
  R.m[8] += 1.0;
  return R;
}

// okvis::kinematics::Transformation as a value type (q normalised on construction)
struct Tf {
  V3 r;
  Q4 q;
  M3 C;
};
__device__ __forceinline__ Tf tf_make(const V3& r, const Q4& q) {
  Tf T;
  T.r = r;
  T.q = qnormalized(q);
  T.C = qrot(T.q);
  return T;
}
__device__ __forceinline__ Tf tf_load(const double* p) { return tf_make(V3{p[0], p[1], p[2]}, Q4{p[3], p[4], p[5], p[6]}); }
__device__ __forceinline__ Tf tf_inverse(const Tf& T) {
  const V3 t = m3v(m3t(T.C), T.r);
  return tf_make(V3{-t.x, -t.y, -t.z}, qinverse(T.q));
}
__device__ __forceinline__ Tf tf_mul(const Tf& A, const Tf& B) {
  const V3 t = m3v(A.C, B.r);
  return tf_make(V3{t.x + A.r.x, t.y + A.r.y, t.z + A.r.z}, qmul(A.q, B.q));
}
// PoseManifold::plus (PoseManifold.cpp:59-82)
__device__ __forceinline__ void pose_plus(const double* x, const double* d, double* out) {
  Tf T = tf_load(x);
  const Q4 dq = delta_q(V3{d[3], d[4], d[5]});
  const Q4 q = qnormalized(qmul(dq, T.q));
  out[0] = T.r.x + d[0];
  out[1] = T.r.y + d[1];
  out[2] = T.r.z + d[2];
  out[3] = q.x;
  out[4] = q.y;
  out[5] = q.z;
  out[6] = q.w;
}

// okvis::Duration::toSec on a nanosecond difference
__device__ __forceinline__ double ns_to_sec(long long ns) {
  long long sec = ns / 1000000000ll;
  long long nsec = ns % 1000000000ll;
  if (nsec < 0) {
    nsec += 1000000000ll;
    sec -= 1;
  }
  return (double)sec + 1e-9 * (double)nsec;
}

// RadTan distort with 2x2 Jacobian; intr = fu fv cu cv k1 k2 p1 p2
__device__ __forceinline__ void radtan_distort(const double* __restrict__ intr, double u0, double u1, double& d0,
                                               double& d1, double& J00, double& J01, double& J10, double& J11) {
  const double k1 = intr[4], k2 = intr[5], p1 = intr[6], p2 = intr[7];
  const double mx_u = u0 * u0, my_u = u1 * u1, mxy_u = u0 * u1;
  const double rho_u = mx_u + my_u;
  const double rad_dist_u = k1 * rho_u + k2 * rho_u * rho_u;
  d0 = u0 + u0 * rad_dist_u + 2.0 * p1 * mxy_u + p2 * (rho_u + 2.0 * mx_u);
  d1 = u1 + u1 * rad_dist_u + 2.0 * p2 * mxy_u + p1 * (rho_u + 2.0 * my_u);
  J00 = 1 + rad_dist_u + k1 * 2.0 * mx_u + k2 * rho_u * 4 * mx_u + 2.0 * p1 * u1 + 6 * p2 * u0;
  J10 = k1 * 2.0 * u0 * u1 + k2 * 4 * rho_u * u0 * u1 + p1 * 2.0 * u0 + 2.0 * p2 * u1;
  J01 = J10;
  J11 = 1 + rad_dist_u + k1 * 2.0 * my_u + k2 * rho_u * 4 * my_u + 6 * p1 * u1 + 2.0 * p2 * u0;
}

// symmetric 3x3 eigenvalues (cyclic Jacobi), ascending
__device__ __forceinline__ void sym3_eigenvalues(const double* A_, double* ev) {
  double A[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) A[i] = A_[i];
  for (int sweep = 0; sweep < 30; ++sweep) {
    // converged when the off-diagonal mass is below 1e-30 of the diagonal's: Jacobi converges quadratically, the
    // eigenvalues are then exact to working precision (waiting for an exact 0.0 cost ~30 sweeps of sqrt/div per
    // landmark and made k_quality 4x as expensive as k_linearize)
    const double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    if (off <= 1.0e-30 * (A[0] * A[0] + A[4] * A[4] + A[8] * A[8])) break;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        const double apq = A[p * 3 + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double akp = A[k * 3 + p], akq = A[k * 3 + q];
          A[k * 3 + p] = c * akp - s * akq;
          A[k * 3 + q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double apk = A[p * 3 + k], aqk = A[q * 3 + k];
          A[p * 3 + k] = c * apk - s * aqk;
          A[q * 3 + k] = s * apk + c * aqk;
        }
      }
  }
  double e0 = A[0], e1 = A[4], e2 = A[8], t;
  if (e0 > e1) { t = e0; e0 = e1; e1 = t; }
  if (e1 > e2) { t = e1; e1 = e2; e2 = t; }
  if (e0 > e1) { t = e0; e0 = e1; e1 = t; }
  ev[0] = e0;
  ev[1] = e1;
  ev[2] = e2;
}

}  // namespace svin
