// TEST INFRASTRUCTURE — CPU oracle (see orc_math.hpp header note).
//
// fp64 restatement of the reference's error terms and camera model.  Every
// function cites the reference file:line it follows (paths relative to
// okvis_ros/okvis/).  Jacobians are the *minimal* ones (w.r.t. the manifold
// perturbation) which — for unit quaternions — equal the local Jacobians Ceres
// forms as J * PlusJacobian (PoseManifold.cpp:128-140 liftJacobian is the
// pseudo-inverse of Transformation.hpp:232-244 oplusJacobian).
#pragma once
#include <cstdint>
#include <vector>

#include "orc_math.hpp"

namespace orc {

// ---------------------------------------------------------------- camera model
// PinholeCamera<RadialTangentialDistortion>; intr = fu fv cu cv k1 k2 p1 p2.
// RadialTangentialDistortion.hpp impl:96-111 (distort + point Jacobian).
inline void radtan_distort(const double* intr, const double u[2], double d[2], double J[4]) {
  const double k1 = intr[4], k2 = intr[5], p1 = intr[6], p2 = intr[7];
  const double u0 = u[0], u1 = u[1];
  const double mx_u = u0 * u0, my_u = u1 * u1, mxy_u = u0 * u1;
  const double rho_u = mx_u + my_u;
  const double rad_dist_u = k1 * rho_u + k2 * rho_u * rho_u;
  d[0] = u0 + u0 * rad_dist_u + 2.0 * p1 * mxy_u + p2 * (rho_u + 2.0 * mx_u);
  d[1] = u1 + u1 * rad_dist_u + 2.0 * p2 * mxy_u + p1 * (rho_u + 2.0 * my_u);
  if (J) {
    J[0] = 1 + rad_dist_u + k1 * 2.0 * mx_u + k2 * rho_u * 4 * mx_u + 2.0 * p1 * u1 + 6 * p2 * u0;
    J[2] = k1 * 2.0 * u0 * u1 + k2 * 4 * rho_u * u0 * u1 + p1 * 2.0 * u0 + 2.0 * p2 * u1;
    J[1] = J[2];
    J[3] = 1 + rad_dist_u + k1 * 2.0 * my_u + k2 * rho_u * 4 * my_u + 6 * p1 * u1 + 2.0 * p2 * u0;
  }
}

enum ProjStatus { PROJ_SUCCESSFUL = 0, PROJ_OUTSIDE = 1, PROJ_MASKED = 2, PROJ_BEHIND = 3, PROJ_INVALID = 4 };

// PinholeCamera.hpp impl:143-212 project(point, imagePoint, pointJacobian).
// Image bounds are not known to the BA terms (status is unused there); pass w,h<=0 to skip.
inline int pinhole_project(const double* intr, const double p[3], double ip[2], double J[6], int w = 0, int h = 0) {
  if (std::fabs(p[2]) < 1.0e-12) return PROJ_INVALID;
  const double fu = intr[0], fv = intr[1], cu = intr[2], cv = intr[3];
  const double rz = 1.0 / p[2];
  const double rz2 = rz * rz;
  double und[2] = {p[0] * rz, p[1] * rz};
  double d[2], Jd[4];
  radtan_distort(intr, und, d, Jd);
  if (J) {
    J[0] = fu * Jd[0] * rz;
    J[1] = fu * Jd[1] * rz;
    J[2] = -fu * (p[0] * Jd[0] + p[1] * Jd[1]) * rz2;
    J[3] = fv * Jd[2] * rz;
    J[4] = fv * Jd[3] * rz;
    J[5] = -fv * (p[0] * Jd[2] + p[1] * Jd[3]) * rz2;
  }
  ip[0] = fu * d[0] + cu;
  ip[1] = fv * d[1] + cv;
  if (w > 0) {  // CameraBase::isInImage (CameraBase.hpp impl:86-94)
    if (ip[0] < 0.0 || ip[1] < 0.0) return PROJ_OUTSIDE;
    if (ip[0] >= w || ip[1] >= h) return PROJ_OUTSIDE;
  }
  return p[2] > 0.0 ? PROJ_SUCCESSFUL : PROJ_BEHIND;
}
// PinholeCamera.hpp impl:320-348 projectHomogeneous (2x4 Jacobian, last column zero).
inline int pinhole_project_h(const double* intr, const double hp[4], double ip[2], double J24[8], int w = 0, int h = 0) {
  double head[3] = {hp[0], hp[1], hp[2]};
  if (hp[3] < 0) {
    head[0] = -head[0];
    head[1] = -head[1];
    head[2] = -head[2];
  }
  double J3[6];
  int st = pinhole_project(intr, head, ip, J24 ? J3 : nullptr, w, h);
  if (J24) {
    if (st == PROJ_INVALID)
      for (int i = 0; i < 6; ++i) J3[i] = 0.0;  // reference leaves it uninitialised; never consumed
    J24[0] = J3[0]; J24[1] = J3[1]; J24[2] = J3[2]; J24[3] = 0.0;
    J24[4] = J3[3]; J24[5] = J3[4]; J24[6] = J3[5]; J24[7] = 0.0;
  }
  return st;
}
// RadialTangentialDistortion.hpp impl:183-218 undistort (5 Gauss-Newton iterations)
inline bool radtan_undistort(const double* intr, const double y[2], double x_bar[2]) {
  x_bar[0] = y[0];
  x_bar[1] = y[1];
  bool success = false;
  for (int i = 0; i < 5; ++i) {
    double x_tmp[2], E[4];
    radtan_distort(intr, x_bar, x_tmp, E);
    const double e[2] = {y[0] - x_tmp[0], y[1] - x_tmp[1]};
    // du = (E^T E)^-1 E^T e
    const double a = E[0] * E[0] + E[2] * E[2], b = E[0] * E[1] + E[2] * E[3], c = E[1] * E[1] + E[3] * E[3];
    const double det = a * c - b * b;
    const double inv[4] = {c / det, -b / det, -b / det, a / det};
    // M = inv * E^T (2x2), du = M e  (Eigen evaluates (E2.inverse()*E^T)*e left to right)
    const double M[4] = {inv[0] * E[0] + inv[1] * E[1], inv[0] * E[2] + inv[1] * E[3],
                         inv[2] * E[0] + inv[3] * E[1], inv[2] * E[2] + inv[3] * E[3]};
    x_bar[0] += M[0] * e[0] + M[1] * e[1];
    x_bar[1] += M[2] * e[0] + M[3] * e[1];
    const double chi2 = e[0] * e[0] + e[1] * e[1];
    if (chi2 < 1e-4) success = true;
    if (chi2 < 1e-15) {
      success = true;
      break;
    }
  }
  return success;
}
// PinholeCamera.hpp impl:391-408 backProject
inline bool pinhole_backproject(const double* intr, const double ip[2], double dir[3]) {
  const double y[2] = {(ip[0] - intr[2]) * (1.0 / intr[0]), (ip[1] - intr[3]) * (1.0 / intr[1])};
  double u[2];
  bool ok = radtan_undistort(intr, y, u);
  dir[0] = u[0];
  dir[1] = u[1];
  dir[2] = 1.0;
  return ok;
}

// ---------------------------------------------------------------- loss functions
// ceres::CauchyLoss / HuberLoss (Ceres 2.2.0 loss_function.cc, un-vendored) and the
// Corrector the reference restates in-tree at MarginalizationError.cpp:283-330.
inline void loss_evaluate(int type, double a, double s, double rho[3]) {
  if (type == 1) {  // Cauchy: b = a^2, c = 1/b
    const double b = a * a, c = 1.0 / b;
    const double sum = 1.0 + s * c;
    const double inv = 1.0 / sum;
    rho[0] = b * std::log(sum);
    rho[1] = inv > 2.2250738585072014e-308 ? inv : 2.2250738585072014e-308;
    rho[2] = -c * (inv * inv);
  } else if (type == 2) {  // Huber
    const double b = a * a;
    if (s > b) {
      const double r = std::sqrt(s);
      rho[0] = 2.0 * a * r - b;
      rho[1] = a / r > 2.2250738585072014e-308 ? a / r : 2.2250738585072014e-308;
      rho[2] = -rho[1] / (2.0 * s);
    } else {
      rho[0] = s;
      rho[1] = 1.0;
      rho[2] = 0.0;
    }
  } else {
    rho[0] = s;
    rho[1] = 1.0;
    rho[2] = 0.0;
  }
}
struct Corrector {
  double sqrt_rho1, residual_scaling, alpha_sq_norm;
  Corrector(double sq_norm, const double rho[3]) {
    sqrt_rho1 = std::sqrt(rho[1]);
    if (sq_norm == 0.0 || rho[2] <= 0.0) {
      residual_scaling = sqrt_rho1;
      alpha_sq_norm = 0.0;
      return;
    }
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling = sqrt_rho1 / (1 - alpha);
    alpha_sq_norm = alpha / sq_norm;
  }
  // J (m x n row-major) <- sqrt(rho') (J - alpha/|r|^2 r r^T J), with the UNcorrected residuals
  void correct_jacobian(int m, int n, const double* r, double* J) const {
    if (alpha_sq_norm == 0.0) {
      for (int i = 0; i < m * n; ++i) J[i] *= sqrt_rho1;
      return;
    }
    for (int c = 0; c < n; ++c) {
      double rtj = 0;
      for (int i = 0; i < m; ++i) rtj += J[i * n + c] * r[i];
      for (int i = 0; i < m; ++i) J[i * n + c] = sqrt_rho1 * (J[i * n + c] - alpha_sq_norm * r[i] * rtj);
    }
  }
  void correct_residuals(int m, double* r) const {
    for (int i = 0; i < m; ++i) r[i] *= residual_scaling;
  }
};

// ---------------------------------------------------------------- ReprojectionError
// ReprojectionError.hpp impl:85-229.  U = squareRootInformation_ (2x2 row-major, upper).
// Outputs r[2], J0[2x6] (pose), J1[2x3] (landmark, Euclidean), J2[2x6] (extrinsics); any J may be null.
inline bool reprojection_evaluate(const double* pose, const double* hp_W, const double* extr, const double* intr,
                                  const double z[2], const double U[4], double r[2], double* J0, double* J1,
                                  double* J2) {
  const double* t_WS_W = pose;
  const Quat q_WS{pose[3], pose[4], pose[5], pose[6]};
  const double* t_SC_S = extr;
  const Quat q_SC{extr[3], extr[4], extr[5], extr[6]};
  double C_SC[9], C_CS[9], C_WS[9], C_SW[9];
  quat_to_rot(q_SC, C_SC);
  mat3_t(C_SC, C_CS);
  quat_to_rot(q_WS, C_WS);
  mat3_t(C_WS, C_SW);
  // T_CS, T_SW as 4x4
  double T_CS[16] = {0}, T_SW[16] = {0};
  double t1[3], t2[3];
  mat3_vec(C_CS, t_SC_S, t1);
  mat3_vec(C_SW, t_WS_W, t2);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      T_CS[i * 4 + j] = C_CS[i * 3 + j];
      T_SW[i * 4 + j] = C_SW[i * 3 + j];
    }
    T_CS[i * 4 + 3] = -t1[i];
    T_SW[i * 4 + 3] = -t2[i];
  }
  T_CS[15] = 1.0;
  T_SW[15] = 1.0;
  double hp_S[4], hp_C[4];
  mv(T_SW, hp_W, hp_S, 4, 4);
  mv(T_CS, hp_S, hp_C, 4, 4);

  const bool want_J = (J0 || J1 || J2);
  double kp[2] = {0.0, 0.0}, Jh[8], Jhw[8];
  pinhole_project_h(intr, hp_C, kp, want_J ? Jh : nullptr);
  if (want_J) mm(U, Jh, Jhw, 2, 2, 4);
  const double e[2] = {z[0] - kp[0], z[1] - kp[1]};
  r[0] = U[0] * e[0] + U[1] * e[1];
  r[1] = U[2] * e[0] + U[3] * e[1];

  bool valid = true;
  if (std::fabs(hp_C[3]) > 1.0e-8) {
    if (hp_C[2] / hp_C[3] < 0.2) valid = false;
  }
  if (J0) {
    const double p[3] = {hp_W[0] - t_WS_W[0] * hp_W[3], hp_W[1] - t_WS_W[1] * hp_W[3], hp_W[2] - t_WS_W[2] * hp_W[3]};
    double J[24] = {0};  // 4x6
    double px[9], Cpx[9];
    cross_mx(p, px);
    mat3_mul(C_SW, px, Cpx);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        J[i * 6 + j] = C_SW[i * 3 + j] * hp_W[3];
        J[i * 6 + 3 + j] = -Cpx[i * 3 + j];
      }
    double A[8];  // Jh_weighted * T_CS (2x4)
    mm(Jhw, T_CS, A, 2, 4, 4);
    mm(A, J, J0, 2, 4, 6);
    if (!valid)
      for (int i = 0; i < 12; ++i) J0[i] = 0.0;
  }
  if (J1) {
    double T_CW[16], Jfull[8];
    mm(T_CS, T_SW, T_CW, 4, 4, 4);
    mm(Jhw, T_CW, Jfull, 2, 4, 4);
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 3; ++j) J1[i * 3 + j] = valid ? -Jfull[i * 4 + j] : 0.0;
  }
  if (J2) {
    const double p[3] = {hp_S[0] - t_SC_S[0] * hp_S[3], hp_S[1] - t_SC_S[1] * hp_S[3], hp_S[2] - t_SC_S[2] * hp_S[3]};
    double J[24] = {0};
    double px[9], Cpx[9];
    cross_mx(p, px);
    mat3_mul(C_CS, px, Cpx);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        J[i * 6 + j] = C_CS[i * 3 + j] * hp_S[3];
        J[i * 6 + 3 + j] = -Cpx[i * 3 + j];
      }
    mm(Jhw, J, J2, 2, 4, 6);
    if (!valid)
      for (int i = 0; i < 12; ++i) J2[i] = 0.0;
  }
  return valid;
}

// ---------------------------------------------------------------- ImuError
struct ImuParams {
  double sigma_g_c, sigma_a_c, sigma_gw_c, sigma_aw_c, g, g_max, a_max;
};
// The mutable pre-integration cache of ImuError (ImuError.hpp:239-270).
struct ImuState {
  Quat Delta_q{0, 0, 0, 1};
  double C_integral[9] = {0}, C_doubleintegral[9] = {0};
  double acc_integral[3] = {0}, acc_doubleintegral[3] = {0};
  double cross[9] = {0};
  double dalpha_db_g[9] = {0}, dv_db_g[9] = {0}, dp_db_g[9] = {0};
  double P_delta[225] = {0};
  double sb_ref[9] = {0};
  double information[225] = {0};
  double sqrt_information[225] = {0};
  bool redo = true;
  int redo_counter = 0;
};
struct ImuMeasView {
  int n;
  const int64_t* t_ns;
  const double* gyro;   // [n][3]
  const double* accel;  // [n][3]
};

// ImuError::redoPreintegration (ImuError.cpp:76-263)
inline int imu_redo_preintegration(ImuState& S, const ImuMeasView& M, const ImuParams& P, int64_t t0, int64_t t1,
                                   const double sb[9]) {
  int64_t time = t0;
  const int64_t end = t1;
  if (!(M.t_ns[M.n - 1] >= end)) return -1;
  S.Delta_q = Quat{0, 0, 0, 1};
  std::memset(S.C_integral, 0, sizeof S.C_integral);
  std::memset(S.C_doubleintegral, 0, sizeof S.C_doubleintegral);
  std::memset(S.acc_integral, 0, sizeof S.acc_integral);
  std::memset(S.acc_doubleintegral, 0, sizeof S.acc_doubleintegral);
  std::memset(S.cross, 0, sizeof S.cross);
  std::memset(S.dalpha_db_g, 0, sizeof S.dalpha_db_g);
  std::memset(S.dv_db_g, 0, sizeof S.dv_db_g);
  std::memset(S.dp_db_g, 0, sizeof S.dp_db_g);
  std::memset(S.P_delta, 0, sizeof S.P_delta);
  bool hasStarted = false;
  int i = 0;
  for (int it = 0; it < M.n; ++it) {
    const int nx = (it + 1 < M.n) ? it + 1 : it;  // reference dereferences end(); never reached with covering data
    double omega_S_0[3], acc_S_0[3], omega_S_1[3], acc_S_1[3];
    for (int k = 0; k < 3; ++k) {
      omega_S_0[k] = M.gyro[it * 3 + k];
      acc_S_0[k] = M.accel[it * 3 + k];
      omega_S_1[k] = M.gyro[nx * 3 + k];
      acc_S_1[k] = M.accel[nx * 3 + k];
    }
    int64_t nexttime = (it + 1 == M.n) ? t1 : M.t_ns[it + 1];
    double dt = ns_to_sec(nexttime - time);
    if (end < nexttime) {
      const double interval = ns_to_sec(nexttime - M.t_ns[it]);
      nexttime = t1;
      dt = ns_to_sec(nexttime - time);
      const double r = dt / interval;
      for (int k = 0; k < 3; ++k) {
        omega_S_1[k] = (1.0 - r) * omega_S_0[k] + r * omega_S_1[k];
        acc_S_1[k] = (1.0 - r) * acc_S_0[k] + r * acc_S_1[k];
      }
    }
    if (dt <= 0.0) continue;
    if (!hasStarted) {
      hasStarted = true;
      const double r = dt / ns_to_sec(nexttime - M.t_ns[it]);
      for (int k = 0; k < 3; ++k) {
        omega_S_0[k] = r * omega_S_0[k] + (1.0 - r) * omega_S_1[k];
        acc_S_0[k] = r * acc_S_0[k] + (1.0 - r) * acc_S_1[k];
      }
    }
    double sigma_g_c = P.sigma_g_c, sigma_a_c = P.sigma_a_c;
    bool gsat = false, asat = false;
    for (int k = 0; k < 3; ++k) {
      if (std::fabs(omega_S_0[k]) > P.g_max || std::fabs(omega_S_1[k]) > P.g_max) gsat = true;
      if (std::fabs(acc_S_0[k]) > P.a_max || std::fabs(acc_S_1[k]) > P.a_max) asat = true;
    }
    if (gsat) sigma_g_c *= 100;
    if (asat) sigma_a_c *= 100;

    double omega_S_true[3], acc_S_true[3];
    for (int k = 0; k < 3; ++k) {
      omega_S_true[k] = 0.5 * (omega_S_0[k] + omega_S_1[k]) - sb[3 + k];
      acc_S_true[k] = 0.5 * (acc_S_0[k] + acc_S_1[k]) - sb[6 + k];
    }
    const double wn = std::sqrt(omega_S_true[0] * omega_S_true[0] + omega_S_true[1] * omega_S_true[1] +
                                omega_S_true[2] * omega_S_true[2]);
    const double theta_half = wn * 0.5 * dt;
    const double sinc_theta_half = sinc(theta_half);  // ode::sinc == kinematics::sinc
    const double cos_theta_half = std::cos(theta_half);
    Quat dq{sinc_theta_half * omega_S_true[0] * 0.5 * dt, sinc_theta_half * omega_S_true[1] * 0.5 * dt,
            sinc_theta_half * omega_S_true[2] * 0.5 * dt, cos_theta_half};
    const Quat Delta_q_1 = quat_mul(S.Delta_q, dq);
    double C[9], C_1[9], CC[9];
    quat_to_rot(S.Delta_q, C);
    quat_to_rot(Delta_q_1, C_1);
    for (int k = 0; k < 9; ++k) CC[k] = C[k] + C_1[k];
    double CCa[3];
    mat3_vec(CC, acc_S_true, CCa);
    double C_integral_1[9], acc_integral_1[3];
    for (int k = 0; k < 9; ++k) C_integral_1[k] = S.C_integral[k] + 0.5 * CC[k] * dt;
    for (int k = 0; k < 3; ++k) acc_integral_1[k] = S.acc_integral[k] + 0.5 * CCa[k] * dt;
    for (int k = 0; k < 9; ++k) S.C_doubleintegral[k] += S.C_integral[k] * dt + 0.25 * CC[k] * dt * dt;
    for (int k = 0; k < 3; ++k) S.acc_doubleintegral[k] += S.acc_integral[k] * dt + 0.25 * CCa[k] * dt * dt;

    double wdt[3] = {omega_S_true[0] * dt, omega_S_true[1] * dt, omega_S_true[2] * dt};
    double Jr[9], C1Jr[9];
    right_jacobian(wdt, Jr);
    mat3_mul(C_1, Jr, C1Jr);
    for (int k = 0; k < 9; ++k) S.dalpha_db_g[k] += C1Jr[k] * dt;
    double Rdq_inv[9], cross_1[9], t9[9];
    quat_to_rot(quat_inverse(dq), Rdq_inv);
    mat3_mul(Rdq_inv, S.cross, t9);
    for (int k = 0; k < 9; ++k) cross_1[k] = t9[k] + Jr[k] * dt;
    double acc_S_x[9], A0[9], A1[9], B0[9], B1[9], G[9];
    cross_mx(acc_S_true, acc_S_x);
    mat3_mul(C, acc_S_x, A0);
    mat3_mul(A0, S.cross, B0);
    mat3_mul(C_1, acc_S_x, A1);
    mat3_mul(A1, cross_1, B1);
    for (int k = 0; k < 9; ++k) G[k] = B0[k] + B1[k];
    double dv_db_g_1[9];
    for (int k = 0; k < 9; ++k) dv_db_g_1[k] = S.dv_db_g[k] + 0.5 * dt * G[k];
    double F09[9];  // dt*dv_db_g_ + 0.25 dt^2 G  (uses OLD dv_db_g_)
    for (int k = 0; k < 9; ++k) F09[k] = dt * S.dv_db_g[k] + 0.25 * dt * dt * G[k];
    for (int k = 0; k < 9; ++k) S.dp_db_g[k] += F09[k];

    // covariance propagation
    double F[225];
    for (int k = 0; k < 225; ++k) F[k] = 0.0;
    for (int k = 0; k < 15; ++k) F[k * 15 + k] = 1.0;
    double v1[3], X[9];
    for (int k = 0; k < 3; ++k) v1[k] = S.acc_integral[k] * dt + 0.25 * CCa[k] * dt * dt;
    cross_mx(v1, X);
    auto setblk = [&](int r0, int c0, const double* Bm, double sc) {
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) F[(r0 + a) * 15 + c0 + b] = sc * Bm[a * 3 + b];
    };
    setblk(0, 3, X, -1.0);
    const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    setblk(0, 6, I3, dt);
    setblk(0, 9, F09, 1.0);
    double F012[9];
    for (int k = 0; k < 9; ++k) F012[k] = -S.C_integral[k] * dt + 0.25 * CC[k] * dt * dt;
    setblk(0, 12, F012, 1.0);
    setblk(3, 9, C_1, -dt);
    double v2[3];
    for (int k = 0; k < 3; ++k) v2[k] = 0.5 * CCa[k] * dt;
    cross_mx(v2, X);
    setblk(6, 3, X, -1.0);
    setblk(6, 9, G, 0.5 * dt);
    setblk(6, 12, CC, -0.5 * dt);
    double FP[225], FPFt[225];
    mm(F, S.P_delta, FP, 15, 15, 15);
    mmt(FP, F, FPFt, 15, 15, 15);
    std::memcpy(S.P_delta, FPFt, sizeof FPFt);
    const double sigma2_dalpha = dt * sigma_g_c * sigma_g_c;
    const double sigma2_v = dt * sigma_a_c * sigma_a_c;
    const double sigma2_p = 0.5 * dt * dt * sigma2_v;
    const double sigma2_b_g = dt * P.sigma_gw_c * P.sigma_gw_c;
    const double sigma2_b_a = dt * P.sigma_aw_c * P.sigma_aw_c;
    for (int k = 0; k < 3; ++k) {
      S.P_delta[(3 + k) * 15 + 3 + k] += sigma2_dalpha;
      S.P_delta[(6 + k) * 15 + 6 + k] += sigma2_v;
      S.P_delta[(0 + k) * 15 + 0 + k] += sigma2_p;
      S.P_delta[(9 + k) * 15 + 9 + k] += sigma2_b_g;
      S.P_delta[(12 + k) * 15 + 12 + k] += sigma2_b_a;
    }
    // memory shift
    S.Delta_q = Delta_q_1;
    std::memcpy(S.C_integral, C_integral_1, sizeof C_integral_1);
    std::memcpy(S.acc_integral, acc_integral_1, sizeof acc_integral_1);
    std::memcpy(S.cross, cross_1, sizeof cross_1);
    std::memcpy(S.dv_db_g, dv_db_g_1, sizeof dv_db_g_1);
    time = nexttime;
    ++i;
    if (nexttime == t1) break;
  }
  std::memcpy(S.sb_ref, sb, 9 * sizeof(double));
  // symmetrise, invert, symmetrise, LLT
  double Ps[225];
  for (int a = 0; a < 15; ++a)
    for (int b = 0; b < 15; ++b) Ps[a * 15 + b] = 0.5 * S.P_delta[a * 15 + b] + 0.5 * S.P_delta[b * 15 + a];
  std::memcpy(S.P_delta, Ps, sizeof Ps);
  double inv[225];
  lu_inverse(S.P_delta, inv, 15);
  for (int a = 0; a < 15; ++a)
    for (int b = 0; b < 15; ++b) S.information[a * 15 + b] = 0.5 * inv[a * 15 + b] + 0.5 * inv[b * 15 + a];
  sqrt_information(S.information, S.sqrt_information, 15);
  return i;
}

// ImuError::EvaluateWithMinimalJacobians (ImuError.cpp:706-866).
// Outputs r[15]; J0[15x6], J1[15x9], J2[15x6], J3[15x9] (any may be null).
inline void imu_evaluate(ImuState& S, const ImuMeasView& M, const ImuParams& P, int64_t t0, int64_t t1,
                         const double* pose0, const double* sb0, const double* pose1, const double* sb1, double* r,
                         double* J0, double* J1, double* J2, double* J3) {
  const Transform T_WS_0 = Transform::from_params(pose0);
  const Transform T_WS_1 = Transform::from_params(pose1);
  double C_S0_W[9];
  mat3_t(T_WS_0.C, C_S0_W);
  const double Delta_t = ns_to_sec(t1 - t0);
  double Delta_b[6];
  for (int k = 0; k < 6; ++k) Delta_b[k] = sb0[3 + k] - S.sb_ref[3 + k];
  const double nb = std::sqrt(Delta_b[0] * Delta_b[0] + Delta_b[1] * Delta_b[1] + Delta_b[2] * Delta_b[2]);
  S.redo = S.redo || (nb * Delta_t > 0.0001);
  if (S.redo) {
    imu_redo_preintegration(S, M, P, t0, t1, sb0);
    S.redo_counter++;
    for (int k = 0; k < 6; ++k) Delta_b[k] = 0.0;
    S.redo = false;
  }
  const double g_W[3] = {0.0, 0.0, P.g};  // g * (0,0,6371009).normalized()
  double F0[225], F1[225];
  for (int k = 0; k < 225; ++k) {
    F0[k] = 0.0;
    F1[k] = 0.0;
  }
  for (int k = 0; k < 15; ++k) {
    F0[k * 15 + k] = 1.0;
    F1[k * 15 + k] = -1.0;
  }
  double delta_p_est_W[3], delta_v_est_W[3];
  for (int k = 0; k < 3; ++k) {
    delta_p_est_W[k] = T_WS_0.r[k] - T_WS_1.r[k] + sb0[k] * Delta_t - 0.5 * g_W[k] * Delta_t * Delta_t;
    delta_v_est_W[k] = sb0[k] - sb1[k] - g_W[k] * Delta_t;
  }
  double mdb[3], tmp3[3];
  mat3_vec(S.dalpha_db_g, Delta_b, tmp3);
  for (int k = 0; k < 3; ++k) mdb[k] = -tmp3[k];
  const Quat Dq = quat_mul(delta_q(mdb), S.Delta_q);
  auto setblk = [](double* F, int r0, int c0, const double* Bm, double sc) {
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) F[(r0 + a) * 15 + c0 + b] = sc * Bm[a * 3 + b];
  };
  double X[9], T9[9];
  setblk(F0, 0, 0, C_S0_W, 1.0);
  cross_mx(delta_p_est_W, X);
  mat3_mul(C_S0_W, X, T9);
  setblk(F0, 0, 3, T9, 1.0);
  setblk(F0, 0, 6, C_S0_W, Delta_t);
  setblk(F0, 0, 9, S.dp_db_g, 1.0);
  setblk(F0, 0, 12, S.C_doubleintegral, -1.0);
  const Quat q1inv = quat_inverse(T_WS_1.q);
  double Qa[16], Qb[16], Qc[16], Qd[16];
  quat_plus(quat_mul(Dq, q1inv), Qa);
  quat_oplus(T_WS_0.q, Qb);
  mm(Qa, Qb, Qc, 4, 4, 4);
  double B33[9];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) B33[a * 3 + b] = Qc[a * 4 + b];
  setblk(F0, 3, 3, B33, 1.0);
  quat_oplus(quat_mul(q1inv, T_WS_0.q), Qa);
  quat_oplus(Dq, Qb);
  mm(Qa, Qb, Qc, 4, 4, 4);
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) B33[a * 3 + b] = Qc[a * 4 + b];
  double nd[9];
  for (int k = 0; k < 9; ++k) nd[k] = -S.dalpha_db_g[k];
  mat3_mul(B33, nd, T9);
  setblk(F0, 3, 9, T9, 1.0);
  cross_mx(delta_v_est_W, X);
  mat3_mul(C_S0_W, X, T9);
  setblk(F0, 6, 3, T9, 1.0);
  setblk(F0, 6, 6, C_S0_W, 1.0);
  setblk(F0, 6, 9, S.dv_db_g, 1.0);
  setblk(F0, 6, 12, S.C_integral, -1.0);

  setblk(F1, 0, 0, C_S0_W, -1.0);
  quat_plus(Dq, Qa);
  quat_oplus(T_WS_0.q, Qb);
  quat_plus(q1inv, Qd);
  mm(Qa, Qb, Qc, 4, 4, 4);
  mm(Qc, Qd, Qa, 4, 4, 4);
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) B33[a * 3 + b] = Qa[a * 4 + b];
  setblk(F1, 3, 3, B33, -1.0);
  setblk(F1, 6, 6, C_S0_W, -1.0);

  double error[15];
  double a3[3];
  mat3_vec(C_S0_W, delta_p_est_W, a3);
  for (int k = 0; k < 3; ++k) {
    double s = 0;
    for (int c = 0; c < 6; ++c) s += F0[(0 + k) * 15 + 9 + c] * Delta_b[c];
    error[k] = a3[k] + S.acc_doubleintegral[k] + s;
  }
  const Quat qe = quat_mul(Dq, quat_mul(q1inv, T_WS_0.q));
  error[3] = 2 * qe.x;
  error[4] = 2 * qe.y;
  error[5] = 2 * qe.z;
  mat3_vec(C_S0_W, delta_v_est_W, a3);
  for (int k = 0; k < 3; ++k) {
    double s = 0;
    for (int c = 0; c < 6; ++c) s += F0[(6 + k) * 15 + 9 + c] * Delta_b[c];
    error[6 + k] = a3[k] + S.acc_integral[k] + s;
  }
  for (int k = 0; k < 6; ++k) error[9 + k] = sb0[3 + k] - sb1[3 + k];
  mv(S.sqrt_information, error, r, 15, 15);
  auto wblock = [&](const double* F, int c0, int nc, double* J) {
    if (!J) return;
    for (int a = 0; a < 15; ++a)
      for (int b = 0; b < nc; ++b) {
        double s = 0;
        for (int k = 0; k < 15; ++k) s += S.sqrt_information[a * 15 + k] * F[k * 15 + c0 + b];
        J[a * nc + b] = s;
      }
  };
  wblock(F0, 0, 6, J0);
  wblock(F0, 6, 9, J1);
  wblock(F1, 0, 6, J2);
  wblock(F1, 6, 9, J3);
}

// ---------------------------------------------------------------- small terms
// PoseError::EvaluateWithMinimalJacobians (PoseError.cpp:85-132); U = sqrt information (6x6).
inline void pose_error_evaluate(const double* meas7, const double* U, const double* pose, double r[6], double* J) {
  const Transform T_meas = Transform::from_params(meas7);
  const Transform T_WS = Transform::from_params(pose);
  const Transform dp = T_meas * T_WS.inverse();
  double e[6] = {T_meas.r[0] - T_WS.r[0], T_meas.r[1] - T_WS.r[1], T_meas.r[2] - T_WS.r[2],
                 2 * dp.q.x, 2 * dp.q.y, 2 * dp.q.z};
  mv(U, e, r, 6, 6);
  if (J) {
    double Jm[36] = {0};
    for (int k = 0; k < 6; ++k) Jm[k * 6 + k] = -1.0;
    double Qp[16];
    quat_plus(dp.q, Qp);
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) Jm[(3 + a) * 6 + 3 + b] = -Qp[a * 4 + b];
    mm(U, Jm, J, 6, 6, 6);
  }
}
// SpeedAndBiasError (SpeedAndBiasError.cpp:84-111)
inline void speedbias_error_evaluate(const double* meas9, const double* U, const double* sb, double r[9], double* J) {
  double e[9];
  for (int k = 0; k < 9; ++k) e[k] = meas9[k] - sb[k];
  mv(U, e, r, 9, 9);
  if (J)
    for (int k = 0; k < 81; ++k) J[k] = -U[k];
}
// RelativePoseError (RelativePoseError.cpp:76-147)
inline void relative_pose_error_evaluate(const double* U, const double* pose0, const double* pose1, double r[6],
                                         double* J0, double* J1) {
  const Transform T0 = Transform::from_params(pose0);
  const Transform T1 = Transform::from_params(pose1);
  const Transform dp = T1 * T0.inverse();
  double e[6] = {T1.r[0] - T0.r[0], T1.r[1] - T0.r[1], T1.r[2] - T0.r[2], 2 * dp.q.x, 2 * dp.q.y, 2 * dp.q.z};
  mv(U, e, r, 6, 6);
  if (J0) {
    double Jm[36] = {0};
    for (int k = 0; k < 6; ++k) Jm[k * 6 + k] = -1.0;
    double Qp[16];
    quat_plus(dp.q, Qp);
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) Jm[(3 + a) * 6 + 3 + b] = -Qp[a * 4 + b];
    mm(U, Jm, J0, 6, 6, 6);
  }
  if (J1) {
    double Jm[36] = {0};
    for (int k = 0; k < 6; ++k) Jm[k * 6 + k] = 1.0;
    double Qo[16];
    quat_oplus(dp.q, Qo);
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) Jm[(3 + a) * 6 + 3 + b] = Qo[a * 4 + b];
    mm(U, Jm, J1, 6, 6, 6);
  }
}
// SonarError (SonarError.cpp:113-183).  The Jacobian is reproduced AS WRITTEN in the
// reference (direction towards the sonar point, not d r/d t_WS); Ceres multiplies the
// 1x7 block by the 7x6 PlusJacobian, which keeps columns 0..2 and maps the four zero
// quaternion columns to zero, giving the local 1x6 below.
inline void sonar_error_evaluate(double range, double heading, double sqrt_info, const double mean[3],
                                 const double* T_SSo7, const double* pose, double r[1], double* J) {
  const Transform T_WS = Transform::from_params(pose);
  const double d[3] = {T_WS.r[0] - mean[0], T_WS.r[1] - mean[1], T_WS.r[2] - mean[2]};
  const double range_corrected = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  r[0] = sqrt_info * (range - range_corrected);
  if (J) {
    const Transform T_SSo = Transform::from_params(T_SSo7);
    const Transform T_WSo = T_WS * T_SSo;
    const double sp[3] = {range * std::cos(heading), range * std::sin(heading), 0.0};
    const Transform sonar_point(sp, Quat{0, 0, 0, 1});
    const Transform T_WSo_point = T_WSo * sonar_point;
    for (int k = 0; k < 3; ++k) J[k] = sqrt_info * ((T_WS.r[k] - T_WSo_point.r[k]) / range);
    J[3] = J[4] = J[5] = 0.0;
  }
}
// DepthError (DepthError.cpp:70-139): error = z_WS - (-depth + first_depth)
inline void depth_error_evaluate(double depth, double first_depth, double sqrt_info, const double* pose, double r[1],
                                 double* J) {
  r[0] = sqrt_info * (pose[2] - (-1 * depth + first_depth));
  if (J) {
    for (int k = 0; k < 6; ++k) J[k] = 0.0;
    J[2] = sqrt_info * 1.0;
  }
}

// PoseManifold::minus (PoseManifold.cpp:92-102)
inline void pose_minus(const double* x_plus_delta, const double* x, double delta[6]) {
  delta[0] = x_plus_delta[0] - x[0];
  delta[1] = x_plus_delta[1] - x[1];
  delta[2] = x_plus_delta[2] - x[2];
  const Quat qp{x_plus_delta[3], x_plus_delta[4], x_plus_delta[5], x_plus_delta[6]};
  const Quat q{x[3], x[4], x[5], x[6]};
  const Quat d = quat_mul(qp, quat_inverse(q));
  delta[3] = 2 * d.x;
  delta[4] = 2 * d.y;
  delta[5] = 2 * d.z;
}
// PoseManifold::plus (PoseManifold.cpp:59-82)
inline void pose_plus(const double* x, const double delta[6], double* x_plus_delta) {
  Transform T = Transform::from_params(x);
  T.oplus(delta);
  T.to_params(x_plus_delta);
}
// Local (6-col) factor that the marginalisation prior's pose-block Jacobian picks up:
// J_lift(x_lin) * J_plus(x)  (MarginalizationError.cpp:822-835 then Ceres' PlusJacobian):
// rotation 3x3 = top-left of oplus(conj(q_lin)) * oplus(normalized(q)).
inline void marg_pose_rotation_factor(const double* x_lin, const double* x, double M[9]) {
  const Quat ql_inv{-x_lin[3], -x_lin[4], -x_lin[5], x_lin[6]};
  const Quat q = quat_normalized(Quat{x[3], x[4], x[5], x[6]});
  double A[16], B[16], Cm[16];
  quat_oplus(ql_inv, A);
  quat_oplus(q, B);
  mm(A, B, Cm, 4, 4, 4);
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) M[a * 3 + b] = Cm[a * 4 + b];
}

}  // namespace orc
