// TEST INFRASTRUCTURE — CPU oracle for hot path (B): sliding-window bundle adjustment.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference
// legs may load this library; the product (svin_b200/csrc) never does.
//
// What it restates, and where the source is:
//   * error terms, manifolds, camera model: in-tree reference code (see orc_terms.hpp).
//   * the solver Estimator::optimize configures (Estimator.cpp:876-899: SPARSE_SCHUR +
//     DOGLEG, default options otherwise) is Ceres Solver 2.2.0 (f3356504, pinned at
//     okvis_ros/okvis/CMakeLists.txt:124-129), which is NOT vendored under /root/reference.
//     The trust-region loop below restates Ceres' published algorithm
//     (trust_region_minimizer.cc, dogleg_strategy.cc TRADITIONAL_DOGLEG,
//     schur_eliminator_impl.h, corrector.cc, Jacobi column scaling 1/(1+||J_j||)).
//     PARITY UNPINNED for the per-iteration trajectory: the reference's own tests
//     (TestEstimator.cpp:209-212 etc.) only pin convergence tolerances, which
//     tests/test_oracle_ba.py reproduces.
//   * landmark quality post-pass: Estimator.cpp:903-922 + Map.cpp:105-150.
//
// Summation is sequential in term order; the linear system is solved by eliminating
// every non-fixed landmark (3x3 blocks), dense Cholesky of the reduced system, and
// back-substitution — mathematically the solve Ceres' SchurComplementSolver performs.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/svin_b200.h"
#include "orc_terms.hpp"

namespace {
using namespace orc;

struct DenseBlock {
  int off;  // offset in the reduced (dense) vector, -1 if the block is fixed
  int ld;   // local dimension
  std::vector<double> J;  // m x ld row-major
};
struct DenseTerm {
  int m;
  std::vector<double> r;
  std::vector<DenseBlock> blocks;
};

struct Lin {
  double cost = 0;
  // reprojection terms, loss-corrected
  std::vector<double> r, Jp, Jl, Je;
  std::vector<DenseTerm> dense;
};

struct Oracle {
  const SvinBaWindow* w;
  std::vector<double> poses, sbs, lms;  // current estimate
  std::vector<ImuState> imu;
  std::vector<double> obs_U, pose_prior_U, sb_prior_U, rel_U;
  std::vector<int> pose_off, sb_off;
  int n_dense = 0;
  int imu_redo_total = 0;

  explicit Oracle(const SvinBaWindow* win) : w(win) {
    poses.assign(w->pose_blocks, w->pose_blocks + 7 * w->num_pose_blocks);
    sbs.assign(w->speedbias, w->speedbias + 9 * w->num_speedbias);
    lms.assign(w->landmarks, w->landmarks + 4 * w->num_landmarks);
    imu.resize(w->num_imu);
    obs_U.resize(4 * (size_t)w->num_obs);
    for (int i = 0; i < w->num_obs; ++i) sqrt_information(w->obs_information + 4 * i, &obs_U[4 * i], 2);
    pose_prior_U.resize(36 * (size_t)w->num_pose_priors);
    for (int i = 0; i < w->num_pose_priors; ++i)
      sqrt_information(w->pose_prior_information + 36 * i, &pose_prior_U[36 * i], 6);
    sb_prior_U.resize(81 * (size_t)w->num_speedbias_priors);
    for (int i = 0; i < w->num_speedbias_priors; ++i)
      sqrt_information(w->speedbias_prior_information + 81 * i, &sb_prior_U[81 * i], 9);
    rel_U.resize(36 * (size_t)w->num_relative_pose);
    for (int i = 0; i < w->num_relative_pose; ++i)
      sqrt_information(w->relative_pose_information + 36 * i, &rel_U[36 * i], 6);
    pose_off.assign(w->num_pose_blocks, -1);
    sb_off.assign(w->num_speedbias, -1);
    n_dense = 0;
    for (int i = 0; i < w->num_pose_blocks; ++i)
      if (!w->pose_fixed[i]) {
        pose_off[i] = n_dense;
        n_dense += 6;
      }
    for (int i = 0; i < w->num_speedbias; ++i)
      if (!w->speedbias_fixed[i]) {
        sb_off[i] = n_dense;
        n_dense += 9;
      }
  }
  bool lm_fixed(int l) const { return w->landmark_fixed && w->landmark_fixed[l]; }

  ImuMeasView imu_view(int i) const {
    const int a = w->imu_meas_offset[i], b = w->imu_meas_offset[i + 1];
    return ImuMeasView{b - a, w->imu_meas_t_ns + a, w->imu_meas_gyro + 3 * a, w->imu_meas_accel + 3 * a};
  }
  ImuParams imu_params() const {
    const SvinImuParams& p = w->imu_params;
    return ImuParams{p.sigma_g_c, p.sigma_a_c, p.sigma_gw_c, p.sigma_aw_c, p.g, p.g_max, p.a_max};
  }

  // Evaluate every term at (P, SB, LM).  want_jac=false: cost only (still mutates IMU state).
  double evaluate(const std::vector<double>& P, const std::vector<double>& SB, const std::vector<double>& LM,
                  bool want_jac, Lin* lin) {
    double cost = 0;
    const int N = w->num_obs;
    if (want_jac) {
      lin->r.assign(2 * (size_t)N, 0);
      lin->Jp.assign(12 * (size_t)N, 0);
      lin->Jl.assign(6 * (size_t)N, 0);
      lin->Je.assign(12 * (size_t)N, 0);
      lin->dense.clear();
    }
    for (int o = 0; o < N; ++o) {
      double r[2], J0[12], J1[6], J2[12];
      reprojection_evaluate(&P[7 * w->obs_pose[o]], &LM[4 * w->obs_landmark[o]], &P[7 * w->obs_extrinsics[o]],
                            w->intrinsics + 8 * w->obs_camera[o], w->obs_measurement + 2 * o, &obs_U[4 * o], r,
                            want_jac ? J0 : nullptr, want_jac ? J1 : nullptr, want_jac ? J2 : nullptr);
      const double sq = r[0] * r[0] + r[1] * r[1];
      double rho[3];
      loss_evaluate(w->loss_type, w->loss_scale, sq, rho);
      cost += 0.5 * rho[0];
      if (want_jac) {
        if (w->loss_type != SVIN_LOSS_NONE) {
          Corrector c(sq, rho);
          c.correct_jacobian(2, 6, r, J0);
          c.correct_jacobian(2, 3, r, J1);
          c.correct_jacobian(2, 6, r, J2);
          c.correct_residuals(2, r);
        }
        std::memcpy(&lin->r[2 * o], r, sizeof r);
        std::memcpy(&lin->Jp[12 * o], J0, sizeof J0);
        std::memcpy(&lin->Jl[6 * o], J1, sizeof J1);
        std::memcpy(&lin->Je[12 * o], J2, sizeof J2);
      }
    }
    auto push = [&](DenseTerm&& t) {
      double s = 0;
      for (double v : t.r) s += v * v;
      cost += 0.5 * s;
      if (want_jac) lin->dense.push_back(std::move(t));
    };
    // ImuError
    for (int i = 0; i < w->num_imu; ++i) {
      DenseTerm t;
      t.m = 15;
      t.r.resize(15);
      double J0[90], J1[135], J2[90], J3[135];
      const int before = imu[i].redo_counter;
      imu_evaluate(imu[i], imu_view(i), imu_params(), w->imu_t0_ns[i], w->imu_t1_ns[i], &P[7 * w->imu_pose0[i]],
                   &SB[9 * w->imu_speedbias0[i]], &P[7 * w->imu_pose1[i]], &SB[9 * w->imu_speedbias1[i]], t.r.data(),
                   want_jac ? J0 : nullptr, want_jac ? J1 : nullptr, want_jac ? J2 : nullptr,
                   want_jac ? J3 : nullptr);
      imu_redo_total += imu[i].redo_counter - before;
      if (want_jac) {
        t.blocks.push_back(DenseBlock{pose_off[w->imu_pose0[i]], 6, std::vector<double>(J0, J0 + 90)});
        t.blocks.push_back(DenseBlock{sb_off[w->imu_speedbias0[i]], 9, std::vector<double>(J1, J1 + 135)});
        t.blocks.push_back(DenseBlock{pose_off[w->imu_pose1[i]], 6, std::vector<double>(J2, J2 + 90)});
        t.blocks.push_back(DenseBlock{sb_off[w->imu_speedbias1[i]], 9, std::vector<double>(J3, J3 + 135)});
      }
      push(std::move(t));
    }
    for (int i = 0; i < w->num_pose_priors; ++i) {
      DenseTerm t;
      t.m = 6;
      t.r.resize(6);
      double J[36];
      const int b = w->pose_prior_block[i];
      pose_error_evaluate(w->pose_prior_measurement + 7 * i, &pose_prior_U[36 * i], &P[7 * b], t.r.data(),
                          want_jac ? J : nullptr);
      if (want_jac) t.blocks.push_back(DenseBlock{pose_off[b], 6, std::vector<double>(J, J + 36)});
      push(std::move(t));
    }
    for (int i = 0; i < w->num_speedbias_priors; ++i) {
      DenseTerm t;
      t.m = 9;
      t.r.resize(9);
      double J[81];
      const int b = w->speedbias_prior_block[i];
      speedbias_error_evaluate(w->speedbias_prior_measurement + 9 * i, &sb_prior_U[81 * i], &SB[9 * b], t.r.data(),
                               want_jac ? J : nullptr);
      if (want_jac) t.blocks.push_back(DenseBlock{sb_off[b], 9, std::vector<double>(J, J + 81)});
      push(std::move(t));
    }
    for (int i = 0; i < w->num_relative_pose; ++i) {
      DenseTerm t;
      t.m = 6;
      t.r.resize(6);
      double J0[36], J1[36];
      const int b0 = w->relative_pose_block0[i], b1 = w->relative_pose_block1[i];
      relative_pose_error_evaluate(&rel_U[36 * i], &P[7 * b0], &P[7 * b1], t.r.data(), want_jac ? J0 : nullptr,
                                   want_jac ? J1 : nullptr);
      if (want_jac) {
        t.blocks.push_back(DenseBlock{pose_off[b0], 6, std::vector<double>(J0, J0 + 36)});
        t.blocks.push_back(DenseBlock{pose_off[b1], 6, std::vector<double>(J1, J1 + 36)});
      }
      push(std::move(t));
    }
    for (int i = 0; i < w->num_sonar; ++i) {
      DenseTerm t;
      t.m = 1;
      t.r.resize(1);
      double J[6];
      const int b = w->sonar_pose[i];
      sonar_error_evaluate(w->sonar_range[i], w->sonar_heading[i], std::sqrt(w->sonar_information[i]),
                           w->sonar_landmark_mean + 3 * i, w->sonar_T_SSo, &P[7 * b], t.r.data(),
                           want_jac ? J : nullptr);
      if (want_jac) t.blocks.push_back(DenseBlock{pose_off[b], 6, std::vector<double>(J, J + 6)});
      push(std::move(t));
    }
    for (int i = 0; i < w->num_depth; ++i) {
      DenseTerm t;
      t.m = 1;
      t.r.resize(1);
      double J[6];
      const int b = w->depth_pose[i];
      depth_error_evaluate(w->depth_measurement[i], w->depth_first[i], std::sqrt(w->depth_information[i]), &P[7 * b],
                           t.r.data(), want_jac ? J : nullptr);
      if (want_jac) t.blocks.push_back(DenseBlock{pose_off[b], 6, std::vector<double>(J, J + 6)});
      push(std::move(t));
    }
    if (w->marg_num_blocks > 0) {
      // MarginalizationError::EvaluateWithMinimalJacobians (MarginalizationError.cpp:798-844)
      const int m = w->marg_dim;
      DenseTerm t;
      t.m = m;
      t.r.assign(w->marg_e0, w->marg_e0 + m);
      std::vector<double> dchi(m, 0.0);
      int col = 0;
      const double* lp = w->marg_linearization_points;
      for (int b = 0; b < w->marg_num_blocks; ++b) {
        const int kind = w->marg_block_kind[b], idx = w->marg_block_index[b];
        if (kind == SVIN_BLOCK_POSE) {
          const bool fixed = w->pose_fixed[idx];
          if (!fixed) {
            pose_minus(&P[7 * idx], lp, &dchi[col]);
            if (want_jac) {
              double Mr[9];
              marg_pose_rotation_factor(lp, &P[7 * idx], Mr);
              DenseBlock blk{pose_off[idx], 6, std::vector<double>(6 * (size_t)m)};
              for (int rr = 0; rr < m; ++rr) {
                const double* Jrow = w->marg_J + (size_t)rr * m + col;
                for (int c = 0; c < 3; ++c) blk.J[rr * 6 + c] = Jrow[c];
                for (int c = 0; c < 3; ++c)
                  blk.J[rr * 6 + 3 + c] = Jrow[3] * Mr[0 * 3 + c] + Jrow[4] * Mr[1 * 3 + c] + Jrow[5] * Mr[2 * 3 + c];
              }
              t.blocks.push_back(std::move(blk));
            }
            col += 6;
          }
          lp += 7;
        } else if (kind == SVIN_BLOCK_SPEEDBIAS) {
          const bool fixed = w->speedbias_fixed[idx];
          if (!fixed) {
            for (int c = 0; c < 9; ++c) dchi[col + c] = SB[9 * idx + c] - lp[c];
            if (want_jac) {
              DenseBlock blk{sb_off[idx], 9, std::vector<double>(9 * (size_t)m)};
              for (int rr = 0; rr < m; ++rr)
                for (int c = 0; c < 9; ++c) blk.J[rr * 9 + c] = w->marg_J[(size_t)rr * m + col + c];
              t.blocks.push_back(std::move(blk));
            }
            col += 9;
          }
          lp += 9;
        }
      }
      for (int rr = 0; rr < m; ++rr) {
        double s = 0;
        for (int c = 0; c < m; ++c) s += w->marg_J[(size_t)rr * m + c] * dchi[c];
        t.r[rr] += s;
      }
      push(std::move(t));
    }
    if (lin) lin->cost = cost;
    return cost;
  }

  // ---- products with the (unscaled) Jacobian -------------------------------------
  void col_sqnorm(const Lin& L, std::vector<double>& cd, std::vector<double>& cl) const {
    cd.assign(n_dense, 0.0);
    cl.assign(3 * (size_t)w->num_landmarks, 0.0);
    for (int o = 0; o < w->num_obs; ++o) {
      const int op = pose_off[w->obs_pose[o]], oe = pose_off[w->obs_extrinsics[o]], l = w->obs_landmark[o];
      for (int rr = 0; rr < 2; ++rr) {
        if (op >= 0)
          for (int c = 0; c < 6; ++c) cd[op + c] += L.Jp[12 * o + rr * 6 + c] * L.Jp[12 * o + rr * 6 + c];
        if (oe >= 0)
          for (int c = 0; c < 6; ++c) cd[oe + c] += L.Je[12 * o + rr * 6 + c] * L.Je[12 * o + rr * 6 + c];
        if (!lm_fixed(l))
          for (int c = 0; c < 3; ++c) cl[3 * l + c] += L.Jl[6 * o + rr * 3 + c] * L.Jl[6 * o + rr * 3 + c];
      }
    }
    for (const DenseTerm& t : L.dense)
      for (const DenseBlock& b : t.blocks)
        if (b.off >= 0)
          for (int rr = 0; rr < t.m; ++rr)
            for (int c = 0; c < b.ld; ++c) cd[b.off + c] += b.J[rr * b.ld + c] * b.J[rr * b.ld + c];
  }
  // g = J^T r
  void jt_r(const Lin& L, std::vector<double>& gd, std::vector<double>& gl) const {
    gd.assign(n_dense, 0.0);
    gl.assign(3 * (size_t)w->num_landmarks, 0.0);
    for (int o = 0; o < w->num_obs; ++o) {
      const int op = pose_off[w->obs_pose[o]], oe = pose_off[w->obs_extrinsics[o]], l = w->obs_landmark[o];
      for (int rr = 0; rr < 2; ++rr) {
        const double rv = L.r[2 * o + rr];
        if (op >= 0)
          for (int c = 0; c < 6; ++c) gd[op + c] += L.Jp[12 * o + rr * 6 + c] * rv;
        if (oe >= 0)
          for (int c = 0; c < 6; ++c) gd[oe + c] += L.Je[12 * o + rr * 6 + c] * rv;
        if (!lm_fixed(l))
          for (int c = 0; c < 3; ++c) gl[3 * l + c] += L.Jl[6 * o + rr * 3 + c] * rv;
      }
    }
    for (const DenseTerm& t : L.dense)
      for (const DenseBlock& b : t.blocks)
        if (b.off >= 0)
          for (int rr = 0; rr < t.m; ++rr)
            for (int c = 0; c < b.ld; ++c) gd[b.off + c] += b.J[rr * b.ld + c] * t.r[rr];
  }
  // for m = J v: returns sum m^2 in *sq and sum m*(r + m/2) in *mc
  void j_v(const Lin& L, const std::vector<double>& vd, const std::vector<double>& vl, double* sq, double* mc) const {
    double s2 = 0, sm = 0;
    for (int o = 0; o < w->num_obs; ++o) {
      const int op = pose_off[w->obs_pose[o]], oe = pose_off[w->obs_extrinsics[o]], l = w->obs_landmark[o];
      for (int rr = 0; rr < 2; ++rr) {
        double m = 0;
        if (op >= 0)
          for (int c = 0; c < 6; ++c) m += L.Jp[12 * o + rr * 6 + c] * vd[op + c];
        if (oe >= 0)
          for (int c = 0; c < 6; ++c) m += L.Je[12 * o + rr * 6 + c] * vd[oe + c];
        if (!lm_fixed(l))
          for (int c = 0; c < 3; ++c) m += L.Jl[6 * o + rr * 3 + c] * vl[3 * l + c];
        s2 += m * m;
        sm += m * (L.r[2 * o + rr] + m / 2.0);
      }
    }
    for (const DenseTerm& t : L.dense)
      for (int rr = 0; rr < t.m; ++rr) {
        double m = 0;
        for (const DenseBlock& b : t.blocks)
          if (b.off >= 0)
            for (int c = 0; c < b.ld; ++c) m += b.J[rr * b.ld + c] * vd[b.off + c];
        s2 += m * m;
        sm += m * (t.r[rr] + m / 2.0);
      }
    *sq = s2;
    *mc = sm;
  }

  // Solve (Js^T Js + diag(D^2)) y = Js^T r with Js = J diag(s); s, D split dense/landmark.
  bool schur_solve(const Lin& L, const std::vector<double>& sd, const std::vector<double>& sl,
                   const std::vector<double>& Dd, const std::vector<double>& Dl, std::vector<double>& yd,
                   std::vector<double>& yl) const {
    const int n = n_dense, NL = w->num_landmarks;
    std::vector<double> H((size_t)n * n, 0.0), g(n, 0.0);
    // dense terms
    for (const DenseTerm& t : L.dense) {
      for (size_t a = 0; a < t.blocks.size(); ++a) {
        const DenseBlock& A = t.blocks[a];
        if (A.off < 0) continue;
        for (int rr = 0; rr < t.m; ++rr)
          for (int c = 0; c < A.ld; ++c) g[A.off + c] += sd[A.off + c] * A.J[rr * A.ld + c] * t.r[rr];
        for (size_t b = 0; b < t.blocks.size(); ++b) {
          const DenseBlock& B = t.blocks[b];
          if (B.off < 0) continue;
          for (int i = 0; i < A.ld; ++i)
            for (int j = 0; j < B.ld; ++j) {
              double s = 0;
              for (int rr = 0; rr < t.m; ++rr) s += A.J[rr * A.ld + i] * B.J[rr * B.ld + j];
              H[(size_t)(A.off + i) * n + B.off + j] += sd[A.off + i] * sd[B.off + j] * s;
            }
        }
      }
    }
    // reprojection terms: group by landmark
    std::vector<std::vector<int>> by_lm(NL);
    for (int o = 0; o < w->num_obs; ++o) by_lm[w->obs_landmark[o]].push_back(o);
    struct WBlock {
      int off;
      double W[18];  // 6x3
    };
    std::vector<double> Vinv_all(9 * (size_t)NL, 0.0), bl_all(3 * (size_t)NL, 0.0);
    std::vector<std::vector<WBlock>> W_all(NL);
    for (int l = 0; l < NL; ++l) {
      const bool lfix = lm_fixed(l);
      double V[9] = {0}, bl[3] = {0};
      std::vector<WBlock>& Ws = W_all[l];
      for (int o : by_lm[l]) {
        const int offs[2] = {pose_off[w->obs_pose[o]], pose_off[w->obs_extrinsics[o]]};
        const double* Js[2] = {&L.Jp[12 * o], &L.Je[12 * o]};
        double Jls[6];
        for (int rr = 0; rr < 2; ++rr)
          for (int c = 0; c < 3; ++c) Jls[rr * 3 + c] = L.Jl[6 * o + rr * 3 + c] * sl[3 * l + c];
        double Jds[2][12];
        for (int k = 0; k < 2; ++k)
          if (offs[k] >= 0)
            for (int rr = 0; rr < 2; ++rr)
              for (int c = 0; c < 6; ++c) Jds[k][rr * 6 + c] = Js[k][rr * 6 + c] * sd[offs[k] + c];
        // dense-dense contributions
        for (int a = 0; a < 2; ++a) {
          if (offs[a] < 0) continue;
          for (int c = 0; c < 6; ++c) g[offs[a] + c] += Jds[a][c] * L.r[2 * o] + Jds[a][6 + c] * L.r[2 * o + 1];
          for (int b = 0; b < 2; ++b) {
            if (offs[b] < 0) continue;
            for (int i = 0; i < 6; ++i)
              for (int j = 0; j < 6; ++j)
                H[(size_t)(offs[a] + i) * n + offs[b] + j] += Jds[a][i] * Jds[b][j] + Jds[a][6 + i] * Jds[b][6 + j];
          }
        }
        if (lfix) continue;
        for (int i = 0; i < 3; ++i) {
          bl[i] += Jls[i] * L.r[2 * o] + Jls[3 + i] * L.r[2 * o + 1];
          for (int j = 0; j < 3; ++j) V[i * 3 + j] += Jls[i] * Jls[j] + Jls[3 + i] * Jls[3 + j];
        }
        for (int a = 0; a < 2; ++a) {
          if (offs[a] < 0) continue;
          WBlock* wb = nullptr;
          for (WBlock& x : Ws)
            if (x.off == offs[a]) wb = &x;
          if (!wb) {
            Ws.push_back(WBlock{offs[a], {0}});
            wb = &Ws.back();
          }
          for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 3; ++j) wb->W[i * 3 + j] += Jds[a][i] * Jls[j] + Jds[a][6 + i] * Jls[3 + j];
        }
      }
      if (lfix) continue;
      for (int i = 0; i < 3; ++i) V[i * 3 + i] += Dl[3 * l + i] * Dl[3 * l + i];
      // inverse of SPD 3x3 via Cholesky (Ceres InvertPSDMatrix with assume_full_rank)
      double Lc[9];
      if (eigen_llt_lower(V, Lc, 3) >= 0) return false;
      double Li[9] = {0};  // inverse of lower-triangular
      Li[0] = 1.0 / Lc[0];
      Li[4] = 1.0 / Lc[4];
      Li[8] = 1.0 / Lc[8];
      Li[3] = -Lc[3] * Li[0] / Lc[4];
      Li[7] = -Lc[7] * Li[4] / Lc[8];
      Li[6] = -(Lc[6] * Li[0] + Lc[7] * Li[3]) / Lc[8];
      double* Vinv = &Vinv_all[9 * l];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double s = 0;
          for (int k = 0; k < 3; ++k) s += Li[k * 3 + i] * Li[k * 3 + j];
          Vinv[i * 3 + j] = s;
        }
      std::memcpy(&bl_all[3 * l], bl, sizeof bl);
      // H -= W Vinv W^T ; g -= W Vinv bl
      for (const WBlock& A : Ws) {
        double WV[18];
        mm(A.W, Vinv, WV, 6, 3, 3);
        for (int i = 0; i < 6; ++i) g[A.off + i] -= WV[i * 3] * bl[0] + WV[i * 3 + 1] * bl[1] + WV[i * 3 + 2] * bl[2];
        for (const WBlock& B : Ws)
          for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j)
              H[(size_t)(A.off + i) * n + B.off + j] -=
                  WV[i * 3] * B.W[j * 3] + WV[i * 3 + 1] * B.W[j * 3 + 1] + WV[i * 3 + 2] * B.W[j * 3 + 2];
      }
    }
    for (int i = 0; i < n; ++i) H[(size_t)i * n + i] += Dd[i] * Dd[i];
    // dense Cholesky (lower), failing on a non-positive / non-finite pivot
    for (int k = 0; k < n; ++k) {
      double x = H[(size_t)k * n + k];
      for (int j = 0; j < k; ++j) x -= H[(size_t)k * n + j] * H[(size_t)k * n + j];
      if (!(x > 0.0) || !std::isfinite(x)) return false;
      x = std::sqrt(x);
      H[(size_t)k * n + k] = x;
      for (int i = k + 1; i < n; ++i) {
        double s = H[(size_t)i * n + k];
        for (int j = 0; j < k; ++j) s -= H[(size_t)i * n + j] * H[(size_t)k * n + j];
        H[(size_t)i * n + k] = s / x;
      }
    }
    yd.assign(n, 0.0);
    for (int i = 0; i < n; ++i) {
      double s = g[i];
      for (int j = 0; j < i; ++j) s -= H[(size_t)i * n + j] * yd[j];
      yd[i] = s / H[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = yd[i];
      for (int j = i + 1; j < n; ++j) s -= H[(size_t)j * n + i] * yd[j];
      yd[i] = s / H[(size_t)i * n + i];
    }
    yl.assign(3 * (size_t)NL, 0.0);
    for (int l = 0; l < NL; ++l) {
      if (lm_fixed(l)) continue;
      double rhs[3] = {bl_all[3 * l], bl_all[3 * l + 1], bl_all[3 * l + 2]};
      for (const WBlock& A : W_all[l])
        for (int j = 0; j < 3; ++j)
          for (int i = 0; i < 6; ++i) rhs[j] -= A.W[i * 3 + j] * yd[A.off + i];
      mat3_vec(&Vinv_all[9 * l], rhs, &yl[3 * l]);
    }
    for (double v : yd)
      if (!std::isfinite(v)) return false;
    for (double v : yl)
      if (!std::isfinite(v)) return false;
    return true;
  }

  // x (+) delta over all non-fixed blocks
  void plus(const std::vector<double>& dd, const std::vector<double>& dl, std::vector<double>& P2,
            std::vector<double>& SB2, std::vector<double>& LM2) const {
    P2 = poses;
    SB2 = sbs;
    LM2 = lms;
    for (int i = 0; i < w->num_pose_blocks; ++i)
      if (pose_off[i] >= 0) pose_plus(&poses[7 * i], &dd[pose_off[i]], &P2[7 * i]);
    for (int i = 0; i < w->num_speedbias; ++i)
      if (sb_off[i] >= 0)
        for (int c = 0; c < 9; ++c) SB2[9 * i + c] = sbs[9 * i + c] + dd[sb_off[i] + c];
    for (int l = 0; l < w->num_landmarks; ++l)
      if (!lm_fixed(l))
        for (int c = 0; c < 3; ++c) LM2[4 * l + c] = lms[4 * l + c] + dl[3 * l + c];
  }
  // ambient norms over non-fixed blocks: |x| and |x - y|
  double ambient_norm(const std::vector<double>& P, const std::vector<double>& SB, const std::vector<double>& LM,
                      const std::vector<double>* P2, const std::vector<double>* SB2,
                      const std::vector<double>* LM2) const {
    double s = 0;
    for (int i = 0; i < w->num_pose_blocks; ++i)
      if (pose_off[i] >= 0)
        for (int c = 0; c < 7; ++c) {
          const double d = P[7 * i + c] - (P2 ? (*P2)[7 * i + c] : 0.0);
          s += d * d;
        }
    for (int i = 0; i < w->num_speedbias; ++i)
      if (sb_off[i] >= 0)
        for (int c = 0; c < 9; ++c) {
          const double d = SB[9 * i + c] - (SB2 ? (*SB2)[9 * i + c] : 0.0);
          s += d * d;
        }
    for (int l = 0; l < w->num_landmarks; ++l)
      if (!lm_fixed(l))
        for (int c = 0; c < 4; ++c) {
          const double d = LM[4 * l + c] - (LM2 ? (*LM2)[4 * l + c] : 0.0);
          s += d * d;
        }
    return std::sqrt(s);
  }
};

// symmetric 3x3 eigenvalues by cyclic Jacobi; ascending in ev[3]
void sym3_eigenvalues(const double A_[9], double ev[3]) {
  double A[9];
  std::memcpy(A, A_, sizeof A);
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    if (off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = A[p * 3 + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A * G
          const double akp = A[k * 3 + p], akq = A[k * 3 + q];
          A[k * 3 + p] = c * akp - s * akq;
          A[k * 3 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- G^T * A
          const double apk = A[p * 3 + k], aqk = A[q * 3 + k];
          A[p * 3 + k] = c * apk - s * aqk;
          A[q * 3 + k] = s * apk + c * aqk;
        }
      }
  }
  ev[0] = A[0];
  ev[1] = A[4];
  ev[2] = A[8];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2 - i; ++j)
      if (ev[j] > ev[j + 1]) {
        double t = ev[j];
        ev[j] = ev[j + 1];
        ev[j + 1] = t;
      }
}

}  // namespace

extern "C" {

void svin_oracle_default_options(SvinBaOptions* o) {
  o->max_num_iterations = 10;
  o->min_num_iterations = 3;
  o->time_limit_seconds = -1.0;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
  o->compute_landmark_quality = 1;
}

// Raw term dump (EvaluateWithMinimalJacobians seam), fresh IMU state.
int svin_oracle_ba_evaluate(const SvinBaWindow* w, SvinBaEvaluation* out) {
  Oracle O(w);
  for (int o = 0; o < w->num_obs; ++o) {
    double r[2], J0[12], J1[6], J2[12];
    reprojection_evaluate(&O.poses[7 * w->obs_pose[o]], &O.lms[4 * w->obs_landmark[o]],
                          &O.poses[7 * w->obs_extrinsics[o]], w->intrinsics + 8 * w->obs_camera[o],
                          w->obs_measurement + 2 * o, &O.obs_U[4 * o], r, J0, J1, J2);
    if (out->reproj_residuals) std::memcpy(out->reproj_residuals + 2 * o, r, sizeof r);
    if (out->reproj_J_pose) std::memcpy(out->reproj_J_pose + 12 * o, J0, sizeof J0);
    if (out->reproj_J_landmark) std::memcpy(out->reproj_J_landmark + 6 * o, J1, sizeof J1);
    if (out->reproj_J_extrinsics) std::memcpy(out->reproj_J_extrinsics + 12 * o, J2, sizeof J2);
  }
  for (int i = 0; i < w->num_imu; ++i) {
    double r[15], J0[90], J1[135], J2[90], J3[135];
    ImuState S;
    imu_evaluate(S, O.imu_view(i), O.imu_params(), w->imu_t0_ns[i], w->imu_t1_ns[i], &O.poses[7 * w->imu_pose0[i]],
                 &O.sbs[9 * w->imu_speedbias0[i]], &O.poses[7 * w->imu_pose1[i]], &O.sbs[9 * w->imu_speedbias1[i]], r,
                 J0, J1, J2, J3);
    if (out->imu_residuals) std::memcpy(out->imu_residuals + 15 * i, r, sizeof r);
    if (out->imu_J_pose0) std::memcpy(out->imu_J_pose0 + 90 * i, J0, sizeof J0);
    if (out->imu_J_speedbias0) std::memcpy(out->imu_J_speedbias0 + 135 * i, J1, sizeof J1);
    if (out->imu_J_pose1) std::memcpy(out->imu_J_pose1 + 90 * i, J2, sizeof J2);
    if (out->imu_J_speedbias1) std::memcpy(out->imu_J_speedbias1 + 135 * i, J3, sizeof J3);
  }
  if (out->cost) {
    Oracle O2(w);
    out->cost[0] = O2.evaluate(O2.poses, O2.sbs, O2.lms, false, nullptr);
  }
  return 0;
}

// Trust-region solve; writes the solution back into the window's in/out arrays.
int svin_oracle_ba_solve(SvinBaWindow* w, const SvinBaOptions* opt, SvinBaSummary* sum, double* landmark_quality) {
  for (int b = 0; b < w->marg_num_blocks; ++b)
    if (w->marg_block_kind[b] == SVIN_BLOCK_LANDMARK) return SVIN_ERR_INVALID_ARGUMENT;
  Oracle O(w);
  const auto t_start = std::chrono::steady_clock::now();
  auto elapsed = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };
  const int n = O.n_dense, NL = w->num_landmarks;
  Lin L;
  double x_cost = O.evaluate(O.poses, O.sbs, O.lms, true, &L);
  std::vector<double> sd(n, 1.0), sl(3 * (size_t)NL, 1.0), cd, cl;
  if (opt->jacobi_scaling) {
    O.col_sqnorm(L, cd, cl);
    for (int i = 0; i < n; ++i) sd[i] = 1.0 / (1.0 + std::sqrt(cd[i]));
    for (size_t i = 0; i < sl.size(); ++i) sl[i] = 1.0 / (1.0 + std::sqrt(cl[i]));
  }
  std::vector<double> gd, gl;
  O.jt_r(L, gd, gl);
  auto grad_max = [&]() {
    double m = 0;
    for (double v : gd) m = std::fmax(m, std::fabs(v));
    for (int l = 0; l < NL; ++l)
      if (!O.lm_fixed(l))
        for (int c = 0; c < 3; ++c) m = std::fmax(m, std::fabs(gl[3 * l + c]));
    return m;
  };
  double gmax = grad_max();
  SvinBaSummary S{};
  S.initial_cost = x_cost;
  S.termination = SVIN_TERM_NO_CONVERGENCE;
  double radius = opt->initial_trust_region_radius;
  double mu = 1e-8;
  const double min_mu = 1e-8, max_mu = 1.0, mu_increase_factor = 10.0;
  bool reuse = false;
  int invalid = 0;
  int iter = 0;
  bool last_successful = false;
  double x_norm = O.ambient_norm(O.poses, O.sbs, O.lms, nullptr, nullptr, nullptr);
  double dogleg_step_norm = 0.0;
  // dogleg state (scaled space, all parameters)
  std::vector<double> diag_d, diag_l, grad_d, grad_l, gn_d, gn_l;
  double alpha = 0;
  double last_iter_time = 0, iter_t0 = elapsed();

  bool done = false;
  if (gmax <= opt->gradient_tolerance) {
    S.termination = SVIN_TERM_CONVERGENCE;
    done = true;
  }
  while (!done) {
    // ---- FinalizeIterationAndCheckIfMinimizerCanContinue of the previous iteration
    if (opt->time_limit_seconds >= 0.0 && iter >= opt->min_num_iterations &&
        elapsed() + last_iter_time > opt->time_limit_seconds) {
      S.termination = SVIN_TERM_USER_SUCCESS;
      break;
    }
    if (iter >= opt->max_num_iterations) {
      S.termination = SVIN_TERM_NO_CONVERGENCE;
      break;
    }
    if (last_successful && gmax <= opt->gradient_tolerance) {
      S.termination = SVIN_TERM_CONVERGENCE;
      break;
    }
    if (radius < opt->min_trust_region_radius) {
      S.termination = SVIN_TERM_CONVERGENCE;
      break;
    }
    ++iter;
    iter_t0 = elapsed();
    last_successful = false;

    // ---- DoglegStrategy::ComputeStep
    bool step_valid = true;
    if (!reuse) {
      reuse = true;
      O.col_sqnorm(L, cd, cl);
      diag_d.assign(n, 0.0);
      diag_l.assign(3 * (size_t)NL, 1.0);
      for (int i = 0; i < n; ++i)
        diag_d[i] = std::sqrt(std::fmin(std::fmax(cd[i] * sd[i] * sd[i], opt->min_lm_diagonal), opt->max_lm_diagonal));
      for (size_t i = 0; i < diag_l.size(); ++i)
        diag_l[i] = std::sqrt(std::fmin(std::fmax(cl[i] * sl[i] * sl[i], opt->min_lm_diagonal), opt->max_lm_diagonal));
      // gradient_ = Js^T r / diagonal
      std::vector<double> jd, jl;
      O.jt_r(L, jd, jl);
      grad_d.assign(n, 0.0);
      grad_l.assign(3 * (size_t)NL, 0.0);
      for (int i = 0; i < n; ++i) grad_d[i] = jd[i] * sd[i] / diag_d[i];
      for (int l = 0; l < NL; ++l)
        if (!O.lm_fixed(l))
          for (int c = 0; c < 3; ++c) grad_l[3 * l + c] = jl[3 * l + c] * sl[3 * l + c] / diag_l[3 * l + c];
      // Cauchy point: alpha = |g|^2 / |Js (g / diag)|^2
      std::vector<double> vd(n), vl(3 * (size_t)NL, 0.0);
      double g2 = 0;
      for (int i = 0; i < n; ++i) {
        vd[i] = sd[i] * grad_d[i] / diag_d[i];
        g2 += grad_d[i] * grad_d[i];
      }
      for (size_t i = 0; i < vl.size(); ++i) {
        vl[i] = sl[i] * grad_l[i] / diag_l[i];
        g2 += grad_l[i] * grad_l[i];
      }
      double Jg2, dummy;
      O.j_v(L, vd, vl, &Jg2, &dummy);
      alpha = g2 / Jg2;
      // Gauss-Newton step with D = sqrt(mu) * diagonal, retrying with larger mu on failure
      bool ok = false;
      while (mu < max_mu) {
        std::vector<double> Dd(n), Dl(3 * (size_t)NL);
        const double smu = std::sqrt(mu);
        for (int i = 0; i < n; ++i) Dd[i] = diag_d[i] * smu;
        for (size_t i = 0; i < Dl.size(); ++i) Dl[i] = diag_l[i] * smu;
        ok = O.schur_solve(L, sd, sl, Dd, Dl, gn_d, gn_l);
        if (!ok) {
          mu *= mu_increase_factor;
          continue;
        }
        break;
      }
      if (ok) {
        for (int i = 0; i < n; ++i) gn_d[i] *= -diag_d[i];
        for (size_t i = 0; i < gn_l.size(); ++i) gn_l[i] *= -diag_l[i];
      } else {
        step_valid = false;
      }
    }
    std::vector<double> step_d(n, 0.0), step_l(3 * (size_t)NL, 0.0);
    double model_cost_change = 0;
    if (step_valid) {
      // ComputeTraditionalDoglegStep
      double g2 = 0, n2 = 0, gdot = 0;
      for (int i = 0; i < n; ++i) {
        g2 += grad_d[i] * grad_d[i];
        n2 += gn_d[i] * gn_d[i];
        gdot += grad_d[i] * gn_d[i];
      }
      for (size_t i = 0; i < grad_l.size(); ++i) {
        g2 += grad_l[i] * grad_l[i];
        n2 += gn_l[i] * gn_l[i];
        gdot += grad_l[i] * gn_l[i];
      }
      const double gradient_norm = std::sqrt(g2), gauss_newton_norm = std::sqrt(n2);
      double cg, cn;  // step = cg * gradient + cn * gauss_newton
      if (gauss_newton_norm <= radius) {
        cg = 0.0;
        cn = 1.0;
        dogleg_step_norm = gauss_newton_norm;
      } else if (gradient_norm * alpha >= radius) {
        cg = -(radius / gradient_norm);
        cn = 0.0;
        dogleg_step_norm = radius;
      } else {
        const double b_dot_a = -alpha * gdot;
        const double a_squared_norm = std::pow(alpha * gradient_norm, 2.0);
        const double b_minus_a_squared_norm = a_squared_norm - 2 * b_dot_a + std::pow(gauss_newton_norm, 2);
        const double c = b_dot_a - a_squared_norm;
        const double d = std::sqrt(c * c + b_minus_a_squared_norm * (std::pow(radius, 2.0) - a_squared_norm));
        const double beta = (c <= 0) ? (d - c) / b_minus_a_squared_norm : (radius * radius - a_squared_norm) / (d + c);
        cg = -alpha * (1.0 - beta);
        cn = beta;
        double s2 = 0;
        for (int i = 0; i < n; ++i) {
          const double v = cg * grad_d[i] + cn * gn_d[i];
          s2 += v * v;
        }
        for (size_t i = 0; i < grad_l.size(); ++i) {
          const double v = cg * grad_l[i] + cn * gn_l[i];
          s2 += v * v;
        }
        dogleg_step_norm = std::sqrt(s2);
      }
      for (int i = 0; i < n; ++i) step_d[i] = (cg * grad_d[i] + cn * gn_d[i]) / diag_d[i];
      for (size_t i = 0; i < step_l.size(); ++i) step_l[i] = (cg * grad_l[i] + cn * gn_l[i]) / diag_l[i];
      // model_cost_change = -m.(r + m/2), m = Js step
      std::vector<double> vd(n), vl(step_l.size());
      for (int i = 0; i < n; ++i) vd[i] = sd[i] * step_d[i];
      for (size_t i = 0; i < vl.size(); ++i) vl[i] = sl[i] * step_l[i];
      double sq, mc;
      O.j_v(L, vd, vl, &sq, &mc);
      model_cost_change = -mc;
      if (!(model_cost_change > 0.0)) step_valid = false;
      if (step_valid) {
        invalid = 0;  // num_consecutive_invalid_steps_ resets on every valid step
        // delta = step .* scale (already in vd, vl)
        std::vector<double> P2, SB2, LM2;
        O.plus(vd, vl, P2, SB2, LM2);
        const double cand_cost = O.evaluate(P2, SB2, LM2, false, nullptr);
        // ParameterToleranceReached
        const double step_norm = O.ambient_norm(O.poses, O.sbs, O.lms, &P2, &SB2, &LM2);
        if (step_norm <= opt->parameter_tolerance * (x_norm + opt->parameter_tolerance)) {
          S.termination = SVIN_TERM_CONVERGENCE;
          break;
        }
        // FunctionToleranceReached
        const double cost_change = x_cost - cand_cost;
        if (std::fabs(cost_change) <= opt->function_tolerance * x_cost) {
          S.termination = SVIN_TERM_CONVERGENCE;
          break;
        }
        const double relative_decrease = cost_change / model_cost_change;
        if (relative_decrease > opt->min_relative_decrease) {
          // HandleSuccessfulStep
          O.poses = P2;
          O.sbs = SB2;
          O.lms = LM2;
          x_norm = O.ambient_norm(O.poses, O.sbs, O.lms, nullptr, nullptr, nullptr);
          x_cost = O.evaluate(O.poses, O.sbs, O.lms, true, &L);
          O.jt_r(L, gd, gl);
          gmax = grad_max();
          last_successful = true;
          ++S.num_successful_steps;
          // DoglegStrategy::StepAccepted
          if (relative_decrease < 0.25) radius *= 0.5;
          if (relative_decrease > 0.75) radius = std::fmax(radius, 3.0 * dogleg_step_norm);
          mu = std::fmax(min_mu, 2.0 * mu / mu_increase_factor);
          reuse = false;
        } else {
          // StepRejected
          radius *= 0.5;
          reuse = true;
        }
      }
    }
    if (!step_valid) {
      // HandleInvalidStep + DoglegStrategy::StepIsInvalid
      ++invalid;
      if (invalid >= opt->max_num_consecutive_invalid_steps) {
        S.termination = SVIN_TERM_FAILURE;
        break;
      }
      mu *= mu_increase_factor;
      reuse = false;
    }
    last_iter_time = elapsed() - iter_t0;
  }
  S.iterations = iter;
  S.final_cost = x_cost;
  S.final_trust_region_radius = radius;
  S.imu_repropagations = O.imu_redo_total;
  std::memcpy(w->pose_blocks, O.poses.data(), O.poses.size() * sizeof(double));
  std::memcpy(w->speedbias, O.sbs.data(), O.sbs.size() * sizeof(double));
  std::memcpy(w->landmarks, O.lms.data(), O.lms.size() * sizeof(double));
  if (landmark_quality && opt->compute_landmark_quality) {
    // Estimator.cpp:903-922: H = sum J1min^T J1min over the landmark's residuals (no loss), eigenvalues
    std::vector<double> H(9 * (size_t)NL, 0.0);
    for (int o = 0; o < w->num_obs; ++o) {
      double r[2], J1[6];
      reprojection_evaluate(&O.poses[7 * w->obs_pose[o]], &O.lms[4 * w->obs_landmark[o]],
                            &O.poses[7 * w->obs_extrinsics[o]], w->intrinsics + 8 * w->obs_camera[o],
                            w->obs_measurement + 2 * o, &O.obs_U[4 * o], r, nullptr, J1, nullptr);
      double* Hl = &H[9 * (size_t)w->obs_landmark[o]];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Hl[i * 3 + j] += J1[i] * J1[j] + J1[3 + i] * J1[3 + j];
    }
    for (int l = 0; l < NL; ++l) {
      double ev[3];
      sym3_eigenvalues(&H[9 * (size_t)l], ev);
      landmark_quality[l] = (ev[0] < 1.0e-12) ? 0.0 : std::sqrt(ev[0]) / std::sqrt(ev[2]);
    }
  }
  if (sum) *sum = S;
  return 0;
}

// ---- single-term entry points used by the unit tests ---------------------------------
void svin_oracle_reprojection(const double* pose, const double* hp_W, const double* extr, const double* intr,
                              const double* z, const double* information, double* r, double* J0, double* J1,
                              double* J2, int* valid) {
  double U[4];
  sqrt_information(information, U, 2);
  bool v = reprojection_evaluate(pose, hp_W, extr, intr, z, U, r, J0, J1, J2);
  if (valid) *valid = v ? 1 : 0;
}
int svin_oracle_project(const double* intr, const double* p3, double* ip, double* J23, int w, int h) {
  return pinhole_project(intr, p3, ip, J23, w, h);
}
int svin_oracle_backproject(const double* intr, const double* ip, double* dir) {
  return pinhole_backproject(intr, ip, dir) ? 1 : 0;
}
void svin_oracle_imu(int n, const int64_t* t_ns, const double* gyro, const double* accel, const SvinImuParams* p,
                     int64_t t0, int64_t t1, const double* pose0, const double* sb0, const double* pose1,
                     const double* sb1, double* r, double* J0, double* J1, double* J2, double* J3) {
  ImuState S;
  ImuMeasView M{n, t_ns, gyro, accel};
  ImuParams P{p->sigma_g_c, p->sigma_a_c, p->sigma_gw_c, p->sigma_aw_c, p->g, p->g_max, p->a_max};
  imu_evaluate(S, M, P, t0, t1, pose0, sb0, pose1, sb1, r, J0, J1, J2, J3);
}
void svin_oracle_pose_plus(const double* x, const double* delta, double* out) { pose_plus(x, delta, out); }
void svin_oracle_pose_minus(const double* xpd, const double* x, double* delta) { pose_minus(xpd, x, delta); }
void svin_oracle_sqrt_information(const double* info, double* U, int n) { sqrt_information(info, U, n); }
void svin_oracle_pose_error(const double* meas7, const double* info36, const double* pose, double* r, double* J) {
  double U[36];
  sqrt_information(info36, U, 6);
  pose_error_evaluate(meas7, U, pose, r, J);
}
void svin_oracle_relative_pose_error(const double* info36, const double* p0, const double* p1, double* r, double* J0,
                                     double* J1) {
  double U[36];
  sqrt_information(info36, U, 6);
  relative_pose_error_evaluate(U, p0, p1, r, J0, J1);
}
void svin_oracle_sonar_error(double range, double heading, double information, const double* mean, const double* T_SSo,
                             const double* pose, double* r, double* J) {
  sonar_error_evaluate(range, heading, std::sqrt(information), mean, T_SSo, pose, r, J);
}
void svin_oracle_sym3_eigenvalues(const double* A, double* ev) { sym3_eigenvalues(A, ev); }

}  // extern "C"

// ======================================================================================================
// Marginalisation: MarginalizationError::addResidualBlock (MarginalizationError.cpp:126-397), marginalizeOut
// (:463-721: landmark part :556-619 with pseudoInverseSymmSqrt on 3x3 blocks, dense part :621-667) and
// updateErrorComputation (:725-758).  Eigen's SelfAdjointEigenSolver is replaced by a cyclic Jacobi
// eigen-solver (same spectrum / invariants; eigenvector signs and order are not unique in either).
// ======================================================================================================
namespace {

// symmetric eigen-decomposition A = U diag(ev) U^T by cyclic two-sided Jacobi with round-robin pair order
void jacobi_eigh(std::vector<double>& A, int n, std::vector<double>& U, std::vector<double>& ev) {
  U.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) U[(size_t)i * n + i] = 1.0;
  const int m = n + (n & 1);  // players of the round-robin tournament (one dummy when n is odd)
  std::vector<int> pl(m);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) (i == j ? diag : off) += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    if (off <= 1e-30 * diag || off == 0.0) break;
    for (int i = 0; i < m; ++i) pl[i] = i;
    for (int round = 0; round < m - 1; ++round) {
      for (int k = 0; k < m / 2; ++k) {
        int p = pl[k], q = pl[m - 1 - k];
        if (p >= n || q >= n) continue;
        if (p > q) std::swap(p, q);
        const double apq = A[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double theta = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int i = 0; i < n; ++i) {  // columns: A <- A G
          const double aip = A[(size_t)i * n + p], aiq = A[(size_t)i * n + q];
          A[(size_t)i * n + p] = c * aip - s * aiq;
          A[(size_t)i * n + q] = s * aip + c * aiq;
        }
        for (int i = 0; i < n; ++i) {  // rows: A <- G^T A
          const double api = A[(size_t)p * n + i], aqi = A[(size_t)q * n + i];
          A[(size_t)p * n + i] = c * api - s * aqi;
          A[(size_t)q * n + i] = s * api + c * aqi;
        }
        for (int i = 0; i < n; ++i) {  // U <- U G
          const double uip = U[(size_t)i * n + p], uiq = U[(size_t)i * n + q];
          U[(size_t)i * n + p] = c * uip - s * uiq;
          U[(size_t)i * n + q] = s * uip + c * uiq;
        }
      }
      // rotate players 1..m-1
      const int last = pl[m - 1];
      for (int i = m - 1; i > 1; --i) pl[i] = pl[i - 1];
      pl[1] = last;
    }
  }
  ev.resize(n);
  for (int i = 0; i < n; ++i) ev[i] = A[(size_t)i * n + i];
}

// Vp = pseudo-inverse of symmetric V (n x n) with tolerance eps * n * lambda_max (pseudoInverseSymm[Sqrt])
void pinv_symm(const std::vector<double>& V, int n, std::vector<double>& Vp) {
  std::vector<double> A = V, U, ev;
  jacobi_eigh(A, n, U, ev);
  double lmax = -1e300;
  for (double v : ev) lmax = std::max(lmax, v);
  const double tol = std::numeric_limits<double>::epsilon() * n * lmax;
  Vp.assign((size_t)n * n, 0.0);
  for (int k = 0; k < n; ++k) {
    if (!(ev[k] > tol)) continue;
    const double inv = 1.0 / ev[k];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) Vp[(size_t)i * n + j] += U[(size_t)i * n + k] * inv * U[(size_t)j * n + k];
  }
}

}  // namespace

extern "C" int svin_oracle_marginalize(const SvinBaWindow* w, const SvinMargSpec* spec, SvinMargResult* out) {
  Oracle O(w);
  Lin L;
  O.evaluate(O.poses, O.sbs, O.lms, true, &L);  // linearise at the supplied (first-estimate) points, loss-corrected
  const int n = O.n_dense, NL = w->num_landmarks;
  std::vector<double> H((size_t)n * n, 0.0), b(n, 0.0);
  // existing prior, embedded
  {
    std::vector<int> map;
    for (int k = 0; k < spec->prior_num_blocks; ++k) {
      const int kind = spec->prior_block_kind[k], idx = spec->prior_block_index[k];
      const int off = kind == SVIN_BLOCK_POSE ? O.pose_off[idx] : O.sb_off[idx];
      const int dim = kind == SVIN_BLOCK_POSE ? 6 : 9;
      for (int c = 0; c < dim; ++c) map.push_back(off + c);
    }
    for (int i = 0; i < spec->prior_dim; ++i)
      for (int j = 0; j < spec->prior_dim; ++j) H[(size_t)map[i] * n + map[j]] += spec->prior_H[(size_t)i * spec->prior_dim + j];
    for (int i = 0; i < spec->prior_dim; ++i) b[map[i]] += spec->prior_b0[i];
  }
  // dense terms: H += J^T J, b0 -= J^T r   (MarginalizationError.cpp:333-383)
  for (const DenseTerm& t : L.dense)
    for (const DenseBlock& A : t.blocks) {
      if (A.off < 0) continue;
      for (int rr = 0; rr < t.m; ++rr)
        for (int c = 0; c < A.ld; ++c) b[A.off + c] -= A.J[rr * A.ld + c] * t.r[rr];
      for (const DenseBlock& B : t.blocks) {
        if (B.off < 0) continue;
        for (int i = 0; i < A.ld; ++i)
          for (int j = 0; j < B.ld; ++j) {
            double s = 0;
            for (int rr = 0; rr < t.m; ++rr) s += A.J[rr * A.ld + i] * B.J[rr * B.ld + j];
            H[(size_t)(A.off + i) * n + B.off + j] += s;
          }
      }
    }
  // reprojection terms, landmark by landmark, eliminated with the preconditioned 3x3 pseudo-inverse
  std::vector<std::vector<int>> by_lm(NL);
  for (int o = 0; o < w->num_obs; ++o) by_lm[w->obs_landmark[o]].push_back(o);
  for (int l = 0; l < NL; ++l) {
    double V[9] = {0}, bl[3] = {0};
    struct WB { int off; double W[18]; };
    std::vector<WB> Ws;
    for (int o : by_lm[l]) {
      const int offs[2] = {O.pose_off[w->obs_pose[o]], O.pose_off[w->obs_extrinsics[o]]};
      const double* Js[2] = {&L.Jp[12 * o], &L.Je[12 * o]};
      const double* Jl = &L.Jl[6 * o];
      const double r0 = L.r[2 * o], r1 = L.r[2 * o + 1];
      for (int a = 0; a < 2; ++a) {
        if (offs[a] < 0) continue;
        for (int i = 0; i < 6; ++i) b[offs[a] + i] -= Js[a][i] * r0 + Js[a][6 + i] * r1;
        for (int c = 0; c < 2; ++c) {
          if (offs[c] < 0) continue;
          for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j)
              H[(size_t)(offs[a] + i) * n + offs[c] + j] += Js[a][i] * Js[c][j] + Js[a][6 + i] * Js[c][6 + j];
        }
        WB* wb = nullptr;
        for (WB& x : Ws)
          if (x.off == offs[a]) wb = &x;
        if (!wb) {
          Ws.push_back(WB{offs[a], {0}});
          wb = &Ws.back();
        }
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 3; ++j) wb->W[i * 3 + j] += Js[a][i] * Jl[j] + Js[a][6 + i] * Jl[3 + j];
      }
      for (int i = 0; i < 3; ++i) {
        bl[i] -= Jl[i] * r0 + Jl[3 + i] * r1;
        for (int j = 0; j < 3; ++j) V[i * 3 + j] += Jl[i] * Jl[j] + Jl[3 + i] * Jl[3 + j];
      }
    }
    // preconditioner of the landmark columns (MarginalizationError.cpp:558-563), pinv in the scaled space
    double pl[3];
    for (int i = 0; i < 3; ++i) pl[i] = V[i * 3 + i] > 1.0e-9 ? std::sqrt(V[i * 3 + i]) : 1.0e-3;
    std::vector<double> Vs(9), Vp;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Vs[i * 3 + j] = V[i * 3 + j] / (pl[i] * pl[j]);
    pinv_symm(Vs, 3, Vp);
    double Veff[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Veff[i * 3 + j] = Vp[i * 3 + j] / (pl[i] * pl[j]);
    for (const WB& A : Ws) {
      double WV[18];
      mm(A.W, Veff, WV, 6, 3, 3);
      for (int i = 0; i < 6; ++i) b[A.off + i] -= WV[i * 3] * bl[0] + WV[i * 3 + 1] * bl[1] + WV[i * 3 + 2] * bl[2];
      for (const WB& B : Ws)
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j)
            H[(size_t)(A.off + i) * n + B.off + j] -=
                WV[i * 3] * B.W[j * 3] + WV[i * 3 + 1] * B.W[j * 3 + 1] + WV[i * 3 + 2] * B.W[j * 3 + 2];
    }
  }
  // dense part (MarginalizationError.cpp:621-667)
  std::vector<int> keep_idx, marg_idx;
  out->num_blocks = 0;
  for (int i = 0; i < w->num_pose_blocks; ++i)
    if (O.pose_off[i] >= 0) {
      const bool mg = spec->marginalize_pose && spec->marginalize_pose[i];
      for (int c = 0; c < 6; ++c) (mg ? marg_idx : keep_idx).push_back(O.pose_off[i] + c);
      if (!mg) {
        out->block_kind[out->num_blocks] = SVIN_BLOCK_POSE;
        out->block_index[out->num_blocks++] = i;
      }
    }
  for (int i = 0; i < w->num_speedbias; ++i)
    if (O.sb_off[i] >= 0) {
      const bool mg = spec->marginalize_speedbias && spec->marginalize_speedbias[i];
      for (int c = 0; c < 9; ++c) (mg ? marg_idx : keep_idx).push_back(O.sb_off[i] + c);
      if (!mg) {
        out->block_kind[out->num_blocks] = SVIN_BLOCK_SPEEDBIAS;
        out->block_index[out->num_blocks++] = i;
      }
    }
  const int nk = (int)keep_idx.size(), nm = (int)marg_idx.size();
  std::vector<double> Hk((size_t)nk * nk), bk(nk);
  if (nm > 0) {
    std::vector<double> p(n);
    for (int i = 0; i < n; ++i) p[i] = H[(size_t)i * n + i] > 1.0e-9 ? std::sqrt(H[(size_t)i * n + i]) : 1.0e-3;
    std::vector<double> V((size_t)nm * nm), Vp;
    for (int i = 0; i < nm; ++i)
      for (int j = 0; j < nm; ++j) {
        const int gi = marg_idx[i], gj = marg_idx[j];
        V[(size_t)i * nm + j] = 0.5 * (H[(size_t)gi * n + gj] + H[(size_t)gj * n + gi]) / (p[gi] * p[gj]);
      }
    pinv_symm(V, nm, Vp);
    // W (kept x marg) in the scaled space, b_b scaled
    std::vector<double> Wm((size_t)nk * nm), WV((size_t)nk * nm), bb(nm);
    for (int i = 0; i < nk; ++i)
      for (int j = 0; j < nm; ++j) Wm[(size_t)i * nm + j] = H[(size_t)keep_idx[i] * n + marg_idx[j]] / (p[keep_idx[i]] * p[marg_idx[j]]);
    for (int j = 0; j < nm; ++j) bb[j] = b[marg_idx[j]] / p[marg_idx[j]];
    mm(Wm.data(), Vp.data(), WV.data(), nk, nm, nm);
    for (int i = 0; i < nk; ++i) {
      double s = b[keep_idx[i]] / p[keep_idx[i]];
      for (int j = 0; j < nm; ++j) s -= WV[(size_t)i * nm + j] * bb[j];
      bk[i] = s * p[keep_idx[i]];
      for (int j = 0; j < nk; ++j) {
        double h = H[(size_t)keep_idx[i] * n + keep_idx[j]] / (p[keep_idx[i]] * p[keep_idx[j]]);
        for (int k = 0; k < nm; ++k) h -= WV[(size_t)i * nm + k] * Wm[(size_t)j * nm + k];
        Hk[(size_t)i * nk + j] = h * p[keep_idx[i]] * p[keep_idx[j]];
      }
    }
  } else {
    for (int i = 0; i < nk; ++i) {
      bk[i] = b[keep_idx[i]];
      for (int j = 0; j < nk; ++j) Hk[(size_t)i * nk + j] = H[(size_t)keep_idx[i] * n + keep_idx[j]];
    }
  }
  // updateErrorComputation (MarginalizationError.cpp:725-758)
  std::vector<double> p(nk), Hs((size_t)nk * nk), U, ev;
  for (int i = 0; i < nk; ++i) p[i] = Hk[(size_t)i * nk + i] > 1.0e-9 ? std::sqrt(Hk[(size_t)i * nk + i]) : 1.0e-3;
  for (int i = 0; i < nk; ++i)
    for (int j = 0; j < nk; ++j) Hs[(size_t)i * nk + j] = 0.5 * (Hk[(size_t)i * nk + j] + Hk[(size_t)j * nk + i]) / (p[i] * p[j]);
  jacobi_eigh(Hs, nk, U, ev);
  double lmax = -1e300;
  for (double v : ev) lmax = std::max(lmax, v);
  const double tol = std::numeric_limits<double>::epsilon() * nk * lmax;
  out->dim = nk;
  for (int k = 0; k < nk; ++k) {
    const double S = ev[k] > tol ? ev[k] : 0.0;
    const double Sp = ev[k] > tol ? 1.0 / ev[k] : 0.0;
    const double ss = std::sqrt(S), sps = std::sqrt(Sp);
    double e = 0;
    for (int i = 0; i < nk; ++i) {
      out->J[(size_t)k * nk + i] = p[i] * U[(size_t)i * nk + k] * ss;       // J_ = (p U sqrt(S))^T
      e += sps * U[(size_t)i * nk + k] * (1.0 / p[i]) * bk[i];
    }
    out->e0[k] = -e;                                                       // e0_ = -sqrt(S+) U^T p^-1 b0_
  }
  std::memcpy(out->H, Hk.data(), sizeof(double) * Hk.size());
  std::memcpy(out->b0, bk.data(), sizeof(double) * bk.size());
  return 0;
}
