// CUDA kernels of the sliding-window BA engine (sm_100a, fp64).
//
// One batch = many independent windows solved concurrently by the same launches; every
// kernel is keyed by tiles that never straddle a window, and consults the window's
// trust-region state (WinState) to decide whether it has work in this slot.
//
// Kernel family           work item          reference arithmetic it carries
//   k_linearize           observation        ReprojectionError::EvaluateWithMinimalJacobians
//                                            (okvis_ceres/.../implementation/ReprojectionError.hpp:85-229)
//                                            + Ceres loss corrector (restated in-tree at
//                                            okvis_ceres/src/MarginalizationError.cpp:283-330)
//   k_dense_eval          window (CTA)       ImuError (ImuError.cpp:76-263,706-866), PoseError, SpeedAndBiasError,
//                                            RelativePoseError, SonarError, DepthError, MarginalizationError::Evaluate
//   k_schur               landmark           landmark-block elimination (Ceres SchurEliminator; the in-tree
//                                            analogue is MarginalizationError.cpp:556-619)
//   k_dense_solve         window (CTA)       reduced camera system: Jacobi scaling, LM diagonal, Cholesky, solve
//   k_backsub             landmark           landmark back-substitution + Cauchy-point accumulation
//   k_step_dense/k_step_lm                   dogleg interpolation (Ceres DoglegStrategy), candidate x (+) delta
//                                            (PoseManifold.cpp:59-82, HomogeneousPointManifold.cpp:57-66), model cost
//   k_decide              window             Ceres TrustRegionMinimizer accept/reject + termination tests
//   k_quality             landmark           Estimator.cpp:903-922 landmark quality
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <cstdlib>

#include "ba_device_utils.cuh"
#include "ba_kernels.cuh"
#include "ba_math.cuh"

namespace svin {

// ------------------------------------------------------------------------------------------ reprojection
struct Reproj {
  double r0, r1;
  double Jp[12], Jl[6], Je[12];
  double cost;
};

// ReprojectionError::EvaluateWithMinimalJacobians for PinholeCamera<RadialTangentialDistortion>.
// raw=true returns the weighted residual and minimal Jacobians before loss correction.
template <bool WANT_J, bool WANT_E>
__device__ __forceinline__ void reproj_eval(const double* __restrict__ pose, const double* __restrict__ lm,
                                            const double* __restrict__ ext, const double* __restrict__ intr, double zx,
                                            double zy, double u00, double u01, double u11, int loss_type,
                                            double loss_a, bool raw, Reproj& o) {
  const V3 t_WS{pose[0], pose[1], pose[2]};
  const M3 C_SW = m3t(qrot(Q4{pose[3], pose[4], pose[5], pose[6]}));
  const V3 t_SC{ext[0], ext[1], ext[2]};
  const M3 C_CS = m3t(qrot(Q4{ext[3], ext[4], ext[5], ext[6]}));
  const V3 hw{lm[0], lm[1], lm[2]};
  const double w = lm[3];
  const V3 cst = m3v(C_SW, t_WS);
  V3 hS = m3v(C_SW, hw);
  hS.x -= cst.x * w;
  hS.y -= cst.y * w;
  hS.z -= cst.z * w;
  const V3 cct = m3v(C_CS, t_SC);
  V3 hC = m3v(C_CS, hS);
  hC.x -= cct.x * w;
  hC.y -= cct.y * w;
  hC.z -= cct.z * w;
  // projectHomogeneous: flip the head when w < 0 (the Jacobian is NOT flipped back in the reference)
  V3 hd = hC;
  if (w < 0) {
    hd.x = -hd.x;
    hd.y = -hd.y;
    hd.z = -hd.z;
  }
  double kx = 0.0, ky = 0.0;
  double Jh[6] = {0, 0, 0, 0, 0, 0};
  if (!(fabs(hd.z) < 1.0e-12)) {
    const double fu = intr[0], fv = intr[1], cu = intr[2], cv = intr[3];
    const double rz = 1.0 / hd.z, rz2 = rz * rz;
    double d0, d1, D00, D01, D10, D11;
    radtan_distort(intr, hd.x * rz, hd.y * rz, d0, d1, D00, D01, D10, D11);
    kx = fu * d0 + cu;
    ky = fv * d1 + cv;
    if (WANT_J) {
      Jh[0] = fu * D00 * rz;
      Jh[1] = fu * D01 * rz;
      Jh[2] = -fu * (hd.x * D00 + hd.y * D01) * rz2;
      Jh[3] = fv * D10 * rz;
      Jh[4] = fv * D11 * rz;
      Jh[5] = -fv * (hd.x * D10 + hd.y * D11) * rz2;
    }
  }
  const double e0 = zx - kx, e1 = zy - ky;
  double r0 = u00 * e0 + u01 * e1;
  double r1 = u11 * e1;
  bool valid = true;
  if (fabs(w) > 1.0e-8) {
    if (hC.z / w < 0.2) valid = false;
  }
  const double sq = r0 * r0 + r1 * r1;
  double rho0, rho1, rho2;
  loss_eval(loss_type, loss_a, sq, rho0, rho1, rho2);
  o.cost = 0.5 * rho0;
  if (WANT_J) {
    // Jhw = U * Jh
    double Jw[6];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      Jw[c] = u00 * Jh[c] + u01 * Jh[3 + c];
      Jw[3 + c] = u11 * Jh[3 + c];
    }
    double A[6], Bm[6];  // A = Jw * C_CS ; Bm = A * C_SW
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        A[i * 3 + j] = Jw[i * 3] * C_CS.m[j] + Jw[i * 3 + 1] * C_CS.m[3 + j] + Jw[i * 3 + 2] * C_CS.m[6 + j];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        Bm[i * 3 + j] = A[i * 3] * C_SW.m[j] + A[i * 3 + 1] * C_SW.m[3 + j] + A[i * 3 + 2] * C_SW.m[6 + j];
    const double vz = valid ? 1.0 : 0.0;
    const V3 p{hw.x - t_WS.x * w, hw.y - t_WS.y * w, hw.z - t_WS.z * w};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double b0 = Bm[i * 3], b1 = Bm[i * 3 + 1], b2 = Bm[i * 3 + 2];
      o.Jp[i * 6 + 0] = vz * b0 * w;
      o.Jp[i * 6 + 1] = vz * b1 * w;
      o.Jp[i * 6 + 2] = vz * b2 * w;
      o.Jp[i * 6 + 3] = -vz * (b1 * p.z - b2 * p.y);
      o.Jp[i * 6 + 4] = -vz * (-b0 * p.z + b2 * p.x);
      o.Jp[i * 6 + 5] = -vz * (b0 * p.y - b1 * p.x);
      o.Jl[i * 3 + 0] = -vz * b0;
      o.Jl[i * 3 + 1] = -vz * b1;
      o.Jl[i * 3 + 2] = -vz * b2;
    }
    if (WANT_E) {
      const V3 pS{hS.x - t_SC.x * w, hS.y - t_SC.y * w, hS.z - t_SC.z * w};
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double a0 = A[i * 3], a1 = A[i * 3 + 1], a2 = A[i * 3 + 2];
        o.Je[i * 6 + 0] = vz * a0 * w;
        o.Je[i * 6 + 1] = vz * a1 * w;
        o.Je[i * 6 + 2] = vz * a2 * w;
        o.Je[i * 6 + 3] = -vz * (a1 * pS.z - a2 * pS.y);
        o.Je[i * 6 + 4] = -vz * (-a0 * pS.z + a2 * pS.x);
        o.Je[i * 6 + 5] = -vz * (a0 * pS.y - a1 * pS.x);
      }
    }
    if (!raw && loss_type != SVIN_LOSS_NONE) {
      // Corrector (corrector.cc): uses the uncorrected residuals for the Jacobian
      const double sqrt_rho1 = sqrt(rho1);
      double residual_scaling = sqrt_rho1, alpha_sq_norm = 0.0;
      if (!(sq == 0.0 || rho2 <= 0.0)) {
        const double D = 1.0 + 2.0 * sq * rho2 / rho1;
        const double alpha = 1.0 - sqrt(D);
        residual_scaling = sqrt_rho1 / (1 - alpha);
        alpha_sq_norm = alpha / sq;
      }
      if (alpha_sq_norm == 0.0) {
#pragma unroll
        for (int i = 0; i < 12; ++i) o.Jp[i] *= sqrt_rho1;
#pragma unroll
        for (int i = 0; i < 6; ++i) o.Jl[i] *= sqrt_rho1;
        if (WANT_E) {
#pragma unroll
          for (int i = 0; i < 12; ++i) o.Je[i] *= sqrt_rho1;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const double rtj = o.Jp[c] * r0 + o.Jp[6 + c] * r1;
          o.Jp[c] = sqrt_rho1 * (o.Jp[c] - alpha_sq_norm * r0 * rtj);
          o.Jp[6 + c] = sqrt_rho1 * (o.Jp[6 + c] - alpha_sq_norm * r1 * rtj);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double rtj = o.Jl[c] * r0 + o.Jl[3 + c] * r1;
          o.Jl[c] = sqrt_rho1 * (o.Jl[c] - alpha_sq_norm * r0 * rtj);
          o.Jl[3 + c] = sqrt_rho1 * (o.Jl[3 + c] - alpha_sq_norm * r1 * rtj);
        }
        if (WANT_E) {
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            const double rtj = o.Je[c] * r0 + o.Je[6 + c] * r1;
            o.Je[c] = sqrt_rho1 * (o.Je[c] - alpha_sq_norm * r0 * rtj);
            o.Je[6 + c] = sqrt_rho1 * (o.Je[6 + c] - alpha_sq_norm * r1 * rtj);
          }
        }
      }
      r0 *= residual_scaling;
      r1 *= residual_scaling;
    }
  }
  o.r0 = r0;
  o.r1 = r1;
}

__device__ __forceinline__ bool step_is_invalid(const WinState& ws) { return ws.gn_failed || !(-ws.acc_mc > 0.0); }

// which: 0 = linearise the current estimate (initialisation), 1 = the candidate.  raw: evaluation dump.
// COMPACT: the pose-Jacobian planes are not written (Batch::fused, see obs_rJ).
template <bool HAS_EXT, int MINB, bool COMPACT = false>
__global__ void __launch_bounds__(kObsTile, MINB) k_linearize(Batch b, int which, int raw) {
  const int tile = blockIdx.x;
  const int w = b.obs_tile_win[tile];
  const int o = b.obs_tile_begin[tile] + threadIdx.x;
  // the observation record does not depend on the window state: issue its loads together with the state's
  // (one dependent-load level less; the kernel is bound by this latency chain, not by bandwidth or fp64)
  const int oc = min(o, b.NOBS - 1);
  const int ip = b.obs_pose[oc], il = b.obs_lm[oc], ie = b.obs_ext[oc], ic = b.obs_cam[oc];
  const double zx = b.obs_zx[oc], zy = b.obs_zy[oc], u00 = b.obs_u00[oc], u01 = b.obs_u01[oc], u11 = b.obs_u11[oc];
  WinState& ws = b.ws[w];
  if (!raw) {
    if (ws.done) return;
    if (which == 1 && (ws.skip_slot || step_is_invalid(ws))) return;
  }
  const WinDesc& wd = b.win[w];
  const int buf = (which == 0) ? ws.cur : 1 - ws.cur;
  const int sbuf = raw ? ws.cur : buf;  // state to read
  double cost[1] = {0.0};
  if (o < wd.obs_end) {
    Reproj R;
    reproj_eval<true, HAS_EXT>(b.pose[sbuf] + 7 * (size_t)ip, b.lm[sbuf] + 4 * (size_t)il,
                               b.pose[sbuf] + 7 * (size_t)ie, b.intr + 8 * (size_t)ic, zx, zy, u00, u01, u11,
                               wd.loss_type, wd.loss_scale, raw != 0, R);
    cost[0] = R.cost;
    const size_t S = b.obs_stride;
    double* r = b.lin_r[buf];
    r[o] = R.r0;
    r[S + o] = R.r1;
    if (!COMPACT) {
      double* Jp = b.lin_Jp[buf];
#pragma unroll
      for (int k = 0; k < 12; ++k) Jp[k * S + o] = R.Jp[k];
    }
    double* Jl = b.lin_Jl[buf];
#pragma unroll
    for (int k = 0; k < 6; ++k) Jl[k * S + o] = R.Jl[k];
    if (HAS_EXT) {
      double* Je = b.lin_Je[buf];
#pragma unroll
      for (int k = 0; k < 12; ++k) Je[k * S + o] = R.Je[k];
    }
  }
  double* const dst[1] = {(b.shard_acc && !raw) ? &b.shard_acc[kShardAcc * (size_t)w + 11] : &ws.cost_cand};
  block_atomic_add<1, kObsTile>(cost, dst);
}

// ------------------------------------------------------------------------------------------ Schur
// 3x3 SPD inverse through Cholesky (Ceres InvertPSDMatrix, assume_full_rank); NaN on failure.
__device__ __forceinline__ void spd3_inverse(const double* V /*6: 00 01 02 11 12 22*/, double* Vi /*6*/) {
  const double l00 = sqrt(V[0]);
  const double l10 = V[1] / l00, l20 = V[2] / l00;
  const double l11 = sqrt(V[3] - l10 * l10);
  const double l21 = (V[4] - l20 * l10) / l11;
  const double l22 = sqrt(V[5] - l20 * l20 - l21 * l21);
  const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
  const double i10 = -l10 * i00 * i11;
  const double i21 = -l21 * i11 * i22;
  const double i20 = -(l20 * i00 + l21 * i10) * i22;
  // Vinv = Li^T Li
  Vi[0] = i00 * i00 + i10 * i10 + i20 * i20;
  Vi[1] = i10 * i11 + i20 * i21;
  Vi[2] = i20 * i22;
  Vi[3] = i11 * i11 + i21 * i21;
  Vi[4] = i21 * i22;
  Vi[5] = i22 * i22;
}

struct ObsJ {
  double r0, r1;
  double Jp[12], Jl[6];
};
__device__ __forceinline__ void load_obs(const Batch& b, int buf, int o, ObsJ& J) {
  const size_t S = b.obs_stride;
  const double* r = b.lin_r[buf];
  J.r0 = r[o];
  J.r1 = r[S + o];
  const double* Jp = b.lin_Jp[buf];
#pragma unroll
  for (int k = 0; k < 12; ++k) J.Jp[k] = Jp[k * S + o];
  const double* Jl = b.lin_Jl[buf];
#pragma unroll
  for (int k = 0; k < 6; ++k) J.Jl[k] = Jl[k * S + o];
}
__device__ __forceinline__ void load_Je(const Batch& b, int buf, int o, double* Je) {
  const size_t S = b.obs_stride;
  const double* p = b.lin_Je[buf];
#pragma unroll
  for (int k = 0; k < 12; ++k) Je[k] = p[k * S + o];
}

// Compact linearisation (Batch::fused, fixed extrinsics): k_linearize stores r (2) and the loss-corrected landmark
// Jacobian Jl (2x3) only - 64 bytes per observation instead of 160.  The pose Jacobian is a linear image of it,
//   J0_min = sqrt(I) Jh C_CW [ w I, -[p]x ]   and   J1 = -sqrt(I) Jh C_CW     (ReprojectionError.hpp impl:153-170),
// and the loss corrector multiplies both from the left by the same 2x2 matrix, so  Jp = [ -w Jl, Jl [p]x ]  with
// p = hw - t_WS w is rebuilt in registers (18 flops) from the landmark and the pose translation of the state buffer.
// (Recomputing r and J from scratch instead - 56 B per observation - was measured slower: the ~600 fp64 instructions
// per observation cost more issue slots than the loads cost bandwidth; profiles/r2a_*.)
struct LmPoint {
  double x, y, z, w;
};
__device__ __forceinline__ LmPoint load_lm(const Batch& b, int buf, int l) {
  const double* p = b.lm[buf] + 4 * (size_t)l;
  return LmPoint{p[0], p[1], p[2], p[3]};
}
template <bool COMPACT>
__device__ __forceinline__ void obs_rJ(const Batch& b, int buf, int o, const LmPoint& lm, double& r0, double& r1,
                                       double (&Jp)[12], double (&Jl)[6]) {
  const size_t S = b.obs_stride;
  const double* r = b.lin_r[buf];
  r0 = r[o];
  r1 = r[S + o];
  const double* pJl = b.lin_Jl[buf];
#pragma unroll
  for (int k = 0; k < 6; ++k) Jl[k] = pJl[k * S + o];
  if (COMPACT) {
    const double* t = b.pose[buf] + 7 * (size_t)b.obs_pose[o];
    const double px = lm.x - t[0] * lm.w, py = lm.y - t[1] * lm.w, pz = lm.z - t[2] * lm.w;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double j0 = Jl[i * 3], j1 = Jl[i * 3 + 1], j2 = Jl[i * 3 + 2];
      Jp[i * 6 + 0] = -j0 * lm.w;
      Jp[i * 6 + 1] = -j1 * lm.w;
      Jp[i * 6 + 2] = -j2 * lm.w;
      Jp[i * 6 + 3] = j1 * pz - j2 * py;
      Jp[i * 6 + 4] = j2 * px - j0 * pz;
      Jp[i * 6 + 5] = j0 * py - j1 * px;
    }
  } else {
    const double* pJp = b.lin_Jp[buf];
#pragma unroll
    for (int k = 0; k < 12; ++k) Jp[k] = pJp[k * S + o];
  }
}

// W(6x3) += Jd^T (2x6)^T * Jls (2x3)
__device__ __forceinline__ void acc_W(const double* Jd, const double* Jls, double* W) {
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) W[i * 3 + j] += Jd[i] * Jls[j] + Jd[6 + i] * Jls[3 + j];
}
// add block -Z * Wq^T at (offp, offq) of the window's upper-triangular reduced matrix
__device__ __forceinline__ void sub_block(double* H, int n, int offp, int offq, const double* Z, const double* Wq) {
  if (offp == offq) {
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        if (j < i) continue;
        // symmetric sum of both orderings is handled by the caller (it visits each unordered pair once)
        const double v = Z[i * 3] * Wq[j * 3] + Z[i * 3 + 1] * Wq[j * 3 + 1] + Z[i * 3 + 2] * Wq[j * 3 + 2];
        atomicAdd(&H[(size_t)(offp + i) * n + offq + j], -v);
      }
  } else if (offp < offq) {
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const double v = Z[i * 3] * Wq[j * 3] + Z[i * 3 + 1] * Wq[j * 3 + 1] + Z[i * 3 + 2] * Wq[j * 3 + 2];
        atomicAdd(&H[(size_t)(offp + i) * n + offq + j], -v);
      }
  } else {
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const double v = Z[i * 3] * Wq[j * 3] + Z[i * 3 + 1] * Wq[j * 3 + 1] + Z[i * 3 + 2] * Wq[j * 3 + 2];
        atomicAdd(&H[(size_t)(offq + j) * n + offp + i], -v);
      }
  }
}

template <bool HAS_EXT>
__global__ void __launch_bounds__(kLmTile) k_schur(Batch b, SvinBaOptions opt) {
  const int tile = blockIdx.x;
  const int w = b.lm_tile_win[tile];
  WinState& ws = b.ws[w];
  if (ws.done || ws.reuse) return;
  const WinDesc& wd = b.win[w];
  const int l = b.lm_tile_begin[tile] + threadIdx.x;
  if (l >= wd.lm_end) return;
  const int buf = ws.cur;
  const int n = wd.n_dense;
  double* H = b.H + wd.H_off;
  double* g_red = b.g_red + wd.d_off;
  double* g_raw = b.g_raw + wd.d_off;
  double* Hdiag = b.Hdiag + wd.d_off;
  const int ob = b.lm_obs_first[l], oe = ob + b.lm_obs_cnt[l];  // this kernel is only used with stride-1 layouts
  const bool lfix = b.lm_fixed[l] != 0;
  const double mu = ws.mu;

  // ---- pass 1: V = sum Jl^T Jl, bl = sum Jl^T r (unscaled)
  double V[6] = {0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};
  {
    const size_t S = b.obs_stride;
    const double* r = b.lin_r[buf];
    const double* Jl = b.lin_Jl[buf];
    for (int o = ob; o < oe; ++o) {
      const double r0 = r[o], r1 = r[S + o];
      double a[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) a[k] = Jl[k * S + o];
      V[0] += a[0] * a[0] + a[3] * a[3];
      V[1] += a[0] * a[1] + a[3] * a[4];
      V[2] += a[0] * a[2] + a[3] * a[5];
      V[3] += a[1] * a[1] + a[4] * a[4];
      V[4] += a[1] * a[2] + a[4] * a[5];
      V[5] += a[2] * a[2] + a[5] * a[5];
      bl[0] += a[0] * r0 + a[3] * r1;
      bl[1] += a[1] * r0 + a[4] * r1;
      bl[2] += a[2] * r0 + a[5] * r1;
    }
  }
  double s[3] = {1.0, 1.0, 1.0}, Vi[6] = {0, 0, 0, 0, 0, 0}, bs[3] = {0, 0, 0};
  if (!lfix) {
    if (ws.iter == 0 && ws.num_successful == 0 && !ws.invalid) {
      // Jacobi scaling from the first Jacobian: 1 / (1 + ||column||)
      if (opt.jacobi_scaling) {
        s[0] = 1.0 / (1.0 + sqrt(V[0]));
        s[1] = 1.0 / (1.0 + sqrt(V[3]));
        s[2] = 1.0 / (1.0 + sqrt(V[5]));
      }
      b.lm_scale[3 * (size_t)l] = s[0];
      b.lm_scale[3 * (size_t)l + 1] = s[1];
      b.lm_scale[3 * (size_t)l + 2] = s[2];
    } else {
      s[0] = b.lm_scale[3 * (size_t)l];
      s[1] = b.lm_scale[3 * (size_t)l + 1];
      s[2] = b.lm_scale[3 * (size_t)l + 2];
    }
    atomic_max_nonneg(&ws.gmax_bits, fmax(fabs(bl[0]), fmax(fabs(bl[1]), fabs(bl[2]))));
    double Vs[6] = {V[0] * s[0] * s[0], V[1] * s[0] * s[1], V[2] * s[0] * s[2],
                    V[3] * s[1] * s[1], V[4] * s[1] * s[2], V[5] * s[2] * s[2]};
    const double d0 = sqrt(fmin(fmax(Vs[0], opt.min_lm_diagonal), opt.max_lm_diagonal));
    const double d1 = sqrt(fmin(fmax(Vs[3], opt.min_lm_diagonal), opt.max_lm_diagonal));
    const double d2 = sqrt(fmin(fmax(Vs[5], opt.min_lm_diagonal), opt.max_lm_diagonal));
    Vs[0] += mu * d0 * d0;
    Vs[3] += mu * d1 * d1;
    Vs[5] += mu * d2 * d2;
    spd3_inverse(Vs, Vi);
    bs[0] = s[0] * bl[0];
    bs[1] = s[1] * bl[1];
    bs[2] = s[2] * bl[2];
    double* p;
    p = b.lm_Vinv + 6 * (size_t)l;
#pragma unroll
    for (int k = 0; k < 6; ++k) p[k] = Vi[k];
    p = b.lm_bs + 3 * (size_t)l;
    p[0] = bs[0]; p[1] = bs[1]; p[2] = bs[2];
    p = b.lm_diag + 3 * (size_t)l;
    p[0] = d0; p[1] = d1; p[2] = d2;
    p = b.lm_grad + 3 * (size_t)l;
    p[0] = bs[0] / d0; p[1] = bs[1] / d1; p[2] = bs[2] / d2;
  }

  // ---- pass 2: dense blocks and the rank-3 update of the reduced system
  if (!HAS_EXT) {
    int i = ob;
    while (i < oe) {
      const int p = b.obs_pose[i];
      int j = i + 1;
      while (j < oe && b.obs_pose[j] == p) ++j;
      const int offp = b.pose_off[p];
      if (offp >= 0) {
        double Hpp[21], gp[6], W[18];
#pragma unroll
        for (int k = 0; k < 21; ++k) Hpp[k] = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) gp[k] = 0;
#pragma unroll
        for (int k = 0; k < 18; ++k) W[k] = 0;
        for (int o = i; o < j; ++o) {
          ObsJ J;
          load_obs(b, buf, o, J);
          int idx = 0;
#pragma unroll
          for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int c = a; c < 6; ++c) Hpp[idx++] += J.Jp[a] * J.Jp[c] + J.Jp[6 + a] * J.Jp[6 + c];
#pragma unroll
          for (int a = 0; a < 6; ++a) gp[a] += J.Jp[a] * J.r0 + J.Jp[6 + a] * J.r1;
          double Jls[6] = {J.Jl[0] * s[0], J.Jl[1] * s[1], J.Jl[2] * s[2], J.Jl[3] * s[0], J.Jl[4] * s[1], J.Jl[5] * s[2]};
          acc_W(J.Jp, Jls, W);
        }
        {
          int idx = 0;
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            atomicAdd(&Hdiag[offp + a], Hpp[idx]);
            idx += 6 - a;
          }
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) atomicAdd(&g_raw[offp + a], gp[a]);
        double Z[18];
        if (!lfix) {
          // Z = W * Vinv
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            const double w0 = W[a * 3], w1 = W[a * 3 + 1], w2 = W[a * 3 + 2];
            Z[a * 3 + 0] = w0 * Vi[0] + w1 * Vi[1] + w2 * Vi[2];
            Z[a * 3 + 1] = w0 * Vi[1] + w1 * Vi[3] + w2 * Vi[4];
            Z[a * 3 + 2] = w0 * Vi[2] + w1 * Vi[4] + w2 * Vi[5];
          }
          int idx = 0;
#pragma unroll
          for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int c = a; c < 6; ++c)
              Hpp[idx++] -= Z[a * 3] * W[c * 3] + Z[a * 3 + 1] * W[c * 3 + 1] + Z[a * 3 + 2] * W[c * 3 + 2];
#pragma unroll
          for (int a = 0; a < 6; ++a) gp[a] -= Z[a * 3] * bs[0] + Z[a * 3 + 1] * bs[1] + Z[a * 3 + 2] * bs[2];
        }
        {
          int idx = 0;
#pragma unroll
          for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int c = a; c < 6; ++c) atomicAdd(&H[(size_t)(offp + a) * n + offp + c], Hpp[idx++]);
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) atomicAdd(&g_red[offp + a], gp[a]);
        if (!lfix) {
          // cross blocks with every later pose run of this landmark
          int i2 = j;
          while (i2 < oe) {
            const int q = b.obs_pose[i2];
            int j2 = i2 + 1;
            while (j2 < oe && b.obs_pose[j2] == q) ++j2;
            const int offq = b.pose_off[q];
            if (offq >= 0) {
              double Wq[18];
#pragma unroll
              for (int k = 0; k < 18; ++k) Wq[k] = 0;
              const size_t S = b.obs_stride;
              const double* Jpp = b.lin_Jp[buf];
              const double* Jlp = b.lin_Jl[buf];
              for (int o = i2; o < j2; ++o) {
                double Jp[12], Jls[6];
#pragma unroll
                for (int k = 0; k < 12; ++k) Jp[k] = Jpp[k * S + o];
#pragma unroll
                for (int k = 0; k < 6; ++k) Jls[k] = Jlp[k * S + o] * s[k % 3];
                acc_W(Jp, Jls, Wq);
              }
              sub_block(H, n, offp, offq, Z, Wq);
            }
            i2 = j2;
          }
        }
      }
      i = j;
    }
  } else {
    // general path: every observation carries a pose block and an extrinsics block
    for (int o = ob; o < oe; ++o) {
      ObsJ J;
      load_obs(b, buf, o, J);
      double Je[12];
      load_Je(b, buf, o, Je);
      const int offs[2] = {b.pose_off[b.obs_pose[o]], b.pose_off[b.obs_ext[o]]};
      const double* Jd[2] = {J.Jp, Je};
      double Jls[6] = {J.Jl[0] * s[0], J.Jl[1] * s[1], J.Jl[2] * s[2], J.Jl[3] * s[0], J.Jl[4] * s[1], J.Jl[5] * s[2]};
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        if (offs[a] < 0) continue;
        const double* Ja = Jd[a];
        // unreduced diagonal blocks, cross block pose-extrinsics, gradient
        for (int x = 0; x < 6; ++x) {
          atomicAdd(&Hdiag[offs[a] + x], Ja[x] * Ja[x] + Ja[6 + x] * Ja[6 + x]);
          const double gv = Ja[x] * J.r0 + Ja[6 + x] * J.r1;
          atomicAdd(&g_raw[offs[a] + x], gv);
          atomicAdd(&g_red[offs[a] + x], gv);
          for (int y = x; y < 6; ++y)
            atomicAdd(&H[(size_t)(offs[a] + x) * n + offs[a] + y], Ja[x] * Ja[y] + Ja[6 + x] * Ja[6 + y]);
        }
        if (a == 0 && offs[1] >= 0) {
          for (int x = 0; x < 6; ++x)
            for (int y = 0; y < 6; ++y) {
              const double v = J.Jp[x] * Je[y] + J.Jp[6 + x] * Je[6 + y];
              if (offs[0] < offs[1])
                atomicAdd(&H[(size_t)(offs[0] + x) * n + offs[1] + y], v);
              else
                atomicAdd(&H[(size_t)(offs[1] + y) * n + offs[0] + x], v);
            }
        }
        if (lfix) continue;
        double W[18], Z[18];
#pragma unroll
        for (int k = 0; k < 18; ++k) W[k] = 0;
        acc_W(Ja, Jls, W);
#pragma unroll
        for (int x = 0; x < 6; ++x) {
          const double w0 = W[x * 3], w1 = W[x * 3 + 1], w2 = W[x * 3 + 2];
          Z[x * 3 + 0] = w0 * Vi[0] + w1 * Vi[1] + w2 * Vi[2];
          Z[x * 3 + 1] = w0 * Vi[1] + w1 * Vi[3] + w2 * Vi[4];
          Z[x * 3 + 2] = w0 * Vi[2] + w1 * Vi[4] + w2 * Vi[5];
        }
        for (int x = 0; x < 6; ++x)
          atomicAdd(&g_red[offs[a] + x], -(Z[x * 3] * bs[0] + Z[x * 3 + 1] * bs[1] + Z[x * 3 + 2] * bs[2]));
        // pair with every (observation, block) of this landmark, each unordered pair of distinct
        // entries once and the self pair once; same-offset pairs of distinct entries need both orderings
        for (int o2 = ob; o2 < oe; ++o2) {
          double Jp2[12], Je2[12], Jl2[6];
          {
            const size_t S = b.obs_stride;
            const double* pp = b.lin_Jp[buf];
            const double* pl = b.lin_Jl[buf];
#pragma unroll
            for (int k = 0; k < 12; ++k) Jp2[k] = pp[k * S + o2];
#pragma unroll
            for (int k = 0; k < 6; ++k) Jl2[k] = pl[k * S + o2] * s[k % 3];
            load_Je(b, buf, o2, Je2);
          }
          const int offs2[2] = {b.pose_off[b.obs_pose[o2]], b.pose_off[b.obs_ext[o2]]};
          for (int a2 = 0; a2 < 2; ++a2) {
            if (offs2[a2] < 0) continue;
            const bool same_entry = (o2 == o && a2 == a);
            // visit ordered pairs (entry, entry2) with entry <= entry2 in (o, a) lexicographic order
            if (o2 < o || (o2 == o && a2 < a)) continue;
            double W2[18];
#pragma unroll
            for (int k = 0; k < 18; ++k) W2[k] = 0;
            acc_W(a2 == 0 ? Jp2 : Je2, Jl2, W2);
            if (same_entry) {
              sub_block(H, n, offs[a], offs2[a2], Z, W2);
            } else if (offs[a] == offs2[a2]) {
              // two distinct entries on the same dense block: Z W2^T + (Z W2^T)^T on the upper triangle
              for (int x = 0; x < 6; ++x)
                for (int y = x; y < 6; ++y) {
                  const double v1 = Z[x * 3] * W2[y * 3] + Z[x * 3 + 1] * W2[y * 3 + 1] + Z[x * 3 + 2] * W2[y * 3 + 2];
                  const double v2 = Z[y * 3] * W2[x * 3] + Z[y * 3 + 1] * W2[x * 3 + 1] + Z[y * 3 + 2] * W2[x * 3 + 2];
                  atomicAdd(&H[(size_t)(offs[a] + x) * n + offs[a] + y], -(v1 + v2));
                }
            } else {
              sub_block(H, n, offs[a], offs2[a2], Z, W2);
            }
          }
        }
      }
    }
  }
}

// Tensor-core landmark elimination (fixed extrinsics): one warp per chunk of <= 32 landmarks that share the exact
// (pose block, camera) observation pattern, so that the pose runs are warp-uniform and the sums over the landmarks
// of a chunk become the K dimension of fp64 tensor-core contractions (mma.sync.m8n8k4.f64, DMMA):
//   per run a and observation:  P (8 x 2c) = [Jp^T ; r^T ; 0] columns (row, landmark)  ->  P P^T accumulates
//        [ Jp^T Jp   Jp^T r ]      the pose-pose block, the raw gradient and diag(J^T J)
//        [ r^T Jp    r^T r  ]
//   per chunk:  Y ((6k+1) x 3c), rows 6a..6a+5 = W_a M^T with V^-1 = M^T M (M = inverse Cholesky factor of the
//        damped landmark block), last row u = M b_l  ->  Y Y^T = [ W V^-1 W^T   W V^-1 b_l ; ... ]
//        i.e. the whole Schur update of the k x k pose blocks and the reduced-gradient correction in one product.
// One fp64 RED per produced entry.  The chunk size is capped at upload so that Y fits kSchurYDoubles.
constexpr int kSchurYDoubles = 1680;   // Y operand per warp (3 CTAs of 4 warps per SM: 3 x 76.3 KB shared memory)
constexpr int kSchurPDoubles = 640;    // P tiles: G x 8 x (2 * pad4(cnt) + 4) doubles, G * pad4(cnt) <= 32
constexpr int kSchurMaxRuns = 64;      // run descriptors per warp: 2 ints each
constexpr int kSchurWarpDoubles = kSchurYDoubles + kSchurPDoubles + kSchurMaxRuns;

__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double bq) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(bq));
}

// spd3_inverse that also returns the inverse Cholesky factor M (lower: 00 10 11 20 21 22), V^-1 = M^T M
__device__ __forceinline__ void spd3_inverse_factor(const double* V, double* Vi, double* M) {
  const double l00 = sqrt(V[0]);
  const double l10 = V[1] / l00, l20 = V[2] / l00;
  const double l11 = sqrt(V[3] - l10 * l10);
  const double l21 = (V[4] - l20 * l10) / l11;
  const double l22 = sqrt(V[5] - l20 * l20 - l21 * l21);
  const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
  const double i10 = -l10 * i00 * i11;
  const double i21 = -l21 * i11 * i22;
  const double i20 = -(l20 * i00 + l21 * i10) * i22;
  Vi[0] = i00 * i00 + i10 * i10 + i20 * i20;
  Vi[1] = i10 * i11 + i20 * i21;
  Vi[2] = i20 * i22;
  Vi[3] = i11 * i11 + i21 * i21;
  Vi[4] = i21 * i22;
  Vi[5] = i22 * i22;
  M[0] = i00; M[1] = i10; M[2] = i11; M[3] = i20; M[4] = i21; M[5] = i22;
}

// One chunk on one warp.  G lanes share a landmark when the chunk holds <= 32 / G landmarks: lane = j * CG + i
// (i = landmark slot, j = sub-lane).  The sub-lanes split the observations in pass 1 (then xor-reduce) and the pose
// runs in pass 2 (run a = t * G + j in round t, one P tile per sub-lane), so that the rare long-track patterns -
// few landmarks, many runs - keep all 32 lanes busy instead of 2..8.
template <int G, bool FUSED>
__device__ __forceinline__ void schur_chunk(const Batch& b, const SvinBaOptions& opt, WinState& ws, const WinDesc& wd,
                                            int chunk, int cnt, int nr, double* Ys, double* Ps, const int* rdesc,
                                            int lane) {
  constexpr int CG = 32 / G;
  const int i = lane & (CG - 1), j = lane / CG;
  const bool active = i < cnt;
  const int l = b.sw_lm_begin[chunk] + (active ? i : 0);
  const int buf = ws.cur;
  const int n = wd.n_dense;
  double* H = b.H + wd.H_off;
  double* g_red = b.g_red + wd.d_off;
  double* g_raw = b.g_raw + wd.d_off;
  double* Hdiag = b.Hdiag + wd.d_off;
  const int ob = b.lm_obs_first[l];
  const int ost = b.lm_obs_stride[l];
  const int nobs = b.lm_obs_cnt[l];
  const LmPoint lmp = load_lm(b, buf, l);
  const bool lfix = b.lm_fixed[l] != 0;
  const double mu = ws.mu;
  const size_t S = b.obs_stride;
  const double* rP = b.lin_r[buf];
  const double* JlP = b.lin_Jl[buf];
  const int cntp = (cnt + 3) & ~3;    // landmark columns of a P tile per residual row
  const int K4p = cntp >> 1;          // k-steps over its 2 * cntp columns
  const int ldp = 2 * cntp + 4;
  const int K4 = (3 * cnt + 3) >> 2;  // k-steps over the 3 * cnt columns of Y
  const int ld = 4 * K4 + 4;
  for (int e = lane; e < G * ldp; e += 32) Ps[(e / ldp) * 8 * ldp + 7 * ldp + (e % ldp)] = 0.0;  // row 7 of every tile

  // ---- pass 1: V = sum Jl^T Jl, bl = sum Jl^T r (unscaled); the sub-lanes take every G-th observation
  double V[6] = {0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};
  for (int k = j; k < nobs; k += 2 * G) {  // two observations in flight per trip
    const int o = ob + k * ost;
    const bool two = k + G < nobs;
    const int ob2 = two ? o + G * ost : o;
    const double r0 = rP[o], r1 = rP[S + o];
    const double q0 = rP[ob2], q1 = rP[S + ob2];
    double a[6], c[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) a[q] = JlP[q * S + o];
#pragma unroll
    for (int q = 0; q < 6; ++q) c[q] = JlP[q * S + ob2];
    const double w2 = two ? 1.0 : 0.0;
    V[0] += a[0] * a[0] + a[3] * a[3];
    V[1] += a[0] * a[1] + a[3] * a[4];
    V[2] += a[0] * a[2] + a[3] * a[5];
    V[3] += a[1] * a[1] + a[4] * a[4];
    V[4] += a[1] * a[2] + a[4] * a[5];
    V[5] += a[2] * a[2] + a[5] * a[5];
    bl[0] += a[0] * r0 + a[3] * r1;
    bl[1] += a[1] * r0 + a[4] * r1;
    bl[2] += a[2] * r0 + a[5] * r1;
    V[0] += w2 * (c[0] * c[0] + c[3] * c[3]);
    V[1] += w2 * (c[0] * c[1] + c[3] * c[4]);
    V[2] += w2 * (c[0] * c[2] + c[3] * c[5]);
    V[3] += w2 * (c[1] * c[1] + c[4] * c[4]);
    V[4] += w2 * (c[1] * c[2] + c[4] * c[5]);
    V[5] += w2 * (c[2] * c[2] + c[5] * c[5]);
    bl[0] += w2 * (c[0] * q0 + c[3] * q1);
    bl[1] += w2 * (c[1] * q0 + c[4] * q1);
    bl[2] += w2 * (c[2] * q0 + c[5] * q1);
  }
  if (G > 1) {
#pragma unroll
    for (int o2 = CG; o2 < 32; o2 <<= 1) {
#pragma unroll
      for (int q = 0; q < 6; ++q) V[q] += __shfl_xor_sync(0xffffffffu, V[q], o2);
#pragma unroll
      for (int q = 0; q < 3; ++q) bl[q] += __shfl_xor_sync(0xffffffffu, bl[q], o2);
    }
  }
  double s[3] = {1.0, 1.0, 1.0}, M[6] = {0, 0, 0, 0, 0, 0}, u[3] = {0, 0, 0};
  if (!lfix) {
    const bool owner = active && j == 0;
    if (ws.iter == 0 && ws.num_successful == 0 && !ws.invalid) {
      if (opt.jacobi_scaling) {
        s[0] = 1.0 / (1.0 + sqrt(V[0]));
        s[1] = 1.0 / (1.0 + sqrt(V[3]));
        s[2] = 1.0 / (1.0 + sqrt(V[5]));
      }
      if (owner) {
        b.lm_scale[3 * (size_t)l] = s[0];
        b.lm_scale[3 * (size_t)l + 1] = s[1];
        b.lm_scale[3 * (size_t)l + 2] = s[2];
      }
    } else {
      s[0] = b.lm_scale[3 * (size_t)l];
      s[1] = b.lm_scale[3 * (size_t)l + 1];
      s[2] = b.lm_scale[3 * (size_t)l + 2];
    }
    double gm = active ? fmax(fabs(bl[0]), fmax(fabs(bl[1]), fabs(bl[2]))) : 0.0;
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, o2));
    if (lane == 0) atomic_max_nonneg(&ws.gmax_bits, gm);
    double Vs[6] = {V[0] * s[0] * s[0], V[1] * s[0] * s[1], V[2] * s[0] * s[2],
                    V[3] * s[1] * s[1], V[4] * s[1] * s[2], V[5] * s[2] * s[2]};
    const double d0 = sqrt(fmin(fmax(Vs[0], opt.min_lm_diagonal), opt.max_lm_diagonal));
    const double d1 = sqrt(fmin(fmax(Vs[3], opt.min_lm_diagonal), opt.max_lm_diagonal));
    const double d2 = sqrt(fmin(fmax(Vs[5], opt.min_lm_diagonal), opt.max_lm_diagonal));
    Vs[0] += mu * d0 * d0;
    Vs[3] += mu * d1 * d1;
    Vs[5] += mu * d2 * d2;
    double Vi[6];
    spd3_inverse_factor(Vs, Vi, M);
    const double bs0 = s[0] * bl[0], bs1 = s[1] * bl[1], bs2 = s[2] * bl[2];
    u[0] = M[0] * bs0;
    u[1] = M[1] * bs0 + M[2] * bs1;
    u[2] = M[3] * bs0 + M[4] * bs1 + M[5] * bs2;
    if (owner) {
      double* p = b.lm_Vinv + 6 * (size_t)l;
#pragma unroll
      for (int k = 0; k < 6; ++k) p[k] = Vi[k];
      p = b.lm_bs + 3 * (size_t)l;
      p[0] = bs0; p[1] = bs1; p[2] = bs2;
      p = b.lm_diag + 3 * (size_t)l;
      p[0] = d0; p[1] = d1; p[2] = d2;
      p = b.lm_grad + 3 * (size_t)l;
      p[0] = bs0 / d0; p[1] = bs1 / d1; p[2] = bs2 / d2;
    }
  }
  __syncwarp();

  const int fr = lane >> 2, fc = lane & 3;
  double* Pt = Ps + j * 8 * ldp;  // this sub-lane's tile
  // ---- pass 2: per pose run: P P^T on the tensor cores -> diagonal block / gradients; Y rows -> shared memory
  for (int t = 0; t * G < nr; ++t) {
    const int a = t * G + j;
    const bool vr = a < nr;
    const int offp = vr ? rdesc[2 * a] : -1;
    const int k0 = vr ? (rdesc[2 * a + 1] >> 8) : 0, m = vr ? (rdesc[2 * a + 1] & 255) : 0;
    int mmax = 0;  // warp-uniform: most observations of any estimated run of this round
#pragma unroll
    for (int jj = 0; jj < G; ++jj) {
      const int a2 = t * G + jj;
      if (a2 < nr && rdesc[2 * a2] >= 0) mmax = max(mmax, rdesc[2 * a2 + 1] & 255);
    }
    double W[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) W[k] = 0;
    double c[G][2];
#pragma unroll
    for (int jj = 0; jj < G; ++jj) c[jj][0] = c[jj][1] = 0.0;
    for (int q = 0; q < mmax; ++q) {
      if (offp >= 0 && q < m) {
        const int o = ob + (k0 + q) * ost;
        double Jp[12], Jls[6], r0, r1;
        obs_rJ<FUSED>(b, buf, o, lmp, r0, r1, Jp, Jls);
#pragma unroll
        for (int e = 0; e < 6; ++e) Jls[e] *= s[e % 3];
        acc_W(Jp, Jls, W);
        if (i < cntp) {
          const double wgt = active ? 1.0 : 0.0;
#pragma unroll
          for (int e = 0; e < 6; ++e) {
            Pt[e * ldp + i] = wgt * Jp[e];
            Pt[e * ldp + cntp + i] = wgt * Jp[6 + e];
          }
          Pt[6 * ldp + i] = wgt * r0;
          Pt[6 * ldp + cntp + i] = wgt * r1;
        }
      } else if (i < cntp) {
#pragma unroll
        for (int e = 0; e < 7; ++e) {
          Pt[e * ldp + i] = 0.0;
          Pt[e * ldp + cntp + i] = 0.0;
        }
      }
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < G; ++jj) {
        const int a2 = t * G + jj;
        if (a2 < nr && rdesc[2 * a2] >= 0 && q < (rdesc[2 * a2 + 1] & 255)) {  // warp-uniform
          // two independent accumulator pairs (K4p is even): the DMMA chain is latency-, not throughput-limited
          const double* pa = Ps + jj * 8 * ldp + fr * ldp + fc;
          double e0 = 0.0, e1 = 0.0;
          for (int ks = 0; ks < K4p; ks += 2) {
            const double x = pa[4 * ks], y = pa[4 * ks + 4];
            dmma8x8x4(c[jj][0], c[jj][1], x, x);
            dmma8x8x4(e0, e1, y, y);
          }
          c[jj][0] += e0;
          c[jj][1] += e1;
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int jj = 0; jj < G; ++jj) {
      const int a2 = t * G + jj;
      if (a2 >= nr) continue;
      const int off2 = rdesc[2 * a2];
      if (off2 < 0 || fr >= 6) continue;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = 2 * fc + e;
        const double val = e ? c[jj][1] : c[jj][0];
        if (col < 6) {
          if (fr <= col) atomicAdd(&H[(size_t)(off2 + fr) * n + off2 + col], val);
          if (fr == col) atomicAdd(&Hdiag[off2 + fr], val);
        } else if (col == 6) {
          atomicAdd(&g_red[off2 + fr], val);
          atomicAdd(&g_raw[off2 + fr], val);
        }
      }
    }
    if (vr && !lfix && active) {
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        const double w0 = W[e * 3], w1 = W[e * 3 + 1], w2 = W[e * 3 + 2];
        double* yr = Ys + (6 * a + e) * ld + 3 * i;
        yr[0] = w0 * M[0];
        yr[1] = w0 * M[1] + w1 * M[2];
        yr[2] = w0 * M[3] + w1 * M[4] + w2 * M[5];
      }
    }
  }
  if (lfix) return;
  const int R = 6 * nr, RY = R + 1, T = (RY + 7) >> 3;
  if (active && j == 0) {
    double* yr = Ys + R * ld + 3 * i;
    yr[0] = u[0]; yr[1] = u[1]; yr[2] = u[2];
  }
  // zero the k padding: columns [3 cnt, 4 K4) of every row
  for (int r = lane; r < RY; r += 32)
    for (int cc = 3 * cnt; cc < 4 * K4; ++cc) Ys[r * ld + cc] = 0.0;
  // row -> dense index of the reduced system (-1: fixed block); lives in the P tiles, which are free now
  int* rowmap = reinterpret_cast<int*>(Ps);
  for (int r = lane; r < R; r += 32) {
    const int op = rdesc[2 * (r / 6)];
    rowmap[r] = op < 0 ? -1 : op + r % 6;
  }
  __syncwarp();
  // ---- C = Y Y^T on the tensor cores, upper tiles only; one RED per upper-triangular entry.
  // Rows >= RY of the last tile read whatever follows in this warp's shared memory: entry (gi, gj) depends on rows
  // gi and gj only and those entries are never written out.
  for (int tm = 0; tm < T; ++tm) {
    const int gi = 8 * tm + fr;
    const double* za = Ys + (gi < RY ? gi : 0) * ld + fc;
    const int ri = gi < R ? rowmap[gi] : -1;
    for (int tn = tm; tn < T; ++tn) {
      const int rb = 8 * tn + fr;
      const double* zb = Ys + (rb < RY ? rb : 0) * ld + fc;
      double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
      int ks = 0;
      for (; ks + 1 < K4; ks += 2) {
        dmma8x8x4(c0, c1, za[4 * ks], zb[4 * ks]);
        dmma8x8x4(e0, e1, za[4 * ks + 4], zb[4 * ks + 4]);
      }
      if (ks < K4) dmma8x8x4(c0, c1, za[4 * ks], zb[4 * ks]);
      c0 += e0;
      c1 += e1;
      if (ri < 0) continue;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gj = 8 * tn + 2 * fc + e;
        if (gj > R || gi > gj) continue;
        const double val = e ? c1 : c0;
        if (gj == R) {
          atomicAdd(&g_red[ri], -val);
        } else {
          const int rj = rowmap[gj];
          if (rj >= 0) atomicAdd(&H[min(ri, rj) * n + max(ri, rj)], -val);
        }
      }
    }
  }
}

// One kernel per lane mapping (chunks are listed by class at upload): keeps each kernel's SASS small enough for
// the instruction cache - the three mappings in one kernel were 120 KB and 18 % of the stalls were "no instruction".
template <int G, bool FUSED>
__global__ void __launch_bounds__(128, 3) k_schur_mma(Batch b, SvinBaOptions opt, const int* list, int count) {
  extern __shared__ double sm_all[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int slot = blockIdx.x * 4 + wid;
  if (slot >= count) return;
  const int chunk = list[slot];
  double* Ys = sm_all + (size_t)wid * kSchurWarpDoubles;
  double* Ps = Ys + kSchurYDoubles;
  int* rdesc = reinterpret_cast<int*>(Ps + kSchurPDoubles);  // [2r] dense offset of the run's pose block, [2r+1] k0<<8|m
  const int w = b.sw_win[chunk];
  WinState& ws = b.ws[w];
  if (ws.done || ws.reuse) return;
  const WinDesc& wd = b.win[w];
  const int cnt = b.sw_count[chunk];
  const int nr = b.sw_nruns[chunk];
  const int rf = b.sw_run_first[chunk];
  for (int r = lane; r < nr; r += 32) {
    rdesc[2 * r] = b.run_off[rf + r];
    rdesc[2 * r + 1] = b.run_k0m[rf + r];
  }
  __syncwarp();
  schur_chunk<G, FUSED>(b, opt, ws, wd, chunk, cnt, nr, Ys, Ps, rdesc, lane);
}

// ---- low-fill chunks: lanes over (pose run, landmark) pairs -------------------------------------------------
// Long tracks come in many distinct patterns with a handful of landmarks each (BASELINE configs[1]: 6 % of the
// landmarks, a third of the chunks, and - with the lane = landmark mappings above - 60 % of the instructions).
// Here a chunk is capped at upload to c <= 32 / runs landmarks and lane = a * c + i works on run a of landmark i:
// every run of every landmark is loaded and contracted at once (one load latency per chunk instead of one per
// run), the per-landmark sums (V, b) and the per-run sums (J_p^T J_p, J_p^T r) are finished through shared memory,
// and only the Y Y^T product runs on the tensor cores as above.
// One CTA of WPC warps per chunk, runs * c <= 32 * WPC units (thread = unit).
template <int WPC>
struct LrCfg {
  // Y operand (6 runs + 1) x (pad4(3 c) + 4), aliased by the reduction scratch (27 doubles per unit)
  static constexpr int kY = WPC == 1 ? 1600 : (WPC == 2 ? 2400 : 3456);
  static constexpr int kIntDoubles = 128;  // 64 ints of run descriptors + 192 ints of row map
  static constexpr int kDoubles = kY + kIntDoubles;
};

// WSM = resident warps per SM the register allocation aims at (12 -> 168 registers, 16 -> 128 with a few spills)
template <int WPC, int WSM = 12, bool FUSED = true>
__global__ void __launch_bounds__(32 * WPC, WSM / WPC) k_schur_lr(Batch b, SvinBaOptions opt, const int* list) {
  extern __shared__ double sm_all[];
  constexpr int NT = 32 * WPC;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, tid = threadIdx.x;
  const int chunk = list[blockIdx.x];
  double* Ys = sm_all;
  int* rdesc = reinterpret_cast<int*>(Ys + LrCfg<WPC>::kY);
  int* rowmap = rdesc + 64;
  const int w = b.sw_win[chunk];
  WinState& ws = b.ws[w];
  if (ws.done || ws.reuse) return;
  const WinDesc& wd = b.win[w];
  const int cnt = b.sw_count[chunk];
  const int nr = b.sw_nruns[chunk];
  const int rf = b.sw_run_first[chunk];
  if (tid < nr) {
    rdesc[2 * tid] = b.run_off[rf + tid];
    rdesc[2 * tid + 1] = b.run_k0m[rf + tid];
  }
  const bool active = tid < cnt * nr;
  const int a = active ? tid / cnt : 0;
  const int i = active ? tid - a * cnt : 0;
  const int l = b.sw_lm_begin[chunk] + i;
  const int buf = ws.cur;
  const int n = wd.n_dense;
  double* H = b.H + wd.H_off;
  double* g_red = b.g_red + wd.d_off;
  double* g_raw = b.g_raw + wd.d_off;
  double* Hdiag = b.Hdiag + wd.d_off;
  const int ob = b.lm_obs_first[l];
  const int ost = b.lm_obs_stride[l];
  const LmPoint lmp = load_lm(b, buf, l);
  const bool lfix = b.lm_fixed[l] != 0;
  const double mu = ws.mu;
  __syncthreads();
  const int offp = rdesc[2 * a];
  const int k0 = rdesc[2 * a + 1] >> 8, m = active ? (rdesc[2 * a + 1] & 255) : 0;

  // ---- this lane's run: partial landmark block, W, diagonal pose block and gradients
  double V[6] = {0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};
  double W[18];
  double D[27];  // upper triangle of J_p^T J_p (21, row-major) | J_p^T r (6)
#pragma unroll
  for (int k = 0; k < 18; ++k) W[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 27; ++k) D[k] = 0.0;
  for (int q = 0; q < m; ++q) {
    const int o = ob + (k0 + q) * ost;
    double Jp[12], Jl[6], r0, r1;
    obs_rJ<FUSED>(b, buf, o, lmp, r0, r1, Jp, Jl);
    V[0] += Jl[0] * Jl[0] + Jl[3] * Jl[3];
    V[1] += Jl[0] * Jl[1] + Jl[3] * Jl[4];
    V[2] += Jl[0] * Jl[2] + Jl[3] * Jl[5];
    V[3] += Jl[1] * Jl[1] + Jl[4] * Jl[4];
    V[4] += Jl[1] * Jl[2] + Jl[4] * Jl[5];
    V[5] += Jl[2] * Jl[2] + Jl[5] * Jl[5];
    bl[0] += Jl[0] * r0 + Jl[3] * r1;
    bl[1] += Jl[1] * r0 + Jl[4] * r1;
    bl[2] += Jl[2] * r0 + Jl[5] * r1;
    if (offp >= 0) {
      acc_W(Jp, Jl, W);
      int k = 0;
#pragma unroll
      for (int x = 0; x < 6; ++x)
#pragma unroll
        for (int y = x; y < 6; ++y) D[k++] += Jp[x] * Jp[y] + Jp[6 + x] * Jp[6 + y];
#pragma unroll
      for (int x = 0; x < 6; ++x) D[21 + x] += Jp[x] * r0 + Jp[6 + x] * r1;
    }
  }
  // ---- per-run sums over the chunk's landmarks -> reduced system (diagonal block, gradients, column norms)
#pragma unroll
  for (int k = 0; k < 27; ++k) Ys[tid * 27 + k] = D[k];
  __syncthreads();
  for (int e = tid; e < nr * 27; e += NT) {
    const int a2 = e / 27, k = e - 27 * a2;
    const int off2 = rdesc[2 * a2];
    if (off2 < 0) continue;
    const double* src = Ys + (size_t)(a2 * cnt) * 27 + k;
    double val = 0.0;
    for (int i2 = 0; i2 < cnt; ++i2) val += src[i2 * 27];
    if (k >= 21) {
      atomicAdd(&g_red[off2 + k - 21], val);
      atomicAdd(&g_raw[off2 + k - 21], val);
    } else {
      // k -> (row, col) of the upper triangle, rows start at 0 6 11 15 18 20
      const int fr = (k >= 6) + (k >= 11) + (k >= 15) + (k >= 18) + (k >= 20);
      const int fc = k - (fr * (13 - fr)) / 2 + fr;
      atomicAdd(&H[(size_t)(off2 + fr) * n + off2 + fc], val);
      if (fr == fc) atomicAdd(&Hdiag[off2 + fr], val);
    }
  }
  __syncthreads();
  // ---- per-landmark sums over the runs: V, b (every unit of a landmark ends up with the full sums)
#pragma unroll
  for (int k = 0; k < 6; ++k) Ys[tid * 9 + k] = V[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) Ys[tid * 9 + 6 + k] = bl[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 6; ++k) V[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) bl[k] = 0.0;
  for (int a2 = 0; a2 < nr; ++a2) {
    const double* src = Ys + (size_t)(a2 * cnt + i) * 9;
#pragma unroll
    for (int k = 0; k < 6; ++k) V[k] += src[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) bl[k] += src[6 + k];
  }
  __syncthreads();
  if (lfix) return;  // chunk-uniform: fixed landmarks are not eliminated
  double s[3] = {1.0, 1.0, 1.0}, M[6], u[3];
  const bool owner = active && a == 0;
  if (ws.iter == 0 && ws.num_successful == 0 && !ws.invalid) {
    if (opt.jacobi_scaling) {
      s[0] = 1.0 / (1.0 + sqrt(V[0]));
      s[1] = 1.0 / (1.0 + sqrt(V[3]));
      s[2] = 1.0 / (1.0 + sqrt(V[5]));
    }
    if (owner) {
      b.lm_scale[3 * (size_t)l] = s[0];
      b.lm_scale[3 * (size_t)l + 1] = s[1];
      b.lm_scale[3 * (size_t)l + 2] = s[2];
    }
  } else {
    s[0] = b.lm_scale[3 * (size_t)l];
    s[1] = b.lm_scale[3 * (size_t)l + 1];
    s[2] = b.lm_scale[3 * (size_t)l + 2];
  }
  {
    double gm = owner ? fmax(fabs(bl[0]), fmax(fabs(bl[1]), fabs(bl[2]))) : 0.0;
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, o2));
    if (lane == 0) atomic_max_nonneg(&ws.gmax_bits, gm);
    double Vs[6] = {V[0] * s[0] * s[0], V[1] * s[0] * s[1], V[2] * s[0] * s[2],
                    V[3] * s[1] * s[1], V[4] * s[1] * s[2], V[5] * s[2] * s[2]};
    const double d0 = sqrt(fmin(fmax(Vs[0], opt.min_lm_diagonal), opt.max_lm_diagonal));
    const double d1 = sqrt(fmin(fmax(Vs[3], opt.min_lm_diagonal), opt.max_lm_diagonal));
    const double d2 = sqrt(fmin(fmax(Vs[5], opt.min_lm_diagonal), opt.max_lm_diagonal));
    Vs[0] += mu * d0 * d0;
    Vs[3] += mu * d1 * d1;
    Vs[5] += mu * d2 * d2;
    double Vi[6];
    spd3_inverse_factor(Vs, Vi, M);
    const double bs0 = s[0] * bl[0], bs1 = s[1] * bl[1], bs2 = s[2] * bl[2];
    u[0] = M[0] * bs0;
    u[1] = M[1] * bs0 + M[2] * bs1;
    u[2] = M[3] * bs0 + M[4] * bs1 + M[5] * bs2;
    if (owner) {
      double* p = b.lm_Vinv + 6 * (size_t)l;
#pragma unroll
      for (int k = 0; k < 6; ++k) p[k] = Vi[k];
      p = b.lm_bs + 3 * (size_t)l;
      p[0] = bs0; p[1] = bs1; p[2] = bs2;
      p = b.lm_diag + 3 * (size_t)l;
      p[0] = d0; p[1] = d1; p[2] = d2;
      p = b.lm_grad + 3 * (size_t)l;
      p[0] = bs0 / d0; p[1] = bs1 / d1; p[2] = bs2 / d2;
    }
  }
  // ---- Y = [W_a diag(s) M^T ; (M b)^T]
  const int K4 = (3 * cnt + 3) >> 2;
  const int ld = 4 * K4 + 4;
  const int R = 6 * nr, RY = R + 1, T = (RY + 7) >> 3;
  if (active) {
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      const double w0 = W[e * 3] * s[0], w1 = W[e * 3 + 1] * s[1], w2 = W[e * 3 + 2] * s[2];
      double* yr = Ys + (6 * a + e) * ld + 3 * i;
      yr[0] = w0 * M[0];
      yr[1] = w0 * M[1] + w1 * M[2];
      yr[2] = w0 * M[3] + w1 * M[4] + w2 * M[5];
    }
    if (a == 0) {
      double* yr = Ys + R * ld + 3 * i;
      yr[0] = u[0]; yr[1] = u[1]; yr[2] = u[2];
    }
  }
  for (int r = tid; r < RY; r += NT)
    for (int cc = 3 * cnt; cc < 4 * K4; ++cc) Ys[r * ld + cc] = 0.0;
  for (int r = tid; r < R; r += NT) {
    const int op = rdesc[2 * (r / 6)];
    rowmap[r] = op < 0 ? -1 : op + r % 6;
  }
  __syncthreads();
  // ---- C = Y Y^T on the tensor cores (upper tiles, dealt out to the warps), one RED per upper-triangular entry.
  // Two column tiles per step share the row fragment and give four independent DMMA chains (K is only 2..12 steps
  // here, the chain latency and the per-pair bookkeeping dominate).
  const int fr = lane >> 2, fc = lane & 3;
  auto emit = [&](int gi, int ri, int tn, double v0, double v1) {
    if (ri < 0) return;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int gj = 8 * tn + 2 * fc + e;
      if (gj > R || gi > gj) continue;
      const double val = e ? v1 : v0;
      if (gj == R) {
        atomicAdd(&g_red[ri], -val);
      } else {
        const int rj = rowmap[gj];
        if (rj >= 0) atomicAdd(&H[min(ri, rj) * n + max(ri, rj)], -val);
      }
    }
  };
  // the (tm, tn >= tm) pairs in row-major order, two at a time; warp wid of a multi-warp chunk takes every WPC-th step
  int step = 0;
  for (int tm = 0; tm < T; ++tm) {
    const int gi = 8 * tm + fr;
    const double* za = Ys + (gi < RY ? gi : 0) * ld + fc;
    const int ri = gi < R ? rowmap[gi] : -1;
    for (int tn = tm; tn < T; tn += 2, ++step) {
      if (WPC > 1 && step % WPC != wid) continue;
      const bool two = tn + 1 < T;
      const int rb0 = 8 * tn + fr, rb1 = 8 * (tn + 1) + fr;
      const double* zb0 = Ys + (rb0 < RY ? rb0 : 0) * ld + fc;
      const double* zb1 = Ys + ((two && rb1 < RY) ? rb1 : 0) * ld + fc;
      double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0, d0 = 0.0, d1 = 0.0, f0 = 0.0, f1 = 0.0;
      int ks = 0;
      for (; ks + 1 < K4; ks += 2) {
        const double a0 = za[4 * ks], a1 = za[4 * ks + 4];
        dmma8x8x4(c0, c1, a0, zb0[4 * ks]);
        dmma8x8x4(d0, d1, a0, zb1[4 * ks]);
        dmma8x8x4(e0, e1, a1, zb0[4 * ks + 4]);
        dmma8x8x4(f0, f1, a1, zb1[4 * ks + 4]);
      }
      if (ks < K4) {
        const double a0 = za[4 * ks];
        dmma8x8x4(c0, c1, a0, zb0[4 * ks]);
        dmma8x8x4(d0, d1, a0, zb1[4 * ks]);
      }
      emit(gi, ri, tn, c0 + e0, c1 + e1);
      if (two) emit(gi, ri, tn + 1, d0 + f0, d1 + f1);
    }
  }
}

// ---- wide chunks of short tracks: one CTA per chunk, one warp per pose run --------------------------------
// Patterns with <= 4 pose runs hold most of the observations (two-view tracks alone are 63 % of the landmarks of
// BASELINE configs[1]).  Warp a takes run a of <= 32 landmarks (lane = landmark): all observations of the chunk are
// in flight at once, the runs' P P^T products (diagonal block, gradients) run concurrently on the tensor cores,
// V and b are summed over the warps through shared memory, and the Y Y^T tile pairs are dealt out to the warps.
template <int NR>
struct WrCfg {
  static constexpr int R = 6 * NR, RY = R + 1, T = (RY + 7) / 8;
  static constexpr int kY = RY * 100;        // Y operand, ld <= 4 * ceil(96 / 4) + 4
  static constexpr int kP = NR * 8 * 68;     // one P tile per warp
  static constexpr int kInts = 8 * NR;       // 2 NR run descriptors + 6 NR row map
  static constexpr int kDoubles = kY + kP + (kInts + 1) / 2;
  static constexpr int kMinBlocks = NR == 2 ? 8 : (NR == 3 ? 5 : 4);
};

template <int NR, bool FUSED = true>
__global__ void __launch_bounds__(32 * NR, WrCfg<NR>::kMinBlocks) k_schur_wr(Batch b, SvinBaOptions opt,
                                                                              const int* list) {
  using C = WrCfg<NR>;
  extern __shared__ double sm_all[];
  double* Ys = sm_all;
  double* Ps = Ys + C::kY;
  int* rdesc = reinterpret_cast<int*>(Ps + C::kP);
  int* rowmap = rdesc + 2 * NR;
  const int lane = threadIdx.x & 31, a = threadIdx.x >> 5;
  const int chunk = list[blockIdx.x];
  const int w = b.sw_win[chunk];
  WinState& ws = b.ws[w];
  if (ws.done || ws.reuse) return;  // CTA-uniform
  const WinDesc& wd = b.win[w];
  const int cnt = b.sw_count[chunk];
  const int rf = b.sw_run_first[chunk];
  const int offp = b.run_off[rf + a];
  const int k0m = b.run_k0m[rf + a];
  const int k0 = k0m >> 8, m = k0m & 255;
  const bool active = lane < cnt;
  const int l = b.sw_lm_begin[chunk] + (active ? lane : 0);
  const int buf = ws.cur;
  const int n = wd.n_dense;
  double* H = b.H + wd.H_off;
  double* g_red = b.g_red + wd.d_off;
  double* g_raw = b.g_raw + wd.d_off;
  double* Hdiag = b.Hdiag + wd.d_off;
  const int ob = b.lm_obs_first[l];
  const int ost = b.lm_obs_stride[l];
  const LmPoint lmp = load_lm(b, buf, l);
  const bool lfix = b.lm_fixed[l] != 0;
  const double mu = ws.mu;
  const int cntp = (cnt + 3) & ~3;
  const int K4p = cntp >> 1;
  const int ldp = 2 * cntp + 4;
  const int K4 = (3 * cnt + 3) >> 2;
  const int ld = 4 * K4 + 4;
  const int fr = lane >> 2, fc = lane & 3;
  double* Pt = Ps + a * (8 * 68);
  if (lane == 0) {
    rdesc[2 * a] = offp;
    rdesc[2 * a + 1] = k0m;
  }
  for (int e = lane; e < ldp; e += 32) Pt[7 * ldp + e] = 0.0;

  // ---- this warp's run: partial V, b; W; P P^T -> diagonal block, gradients
  double V[6] = {0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};
  double W[18];
#pragma unroll
  for (int k = 0; k < 18; ++k) W[k] = 0.0;
  double c0 = 0.0, c1 = 0.0;
  const double wgt = active ? 1.0 : 0.0;
  for (int q = 0; q < m; ++q) {
    const int o = ob + (k0 + q) * ost;
    double Jp[12], Jl[6], r0, r1;
    obs_rJ<FUSED>(b, buf, o, lmp, r0, r1, Jp, Jl);
#pragma unroll
    for (int e = 0; e < 12; ++e) Jp[e] *= wgt;
#pragma unroll
    for (int e = 0; e < 6; ++e) Jl[e] *= wgt;
    r0 *= wgt;
    r1 *= wgt;
    V[0] += Jl[0] * Jl[0] + Jl[3] * Jl[3];
    V[1] += Jl[0] * Jl[1] + Jl[3] * Jl[4];
    V[2] += Jl[0] * Jl[2] + Jl[3] * Jl[5];
    V[3] += Jl[1] * Jl[1] + Jl[4] * Jl[4];
    V[4] += Jl[1] * Jl[2] + Jl[4] * Jl[5];
    V[5] += Jl[2] * Jl[2] + Jl[5] * Jl[5];
    bl[0] += Jl[0] * r0 + Jl[3] * r1;
    bl[1] += Jl[1] * r0 + Jl[4] * r1;
    bl[2] += Jl[2] * r0 + Jl[5] * r1;
    if (offp >= 0) {  // warp-uniform
      acc_W(Jp, Jl, W);
      if (lane < cntp) {
#pragma unroll
        for (int e = 0; e < 6; ++e) {
          Pt[e * ldp + lane] = Jp[e];
          Pt[e * ldp + cntp + lane] = Jp[6 + e];
        }
        Pt[6 * ldp + lane] = r0;
        Pt[6 * ldp + cntp + lane] = r1;
      }
      __syncwarp();
      const double* pa = Pt + fr * ldp + fc;
      double e0 = 0.0, e1 = 0.0;
      for (int ks = 0; ks < K4p; ks += 2) {
        const double x = pa[4 * ks], y = pa[4 * ks + 4];
        dmma8x8x4(c0, c1, x, x);
        dmma8x8x4(e0, e1, y, y);
      }
      c0 += e0;
      c1 += e1;
      __syncwarp();
    }
  }
  if (offp >= 0 && fr < 6) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = 2 * fc + e;
      const double val = e ? c1 : c0;
      if (col < 6) {
        if (fr <= col) atomicAdd(&H[(size_t)(offp + fr) * n + offp + col], val);
        if (fr == col) atomicAdd(&Hdiag[offp + fr], val);
      } else if (col == 6) {
        atomicAdd(&g_red[offp + fr], val);
        atomicAdd(&g_raw[offp + fr], val);
      }
    }
  }
  if (lfix) return;  // chunk-uniform: fixed landmarks are not eliminated
  // ---- V, b over the runs
  {
    double* dst = Ys + (size_t)(a * 32 + lane) * 9;
#pragma unroll
    for (int k = 0; k < 6; ++k) dst[k] = V[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) dst[6 + k] = bl[k];
  }
  __syncthreads();
  // warp 0 finishes the landmark blocks (scale, damping, inverse factor) and hands s, M to the other warps
  double s[3] = {1.0, 1.0, 1.0}, M[6], u[3] = {0.0, 0.0, 0.0};
  double* sM = Ps;  // the P tiles are free now: [32][9]
  if (a == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) V[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) bl[k] = 0.0;
#pragma unroll
    for (int a2 = 0; a2 < NR; ++a2) {
      const double* src = Ys + (size_t)(a2 * 32 + lane) * 9;
#pragma unroll
      for (int k = 0; k < 6; ++k) V[k] += src[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) bl[k] += src[6 + k];
    }
    if (ws.iter == 0 && ws.num_successful == 0 && !ws.invalid) {
      if (opt.jacobi_scaling) {
        s[0] = 1.0 / (1.0 + sqrt(V[0]));
        s[1] = 1.0 / (1.0 + sqrt(V[3]));
        s[2] = 1.0 / (1.0 + sqrt(V[5]));
      }
      if (active) {
        b.lm_scale[3 * (size_t)l] = s[0];
        b.lm_scale[3 * (size_t)l + 1] = s[1];
        b.lm_scale[3 * (size_t)l + 2] = s[2];
      }
    } else {
      s[0] = b.lm_scale[3 * (size_t)l];
      s[1] = b.lm_scale[3 * (size_t)l + 1];
      s[2] = b.lm_scale[3 * (size_t)l + 2];
    }
    double gm = active ? fmax(fabs(bl[0]), fmax(fabs(bl[1]), fabs(bl[2]))) : 0.0;
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, o2));
    if (lane == 0) atomic_max_nonneg(&ws.gmax_bits, gm);
    double Vs[6] = {V[0] * s[0] * s[0], V[1] * s[0] * s[1], V[2] * s[0] * s[2],
                    V[3] * s[1] * s[1], V[4] * s[1] * s[2], V[5] * s[2] * s[2]};
    const double d0 = sqrt(fmin(fmax(Vs[0], opt.min_lm_diagonal), opt.max_lm_diagonal));
    const double d1 = sqrt(fmin(fmax(Vs[3], opt.min_lm_diagonal), opt.max_lm_diagonal));
    const double d2 = sqrt(fmin(fmax(Vs[5], opt.min_lm_diagonal), opt.max_lm_diagonal));
    Vs[0] += mu * d0 * d0;
    Vs[3] += mu * d1 * d1;
    Vs[5] += mu * d2 * d2;
    double Vi[6];
    spd3_inverse_factor(Vs, Vi, M);
    const double bs0 = s[0] * bl[0], bs1 = s[1] * bl[1], bs2 = s[2] * bl[2];
    u[0] = M[0] * bs0;
    u[1] = M[1] * bs0 + M[2] * bs1;
    u[2] = M[3] * bs0 + M[4] * bs1 + M[5] * bs2;
    if (active) {
      double* p = b.lm_Vinv + 6 * (size_t)l;
#pragma unroll
      for (int k = 0; k < 6; ++k) p[k] = Vi[k];
      p = b.lm_bs + 3 * (size_t)l;
      p[0] = bs0; p[1] = bs1; p[2] = bs2;
      p = b.lm_diag + 3 * (size_t)l;
      p[0] = d0; p[1] = d1; p[2] = d2;
      p = b.lm_grad + 3 * (size_t)l;
      p[0] = bs0 / d0; p[1] = bs1 / d1; p[2] = bs2 / d2;
    }
    double* dst = sM + lane * 9;
#pragma unroll
    for (int k = 0; k < 3; ++k) dst[k] = s[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) dst[3 + k] = M[k];
  }
  __syncthreads();
  if (a != 0) {
    const double* src = sM + lane * 9;
#pragma unroll
    for (int k = 0; k < 3; ++k) s[k] = src[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) M[k] = src[3 + k];
  }
  // ---- Y = [W_a diag(s) M^T ; (M b)^T]
  if (active) {
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      const double w0 = W[e * 3] * s[0], w1 = W[e * 3 + 1] * s[1], w2 = W[e * 3 + 2] * s[2];
      double* yr = Ys + (6 * a + e) * ld + 3 * lane;
      yr[0] = w0 * M[0];
      yr[1] = w0 * M[1] + w1 * M[2];
      yr[2] = w0 * M[3] + w1 * M[4] + w2 * M[5];
    }
    if (a == 0) {
      double* yr = Ys + C::R * ld + 3 * lane;
      yr[0] = u[0]; yr[1] = u[1]; yr[2] = u[2];
    }
  }
  for (int r = threadIdx.x; r < C::RY; r += 32 * NR)
    for (int cc = 3 * cnt; cc < 4 * K4; ++cc) Ys[r * ld + cc] = 0.0;
  for (int r = threadIdx.x; r < C::R; r += 32 * NR) {
    const int op = rdesc[2 * (r / 6)];
    rowmap[r] = op < 0 ? -1 : op + r % 6;
  }
  __syncthreads();
  // ---- C = Y Y^T on the tensor cores: the upper tile pairs are dealt out to the warps
  int pidx = 0;
#pragma unroll
  for (int tm = 0; tm < C::T; ++tm) {
#pragma unroll
    for (int tn = tm; tn < C::T; ++tn, ++pidx) {
      if (pidx % NR != a) continue;
      const int gi = 8 * tm + fr;
      const double* za = Ys + (gi < C::RY ? gi : 0) * ld + fc;
      const int ri = gi < C::R ? rowmap[gi] : -1;
      const int rb = 8 * tn + fr;
      const double* zb = Ys + (rb < C::RY ? rb : 0) * ld + fc;
      double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
      int ks = 0;
      for (; ks + 1 < K4; ks += 2) {
        dmma8x8x4(d0, d1, za[4 * ks], zb[4 * ks]);
        dmma8x8x4(e0, e1, za[4 * ks + 4], zb[4 * ks + 4]);
      }
      if (ks < K4) dmma8x8x4(d0, d1, za[4 * ks], zb[4 * ks]);
      d0 += e0;
      d1 += e1;
      if (ri < 0) continue;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gj = 8 * tn + 2 * fc + e;
        if (gj > C::R || gi > gj) continue;
        const double val = e ? d1 : d0;
        if (gj == C::R) {
          atomicAdd(&g_red[ri], -val);
        } else {
          const int rj = rowmap[gj];
          if (rj >= 0) atomicAdd(&H[min(ri, rj) * n + max(ri, rj)], -val);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ dense terms
// Warp-cooperative 15x15 helpers on shared memory (row-major).
__device__ __forceinline__ void warp_mm15(const double* A, const double* Bm, double* C, bool transB, int lane) {
  for (int e = lane; e < 225; e += 32) {
    const int i = e / 15, j = e % 15;
    double s = 0;
    if (!transB) {
#pragma unroll
      for (int k = 0; k < 15; ++k) s += A[i * 15 + k] * Bm[k * 15 + j];
    } else {
#pragma unroll
      for (int k = 0; k < 15; ++k) s += A[i * 15 + k] * Bm[j * 15 + k];
    }
    C[e] = s;
  }
}
__device__ __forceinline__ void set_blk(double* F, int r0, int c0, const M3& Bm, double sc) {
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) F[(r0 + a) * 15 + c0 + c] = sc * Bm.m[a * 3 + c];
}

// ImuError::redoPreintegration (ImuError.cpp:76-263) by one warp; sm = 3*225 doubles scratch (P, F, T).
__device__ void imu_redo_warp(const Batch& b, const ImuTerm& t, ImuCache* c, const ImuP& P, const double* sb0,
                              double* sm, int lane) {
  double* Pm = sm;
  double* F = sm + 225;
  double* T = sm + 450;
  for (int e = lane; e < 225; e += 32) Pm[e] = 0.0;
  __syncwarp();
  // running quantities live in lane 0's registers
  Q4 Delta_q{0, 0, 0, 1};
  M3 C_integral, C_doubleintegral, cross, dalpha_db_g, dv_db_g, dp_db_g;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    C_integral.m[k] = 0; C_doubleintegral.m[k] = 0; cross.m[k] = 0;
    dalpha_db_g.m[k] = 0; dv_db_g.m[k] = 0; dp_db_g.m[k] = 0;
  }
  V3 acc_integral{0, 0, 0}, acc_doubleintegral{0, 0, 0};
  long long time = t.t0;
  const long long end = t.t1;
  const int n = t.meas_end - t.meas_begin;
  const long long* mt = b.imu_meas_t + t.meas_begin;
  const double* mg = b.imu_meas_gyro + 3 * (size_t)t.meas_begin;
  const double* ma = b.imu_meas_accel + 3 * (size_t)t.meas_begin;
  bool hasStarted = false;
  if (!(mt[n - 1] >= end)) return;
  for (int it = 0; it < n; ++it) {
    const int nx = (it + 1 < n) ? it + 1 : it;
    V3 w0{mg[3 * it], mg[3 * it + 1], mg[3 * it + 2]}, a0{ma[3 * it], ma[3 * it + 1], ma[3 * it + 2]};
    V3 w1{mg[3 * nx], mg[3 * nx + 1], mg[3 * nx + 2]}, a1{ma[3 * nx], ma[3 * nx + 1], ma[3 * nx + 2]};
    long long nexttime = (it + 1 == n) ? t.t1 : mt[it + 1];
    double dt = ns_to_sec(nexttime - time);
    if (end < nexttime) {
      const double interval = ns_to_sec(nexttime - mt[it]);
      nexttime = t.t1;
      dt = ns_to_sec(nexttime - time);
      const double r = dt / interval;
      w1 = V3{(1.0 - r) * w0.x + r * w1.x, (1.0 - r) * w0.y + r * w1.y, (1.0 - r) * w0.z + r * w1.z};
      a1 = V3{(1.0 - r) * a0.x + r * a1.x, (1.0 - r) * a0.y + r * a1.y, (1.0 - r) * a0.z + r * a1.z};
    }
    if (dt <= 0.0) continue;  // warp-uniform
    if (!hasStarted) {
      hasStarted = true;
      const double r = dt / ns_to_sec(nexttime - mt[it]);
      w0 = V3{r * w0.x + (1.0 - r) * w1.x, r * w0.y + (1.0 - r) * w1.y, r * w0.z + (1.0 - r) * w1.z};
      a0 = V3{r * a0.x + (1.0 - r) * a1.x, r * a0.y + (1.0 - r) * a1.y, r * a0.z + (1.0 - r) * a1.z};
    }
    double sigma_g_c = P.sigma_g_c, sigma_a_c = P.sigma_a_c;
    if (fabs(w0.x) > P.g_max || fabs(w0.y) > P.g_max || fabs(w0.z) > P.g_max || fabs(w1.x) > P.g_max ||
        fabs(w1.y) > P.g_max || fabs(w1.z) > P.g_max)
      sigma_g_c *= 100;
    if (fabs(a0.x) > P.a_max || fabs(a0.y) > P.a_max || fabs(a0.z) > P.a_max || fabs(a1.x) > P.a_max ||
        fabs(a1.y) > P.a_max || fabs(a1.z) > P.a_max)
      sigma_a_c *= 100;
    if (lane == 0) {
      const V3 wt{0.5 * (w0.x + w1.x) - sb0[3], 0.5 * (w0.y + w1.y) - sb0[4], 0.5 * (w0.z + w1.z) - sb0[5]};
      const V3 at{0.5 * (a0.x + a1.x) - sb0[6], 0.5 * (a0.y + a1.y) - sb0[7], 0.5 * (a0.z + a1.z) - sb0[8]};
      const double theta_half = sqrt(wt.x * wt.x + wt.y * wt.y + wt.z * wt.z) * 0.5 * dt;
      const double sth = sinc_okvis(theta_half), cth = cos(theta_half);
      const Q4 dq{sth * wt.x * 0.5 * dt, sth * wt.y * 0.5 * dt, sth * wt.z * 0.5 * dt, cth};
      const Q4 Delta_q_1 = qmul(Delta_q, dq);
      const M3 C = qrot(Delta_q), C_1 = qrot(Delta_q_1);
      M3 CC;
#pragma unroll
      for (int k = 0; k < 9; ++k) CC.m[k] = C.m[k] + C_1.m[k];
      const V3 CCa = m3v(CC, at);
      M3 C_integral_1;
#pragma unroll
      for (int k = 0; k < 9; ++k) C_integral_1.m[k] = C_integral.m[k] + 0.5 * CC.m[k] * dt;
      const V3 acc_integral_1{acc_integral.x + 0.5 * CCa.x * dt, acc_integral.y + 0.5 * CCa.y * dt,
                              acc_integral.z + 0.5 * CCa.z * dt};
#pragma unroll
      for (int k = 0; k < 9; ++k) C_doubleintegral.m[k] += C_integral.m[k] * dt + 0.25 * CC.m[k] * dt * dt;
      acc_doubleintegral.x += acc_integral.x * dt + 0.25 * CCa.x * dt * dt;
      acc_doubleintegral.y += acc_integral.y * dt + 0.25 * CCa.y * dt * dt;
      acc_doubleintegral.z += acc_integral.z * dt + 0.25 * CCa.z * dt * dt;
      const M3 Jr = right_jacobian(V3{wt.x * dt, wt.y * dt, wt.z * dt});
      const M3 C1Jr = m3mul(C_1, Jr);
#pragma unroll
      for (int k = 0; k < 9; ++k) dalpha_db_g.m[k] += C1Jr.m[k] * dt;
      const M3 t9 = m3mul(qrot(qinverse(dq)), cross);
      M3 cross_1;
#pragma unroll
      for (int k = 0; k < 9; ++k) cross_1.m[k] = t9.m[k] + Jr.m[k] * dt;
      const M3 ax = crossmx(at);
      const M3 B0 = m3mul(m3mul(C, ax), cross);
      const M3 B1 = m3mul(m3mul(C_1, ax), cross_1);
      M3 G;
#pragma unroll
      for (int k = 0; k < 9; ++k) G.m[k] = B0.m[k] + B1.m[k];
      M3 dv_db_g_1, F09, F012;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        dv_db_g_1.m[k] = dv_db_g.m[k] + 0.5 * dt * G.m[k];
        F09.m[k] = dt * dv_db_g.m[k] + 0.25 * dt * dt * G.m[k];
        dp_db_g.m[k] += F09.m[k];
        F012.m[k] = -C_integral.m[k] * dt + 0.25 * CC.m[k] * dt * dt;
      }
      // F_delta
      for (int k = 0; k < 225; ++k) F[k] = 0.0;
      for (int k = 0; k < 15; ++k) F[k * 15 + k] = 1.0;
      set_blk(F, 0, 3,
              crossmx(V3{acc_integral.x * dt + 0.25 * CCa.x * dt * dt, acc_integral.y * dt + 0.25 * CCa.y * dt * dt,
                         acc_integral.z * dt + 0.25 * CCa.z * dt * dt}),
              -1.0);
      M3 I3;
#pragma unroll
      for (int k = 0; k < 9; ++k) I3.m[k] = 0;
      I3.m[0] = I3.m[4] = I3.m[8] = 1.0;
      set_blk(F, 0, 6, I3, dt);
      set_blk(F, 0, 9, F09, 1.0);
      set_blk(F, 0, 12, F012, 1.0);
      set_blk(F, 3, 9, C_1, -dt);
      set_blk(F, 6, 3, crossmx(V3{0.5 * CCa.x * dt, 0.5 * CCa.y * dt, 0.5 * CCa.z * dt}), -1.0);
      set_blk(F, 6, 9, G, 0.5 * dt);
      set_blk(F, 6, 12, CC, -0.5 * dt);
      Delta_q = Delta_q_1;
      C_integral = C_integral_1;
      acc_integral = acc_integral_1;
      cross = cross_1;
      dv_db_g = dv_db_g_1;
    }
    __syncwarp();
    warp_mm15(F, Pm, T, false, lane);
    __syncwarp();
    warp_mm15(T, F, Pm, true, lane);
    __syncwarp();
    if (lane == 0) {
      const double sigma2_dalpha = dt * sigma_g_c * sigma_g_c;
      const double sigma2_v = dt * sigma_a_c * sigma_a_c;
      const double sigma2_p = 0.5 * dt * dt * sigma2_v;
      const double sigma2_b_g = dt * P.sigma_gw_c * P.sigma_gw_c;
      const double sigma2_b_a = dt * P.sigma_aw_c * P.sigma_aw_c;
      for (int k = 0; k < 3; ++k) {
        Pm[(3 + k) * 15 + 3 + k] += sigma2_dalpha;
        Pm[(6 + k) * 15 + 6 + k] += sigma2_v;
        Pm[(0 + k) * 15 + 0 + k] += sigma2_p;
        Pm[(9 + k) * 15 + 9 + k] += sigma2_b_g;
        Pm[(12 + k) * 15 + 12 + k] += sigma2_b_a;
      }
    }
    __syncwarp();
    time = nexttime;
    if (nexttime == t.t1) break;
  }
  if (lane == 0) {
    c->Delta_q[0] = Delta_q.x; c->Delta_q[1] = Delta_q.y; c->Delta_q[2] = Delta_q.z; c->Delta_q[3] = Delta_q.w;
    for (int k = 0; k < 9; ++k) {
      c->C_integral[k] = C_integral.m[k];
      c->C_doubleintegral[k] = C_doubleintegral.m[k];
      c->dalpha_db_g[k] = dalpha_db_g.m[k];
      c->dv_db_g[k] = dv_db_g.m[k];
      c->dp_db_g[k] = dp_db_g.m[k];
      c->sb_ref[k] = sb0[k];
    }
    c->acc_integral[0] = acc_integral.x; c->acc_integral[1] = acc_integral.y; c->acc_integral[2] = acc_integral.z;
    c->acc_doubleintegral[0] = acc_doubleintegral.x;
    c->acc_doubleintegral[1] = acc_doubleintegral.y;
    c->acc_doubleintegral[2] = acc_doubleintegral.z;
  }
  // symmetrise P -> F ; invert (LU, partial pivoting) ; symmetrise ; LLT
  for (int e = lane; e < 225; e += 32) {
    const int i = e / 15, j = e % 15;
    F[e] = 0.5 * Pm[i * 15 + j] + 0.5 * Pm[j * 15 + i];
  }
  __syncwarp();
  // LU in F (in place), pivots tracked by lane 0 in T[0..15) as doubles
  for (int k = 0; k < 15; ++k) {
    int p = k;
    if (lane == 0) {
      double best = fabs(F[k * 15 + k]);
      for (int i = k + 1; i < 15; ++i)
        if (fabs(F[i * 15 + k]) > best) {
          best = fabs(F[i * 15 + k]);
          p = i;
        }
    }
    p = __shfl_sync(0xffffffffu, p, 0);
    if (k == 0 && lane < 15) T[lane] = (double)lane;
    __syncwarp();
    if (p != k) {
      if (lane < 15) {
        const double tmp = F[k * 15 + lane];
        F[k * 15 + lane] = F[p * 15 + lane];
        F[p * 15 + lane] = tmp;
      }
      if (lane == 0) {
        const double tp = T[k];
        T[k] = T[p];
        T[p] = tp;
      }
    }
    __syncwarp();
    const double piv = F[k * 15 + k];
    if (lane > k && lane < 15) F[lane * 15 + k] /= piv;
    __syncwarp();
    // trailing update: element (i, j), i,j > k
    for (int e = lane; e < 225; e += 32) {
      const int i = e / 15, j = e % 15;
      if (i > k && j > k) F[e] -= F[i * 15 + k] * F[k * 15 + j];
    }
    __syncwarp();
  }
  // columns of the inverse: lane c solves A x = e_c  (result into Pm, row-major)
  if (lane < 15) {
    double y[15];
    for (int i = 0; i < 15; ++i) {
      double s = ((int)T[i] == lane) ? 1.0 : 0.0;
      for (int j = 0; j < i; ++j) s -= F[i * 15 + j] * y[j];
      y[i] = s;
    }
    for (int i = 14; i >= 0; --i) {
      double s = y[i];
      for (int j = i + 1; j < 15; ++j) s -= F[i * 15 + j] * y[j];
      y[i] = s / F[i * 15 + i];
    }
    for (int i = 0; i < 15; ++i) Pm[i * 15 + lane] = y[i];
  }
  __syncwarp();
  for (int e = lane; e < 225; e += 32) {
    const int i = e / 15, j = e % 15;
    F[e] = 0.5 * Pm[i * 15 + j] + 0.5 * Pm[j * 15 + i];  // information_
  }
  __syncwarp();
  // Eigen LLT (lower, left-looking), in F; early exit on a non-positive pivot like Eigen
  for (int k = 0; k < 15; ++k) {
    double x = F[k * 15 + k];
    for (int j = 0; j < k; ++j) x -= F[k * 15 + j] * F[k * 15 + j];
    if (x <= 0.0) break;
    x = sqrt(x);
    __syncwarp();
    if (lane == 0) F[k * 15 + k] = x;
    if (lane > k && lane < 15) {
      double s = F[lane * 15 + k];
      for (int j = 0; j < k; ++j) s -= F[lane * 15 + j] * F[k * 15 + j];
      F[lane * 15 + k] = s / x;
    }
    __syncwarp();
  }
  // squareRootInformation_ = L^T
  for (int e = lane; e < 225; e += 32) {
    const int i = e / 15, j = e % 15;
    c->sqrt_info[e] = (j >= i) ? F[j * 15 + i] : 0.0;
  }
  __syncwarp();
}

struct ImuEvalOut {  // optional raw dump (svin_ba_evaluate)
  double *r, *J0, *J1, *J2, *J3;
};

// ImuError::EvaluateWithMinimalJacobians (ImuError.cpp:706-866) by one warp.
__device__ void imu_eval_warp(const Batch& b, int ti, int sbuf, double* Jd, double* rd, int n, bool want_jac,
                              double* sm, int lane, double* cost_out, ImuEvalOut* dump) {
  const ImuTerm& t = b.imu[ti];
  ImuCache* c = b.imu_cache + ti;
  const WinDesc& wd = b.win[t.win];
  const double* pose0 = b.pose[sbuf] + 7 * (size_t)t.pose0;
  const double* pose1 = b.pose[sbuf] + 7 * (size_t)t.pose1;
  const double* sb0 = b.sb[sbuf] + 9 * (size_t)t.sb0;
  const double* sb1 = b.sb[sbuf] + 9 * (size_t)t.sb1;
  const double Delta_t = ns_to_sec(t.t1 - t.t0);
  double Delta_b[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) Delta_b[k] = sb0[3 + k] - c->sb_ref[3 + k];
  const double nb = sqrt(Delta_b[0] * Delta_b[0] + Delta_b[1] * Delta_b[1] + Delta_b[2] * Delta_b[2]);
  const bool redo = (c->redo != 0) || (nb * Delta_t > 0.0001);
  __syncwarp();
  if (redo) {
    imu_redo_warp(b, t, c, wd.imu, sb0, sm, lane);
    if (lane == 0) {
      c->redo_counter++;
      c->redo = 0;
      atomicAdd(&b.ws[t.win].imu_redo, 1);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) Delta_b[k] = 0.0;
    __syncwarp();
  }
  double* F0 = sm;          // 225
  double* F1 = sm + 225;    // 225
  double* err = sm + 450;   // 15
  if (lane == 0) {
    const Tf T0 = tf_load(pose0), T1 = tf_load(pose1);
    const M3 C_S0_W = m3t(T0.C);
    const V3 g_W{0.0, 0.0, wd.imu.g};
    for (int k = 0; k < 225; ++k) {
      F0[k] = 0.0;
      F1[k] = 0.0;
    }
    for (int k = 0; k < 15; ++k) {
      F0[k * 15 + k] = 1.0;
      F1[k * 15 + k] = -1.0;
    }
    const V3 dp{T0.r.x - T1.r.x + sb0[0] * Delta_t - 0.5 * g_W.x * Delta_t * Delta_t,
                T0.r.y - T1.r.y + sb0[1] * Delta_t - 0.5 * g_W.y * Delta_t * Delta_t,
                T0.r.z - T1.r.z + sb0[2] * Delta_t - 0.5 * g_W.z * Delta_t * Delta_t};
    const V3 dv{sb0[0] - sb1[0] - g_W.x * Delta_t, sb0[1] - sb1[1] - g_W.y * Delta_t,
                sb0[2] - sb1[2] - g_W.z * Delta_t};
    M3 dalpha, dvg, dpg, Cint, Cdint;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      dalpha.m[k] = c->dalpha_db_g[k];
      dvg.m[k] = c->dv_db_g[k];
      dpg.m[k] = c->dp_db_g[k];
      Cint.m[k] = c->C_integral[k];
      Cdint.m[k] = c->C_doubleintegral[k];
    }
    const V3 adb = m3v(dalpha, V3{Delta_b[0], Delta_b[1], Delta_b[2]});
    const Q4 Dq = qmul(delta_q(V3{-adb.x, -adb.y, -adb.z}),
                       Q4{c->Delta_q[0], c->Delta_q[1], c->Delta_q[2], c->Delta_q[3]});
    set_blk(F0, 0, 0, C_S0_W, 1.0);
    set_blk(F0, 0, 3, m3mul(C_S0_W, crossmx(dp)), 1.0);
    set_blk(F0, 0, 6, C_S0_W, Delta_t);
    set_blk(F0, 0, 9, dpg, 1.0);
    set_blk(F0, 0, 12, Cdint, -1.0);
    const Q4 q1inv = qinverse(T1.q);
    double Qa[16], Qb[16], Qc[16], Qd[16];
    qplus44(qmul(Dq, q1inv), Qa);
    qoplus44(T0.q, Qb);
    mat44mul(Qa, Qb, Qc);
    M3 B33;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) B33.m[a * 3 + cc] = Qc[a * 4 + cc];
    set_blk(F0, 3, 3, B33, 1.0);
    qoplus44(qmul(q1inv, T0.q), Qa);
    qoplus44(Dq, Qb);
    mat44mul(Qa, Qb, Qc);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) B33.m[a * 3 + cc] = Qc[a * 4 + cc];
    M3 nd;
#pragma unroll
    for (int k = 0; k < 9; ++k) nd.m[k] = -dalpha.m[k];
    set_blk(F0, 3, 9, m3mul(B33, nd), 1.0);
    set_blk(F0, 6, 3, m3mul(C_S0_W, crossmx(dv)), 1.0);
    set_blk(F0, 6, 6, C_S0_W, 1.0);
    set_blk(F0, 6, 9, dvg, 1.0);
    set_blk(F0, 6, 12, Cint, -1.0);
    set_blk(F1, 0, 0, C_S0_W, -1.0);
    qplus44(Dq, Qa);
    qoplus44(T0.q, Qb);
    qplus44(q1inv, Qd);
    mat44mul(Qa, Qb, Qc);
    mat44mul(Qc, Qd, Qa);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) B33.m[a * 3 + cc] = Qa[a * 4 + cc];
    set_blk(F1, 3, 3, B33, -1.0);
    set_blk(F1, 6, 6, C_S0_W, -1.0);
    const V3 a3 = m3v(C_S0_W, dp);
    const double a3v[3] = {a3.x, a3.y, a3.z};
    for (int k = 0; k < 3; ++k) {
      double s = 0;
      for (int cc = 0; cc < 6; ++cc) s += F0[(0 + k) * 15 + 9 + cc] * Delta_b[cc];
      err[k] = a3v[k] + c->acc_doubleintegral[k] + s;
    }
    const Q4 qe = qmul(Dq, qmul(q1inv, T0.q));
    err[3] = 2 * qe.x;
    err[4] = 2 * qe.y;
    err[5] = 2 * qe.z;
    const V3 b3 = m3v(C_S0_W, dv);
    const double b3v[3] = {b3.x, b3.y, b3.z};
    for (int k = 0; k < 3; ++k) {
      double s = 0;
      for (int cc = 0; cc < 6; ++cc) s += F0[(6 + k) * 15 + 9 + cc] * Delta_b[cc];
      err[6 + k] = b3v[k] + c->acc_integral[k] + s;
    }
    for (int k = 0; k < 6; ++k) err[9 + k] = sb0[3 + k] - sb1[3 + k];
  }
  __syncwarp();
  const double* U = c->sqrt_info;
  double cst = 0.0;
  if (lane < 15) {
    double s = 0;
    for (int k = lane; k < 15; ++k) s += U[lane * 15 + k] * err[k];
    if (rd) rd[t.row0 + lane] = s;
    if (dump && dump->r) dump->r[15 * (size_t)ti + lane] = s;
    cst = 0.5 * s * s;
  }
  cst = warp_sum(cst);
  if (lane == 0 && cost_out) atomicAdd(cost_out, cst);
  if (want_jac) {
    const int offs[4] = {b.pose_off[t.pose0], b.sb_off[t.sb0], b.pose_off[t.pose1], b.sb_off[t.sb1]};
    // 15 x 30 outputs: columns 0..5 F0[:,0:6], 6..14 F0[:,6:15], 15..20 F1[:,0:6], 21..29 F1[:,6:15]
    for (int e = lane; e < 450; e += 32) {
      const int a = e / 30, col = e % 30;
      const double* F = (col < 15) ? F0 : F1;
      const int fc = (col < 15) ? col : col - 15;
      double s = 0;
      for (int k = a; k < 15; ++k) s += U[a * 15 + k] * F[k * 15 + fc];
      const int blk = (col < 6) ? 0 : (col < 15) ? 1 : (col < 21) ? 2 : 3;
      const int cc = (blk == 0) ? col : (blk == 1) ? col - 6 : (blk == 2) ? col - 15 : col - 21;
      if (Jd && offs[blk] >= 0) Jd[(size_t)(t.row0 + a) * n + offs[blk] + cc] = s;
      if (dump) {
        if (blk == 0 && dump->J0) dump->J0[90 * (size_t)ti + a * 6 + cc] = s;
        if (blk == 1 && dump->J1) dump->J1[135 * (size_t)ti + a * 9 + cc] = s;
        if (blk == 2 && dump->J2) dump->J2[90 * (size_t)ti + a * 6 + cc] = s;
        if (blk == 3 && dump->J3) dump->J3[135 * (size_t)ti + a * 9 + cc] = s;
      }
    }
  }
  __syncwarp();
}

// pose-type error  r = U * [t_m - t ; 2 vec(q_m * q^-1)]   (PoseError.cpp:85-132)
__device__ void pose_error_eval(const double* meas, const double* U, const double* pose, double* r, double* J) {
  const Tf Tm = tf_load(meas), T = tf_load(pose);
  const Tf dp = tf_mul(Tm, tf_inverse(T));
  const double e[6] = {Tm.r.x - T.r.x, Tm.r.y - T.r.y, Tm.r.z - T.r.z, 2 * dp.q.x, 2 * dp.q.y, 2 * dp.q.z};
  for (int i = 0; i < 6; ++i) {
    double s = 0;
    for (int k = 0; k < 6; ++k) s += U[i * 6 + k] * e[k];
    r[i] = s;
  }
  if (J) {
    double Jm[36];
    for (int k = 0; k < 36; ++k) Jm[k] = 0;
    for (int k = 0; k < 6; ++k) Jm[k * 6 + k] = -1.0;
    const M3 Qp = qplus33(dp.q);
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) Jm[(3 + a) * 6 + 3 + c] = -Qp.m[a * 3 + c];
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += U[i * 6 + k] * Jm[k * 6 + j];
        J[i * 6 + j] = s;
      }
  }
}

// One CTA per window evaluates every non-reprojection term at state buffer (which: 0 cur, 1 candidate),
// writing residuals into rd[buf] and local Jacobians into Jd[buf] (dense rows x n_dense).
// MINB CTAs per SM: at 255 registers two CTAs take a whole SM's register file and lock the concurrently running
// k_linearize out of it; a tighter cap spills in the serial IMU sections but leaves room for the main stream.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_dense_eval(Batch b, int which, int raw, ImuEvalOut dump) {
  const int w = blockIdx.x;
  WinState& ws = b.ws[w];
  if (!raw) {
    if (ws.done) return;
    if (which == 1 && (ws.skip_slot || step_is_invalid(ws))) return;
  }
  const WinDesc& wd = b.win[w];
  const int buf = (which == 0) ? ws.cur : 1 - ws.cur;
  const int sbuf = raw ? ws.cur : buf;
  const int n = wd.n_dense;
  double* Jd = b.Jd[buf] + wd.Jd_off;
  double* rd = b.rd[buf] + wd.rd_off;
  __shared__ double sm[4][3 * 225];
  __shared__ double cost_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) cost_s = 0.0;
  __syncthreads();
  ImuEvalOut* dp = raw ? &dump : nullptr;
  for (int ti = wd.imu_begin + wid; ti < wd.imu_end; ti += 4)
    imu_eval_warp(b, ti, sbuf, raw ? nullptr : Jd, raw ? nullptr : rd, n, true, sm[wid], lane, &cost_s, dp);
  // small single-thread terms, spread over threads
  const int n_pp = wd.pp_end - wd.pp_begin, n_sp = wd.sp_end - wd.sp_begin, n_rp = wd.rp_end - wd.rp_begin;
  const int n_so = wd.so_end - wd.so_begin, n_de = wd.de_end - wd.de_begin;
  const int n_small = raw ? 0 : n_pp + n_sp + n_rp + n_so + n_de;
  for (int k = threadIdx.x; k < n_small; k += blockDim.x) {
    double cst = 0;
    if (k < n_pp) {
      const PosePrior& t = b.pp[wd.pp_begin + k];
      double r[6], J[36];
      pose_error_eval(t.meas, t.U, b.pose[sbuf] + 7 * (size_t)t.block, r, J);
      const int off = b.pose_off[t.block];
      for (int i = 0; i < 6; ++i) {
        rd[t.row0 + i] = r[i];
        cst += 0.5 * r[i] * r[i];
        if (off >= 0)
          for (int j = 0; j < 6; ++j) Jd[(size_t)(t.row0 + i) * n + off + j] = J[i * 6 + j];
      }
    } else if (k < n_pp + n_sp) {
      const SbPrior& t = b.sp[wd.sp_begin + k - n_pp];
      const double* x = b.sb[sbuf] + 9 * (size_t)t.block;
      const int off = b.sb_off[t.block];
      for (int i = 0; i < 9; ++i) {
        double s = 0;
        for (int c = 0; c < 9; ++c) s += t.U[i * 9 + c] * (t.meas[c] - x[c]);
        rd[t.row0 + i] = s;
        cst += 0.5 * s * s;
        if (off >= 0)
          for (int j = 0; j < 9; ++j) Jd[(size_t)(t.row0 + i) * n + off + j] = -t.U[i * 9 + j];
      }
    } else if (k < n_pp + n_sp + n_rp) {
      // RelativePoseError.cpp:76-147
      const RelPose& t = b.rp[wd.rp_begin + k - n_pp - n_sp];
      const Tf T0 = tf_load(b.pose[sbuf] + 7 * (size_t)t.block0), T1 = tf_load(b.pose[sbuf] + 7 * (size_t)t.block1);
      const Tf dpq = tf_mul(T1, tf_inverse(T0));
      const double e[6] = {T1.r.x - T0.r.x, T1.r.y - T0.r.y, T1.r.z - T0.r.z, 2 * dpq.q.x, 2 * dpq.q.y, 2 * dpq.q.z};
      const M3 Qp = qplus33(dpq.q), Qo = qoplus33(dpq.q);
      const int off0 = b.pose_off[t.block0], off1 = b.pose_off[t.block1];
      for (int i = 0; i < 6; ++i) {
        double s = 0;
        for (int c = 0; c < 6; ++c) s += t.U[i * 6 + c] * e[c];
        rd[t.row0 + i] = s;
        cst += 0.5 * s * s;
        for (int j = 0; j < 6; ++j) {
          // J0 = U * [-I 0; 0 -plus33], J1 = U * [I 0; 0 oplus33]
          double j0 = 0, j1 = 0;
          if (j < 3) {
            j0 = -t.U[i * 6 + j];
            j1 = t.U[i * 6 + j];
          } else {
            for (int c = 0; c < 3; ++c) {
              j0 -= t.U[i * 6 + 3 + c] * Qp.m[c * 3 + (j - 3)];
              j1 += t.U[i * 6 + 3 + c] * Qo.m[c * 3 + (j - 3)];
            }
          }
          if (off0 >= 0) Jd[(size_t)(t.row0 + i) * n + off0 + j] = j0;
          if (off1 >= 0) Jd[(size_t)(t.row0 + i) * n + off1 + j] = j1;
        }
      }
    } else if (k < n_pp + n_sp + n_rp + n_so) {
      // SonarError.cpp:113-183 (Jacobian reproduced as written in the reference)
      const SonarTerm& t = b.so[wd.so_begin + k - n_pp - n_sp - n_rp];
      const Tf T = tf_load(b.pose[sbuf] + 7 * (size_t)t.pose);
      const double dx = T.r.x - t.mean[0], dy = T.r.y - t.mean[1], dz = T.r.z - t.mean[2];
      const double r = t.sqrt_info * (t.range - sqrt(dx * dx + dy * dy + dz * dz));
      rd[t.row0] = r;
      cst += 0.5 * r * r;
      const int off = b.pose_off[t.pose];
      if (off >= 0) {
        const Tf T_SSo = tf_load(wd.T_SSo);
        const Tf T_WSo = tf_mul(T, T_SSo);
        const Tf sp = tf_make(V3{t.range * cos(t.heading), t.range * sin(t.heading), 0.0}, Q4{0, 0, 0, 1});
        const Tf Tp = tf_mul(T_WSo, sp);
        double* Jr = Jd + (size_t)t.row0 * n + off;
        Jr[0] = t.sqrt_info * ((T.r.x - Tp.r.x) / t.range);
        Jr[1] = t.sqrt_info * ((T.r.y - Tp.r.y) / t.range);
        Jr[2] = t.sqrt_info * ((T.r.z - Tp.r.z) / t.range);
        Jr[3] = 0; Jr[4] = 0; Jr[5] = 0;
      }
    } else {
      // DepthError.cpp:70-139
      const DepthTerm& t = b.de[wd.de_begin + k - n_pp - n_sp - n_rp - n_so];
      const double* x = b.pose[sbuf] + 7 * (size_t)t.pose;
      const double r = t.sqrt_info * (x[2] - (-1 * t.depth + t.first));
      rd[t.row0] = r;
      cst += 0.5 * r * r;
      const int off = b.pose_off[t.pose];
      if (off >= 0) {
        double* Jr = Jd + (size_t)t.row0 * n + off;
        Jr[0] = 0; Jr[1] = 0; Jr[2] = t.sqrt_info; Jr[3] = 0; Jr[4] = 0; Jr[5] = 0;
      }
    }
    if (!raw) atomicAdd(&cost_s, cst);
  }
  // marginalisation prior (MarginalizationError.cpp:798-844): e = e0 + J dchi ; local J per block
  if (wd.marg_dim > 0 && !raw) {
    const int m = wd.marg_dim;
    __shared__ double dchi[512];
    __shared__ double Mrot[64][9];
    const int nb = wd.marg_blk_end - wd.marg_blk_begin;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) {
      const MargBlock& mb = b.marg_blk[wd.marg_blk_begin + k];
      if (mb.col0 < 0) continue;
      const double* lp = b.marg_lin + wd.marg_lin_off + mb.lin_off;
      if (mb.kind == SVIN_BLOCK_POSE) {
        const double* x = b.pose[sbuf] + 7 * (size_t)mb.index;
        // PoseManifold::minus (PoseManifold.cpp:92-102)
        const Q4 d = qmul(Q4{x[3], x[4], x[5], x[6]}, qinverse(Q4{lp[3], lp[4], lp[5], lp[6]}));
        dchi[mb.col0 + 0] = x[0] - lp[0];
        dchi[mb.col0 + 1] = x[1] - lp[1];
        dchi[mb.col0 + 2] = x[2] - lp[2];
        dchi[mb.col0 + 3] = 2 * d.x;
        dchi[mb.col0 + 4] = 2 * d.y;
        dchi[mb.col0 + 5] = 2 * d.z;
        // J_lift(x_lin) * J_plus(x): top-left 3x3 of oplus(conj(q_lin)) * oplus(normalized(q))
        double A[16], Bq[16], Cq[16];
        qoplus44(Q4{-lp[3], -lp[4], -lp[5], lp[6]}, A);
        qoplus44(qnormalized(Q4{x[3], x[4], x[5], x[6]}), Bq);
        mat44mul(A, Bq, Cq);
        for (int a = 0; a < 3; ++a)
          for (int c = 0; c < 3; ++c) Mrot[k][a * 3 + c] = Cq[a * 4 + c];
      } else {
        const double* x = b.sb[sbuf] + 9 * (size_t)mb.index;
        for (int c = 0; c < 9; ++c) dchi[mb.col0 + c] = x[c] - lp[c];
      }
    }
    __syncthreads();
    const double* MJ = b.marg_J + wd.margJ_off;
    const double* e0 = b.marg_e0 + wd.marg_e0_off;
    double cst = 0;
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
      double s = e0[r];
      for (int c = 0; c < m; ++c) s += MJ[(size_t)r * m + c] * dchi[c];
      rd[wd.marg_row0 + r] = s;
      cst += 0.5 * s * s;
    }
    atomicAdd(&cost_s, cst);
    // Jacobian rows
    for (int e = threadIdx.x; e < m * nb; e += blockDim.x) {
      const int r = e / nb, k = e % nb;
      const MargBlock& mb = b.marg_blk[wd.marg_blk_begin + k];
      if (mb.col0 < 0) continue;
      const double* Jrow = MJ + (size_t)r * m + mb.col0;
      if (mb.kind == SVIN_BLOCK_POSE) {
        const int off = b.pose_off[mb.index];
        double* out = Jd + (size_t)(wd.marg_row0 + r) * n + off;
        out[0] = Jrow[0];
        out[1] = Jrow[1];
        out[2] = Jrow[2];
        for (int c = 0; c < 3; ++c)
          out[3 + c] = Jrow[3] * Mrot[k][0 * 3 + c] + Jrow[4] * Mrot[k][1 * 3 + c] + Jrow[5] * Mrot[k][2 * 3 + c];
      } else {
        const int off = b.sb_off[mb.index];
        double* out = Jd + (size_t)(wd.marg_row0 + r) * n + off;
        for (int c = 0; c < 9; ++c) out[c] = Jrow[c];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && !raw) atomicAdd(&ws.cost_cand, cost_s);
}

// ------------------------------------------------------------------------------------------ reduced system
// One CTA per window.  Adds the dense-term normal equations, applies Jacobi scaling and the LM
// diagonal, factorises (Cholesky) and solves; prepares the vectors the landmark kernels need.
__global__ void __launch_bounds__(kDenseThreads) k_dense_solve(Batch b, SvinBaOptions opt, int use_smem) {
  extern __shared__ double smem[];
  const int w = blockIdx.x;
  WinState& ws = b.ws[w];
  if (ws.done || ws.reuse) return;
  const WinDesc& wd = b.win[w];
  const int n = wd.n_dense, M = wd.n_rows, buf = ws.cur;
  const int tid = threadIdx.x, T = blockDim.x;
  double* Hg = b.H + wd.H_off;
  double* A = use_smem ? smem : Hg;  // (n+1) x n when in smem; in global the rhs row lives in g_red
  const double* Jd = b.Jd[buf] + wd.Jd_off;
  const double* rd = b.rd[buf] + wd.rd_off;
  double* g_red = b.g_red + wd.d_off;
  double* g_raw = b.g_raw + wd.d_off;
  double* Hdiag = b.Hdiag + wd.d_off;
  double* scale = b.scale_d + wd.d_off;
  double* diag = b.diag_d + wd.d_off;
  double* grad = b.grad_d + wd.d_off;
  double* gn = b.gn_d + wd.d_off;
  double* u = b.u_d + wd.d_off;
  double* cvec = b.c_d + wd.d_off;
  __shared__ int fail_s;
  // 1. A(upper) = H~ + Jd^T Jd
  for (int e = tid; e < n * n; e += T) {
    const int i = e / n, j = e % n;
    if (j < i) continue;
    double s = Hg[(size_t)i * n + j];
    for (int r = 0; r < M; ++r) s += Jd[(size_t)r * n + i] * Jd[(size_t)r * n + j];
    A[(size_t)i * n + j] = s;
  }
  // 2. vectors
  double gmax_l = 0.0;
  for (int i = tid; i < n; i += T) {
    double hd = Hdiag[i], gr = 0.0;
    for (int r = 0; r < M; ++r) {
      const double jv = Jd[(size_t)r * n + i];
      hd += jv * jv;
      gr += jv * rd[r];
    }
    Hdiag[i] = hd;
    g_raw[i] += gr;
    g_red[i] += gr;
    gmax_l = fmax(gmax_l, fabs(g_raw[i]));
  }
  if (tid == 0) fail_s = 0;
  atomic_max_nonneg(&ws.gmax_bits, gmax_l);
  __syncthreads();
  // 3. gradient tolerance (FinalizeIterationAndCheckIfMinimizerCanContinue order: after max-iterations)
  __shared__ unsigned long long gmax_s;
  if (tid == 0) gmax_s = atomicMax(&ws.gmax_bits, 0ull);  // atomic read: L1 may hold a stale WinState line
  __syncthreads();
  const double gmax = __longlong_as_double((long long)gmax_s);
  if (ws.last_successful && gmax <= opt.gradient_tolerance) {
    __syncthreads();
    if (tid == 0) {
      ws.done = 1;
      ws.termination = SVIN_TERM_CONVERGENCE;
    }
    return;
  }
  const bool first = (ws.iter == 0 && ws.num_successful == 0 && !ws.invalid);
  const double mu = ws.mu;
  for (int i = tid; i < n; i += T) {
    double s;
    if (first) {
      s = opt.jacobi_scaling ? 1.0 / (1.0 + sqrt(Hdiag[i])) : 1.0;
      scale[i] = s;
    } else {
      s = scale[i];
    }
    const double d = sqrt(fmin(fmax(Hdiag[i] * s * s, opt.min_lm_diagonal), opt.max_lm_diagonal));
    diag[i] = d;
    grad[i] = s * g_raw[i] / d;
  }
  __syncthreads();
  // 4. scaled system, full symmetric, + LM diagonal; rhs in row n (smem) or g_red (global)
  for (int e = tid; e < n * n; e += T) {
    const int i = e / n, j = e % n;
    if (j < i) continue;
    double v = A[(size_t)i * n + j] * scale[i] * scale[j];
    if (i == j) v += mu * diag[i] * diag[i];
    A[(size_t)i * n + j] = v;
  }
  __syncthreads();
  for (int e = tid; e < n * n; e += T) {
    const int i = e / n, j = e % n;
    if (j < i) A[(size_t)i * n + j] = A[(size_t)j * n + i];
  }
  double* rhs = use_smem ? (A + (size_t)n * n) : g_red;
  for (int i = tid; i < n; i += T) rhs[i] = scale[i] * g_red[i];
  __syncthreads();
  // 5. right-looking Cholesky (lower) with the rhs carried as an extra row -> forward solve for free
  for (int k = 0; k < n; ++k) {
    const double akk = A[(size_t)k * n + k];
    if (!(akk > 0.0) || !isfinite(akk)) {
      if (tid == 0) fail_s = 1;
      break;
    }
    const double d = sqrt(akk);
    __syncthreads();
    for (int i = k + 1 + tid; i <= n; i += T) {
      if (i < n)
        A[(size_t)i * n + k] /= d;
      else
        rhs[k] /= d;
    }
    if (tid == 0) A[(size_t)k * n + k] = d;
    __syncthreads();
    const int mrem = n - k - 1;
    for (int e = tid; e < (mrem + 1) * mrem; e += T) {
      const int ii = e / mrem, jj = e % mrem;  // ii in [0, mrem], row mrem is the rhs row
      const int j = k + 1 + jj;
      if (ii < mrem) {
        const int i = k + 1 + ii;
        if (j <= i) A[(size_t)i * n + j] -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      } else {
        rhs[j] -= rhs[k] * A[(size_t)j * n + k];
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (fail_s) {
    if (tid == 0) {
      // DoglegStrategy::ComputeGaussNewtonStep retry: mu *= 10 while mu < max_mu (1.0)
      ws.mu *= 10.0;
      if (ws.mu < 1.0) {
        ws.skip_slot = 1;  // redo the elimination in the next slot; no iteration consumed
      } else {
        ws.gn_failed = 1;  // linear solver FAILURE -> invalid step
      }
    }
    return;
  }
  // 6. backward solve L^T y = z by warp 0
  if (tid < 32) {
    for (int k = n - 1; k >= 0; --k) {
      const double yk = rhs[k] / A[(size_t)k * n + k];
      __syncwarp();
      if (tid == 0) rhs[k] = yk;
      for (int i = tid; i < k; i += 32) rhs[i] -= A[(size_t)k * n + i] * yk;
      __syncwarp();
    }
  }
  __syncthreads();
  bool bad = false;
  double g2 = 0, n2 = 0, gd = 0;
  for (int i = tid; i < n; i += T) {
    const double y = rhs[i];
    if (!isfinite(y)) bad = true;
    const double gni = -y * diag[i];
    gn[i] = gni;
    u[i] = scale[i] * y;
    cvec[i] = scale[i] * grad[i] / diag[i];
    g2 += grad[i] * grad[i];
    n2 += gni * gni;
    gd += grad[i] * gni;
  }
  if (__syncthreads_or(bad)) {
    if (tid == 0) {
      ws.mu *= 10.0;
      if (ws.mu < 1.0)
        ws.skip_slot = 1;
      else
        ws.gn_failed = 1;
    }
    return;
  }
  // Cauchy-point accumulation over the dense rows: |Jd c|^2
  double jg2 = 0;
  for (int r = tid; r < M; r += T) {
    double s = 0;
    for (int i = 0; i < n; ++i) s += Jd[(size_t)r * n + i] * cvec[i];
    jg2 += s * s;
  }
  double v[4] = {g2, n2, gd, jg2};
  double* const dst[4] = {&ws.acc_g2, &ws.acc_n2, &ws.acc_gdot, &ws.acc_Jg2};
  block_atomic_add<4, kDenseThreads>(v, dst);
}

// ------------------------------------------------------------------------------------------ back-substitution
// SPLIT threads share a landmark (they take every SPLIT-th observation and add their sums with shuffles): the
// kernel is bound by the per-observation load latency chain, not by bandwidth, so halving the chain pays.
template <bool HAS_EXT, int SPLIT, bool FUSED = false, int MINB = 1>
__global__ void __launch_bounds__(kLmTile * SPLIT, MINB) k_backsub(Batch b) {  // 152 registers, 3 CTAs/SM: a 128 cap measured 4 % slower
  const int tile = blockIdx.x;
  const int w = b.lm_tile_win[tile];
  WinState& ws = b.ws[w];
  if (ws.done || ws.reuse || ws.skip_slot || ws.gn_failed) return;
  const WinDesc& wd = b.win[w];
  const int l = b.lm_tile_begin[tile] + threadIdx.x / SPLIT;
  const int half = threadIdx.x % SPLIT;
  const int buf = ws.cur;
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // g2, n2, gdot, Jg2, acc_A[0..4]
  const bool in_range = l < wd.lm_end;
  const bool lfix = in_range ? b.lm_fixed[l] != 0 : true;
  double s[3] = {1, 1, 1}, cl[3] = {0, 0, 0}, gr[3] = {0, 0, 0}, dg[3] = {1, 1, 1};
  // per landmark: a3 = sum Jl^T t, jg = sum Jl^T mg, jr = sum Jl^T r, V = sum Jl^T Jl   (t = Jp u, mg = J cauchy)
  double a3[3] = {0, 0, 0}, jg[3] = {0, 0, 0}, jr[3] = {0, 0, 0}, V[6] = {0, 0, 0, 0, 0, 0};
  double s_tr = 0, s_tt = 0, s_gt = 0;
  if (in_range) {
    const double* u = b.u_d + wd.d_off;
    const double* cv = b.c_d + wd.d_off;
    if (!lfix) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        s[k] = b.lm_scale[3 * (size_t)l + k];
        gr[k] = b.lm_grad[3 * (size_t)l + k];
        dg[k] = b.lm_diag[3 * (size_t)l + k];
        cl[k] = s[k] * gr[k] / dg[k];
      }
    }
    const LmPoint lmp = load_lm(b, buf, l);
    const int ob = b.lm_obs_first[l], ost = b.lm_obs_stride[l], nobs = b.lm_obs_cnt[l];
    for (int k = half; k < nobs; k += SPLIT) {
      const int o = ob + k * ost;
      ObsJ J;
      obs_rJ<FUSED>(b, buf, o, lmp, J.r0, J.r1, J.Jp, J.Jl);
      const int offp = b.obs_poff[o];
      double t0 = 0, t1 = 0, m0 = 0, m1 = 0;
      if (offp >= 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          t0 += J.Jp[k] * u[offp + k];
          t1 += J.Jp[6 + k] * u[offp + k];
          m0 += J.Jp[k] * cv[offp + k];
          m1 += J.Jp[6 + k] * cv[offp + k];
        }
      }
      if (HAS_EXT) {
        const int offe = b.pose_off[b.obs_ext[o]];
        if (offe >= 0) {
          double Je[12];
          load_Je(b, buf, o, Je);
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            t0 += Je[k] * u[offe + k];
            t1 += Je[6 + k] * u[offe + k];
            m0 += Je[k] * cv[offe + k];
            m1 += Je[6 + k] * cv[offe + k];
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        m0 += J.Jl[k] * cl[k];
        m1 += J.Jl[3 + k] * cl[k];
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        a3[k] += J.Jl[k] * t0 + J.Jl[3 + k] * t1;
        jg[k] += J.Jl[k] * m0 + J.Jl[3 + k] * m1;
        jr[k] += J.Jl[k] * J.r0 + J.Jl[3 + k] * J.r1;
      }
      V[0] += J.Jl[0] * J.Jl[0] + J.Jl[3] * J.Jl[3];
      V[1] += J.Jl[0] * J.Jl[1] + J.Jl[3] * J.Jl[4];
      V[2] += J.Jl[0] * J.Jl[2] + J.Jl[3] * J.Jl[5];
      V[3] += J.Jl[1] * J.Jl[1] + J.Jl[4] * J.Jl[4];
      V[4] += J.Jl[1] * J.Jl[2] + J.Jl[4] * J.Jl[5];
      V[5] += J.Jl[2] * J.Jl[2] + J.Jl[5] * J.Jl[5];
      acc[3] += m0 * m0 + m1 * m1;
      acc[4] += m0 * J.r0 + m1 * J.r1;
      s_tr += t0 * J.r0 + t1 * J.r1;
      s_tt += t0 * t0 + t1 * t1;
      s_gt += m0 * t0 + m1 * t1;
    }
  }
  if (SPLIT > 1) {
#pragma unroll
    for (int o2 = 1; o2 < SPLIT; o2 <<= 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        a3[k] += __shfl_xor_sync(0xffffffffu, a3[k], o2);
        jg[k] += __shfl_xor_sync(0xffffffffu, jg[k], o2);
        jr[k] += __shfl_xor_sync(0xffffffffu, jr[k], o2);
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) V[k] += __shfl_xor_sync(0xffffffffu, V[k], o2);
      s_tr += __shfl_xor_sync(0xffffffffu, s_tr, o2);
      s_tt += __shfl_xor_sync(0xffffffffu, s_tt, o2);
      s_gt += __shfl_xor_sync(0xffffffffu, s_gt, o2);
      acc[3] += __shfl_xor_sync(0xffffffffu, acc[3], o2);
      acc[4] += __shfl_xor_sync(0xffffffffu, acc[4], o2);
    }
    if (half != 0) acc[3] = acc[4] = 0.0;
  }
  if (in_range && half == 0) {
    acc[6] = acc[3];
    double q[3] = {0, 0, 0};  // s o y: minus the landmark's Gauss-Newton step in the original parameters
    if (!lfix) {
      const double* Vi = b.lm_Vinv + 6 * (size_t)l;
      const double* bs = b.lm_bs + 3 * (size_t)l;
      const double r0 = bs[0] - s[0] * a3[0], r1 = bs[1] - s[1] * a3[1], r2 = bs[2] - s[2] * a3[2];
      const double y0 = Vi[0] * r0 + Vi[1] * r1 + Vi[2] * r2;
      const double y1 = Vi[1] * r0 + Vi[3] * r1 + Vi[4] * r2;
      const double y2 = Vi[2] * r0 + Vi[4] * r1 + Vi[5] * r2;
      const double g0 = -y0 * dg[0], g1 = -y1 * dg[1], g2 = -y2 * dg[2];
      double* gn = b.lm_gn + 3 * (size_t)l;
      gn[0] = g0;
      gn[1] = g1;
      gn[2] = g2;
      acc[0] = gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2];
      acc[1] = g0 * g0 + g1 * g1 + g2 * g2;
      acc[2] = gr[0] * g0 + gr[1] * g1 + gr[2] * g2;
      if (!isfinite(y0) || !isfinite(y1) || !isfinite(y2)) acc[1] = nan("");
      q[0] = s[0] * y0; q[1] = s[1] * y1; q[2] = s[2] * y2;
    }
    // an = t + Jl q per observation, summed in closed form
    const double Vq0 = V[0] * q[0] + V[1] * q[1] + V[2] * q[2];
    const double Vq1 = V[1] * q[0] + V[3] * q[1] + V[4] * q[2];
    const double Vq2 = V[2] * q[0] + V[4] * q[1] + V[5] * q[2];
    acc[5] = s_tr + q[0] * jr[0] + q[1] * jr[1] + q[2] * jr[2];
    acc[7] = s_gt + q[0] * jg[0] + q[1] * jg[1] + q[2] * jg[2];
    acc[8] = s_tt + 2.0 * (q[0] * a3[0] + q[1] * a3[1] + q[2] * a3[2]) + q[0] * Vq0 + q[1] * Vq1 + q[2] * Vq2;
  }
  double* sa = b.shard_acc ? b.shard_acc + kShardAcc * (size_t)w : nullptr;
  double* const dst[9] = {sa ? sa + 0 : &ws.acc_g2,   sa ? sa + 1 : &ws.acc_n2,   sa ? sa + 2 : &ws.acc_gdot,
                          sa ? sa + 3 : &ws.acc_Jg2,  sa ? sa + 4 : &ws.acc_A[0], sa ? sa + 5 : &ws.acc_A[1],
                          sa ? sa + 6 : &ws.acc_A[2], sa ? sa + 7 : &ws.acc_A[3], sa ? sa + 8 : &ws.acc_A[4]};
  block_atomic_add<9, kLmTile * SPLIT>(acc, dst);
}

// ------------------------------------------------------------------------------------------ dogleg step
// One CTA per window: iteration accounting, dogleg interpolation coefficients, dense part of the step,
// candidate poses / speed-biases, dense rows of the model cost change.
__global__ void __launch_bounds__(128) k_step_dense(Batch b, SvinBaOptions opt) {
  const int w = blockIdx.x;
  WinState& ws = b.ws[w];
  if (ws.done || ws.skip_slot) return;
  const WinDesc& wd = b.win[w];
  const int tid = threadIdx.x, T = blockDim.x;
  const int n = wd.n_dense, M = wd.n_rows, buf = ws.cur;
  __shared__ double cg_s, cn_s;
  __shared__ int failed_s;
  __syncthreads();
  if (tid == 0) {
    ws.iter += 1;
    ws.last_successful = 0;
    ws.t_iter_start_ns = globaltimer_ns();
    if (!ws.gn_failed) {
      if (!ws.reuse) {
        ws.alpha = ws.acc_g2 / ws.acc_Jg2;
        if (!isfinite(ws.acc_n2)) ws.gn_failed = 1;  // IsArrayValid(gauss_newton_step_) failed
      }
    }
    double cg = 0, cn = 0;
    if (!ws.gn_failed) {
      // DoglegStrategy::ComputeTraditionalDoglegStep
      const double g2 = ws.acc_g2, n2 = ws.acc_n2, gdot = ws.acc_gdot, radius = ws.radius, alpha = ws.alpha;
      const double gradient_norm = sqrt(g2), gauss_newton_norm = sqrt(n2);
      if (gauss_newton_norm <= radius) {
        cg = 0.0;
        cn = 1.0;
        ws.dogleg_step_norm = gauss_newton_norm;
      } else if (gradient_norm * alpha >= radius) {
        cg = -(radius / gradient_norm);
        cn = 0.0;
        ws.dogleg_step_norm = radius;
      } else {
        const double b_dot_a = -alpha * gdot;
        const double a_squared_norm = (alpha * gradient_norm) * (alpha * gradient_norm);
        const double b_minus_a_squared_norm = a_squared_norm - 2 * b_dot_a + gauss_newton_norm * gauss_newton_norm;
        const double c = b_dot_a - a_squared_norm;
        const double d = sqrt(c * c + b_minus_a_squared_norm * (radius * radius - a_squared_norm));
        const double beta = (c <= 0) ? (d - c) / b_minus_a_squared_norm : (radius * radius - a_squared_norm) / (d + c);
        cg = -alpha * (1.0 - beta);
        cn = beta;
        ws.dogleg_step_norm = sqrt(fmax(0.0, cg * cg * g2 + 2.0 * cg * cn * gdot + cn * cn * n2));
      }
    }
    ws.cg = cg;
    ws.cn = cn;
    ws.reuse = 1;
    cg_s = cg;
    cn_s = cn;
    failed_s = ws.gn_failed;
  }
  __syncthreads();
  if (failed_s) return;
  const double cg = cg_s, cn = cn_s;
  const double* scale = b.scale_d + wd.d_off;
  const double* diag = b.diag_d + wd.d_off;
  const double* grad = b.grad_d + wd.d_off;
  const double* gn = b.gn_d + wd.d_off;
  double* delta = b.delta_d + wd.d_off;
  for (int i = tid; i < n; i += T) delta[i] = scale[i] * (cg * grad[i] + cn * gn[i]) / diag[i];
  __syncthreads();
  double acc[3] = {0, 0, 0};  // mc, step2, xnorm2
  // candidate dense state (x (+) delta) into the other buffer; fixed blocks are copied
  const int np = wd.pose_end - wd.pose_begin, ns = wd.sb_end - wd.sb_begin;
  for (int k = tid; k < np + ns; k += T) {
    if (k < np) {
      const int p = wd.pose_begin + k;
      const double* x = b.pose[buf] + 7 * (size_t)p;
      double* y = b.pose[1 - buf] + 7 * (size_t)p;
      const int off = b.pose_off[p];
      if (off >= 0) {
        pose_plus(x, delta + off, y);
        for (int c = 0; c < 7; ++c) {
          acc[1] += (x[c] - y[c]) * (x[c] - y[c]);
          acc[2] += x[c] * x[c];
        }
      } else {
        for (int c = 0; c < 7; ++c) y[c] = x[c];
      }
    } else {
      const int s = wd.sb_begin + k - np;
      const double* x = b.sb[buf] + 9 * (size_t)s;
      double* y = b.sb[1 - buf] + 9 * (size_t)s;
      const int off = b.sb_off[s];
      for (int c = 0; c < 9; ++c) {
        y[c] = (off >= 0) ? x[c] + delta[off + c] : x[c];
        if (off >= 0) {
          acc[1] += (x[c] - y[c]) * (x[c] - y[c]);
          acc[2] += x[c] * x[c];
        }
      }
    }
  }
  // dense rows of the model: m = Jd * delta
  const double* Jd = b.Jd[buf] + wd.Jd_off;
  const double* rd = b.rd[buf] + wd.rd_off;
  for (int r = tid; r < M; r += T) {
    double m = 0;
    for (int i = 0; i < n; ++i) m += Jd[(size_t)r * n + i] * delta[i];
    acc[0] += m * (rd[r] + m / 2.0);
  }
  // reprojection rows of the model from the sums k_backsub left: m = cg * mg - cn * an per observation
  if (tid == 0)
    acc[0] += cg * ws.acc_A[0] - cn * ws.acc_A[1] +
              0.5 * (cg * cg * ws.acc_A[2] - 2.0 * cg * cn * ws.acc_A[3] + cn * cn * ws.acc_A[4]);
  double* const dst[3] = {&ws.acc_mc, &ws.acc_step2, &ws.acc_xnorm2};
  block_atomic_add<3, 128>(acc, dst);
}

template <bool HAS_EXT>
__global__ void __launch_bounds__(kLmTile) k_step_lm(Batch b) {
  const int tile = blockIdx.x;
  const int w = b.lm_tile_win[tile];
  WinState& ws = b.ws[w];
  if (ws.done || ws.skip_slot || ws.gn_failed) return;
  const WinDesc& wd = b.win[w];
  const int l = b.lm_tile_begin[tile] + threadIdx.x;
  const int buf = ws.cur;
  double acc[2] = {0, 0};  // step2, xnorm2 (the model rows of the landmark terms come from k_backsub's sums)
  if (l < wd.lm_end) {
    const double cg = ws.cg, cn = ws.cn;
    const bool lfix = b.lm_fixed[l] != 0;
    const double* x = b.lm[buf] + 4 * (size_t)l;
    double* y = b.lm[1 - buf] + 4 * (size_t)l;
    double dl[3] = {0, 0, 0};
    const double x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3];
    if (!lfix) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
        dl[k] = b.lm_scale[3 * (size_t)l + k] *
                (cg * b.lm_grad[3 * (size_t)l + k] + cn * b.lm_gn[3 * (size_t)l + k]) / b.lm_diag[3 * (size_t)l + k];
      acc[0] = dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2];
      acc[1] = x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3;
    }
    y[0] = x0 + dl[0];
    y[1] = x1 + dl[1];
    y[2] = x2 + dl[2];
    y[3] = x3;
  }
  double* sa = b.shard_acc ? b.shard_acc + kShardAcc * (size_t)w : nullptr;
  double* const dst[2] = {sa ? sa + 9 : &ws.acc_step2, sa ? sa + 10 : &ws.acc_xnorm2};
  block_atomic_add<2, kLmTile>(acc, dst);
}

// ------------------------------------------------------------------------------------------ accept / reject
__device__ __forceinline__ void clear_accumulators(WinState& ws) {
  ws.acc_g2 = ws.acc_n2 = ws.acc_gdot = ws.acc_Jg2 = 0.0;
  ws.acc_A[0] = ws.acc_A[1] = ws.acc_A[2] = ws.acc_A[3] = ws.acc_A[4] = 0.0;
}
__global__ void k_decide(Batch b, SvinBaOptions opt) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= b.B) return;
  WinState& ws = b.ws[w];
  if (ws.done) return;
  if (ws.skip_slot) {
    ws.skip_slot = 0;
    clear_accumulators(ws);
    ws.gmax_bits = 0ull;
    return;
  }
  const double model_cost_change = -ws.acc_mc;
  if (ws.gn_failed || !(model_cost_change > 0.0)) {
    // HandleInvalidStep + DoglegStrategy::StepIsInvalid
    ws.invalid += 1;
    if (ws.invalid >= opt.max_num_consecutive_invalid_steps) {
      ws.done = 1;
      ws.termination = SVIN_TERM_FAILURE;
    }
    ws.mu *= 10.0;
    ws.reuse = 0;
    ws.gn_failed = 0;
  } else {
    ws.invalid = 0;
    const double step_norm = sqrt(ws.acc_step2), x_norm = sqrt(ws.acc_xnorm2);
    const double cost_change = ws.cost_x - ws.cost_cand;
    if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
      ws.done = 1;
      ws.termination = SVIN_TERM_CONVERGENCE;
    } else if (fabs(cost_change) <= opt.function_tolerance * ws.cost_x) {
      ws.done = 1;
      ws.termination = SVIN_TERM_CONVERGENCE;
    } else {
      const double relative_decrease = cost_change / model_cost_change;
      if (relative_decrease > opt.min_relative_decrease) {
        // HandleSuccessfulStep + DoglegStrategy::StepAccepted
        ws.cur ^= 1;
        ws.cost_x = ws.cost_cand;
        ws.last_successful = 1;
        ws.num_successful += 1;
        if (relative_decrease < 0.25) ws.radius *= 0.5;
        if (relative_decrease > 0.75) ws.radius = fmax(ws.radius, 3.0 * ws.dogleg_step_norm);
        ws.mu = fmax(1e-8, 2.0 * ws.mu / 10.0);
        ws.reuse = 0;
      } else {
        ws.radius *= 0.5;  // StepRejected; reuse stays set
      }
    }
  }
  ws.cost_cand = 0.0;
  ws.acc_mc = ws.acc_step2 = ws.acc_xnorm2 = 0.0;
  if (!ws.reuse) {
    clear_accumulators(ws);
    ws.gmax_bits = 0ull;
  }
  if (!ws.done) {
    const unsigned long long now = globaltimer_ns();
    const double iter_time = (double)(now - ws.t_iter_start_ns) * 1e-9;
    const double cumulative = (double)(now - ws.t_start_ns) * 1e-9;
    if (opt.time_limit_seconds >= 0.0 && ws.iter >= opt.min_num_iterations &&
        cumulative + iter_time > opt.time_limit_seconds) {
      ws.done = 1;
      ws.termination = SVIN_TERM_USER_SUCCESS;
    } else if (ws.iter >= opt.max_num_iterations) {
      ws.done = 1;
      ws.termination = SVIN_TERM_NO_CONVERGENCE;
    } else if (ws.radius < opt.min_trust_region_radius) {
      ws.done = 1;
      ws.termination = SVIN_TERM_CONVERGENCE;
    }
  }
}

// after the initial linearisation
__global__ void k_init(Batch b, SvinBaOptions opt) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= b.B) return;
  WinState& ws = b.ws[w];
  ws.cost_x = ws.cost_cand;
  ws.initial_cost = ws.cost_cand;
  ws.cost_cand = 0.0;
  ws.radius = opt.initial_trust_region_radius;
  ws.t_start_ns = globaltimer_ns();
  if (opt.max_num_iterations <= 0) {
    ws.done = 1;
    ws.termination = SVIN_TERM_NO_CONVERGENCE;
  }
}

__global__ void k_count_active(Batch b, int* out) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= b.B) return;
  if (!b.ws[w].done) atomicAdd(out, 1);
}

// Estimator.cpp:903-922: H = sum J1min^T J1min over the landmark's residuals (no loss), quality = sqrt(lmin/lmax)
__global__ void __launch_bounds__(kLmTile) k_quality(Batch b) {
  const int tile = blockIdx.x;
  const int w = b.lm_tile_win[tile];
  const WinDesc& wd = b.win[w];
  const WinState& ws = b.ws[w];
  const int l = b.lm_tile_begin[tile] + threadIdx.x;
  if (l >= wd.lm_end) return;
  const int buf = ws.cur;
  double Hm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const int ob = b.lm_obs_first[l], ost = b.lm_obs_stride[l], nobs = b.lm_obs_cnt[l];
  for (int k = 0; k < nobs; ++k) {
    const int o = ob + k * ost;
    Reproj R;
    reproj_eval<true, false>(b.pose[buf] + 7 * (size_t)b.obs_pose[o], b.lm[buf] + 4 * (size_t)l,
                             b.pose[buf] + 7 * (size_t)b.obs_ext[o], b.intr + 8 * (size_t)b.obs_cam[o], b.obs_zx[o],
                             b.obs_zy[o], b.obs_u00[o], b.obs_u01[o], b.obs_u11[o], wd.loss_type, wd.loss_scale, true,
                             R);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Hm[i * 3 + j] += R.Jl[i] * R.Jl[j] + R.Jl[3 + i] * R.Jl[3 + j];
  }
  double ev[3];
  sym3_eigenvalues(Hm, ev);
  b.lm_quality[l] = (ev[0] < 1.0e-12) ? 0.0 : sqrt(ev[0]) / sqrt(ev[2]);
}

static inline unsigned div_up(unsigned a, unsigned b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------ sharded mode
// stage 2: after k_backsub; 3: after k_linearize (landmark step norms + candidate reprojection cost).  Adds the
// all-reduced landmark-side sums to the replicated dense-side sums in WinState and clears them.
__global__ void k_fold(Batch b, int stage) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= b.B) return;
  WinState& ws = b.ws[w];
  double* sa = b.shard_acc + kShardAcc * (size_t)w;
  // atomics: the dense terms of the candidate are evaluated concurrently on the side stream and add to the same sums
  if (stage == 2) {
    atomicAdd(&ws.acc_g2, sa[0]); atomicAdd(&ws.acc_n2, sa[1]); atomicAdd(&ws.acc_gdot, sa[2]); atomicAdd(&ws.acc_Jg2, sa[3]);
    for (int k = 0; k < 5; ++k) atomicAdd(&ws.acc_A[k], sa[4 + k]);
    for (int k = 0; k < 9; ++k) sa[k] = 0.0;
  } else {
    atomicAdd(&ws.acc_step2, sa[9]); atomicAdd(&ws.acc_xnorm2, sa[10]);
    atomicAdd(&ws.cost_cand, sa[11]);
    sa[9] = sa[10] = sa[11] = 0.0;
  }
}
// unpack = 0: this rank's landmark gradient max into its slot of the (cleared) exchange buffer; 1: max over the ranks' slots
__global__ void k_gmax_pack(Batch b, int unpack) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= b.B) return;
  double* slots = b.gmax_buf + (size_t)w * b.comm_world;
  if (!unpack) {
    slots[b.comm_rank] = __longlong_as_double((long long)b.ws[w].gmax_bits);
  } else {
    double m = 0.0;
    for (int r = 0; r < b.comm_world; ++r) m = fmax(m, slots[r]);
    b.ws[w].gmax_bits = (unsigned long long)__double_as_longlong(m);
  }
}
__global__ void k_obs_poff(Batch b) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o < b.NOBS) b.obs_poff[o] = b.pose_off[b.obs_pose[o]];
}
// Upload-time gather: internal observation g of window w takes caller observation order[g] of that window; indices
// become batch-global, U = upper-triangular sqrt(information) as Eigen's LLT gives it (llt_sqrt_information2 on the
// host before: sqrt / div / mul / sub in the same order, no contraction, so the bits are the same).
__global__ void __launch_bounds__(kObsTile) k_pack_obs(Batch b, RawObs raw) {
  const int tile = blockIdx.x;
  const int w = b.obs_tile_win[tile];
  const int g = b.obs_tile_begin[tile] + threadIdx.x;
  const WinDesc& wd = b.win[w];
  if (g >= wd.obs_end) return;
  const int src = wd.obs_begin + raw.order[g];
  int ip, ie, ic, il;
  if (raw.one_word) {   // landmark | pose << 18 | ext << 24 | cam << 30
    const unsigned v = (unsigned)raw.pec[src];
    il = (int)(v & 0x3ffffu);
    ip = (int)((v >> 18) & 63u);
    ie = (int)((v >> 24) & 63u);
    ic = (int)(v >> 30);
  } else if (raw.pec) {   // one word per observation: pose | ext << 10 | cam << 20
    const int v = raw.pec[src];
    ip = v & 1023;
    ie = (v >> 10) & 1023;
    ic = (v >> 20) & 1023;
    il = raw.lm[src];
  } else {
    ip = raw.pose[src];
    ie = raw.ext[src];
    ic = raw.cam[src];
    il = raw.lm[src];
  }
  b.obs_pose[g] = wd.pose_begin + ip;
  b.obs_lm[g] = wd.lm_begin + raw.lm_inv[wd.lm_begin + il];
  b.obs_ext[g] = wd.pose_begin + ie;
  b.obs_cam[g] = wd.cam_begin + ic;
  if (raw.meas32) {
    const float2 z = reinterpret_cast<const float2*>(raw.meas32)[src];
    b.obs_zx[g] = (double)z.x;
    b.obs_zy[g] = (double)z.y;
  } else {
    b.obs_zx[g] = raw.meas[2 * (size_t)src];
    b.obs_zy[g] = raw.meas[2 * (size_t)src + 1];
  }
  double a0, a2, a3;
  if (wd.info_uniform) {   // one information matrix for the whole window (single-scale detector: one keypoint size)
    a0 = wd.info3[0];
    a2 = wd.info3[1];
    a3 = wd.info3[2];
  } else {
    a0 = raw.info3[3 * (size_t)src];
    a2 = raw.info3[3 * (size_t)src + 1];
    a3 = raw.info3[3 * (size_t)src + 2];
  }
  double l00 = a0, l10 = a2, l11 = a3;
  if (a0 > 0.0) {
    l00 = sqrt(a0);
    l10 = __ddiv_rn(a2, l00);
    const double x = __dsub_rn(a3, __dmul_rn(l10, l10));
    if (x > 0.0) l11 = sqrt(x);
  }
  b.obs_u00[g] = l00;
  b.obs_u01[g] = l10;
  b.obs_u11[g] = l11;
}
void launch_pack_obs(const Batch& b, const RawObs& raw, cudaStream_t st) {
  if (b.n_obs_tiles == 0) return;
  k_pack_obs<<<b.n_obs_tiles, kObsTile, 0, st>>>(b, raw);
}
void launch_obs_poff(const Batch& b, cudaStream_t st) {
  if (b.NOBS) k_obs_poff<<<div_up(b.NOBS, 256), 256, 0, st>>>(b);
}
void launch_fold(const Batch& b, int stage, cudaStream_t st) { k_fold<<<div_up(b.B, 64), 64, 0, st>>>(b, stage); }
void launch_gmax_pack(const Batch& b, int unpack, cudaStream_t st) {
  k_gmax_pack<<<div_up(b.B, 64), 64, 0, st>>>(b, unpack);
}

size_t schur_mma_smem_bytes() { return 4 * (size_t)kSchurWarpDoubles * sizeof(double); }
int schur_lr_max_chunk(int runs, int wpc) {
  // landmarks per run-parallel chunk of wpc warps: runs * c <= 32 * wpc units, c <= 32, and Y fits the operand buffer
  if (runs < 1 || runs > 32) return 0;
  const int ycap = wpc == 1 ? LrCfg<1>::kY : (wpc == 2 ? LrCfg<2>::kY : LrCfg<4>::kY);
  int c = std::min(32, 32 * wpc / runs);
  while (c > 0 && (6 * runs + 1) * (((3 * c + 3) / 4) * 4 + 4) > ycap) --c;
  return c;
}
int schur_mma_max_chunk(int runs) {
  // largest chunk size c <= 32 with (6 runs + 1) * (pad4(3c) + 4) <= kSchurYDoubles
  if (runs > kSchurMaxRuns) return 0;
  const int rows = 6 * runs + 1;
  int best = 0;
  for (int c = 1; c <= 32; ++c)
    if (rows * (((3 * c + 3) / 4) * 4 + 4) <= kSchurYDoubles) best = c;
  return best;
}
cudaError_t configure_schur() {
  const int sm = (int)schur_mma_smem_bytes();
  cudaError_t e = cudaSuccess;
  auto set = [&](const void* f) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
  };
  set((const void*)k_schur_mma<1, false>);
  set((const void*)k_schur_mma<2, false>);
  set((const void*)k_schur_mma<4, false>);
  set((const void*)k_schur_mma<1, true>);
  set((const void*)k_schur_mma<2, true>);
  set((const void*)k_schur_mma<4, true>);
  return e;
}
int schur_chunk_class(int count) { return count <= 8 ? 2 : (count <= 16 ? 1 : 0); }

// ------------------------------------------------------------------------------------------ launchers

void launch_linearize(const Batch& b, int which, int raw, cudaStream_t st, bool compact) {
  if (b.n_obs_tiles == 0) return;
  static const int minb = std::getenv("SVIN_LIN_MINB") ? std::atoi(std::getenv("SVIN_LIN_MINB")) : 5;  // A/B knob: 5 CTAs/SM (96 registers, 100 B spilled) measured fastest
  if (compact && !raw && !b.has_ext) {
    // 8 CTAs/SM (64 registers, 170 B spilled): 2.05 ms per 11 launches vs 2.14 at 5 CTAs/SM and 2.27 at 6 (r2aj)
    k_linearize<false, 8, true><<<b.n_obs_tiles, kObsTile, 0, st>>>(b, which, raw);
    return;
  }
  if (b.has_ext || (raw && b.lin_Je[1] != nullptr))
    k_linearize<true, 4><<<b.n_obs_tiles, kObsTile, 0, st>>>(b, which, raw);
  else if (minb == 5)
    k_linearize<false, 5><<<b.n_obs_tiles, kObsTile, 0, st>>>(b, which, raw);
  else if (minb == 6)
    k_linearize<false, 6><<<b.n_obs_tiles, kObsTile, 0, st>>>(b, which, raw);
  else
    k_linearize<false, 4><<<b.n_obs_tiles, kObsTile, 0, st>>>(b, which, raw);
}
void launch_dense_eval(const Batch& b, int which, int raw, const double* const* dump, cudaStream_t st) {
  ImuEvalOut d{nullptr, nullptr, nullptr, nullptr, nullptr};
  if (dump) {
    d.r = const_cast<double*>(dump[0]);
    d.J0 = const_cast<double*>(dump[1]);
    d.J1 = const_cast<double*>(dump[2]);
    d.J2 = const_cast<double*>(dump[3]);
    d.J3 = const_cast<double*>(dump[4]);
  }
  static const int minb = std::getenv("SVIN_DENSE_MINB") ? std::atoi(std::getenv("SVIN_DENSE_MINB")) : 2;  // A/B knob
  if (minb == 4)
    k_dense_eval<4><<<b.B, 128, 0, st>>>(b, which, raw, d);
  else if (minb == 3)
    k_dense_eval<3><<<b.B, 128, 0, st>>>(b, which, raw, d);
  else
    k_dense_eval<2><<<b.B, 128, 0, st>>>(b, which, raw, d);
}
// The chunk kernels of one slot are independent (they only meet in the fp64 REDs into the reduced system): with
// `par` they run concurrently on the auxiliary streams (forked from / joined to `st` with events, capturable into the
// solver's graph) so that one kernel's tail overlaps the next one's body instead of six drained boundaries per slot.
int launch_schur(const Batch& b, const SvinBaOptions& opt, cudaStream_t st, const SchurStreams* par) {
  if (b.n_lm_tiles == 0) return 0;
  int launched = 1;
  if (b.has_ext)
    k_schur<true><<<b.n_lm_tiles, kLmTile, 0, st>>>(b, opt);
  else if (b.n_schur_warps > 0) {
    // class 0: > 16 landmarks per chunk (lane = landmark), 1: 9..16 (2 lanes / landmark), 2: <= 8 (4 lanes / landmark),
    // 3, 7, 8: run-parallel chunks of 1, 2, 4 warps (k_schur_lr), 4..6: warp-per-run chunks of 2..4 runs (k_schur_wr).
    // The expensive low-fill classes go first so that the cheap full chunks fill the tail of the grid.
    // SVIN_SCHUR_CLASSMASK (diagnostics only: results are wrong when a class is skipped) times the classes separately.
    static const int mask = std::getenv("SVIN_SCHUR_CLASSMASK") ? std::atoi(std::getenv("SVIN_SCHUR_CLASSMASK")) : 511;
    const int* lst[kSchurClasses];
    lst[0] = b.sw_list;
    for (int k = 1; k < kSchurClasses; ++k) lst[k] = lst[k - 1] + b.sw_class_count[k - 1];
    const int* cc = b.sw_class_count;
    const size_t sm = schur_mma_smem_bytes();
    int used = 0;  // kernels launched so far: the first stays on `st`, the others go round-robin over the aux streams
    bool forked = false;
    auto next_stream = [&]() -> cudaStream_t {
      const int k = used++;
      if (!par || par->n < 1 || k == 0) return st;
      if (!forked) {
        cudaEventRecord(par->fork, st);
        forked = true;
      }
      const int a = (k - 1) % par->n;
      if (k - 1 < par->n) cudaStreamWaitEvent(par->aux[a], par->fork, 0);
      return par->aux[a];
    };
    if (b.fused) {
      if (cc[8] && (mask & 256)) k_schur_lr<4, 12, true><<<cc[8], 128, LrCfg<4>::kDoubles * sizeof(double), next_stream()>>>(b, opt, lst[8]);
      if (cc[7] && (mask & 128)) k_schur_lr<2, 12, true><<<cc[7], 64, LrCfg<2>::kDoubles * sizeof(double), next_stream()>>>(b, opt, lst[7]);
      if (cc[3] && (mask & 8)) k_schur_lr<1, 12, true><<<cc[3], 32, LrCfg<1>::kDoubles * sizeof(double), next_stream()>>>(b, opt, lst[3]);
      if (cc[6] && (mask & 64)) k_schur_wr<4, true><<<cc[6], 128, WrCfg<4>::kDoubles * sizeof(double), next_stream()>>>(b, opt, lst[6]);
      if (cc[5] && (mask & 32)) k_schur_wr<3, true><<<cc[5], 96, WrCfg<3>::kDoubles * sizeof(double), next_stream()>>>(b, opt, lst[5]);
      if (cc[4] && (mask & 16)) k_schur_wr<2, true><<<cc[4], 64, WrCfg<2>::kDoubles * sizeof(double), next_stream()>>>(b, opt, lst[4]);
      if (cc[2] && (mask & 4)) k_schur_mma<4, true><<<div_up(cc[2], 4), 128, sm, next_stream()>>>(b, opt, lst[2], cc[2]);
      if (cc[1] && (mask & 2)) k_schur_mma<2, true><<<div_up(cc[1], 4), 128, sm, next_stream()>>>(b, opt, lst[1], cc[1]);
      if (cc[0] && (mask & 1)) k_schur_mma<1, true><<<div_up(cc[0], 4), 128, sm, next_stream()>>>(b, opt, lst[0], cc[0]);
    } else {
      if (cc[8] && (mask & 256)) k_schur_lr<4, 12, false><<<cc[8], 128, LrCfg<4>::kDoubles * sizeof(double), next_stream()>>>(b, opt, lst[8]);
      if (cc[7] && (mask & 128)) k_schur_lr<2, 12, false><<<cc[7], 64, LrCfg<2>::kDoubles * sizeof(double), next_stream()>>>(b, opt, lst[7]);
      if (cc[3] && (mask & 8)) k_schur_lr<1, 12, false><<<cc[3], 32, LrCfg<1>::kDoubles * sizeof(double), next_stream()>>>(b, opt, lst[3]);
      if (cc[6] && (mask & 64)) k_schur_wr<4, false><<<cc[6], 128, WrCfg<4>::kDoubles * sizeof(double), next_stream()>>>(b, opt, lst[6]);
      if (cc[5] && (mask & 32)) k_schur_wr<3, false><<<cc[5], 96, WrCfg<3>::kDoubles * sizeof(double), next_stream()>>>(b, opt, lst[5]);
      if (cc[4] && (mask & 16)) k_schur_wr<2, false><<<cc[4], 64, WrCfg<2>::kDoubles * sizeof(double), next_stream()>>>(b, opt, lst[4]);
      if (cc[2] && (mask & 4)) k_schur_mma<4, false><<<div_up(cc[2], 4), 128, sm, next_stream()>>>(b, opt, lst[2], cc[2]);
      if (cc[1] && (mask & 2)) k_schur_mma<2, false><<<div_up(cc[1], 4), 128, sm, next_stream()>>>(b, opt, lst[1], cc[1]);
      if (cc[0] && (mask & 1)) k_schur_mma<1, false><<<div_up(cc[0], 4), 128, sm, next_stream()>>>(b, opt, lst[0], cc[0]);
    }
    if (forked) {
      const int touched = std::min(used - 1, par->n);
      for (int a = 0; a < touched; ++a) {
        cudaEventRecord(par->join[a], par->aux[a]);
        cudaStreamWaitEvent(st, par->join[a], 0);
      }
    }
    launched = used;
  }
  else
    k_schur<false><<<b.n_lm_tiles, kLmTile, 0, st>>>(b, opt);
  return launched;  // kernels launched (for the launch accounting of svin_ba_timings)
}
// generic fallback (reduced system stays in global memory); the fast path is in ba_dense_solve.cu
void launch_dense_solve_generic(const Batch& b, const SvinBaOptions& opt, cudaStream_t st) {
  k_dense_solve<<<b.B, kDenseThreads, 0, st>>>(b, opt, 0);
}
void launch_backsub(const Batch& b, cudaStream_t st) {
  if (b.n_lm_tiles == 0) return;
  if (b.has_ext)
    k_backsub<true, 1, false><<<b.n_lm_tiles, kLmTile, 0, st>>>(b);
  else if (b.fused)
    // 4 CTAs/SM (128 registers): 1.30 ms per 10 launches; 168 registers 1.72, 96 registers 1.30, 80 registers 1.41 (r2aj)
    k_backsub<false, 1, true, 4><<<b.n_lm_tiles, kLmTile, 0, st>>>(b);
  else
    k_backsub<false, 1, false><<<b.n_lm_tiles, kLmTile, 0, st>>>(b);
}
void launch_step_dense(const Batch& b, const SvinBaOptions& opt, cudaStream_t st) {
  k_step_dense<<<b.B, 128, 0, st>>>(b, opt);
}
void launch_step_lm(const Batch& b, cudaStream_t st) {
  if (b.n_lm_tiles == 0) return;
  if (b.has_ext)
    k_step_lm<true><<<b.n_lm_tiles, kLmTile, 0, st>>>(b);
  else
    k_step_lm<false><<<b.n_lm_tiles, kLmTile, 0, st>>>(b);
}
void launch_decide(const Batch& b, const SvinBaOptions& opt, cudaStream_t st) {
  k_decide<<<div_up(b.B, 64), 64, 0, st>>>(b, opt);
}
void launch_init(const Batch& b, const SvinBaOptions& opt, cudaStream_t st) {
  k_init<<<div_up(b.B, 64), 64, 0, st>>>(b, opt);
}
void launch_count_active(const Batch& b, int* out, cudaStream_t st) {
  k_count_active<<<div_up(b.B, 64), 64, 0, st>>>(b, out);
}
// copy the accepted estimate (buffer ws.cur of each window) into contiguous output arrays
__global__ void k_gather_state(Batch b, const int* pose_win, const int* sb_win, double* pose_out, double* sb_out,
                               double* lm_out) {
  const int stride = gridDim.x * blockDim.x;
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = t0; i < b.NPB * 7; i += stride) pose_out[i] = b.pose[b.ws[pose_win[i / 7]].cur][i];
  for (int i = t0; i < b.NSB * 9; i += stride) sb_out[i] = b.sb[b.ws[sb_win[i / 9]].cur][i];
  for (int i = t0; i < b.NL * 4; i += stride) lm_out[i] = b.lm[b.ws[b.lm_win[i / 4]].cur][i];
}
void launch_gather_state(const Batch& b, const int* pose_win, const int* sb_win, double* pose_out, double* sb_out,
                         double* lm_out, cudaStream_t st) {
  const int n = max(b.NPB * 7, max(b.NSB * 9, b.NL * 4));
  k_gather_state<<<max(1u, min(div_up(n, 256), 1184u)), 256, 0, st>>>(b, pose_win, sb_win, pose_out, sb_out, lm_out);
}

void launch_quality(const Batch& b, cudaStream_t st) {
  if (b.n_lm_tiles == 0) return;
  k_quality<<<b.n_lm_tiles, kLmTile, 0, st>>>(b);
}

}  // namespace svin
