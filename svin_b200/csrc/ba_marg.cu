// Marginalisation core on the device.
//   k_marg_landmarks : thread per landmark — JtJ / Jtr of the reprojection terms at the linearisation point and the
//                      landmark Schur with the preconditioned 3x3 pseudo-inverse
//                      (okvis_ceres/src/MarginalizationError.cpp:556-619, pseudoInverseSymmSqrt
//                       okvis_ceres/include/okvis/ceres/implementation/MarginalizationError.hpp:193-220)
//   k_marg_dense     : one CTA — existing prior + dense terms (MarginalizationError.cpp:333-383), dense Schur with a
//                      pseudo-inverse from a Jacobi eigen-decomposition (:621-667), updateErrorComputation (:725-758)
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>

#include "ba_device_utils.cuh"
#include "ba_kernels.cuh"
#include "ba_math.cuh"

namespace svin {

// symmetric 3x3 eigen-decomposition (cyclic Jacobi) with eigenvectors; A row-major
__device__ void sym3_eigh(const double* A_, double* ev, double* U) {
  double A[9];
  for (int i = 0; i < 9; ++i) {
    A[i] = A_[i];
    U[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }
  for (int sweep = 0; sweep < 40; ++sweep) {
    const double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    if (off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = A[p * 3 + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = A[k * 3 + p], akq = A[k * 3 + q];
          A[k * 3 + p] = c * akp - s * akq;
          A[k * 3 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = A[p * 3 + k], aqk = A[q * 3 + k];
          A[p * 3 + k] = c * apk - s * aqk;
          A[q * 3 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double ukp = U[k * 3 + p], ukq = U[k * 3 + q];
          U[k * 3 + p] = c * ukp - s * ukq;
          U[k * 3 + q] = s * ukp + c * ukq;
        }
      }
  }
  ev[0] = A[0];
  ev[1] = A[4];
  ev[2] = A[8];
}

template <bool HAS_EXT>
__global__ void __launch_bounds__(128) k_marg_landmarks(Batch b, MargArgs m) {
  const WinDesc& wd = b.win[m.w];
  const int l = wd.lm_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= wd.lm_end) return;
  const int buf = b.ws[m.w].cur;
  const int n = m.n;
  const int ob = b.lm_obs_first[l], ost = b.lm_obs_stride[l], nobs = b.lm_obs_cnt[l];
  const size_t S = b.obs_stride;
  const double* rP = b.lin_r[buf];
  const double* JpP = b.lin_Jp[buf];
  const double* JlP = b.lin_Jl[buf];
  const double* JeP = HAS_EXT ? b.lin_Je[buf] : nullptr;
  // V, b_l (b0 -= J^T r)
  double V[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};
  for (int k = 0; k < nobs; ++k) {
    const int o = ob + k * ost;
    const double r0 = rP[o], r1 = rP[S + o];
    double a[6];
    for (int q = 0; q < 6; ++q) a[q] = JlP[q * S + o];
    for (int i = 0; i < 3; ++i) {
      bl[i] -= a[i] * r0 + a[3 + i] * r1;
      for (int j = 0; j < 3; ++j) V[i * 3 + j] += a[i] * a[j] + a[3 + i] * a[3 + j];
    }
  }
  double pl[3], Vs[9], ev[3], U[9], Veff[9];
  for (int i = 0; i < 3; ++i) pl[i] = V[i * 3 + i] > 1.0e-9 ? sqrt(V[i * 3 + i]) : 1.0e-3;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Vs[i * 3 + j] = V[i * 3 + j] / (pl[i] * pl[j]);
  sym3_eigh(Vs, ev, U);
  const double lmax = fmax(ev[0], fmax(ev[1], ev[2]));
  const double tol = 2.220446049250313e-16 * 3 * lmax;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k)
        if (ev[k] > tol) s += U[i * 3 + k] * (1.0 / ev[k]) * U[j * 3 + k];
      Veff[i * 3 + j] = s / (pl[i] * pl[j]);
    }
  const int nblk = HAS_EXT ? 2 : 1;
  for (int k = 0; k < nobs; ++k) {
    const int o = ob + k * ost;
    const double r0 = rP[o], r1 = rP[S + o];
    double Jd[2][12], Jl[6];
    for (int q = 0; q < 12; ++q) Jd[0][q] = JpP[q * S + o];
    if (HAS_EXT)
      for (int q = 0; q < 12; ++q) Jd[1][q] = JeP[q * S + o];
    for (int q = 0; q < 6; ++q) Jl[q] = JlP[q * S + o];
    const int offs[2] = {b.pose_off[b.obs_pose[o]], HAS_EXT ? b.pose_off[b.obs_ext[o]] : -1};
    for (int a = 0; a < nblk; ++a) {
      if (offs[a] < 0) continue;
      for (int i = 0; i < 6; ++i) atomicAdd(&m.b[offs[a] + i], -(Jd[a][i] * r0 + Jd[a][6 + i] * r1));
      for (int c = 0; c < nblk; ++c) {
        if (offs[c] < 0) continue;
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j)
            atomicAdd(&m.H[(size_t)(offs[a] + i) * n + offs[c] + j], Jd[a][i] * Jd[c][j] + Jd[a][6 + i] * Jd[c][6 + j]);
      }
      // W_e Veff
      double W[18] = {0}, WV[18];
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 3; ++j) W[i * 3 + j] = Jd[a][i] * Jl[j] + Jd[a][6 + i] * Jl[3 + j];
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 3; ++j)
          WV[i * 3 + j] = W[i * 3] * Veff[j] + W[i * 3 + 1] * Veff[3 + j] + W[i * 3 + 2] * Veff[6 + j];
      for (int i = 0; i < 6; ++i)
        atomicAdd(&m.b[offs[a] + i], -(WV[i * 3] * bl[0] + WV[i * 3 + 1] * bl[1] + WV[i * 3 + 2] * bl[2]));
      for (int k2 = 0; k2 < nobs; ++k2) {
        const int o2 = ob + k2 * ost;
        double Jl2[6];
        for (int q = 0; q < 6; ++q) Jl2[q] = JlP[q * S + o2];
        const int offs2[2] = {b.pose_off[b.obs_pose[o2]], HAS_EXT ? b.pose_off[b.obs_ext[o2]] : -1};
        for (int a2 = 0; a2 < nblk; ++a2) {
          if (offs2[a2] < 0) continue;
          double J2[12];
          const double* src = (a2 == 0) ? JpP : JeP;
          for (int q = 0; q < 12; ++q) J2[q] = src[q * S + o2];
          for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
              double v = 0;
              for (int c = 0; c < 3; ++c) v += WV[i * 3 + c] * (J2[j] * Jl2[c] + J2[6 + j] * Jl2[3 + c]);
              atomicAdd(&m.H[(size_t)(offs[a] + i) * n + offs2[a2] + j], -v);
            }
        }
      }
    }
  }
}

// ---- symmetric eigen-decomposition on a thread-block CLUSTER: one-sided (Hestenes) Jacobi, one warp per column pair ----
// A (n x n, symmetric; row-major == column-major) is overwritten by G = A V with mutually orthogonal columns, V (the
// eigenvectors, column k contiguous) is accumulated in U and transposed at the end into the layout the callers use
// (U[i * n + k] = component i of eigenvector k); ev[k] = v_k . g_k = v_k^T A v_k keeps the sign of tiny negative
// eigenvalues.  A rotation touches two contiguous columns only (three warp-reduced dot products, then 2 x 2n updates); the
// n/2 pairs of a round-robin round are independent and go to the 4 x 32 warps of a 4-CTA cluster (n ~ 180: one pair per
// warp), one cluster barrier per round; the columns live in global memory (L2), read and written with .cg so that no SM's
// L1 holds a stale copy.  History of the B9 call of bench.py's 10-KF window (195-dim): two-sided Jacobi, one 256-thread
// CTA, matrices in global memory, three barriers + a serial reshuffle per round 135 ms (r1) -> one-sided, one 1024-thread
// CTA 8.0 ms (r2p) -> this kernel.
constexpr int kJacobiCluster = 4;
__global__ void __cluster_dims__(kJacobiCluster, 1, 1) __launch_bounds__(1024) k_jacobi_cluster(double* A, int n, double* U,
                                                                                                double* ev, double* conv) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int T = kJacobiCluster * 1024, gtid = crank * 1024 + tid;
  const int gw = crank * 32 + wid, nw = kJacobiCluster * 32;
  const int mm_ = n + (n & 1);   // players (one dummy when n is odd)
  const int half = mm_ / 2;
  __shared__ double worst_s[32];
  for (int e = gtid; e < n * n; e += T) __stcg(&U[e], (e / n == e % n) ? 1.0 : 0.0);
  cluster.sync();
  for (int sweep = 0; sweep < 40; ++sweep) {
    double worst = 0.0;   // largest |cos angle| between two columns seen by this warp in the sweep
    for (int round = 0; round < mm_ - 1; ++round) {
      for (int k = gw; k < half; k += nw) {
        // circle method: player mm_-1 stays, the others rotate
        int p = (k == 0) ? mm_ - 1 : (round + k) % (mm_ - 1);
        int q = (round + mm_ - 1 - k) % (mm_ - 1);
        if (p >= n || q >= n) continue;   // the dummy
        if (p > q) { const int t = p; p = q; q = t; }
        double* gp = A + (size_t)p * n;
        double* gq = A + (size_t)q * n;
        double al = 0.0, be = 0.0, ga = 0.0;
        for (int i = lane; i < n; i += 32) {
          const double x = __ldcg(&gp[i]), y = __ldcg(&gq[i]);
          al += x * x;
          be += y * y;
          ga += x * y;
        }
        al = warp_sum(al);
        be = warp_sum(be);
        ga = warp_sum(ga);
        if (ga == 0.0 || al == 0.0 || be == 0.0) continue;
        const double cosang = fabs(ga) / sqrt(al * be);
        worst = fmax(worst, cosang);
        if (cosang < 1.0e-15) continue;
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
        double* vp = U + (size_t)p * n;
        double* vq = U + (size_t)q * n;
        for (int i = lane; i < n; i += 32) {
          const double x = __ldcg(&gp[i]), y = __ldcg(&gq[i]);
          __stcg(&gp[i], c * x - sn * y);
          __stcg(&gq[i], sn * x + c * y);
          const double u = __ldcg(&vp[i]), v = __ldcg(&vq[i]);
          __stcg(&vp[i], c * u - sn * v);
          __stcg(&vq[i], sn * u + c * v);
        }
      }
      cluster.sync();
    }
    if (lane == 0) worst_s[wid] = worst;
    __syncthreads();
    if (tid == 0) {
      double w = 0.0;
      for (int i = 0; i < 32; ++i) w = fmax(w, worst_s[i]);
      __stcg(&conv[crank], w);
    }
    cluster.sync();
    double w = 0.0;
    for (int r = 0; r < kJacobiCluster; ++r) w = fmax(w, __ldcg(&conv[r]));
    cluster.sync();   // everyone has read conv before the next sweep may overwrite it
    if (w < 1.0e-14) break;
  }
  for (int k = gw; k < n; k += nw) {
    double sdot = 0.0;
    for (int i = lane; i < n; i += 32) sdot += __ldcg(&U[(size_t)k * n + i]) * __ldcg(&A[(size_t)k * n + i]);
    sdot = warp_sum(sdot);
    if (lane == 0) ev[k] = sdot;
  }
  cluster.sync();
  // V columns -> U[i][k]
  for (int e = gtid; e < n * n; e += T) {
    const int i = e / n, k = e % n;
    if (i < k) {
      const double x = __ldcg(&U[(size_t)i * n + k]), y = __ldcg(&U[(size_t)k * n + i]);
      __stcg(&U[(size_t)i * n + k], y);
      __stcg(&U[(size_t)k * n + i], x);
    }
  }
}

// stage 0: H += prior + Jd^T Jd, b likewise; scaled marginalised block -> A.   [k_jacobi_cluster on nm]
// stage 1: pseudo-inverse, dense Schur complement -> Hk, bk; scaled Hk -> A.    [k_jacobi_cluster on nk]
// stage 2: J = (p o U) S^1/2, e0 (MarginalizationError.cpp:725-758)
__global__ void __launch_bounds__(1024) k_marg_dense(Batch b, MargArgs m, int stage) {
  const WinDesc& wd = b.win[m.w];
  const int buf = b.ws[m.w].cur;
  const int tid = threadIdx.x, T = blockDim.x;
  const int n = m.n, nk = m.nk, nm = m.nm, M = wd.n_rows;
  const double* Jd = b.Jd[buf] + wd.Jd_off;
  const double* rd = b.rd[buf] + wd.rd_off;
  __shared__ double lmax_s;
  if (stage == 0) {
    // 1. H += prior + Jd^T Jd ; b += prior_b - Jd^T rd
    for (int e = tid; e < n * n; e += T) {
      const int i = e / n, j = e % n;
      double s = m.H[e];
      for (int r = 0; r < M; ++r) s += Jd[(size_t)r * n + i] * Jd[(size_t)r * n + j];
      m.H[e] = s;
    }
    for (int i = tid; i < n; i += T) {
      double s = m.b[i];
      for (int r = 0; r < M; ++r) s -= Jd[(size_t)r * n + i] * rd[r];
      m.b[i] = s;
    }
    __syncthreads();
    for (int e = tid; e < m.prior_dim * m.prior_dim; e += T) {
      const int i = e / m.prior_dim, j = e % m.prior_dim;
      m.H[(size_t)m.prior_map[i] * n + m.prior_map[j]] += m.prior_H[e];  // distinct (i,j) -> distinct entries
    }
    for (int i = tid; i < m.prior_dim; i += T) m.b[m.prior_map[i]] += m.prior_b[i];
    __syncthreads();
    if (nm > 0) {
      for (int i = tid; i < n; i += T) m.pvec[i] = m.H[(size_t)i * n + i] > 1.0e-9 ? sqrt(m.H[(size_t)i * n + i]) : 1.0e-3;
      __syncthreads();
      for (int e = tid; e < nm * nm; e += T) {
        const int i = e / nm, j = e % nm;
        const int gi = m.marg_idx[i], gj = m.marg_idx[j];
        m.A[e] = 0.5 * (m.H[(size_t)gi * n + gj] + m.H[(size_t)gj * n + gi]) / (m.pvec[gi] * m.pvec[gj]);
      }
    }
    return;
  }
  if (stage == 1) {
    // 2. dense Schur
    if (nm > 0) {
      if (tid == 0) {
        double lm = -1e300;
        for (int i = 0; i < nm; ++i) lm = fmax(lm, m.ev[i]);
        lmax_s = lm;
      }
      __syncthreads();
      const double tol = 2.220446049250313e-16 * nm * lmax_s;
      for (int e = tid; e < nm * nm; e += T) {
        const int i = e / nm, j = e % nm;
        double s = 0;
        for (int k = 0; k < nm; ++k)
          if (m.ev[k] > tol) s += m.U[(size_t)i * nm + k] * (1.0 / m.ev[k]) * m.U[(size_t)j * nm + k];
        m.Vp[e] = s;
      }
      for (int e = tid; e < nk * nm; e += T) {
        const int i = e / nm, j = e % nm;
        m.Wm[e] = m.H[(size_t)m.keep_idx[i] * n + m.marg_idx[j]] / (m.pvec[m.keep_idx[i]] * m.pvec[m.marg_idx[j]]);
      }
      __syncthreads();
      for (int e = tid; e < nk * nm; e += T) {
        const int i = e / nm, j = e % nm;
        double s = 0;
        for (int k = 0; k < nm; ++k) s += m.Wm[(size_t)i * nm + k] * m.Vp[(size_t)k * nm + j];
        m.WV[e] = s;
      }
      __syncthreads();
      for (int i = tid; i < nk; i += T) {
        const int gi = m.keep_idx[i];
        double s = m.b[gi] / m.pvec[gi];
        for (int j = 0; j < nm; ++j) s -= m.WV[(size_t)i * nm + j] * (m.b[m.marg_idx[j]] / m.pvec[m.marg_idx[j]]);
        m.bk[i] = s * m.pvec[gi];
      }
      for (int e = tid; e < nk * nk; e += T) {
        const int i = e / nk, j = e % nk;
        const int gi = m.keep_idx[i], gj = m.keep_idx[j];
        double h = m.H[(size_t)gi * n + gj] / (m.pvec[gi] * m.pvec[gj]);
        for (int k = 0; k < nm; ++k) h -= m.WV[(size_t)i * nm + k] * m.Wm[(size_t)j * nm + k];
        m.Hk[e] = h * m.pvec[gi] * m.pvec[gj];
      }
    } else {
      for (int i = tid; i < nk; i += T) m.bk[i] = m.b[m.keep_idx[i]];
      for (int e = tid; e < nk * nk; e += T) m.Hk[e] = m.H[(size_t)m.keep_idx[e / nk] * n + m.keep_idx[e % nk]];
    }
    __syncthreads();
    // 3. updateErrorComputation: scaled Hk -> A
    for (int i = tid; i < nk; i += T) m.pvec[i] = m.Hk[(size_t)i * nk + i] > 1.0e-9 ? sqrt(m.Hk[(size_t)i * nk + i]) : 1.0e-3;
    __syncthreads();
    for (int e = tid; e < nk * nk; e += T) {
      const int i = e / nk, j = e % nk;
      m.A[e] = 0.5 * (m.Hk[(size_t)i * nk + j] + m.Hk[(size_t)j * nk + i]) / (m.pvec[i] * m.pvec[j]);
    }
    return;
  }
  if (tid == 0) {
    double lm = -1e300;
    for (int i = 0; i < nk; ++i) lm = fmax(lm, m.ev[i]);
    lmax_s = lm;
  }
  __syncthreads();
  const double tol2 = 2.220446049250313e-16 * nk * lmax_s;
  for (int e = tid; e < nk * nk; e += T) {
    const int k = e / nk, i = e % nk;
    const double S = m.ev[k] > tol2 ? m.ev[k] : 0.0;
    m.J[e] = m.pvec[i] * m.U[(size_t)i * nk + k] * sqrt(S);
  }
  for (int k = tid; k < nk; k += T) {
    const double Sp = m.ev[k] > tol2 ? 1.0 / m.ev[k] : 0.0;
    const double sps = sqrt(Sp);
    double e = 0;
    for (int i = 0; i < nk; ++i) e += sps * m.U[(size_t)i * nk + k] * (1.0 / m.pvec[i]) * m.bk[i];
    m.e0[k] = -e;
  }
}

void launch_marg(const Batch& b, const MargArgs& m, int num_landmarks, cudaStream_t st) {
  if (num_landmarks > 0) {
    const int grid = (num_landmarks + 127) / 128;
    if (b.has_ext)
      k_marg_landmarks<true><<<grid, 128, 0, st>>>(b, m);
    else
      k_marg_landmarks<false><<<grid, 128, 0, st>>>(b, m);
  }
  double* conv = m.pvec + m.n;   // kJacobiCluster doubles of scratch behind pvec (see the arena in svin_ba_marginalize)
  k_marg_dense<<<1, 1024, 0, st>>>(b, m, 0);
  if (m.nm > 0) k_jacobi_cluster<<<kJacobiCluster, 1024, 0, st>>>(m.A, m.nm, m.U, m.ev, conv);
  k_marg_dense<<<1, 1024, 0, st>>>(b, m, 1);
  k_jacobi_cluster<<<kJacobiCluster, 1024, 0, st>>>(m.A, m.nk, m.U, m.ev, conv);
  k_marg_dense<<<1, 1024, 0, st>>>(b, m, 2);
}

}  // namespace svin
