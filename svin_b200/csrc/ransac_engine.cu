// RANSAC row: batched-hypothesis absolute-pose (3D-2D), relative-pose and rotation-only (2D-2D)
// consensus on the device, behind svin_ransac_* (include/svin_b200.h).
//
// Reference seams: Frontend::runRansac3d2d (okvis_frontend/src/Frontend.cpp:617-676), runRansac2d2d (:832-980) and
// SVIn's runRansac2d2dToRefineScale (:680-830).  The consensus scores are the in-tree ones
// (FrameAbsolutePoseSacProblem.hpp:131-160, FrameRelativePoseSacProblem.hpp:121-157,
// FrameRotationOnlySacProblem.hpp:116-137); the RANSAC loop and the minimal solvers live in OpenGV (un-vendored) and are
// restated from their published form - see the header comment of svin_ransac_absolute for the declared differences.
//
// Mapping: every hypothesis of every problem of a call is independent.
//   k_hypotheses  one thread per hypothesis      minimal solver (P3P quartic / 8-point / 2-point rotation) -> 3x4 model
//   k_consensus   one CTA (4 warps) per hypothesis, lanes over the correspondences: score < threshold, ballot + popc
//   k_select      one CTA per problem: thread 0 replays sac::Ransac::computeModel's sequential bookkeeping over the
//                 per-hypothesis counts (adaptive iteration bound k, skipped models), then all threads write the
//                 winning model's inlier mask
// The sequential reference evaluates hypotheses one after the other and stops early; evaluating all of them at once and
// replaying the stop rule gives the same winner.
#include <cuda_runtime.h>
#include <math.h>

#include <cstring>
#include <string>
#include <vector>

#include "common.hpp"

using namespace svin;

namespace {

enum { KIND_ABS = 0, KIND_REL = 1, KIND_ROT = 2 };

struct RProb {
  int kind, n, ns, sample_size;
  int corr_off;   // first correspondence in the concatenated arrays
  int samp_off;   // first sample index
  int hyp_off;    // first hypothesis
  int cam_off;    // first camera (absolute pose)
  int num_cams, max_iterations;
  double threshold;
};
struct RData {
  const RProb* prob;
  const double *A3, *B3;       // [N][3] world points / bearings 1 ; bearings / bearings 2
  const double *s1, *s2;       // sigma angles
  const int* cam;              // camera index (absolute pose)
  const double *camR, *camT;   // [C][9], [C][3]
  const int* samples;
  const int* hyp_prob;         // [H] problem of each hypothesis
  double* models;              // [H][12] row-major [R | t]
  int *valid, *counts;         // [H]
  int *best, *ninl, *iters;    // [P]
  unsigned char* mask;         // [N]
};

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ bool normalize3(double* a) {
  const double n = sqrt(dot3(a, a));
  if (!(n > 1e-14)) return false;
  a[0] /= n; a[1] /= n; a[2] /= n;
  return true;
}
// orthonormal triad with e1 = (b - a) / |b - a|, e3 normal of the triangle (a, b, c): columns of F (row-major 3x3)
__device__ bool triangle_frame(const double* a, const double* b, const double* c, double* F) {
  double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
  if (!normalize3(e1)) return false;
  const double d[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
  double e3[3], e2[3];
  cross3(e1, d, e3);
  if (!normalize3(e3)) return false;
  cross3(e3, e1, e2);
  for (int i = 0; i < 3; ++i) {
    F[i * 3 + 0] = e1[i];
    F[i * 3 + 1] = e2[i];
    F[i * 3 + 2] = e3[i];
  }
  return true;
}
// C = A B^T (3x3 row-major)
__device__ __forceinline__ void mul_abt(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j * 3] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
}
__device__ __forceinline__ void mul_ab(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

// ---- scores (in-tree) -------------------------------------------------------------------------------------
// FrameAbsolutePoseSacProblem.hpp:131-160; model = body pose in the world
__device__ double score_abs(const double* M, const double* p, const double* f, const double* Rc, const double* tc, double sigma) {
  const double d[3] = {p[0] - M[3], p[1] - M[7], p[2] - M[11]};
  double body[3], rep[3];
  for (int j = 0; j < 3; ++j) body[j] = M[j] * d[0] + M[4 + j] * d[1] + M[8 + j] * d[2] - tc[j];   // R^T (p - t) - t_c
  for (int j = 0; j < 3; ++j) rep[j] = Rc[j] * body[0] + Rc[3 + j] * body[1] + Rc[6 + j] * body[2];  // R_c^T (.)
  const double n = sqrt(dot3(rep, rep));
  const double e[3] = {rep[0] / n - f[0], rep[1] / n - f[1], rep[2] / n - f[2]};
  return dot3(e, e) / sigma;
}
// opengv::triangulation::triangulate2 (mid-point between the two rays), point in frame 1
__device__ void triangulate2(const double* M, const double* f1, const double* f2, double* p) {
  const double t[3] = {M[3], M[7], M[11]};
  double f2u[3];
  for (int i = 0; i < 3; ++i) f2u[i] = M[i * 4] * f2[0] + M[i * 4 + 1] * f2[1] + M[i * 4 + 2] * f2[2];
  const double b0 = dot3(t, f1), b1 = dot3(t, f2u);
  const double a00 = dot3(f1, f1), a10 = dot3(f1, f2u), a01 = -a10, a11 = -dot3(f2u, f2u);
  const double det = a00 * a11 - a01 * a10;
  const double l0 = (a11 * b0 - a01 * b1) / det, l1 = (-a10 * b0 + a00 * b1) / det;
  for (int i = 0; i < 3; ++i) p[i] = 0.5 * (l0 * f1[i] + t[i] + l1 * f2u[i]);
}
// FrameRelativePoseSacProblem.hpp:121-157
__device__ double score_rel(const double* M, const double* f1, const double* f2, double s1, double s2) {
  double p[3];
  triangulate2(M, f1, f2, p);
  const double n1 = sqrt(dot3(p, p));
  const double d[3] = {p[0] - M[3], p[1] - M[7], p[2] - M[11]};
  double q[3];
  for (int j = 0; j < 3; ++j) q[j] = M[j] * d[0] + M[4 + j] * d[1] + M[8 + j] * d[2];   // R12^T (p - t12)
  const double n2 = sqrt(dot3(q, q));
  const double e1[3] = {p[0] / n1 - f1[0], p[1] / n1 - f1[1], p[2] / n1 - f1[2]};
  const double e2[3] = {q[0] / n2 - f2[0], q[1] / n2 - f2[1], q[2] / n2 - f2[2]};
  return dot3(e1, e1) * 0.5 / s1 + dot3(e2, e2) * 0.5 / s2;
}
// FrameRotationOnlySacProblem.hpp:116-137
__device__ double score_rot(const double* M, const double* f1, const double* f2, double s1, double s2) {
  double e1[3], e2[3];
  for (int i = 0; i < 3; ++i) e1[i] = M[i * 4] * f2[0] + M[i * 4 + 1] * f2[1] + M[i * 4 + 2] * f2[2] - f1[i];
  for (int j = 0; j < 3; ++j) e2[j] = M[j] * f1[0] + M[4 + j] * f1[1] + M[8 + j] * f1[2] - f2[j];
  return dot3(e1, e1) * 0.5 / s1 + dot3(e2, e2) * 0.5 / s2;
}

// ---- quartic roots: Aberth-Ehrlich iteration in complex arithmetic (fixed count, no data-dependent branching on the
// convergence path), then a Newton polish of the nearly-real roots on the real polynomial ---------------------------
__device__ int quartic_real_roots(const double* c /*c[0] x^4 + ... + c[4]*/, double* roots) {
  if (fabs(c[0]) < 1e-300) return 0;
  const double a[4] = {c[1] / c[0], c[2] / c[0], c[3] / c[0], c[4] / c[0]};   // monic
  const double rad = 1.0 + fmax(fmax(fabs(a[0]), fabs(a[1])), fmax(fabs(a[2]), fabs(a[3])));
  double zr[4], zi[4];
  for (int k = 0; k < 4; ++k) {
    const double ang = 0.4 + 1.5707963267948966 * k;
    zr[k] = 0.5 * rad * cos(ang);
    zi[k] = 0.5 * rad * sin(ang);
  }
  for (int it = 0; it < 60; ++it) {
    for (int k = 0; k < 4; ++k) {
      // p(z), p'(z) by Horner
      double pr = 1.0, pi = 0.0, dr = 0.0, di = 0.0;
      for (int j = 0; j < 4; ++j) {
        const double ndr = dr * zr[k] - di * zi[k] + pr, ndi = dr * zi[k] + di * zr[k] + pi;
        dr = ndr; di = ndi;
        const double npr = pr * zr[k] - pi * zi[k] + a[j], npi = pr * zi[k] + pi * zr[k];
        pr = npr; pi = npi;
      }
      const double dd = dr * dr + di * di;
      if (dd == 0.0) continue;
      double wr = (pr * dr + pi * di) / dd, wi = (pi * dr - pr * di) / dd;   // p / p'
      double sr = 0.0, si = 0.0;
      for (int j = 0; j < 4; ++j) {
        if (j == k) continue;
        const double xr = zr[k] - zr[j], xi = zi[k] - zi[j];
        const double xx = xr * xr + xi * xi;
        if (xx == 0.0) continue;
        sr += xr / xx;
        si -= xi / xx;
      }
      // z -= w / (1 - w s)
      const double er = 1.0 - (wr * sr - wi * si), ei = -(wr * si + wi * sr);
      const double ee = er * er + ei * ei;
      if (ee == 0.0) continue;
      zr[k] -= (wr * er + wi * ei) / ee;
      zi[k] -= (wi * er - wr * ei) / ee;
    }
  }
  int n = 0;
  for (int k = 0; k < 4; ++k) {
    if (fabs(zi[k]) > 1e-9 * fmax(1.0, fabs(zr[k]))) continue;
    double x = zr[k];
    for (int it = 0; it < 3; ++it) {
      const double p = (((x + a[0]) * x + a[1]) * x + a[2]) * x + a[3];
      const double d = ((4.0 * x + 3.0 * a[0]) * x + 2.0 * a[1]) * x + a[2];
      if (d == 0.0) break;
      x -= p / d;
    }
    roots[n++] = x;
  }
  return n;
}
__device__ __forceinline__ void polymul(const double* a, int na, const double* b, int nb, double* c) {  // degrees na, nb
  for (int i = 0; i <= na + nb; ++i) c[i] = 0.0;
  for (int i = 0; i <= na; ++i)
    for (int j = 0; j <= nb; ++j) c[i + j] += a[i] * b[j];
}

// Central P3P in camera `c` (Grunert's quartic), the sample's 4th point picks among the <= 4 solutions.
__device__ bool solve_abs(const RData& d, const RProb& P, const int* smp, double* M) {
  const int i0 = P.corr_off + smp[0], i1 = P.corr_off + smp[1], i2 = P.corr_off + smp[2], i3 = P.corr_off + smp[3];
  const int c = d.cam[i0];
  if (d.cam[i1] != c || d.cam[i2] != c) return false;
  if (smp[0] == smp[1] || smp[0] == smp[2] || smp[1] == smp[2] || smp[3] == smp[0] || smp[3] == smp[1] || smp[3] == smp[2])
    return false;
  const double* X0 = d.A3 + 3 * (size_t)i0; const double* X1 = d.A3 + 3 * (size_t)i1; const double* X2 = d.A3 + 3 * (size_t)i2;
  const double* f0 = d.B3 + 3 * (size_t)i0; const double* f1 = d.B3 + 3 * (size_t)i1; const double* f2 = d.B3 + 3 * (size_t)i2;
  auto dist = [](const double* a, const double* b) {
    const double e[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    return sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
  };
  const double a = dist(X1, X2), b = dist(X0, X2), cc = dist(X0, X1);
  if (fmin(a, fmin(b, cc)) < 1e-12) return false;
  const double ca = dot3(f1, f2), cb = dot3(f0, f2), cg = dot3(f0, f1);
  const double K = (a * a - cc * cc) / (b * b), c2b2 = (cc * cc) / (b * b);
  // s2 = u s1, s3 = v s1, u = N(v) / D(v); highest power first
  const double N[3] = {K - 1.0, -2.0 * K * cb, 1.0 + K};
  const double D[2] = {-2.0 * ca, 2.0 * cg};
  const double Q[3] = {1.0, -2.0 * cb, 1.0};
  double D2[3], N2[5], ND[4], QD2[5], Pq[5];
  polymul(D, 1, D, 1, D2);
  polymul(N, 2, N, 2, N2);
  polymul(N, 2, D, 1, ND);
  polymul(Q, 2, D2, 2, QD2);
  for (int i = 0; i < 5; ++i) Pq[i] = N2[i] - c2b2 * QD2[i];
  for (int i = 0; i < 4; ++i) Pq[i + 1] += -2.0 * cg * ND[i];
  for (int i = 0; i < 3; ++i) Pq[i + 2] += D2[i];
  double roots[4];
  const int nr = quartic_real_roots(Pq, roots);
  const double* Rc = d.camR + 9 * (size_t)(P.cam_off + c);
  const double* tc = d.camT + 3 * (size_t)(P.cam_off + c);
  const double* X3 = d.A3 + 3 * (size_t)i3; const double* f3 = d.B3 + 3 * (size_t)i3;
  const int c3 = d.cam[i3];
  const double* Rc3 = d.camR + 9 * (size_t)(P.cam_off + c3);
  const double* tc3 = d.camT + 3 * (size_t)(P.cam_off + c3);
  const double sg3 = d.s1[i3];
  double Fw[9];
  if (!triangle_frame(X0, X1, X2, Fw)) return false;
  bool found = false;
  double best = 1e300;
  for (int k = 0; k < nr; ++k) {
    const double v = roots[k];
    if (!(v > 0.0)) continue;
    const double den = D[0] * v + D[1];
    if (fabs(den) < 1e-14) continue;
    const double u = ((N[0] * v + N[1]) * v + N[2]) / den;
    const double q = 1.0 + v * v - 2.0 * v * cb;
    if (!(u > 0.0) || !(q > 0.0)) continue;
    const double s1 = b / sqrt(q), s2 = u * s1, s3 = v * s1;
    const double Y0[3] = {s1 * f0[0], s1 * f0[1], s1 * f0[2]};
    const double Y1[3] = {s2 * f1[0], s2 * f1[1], s2 * f1[2]};
    const double Y2[3] = {s3 * f2[0], s3 * f2[1], s3 * f2[2]};
    double Fc[9], Rcw[9], tcw[3], Rbw[9], tbw[3], Mc[12];
    if (!triangle_frame(Y0, Y1, Y2, Fc)) continue;
    mul_abt(Fc, Fw, Rcw);                                         // camera-from-world rotation
    for (int i = 0; i < 3; ++i) tcw[i] = Y0[i] - (Rcw[i * 3] * X0[0] + Rcw[i * 3 + 1] * X0[1] + Rcw[i * 3 + 2] * X0[2]);
    mul_ab(Rc, Rcw, Rbw);                                         // body-from-world
    for (int i = 0; i < 3; ++i) tbw[i] = Rc[i * 3] * tcw[0] + Rc[i * 3 + 1] * tcw[1] + Rc[i * 3 + 2] * tcw[2] + tc[i];
    for (int i = 0; i < 3; ++i) {                                 // world-from-body: R = Rbw^T, t = -Rbw^T tbw
      for (int j = 0; j < 3; ++j) Mc[i * 4 + j] = Rbw[j * 3 + i];
      Mc[i * 4 + 3] = -(Rbw[i] * tbw[0] + Rbw[3 + i] * tbw[1] + Rbw[6 + i] * tbw[2]);
    }
    const double sc = score_abs(Mc, X3, f3, Rc3, tc3, sg3);
    if (sc < best) {
      best = sc;
      found = true;
      for (int i = 0; i < 12; ++i) M[i] = Mc[i];
    }
  }
  return found;
}

// f1 = R12 f2 from two bearing pairs (orthonormal triads)
__device__ bool solve_rot(const RData& d, const RProb& P, const int* smp, double* M) {
  if (smp[0] == smp[1]) return false;
  const double* a1 = d.A3 + 3 * (size_t)(P.corr_off + smp[0]); const double* b1 = d.A3 + 3 * (size_t)(P.corr_off + smp[1]);
  const double* a2 = d.B3 + 3 * (size_t)(P.corr_off + smp[0]); const double* b2 = d.B3 + 3 * (size_t)(P.corr_off + smp[1]);
  auto triad = [](const double* x, const double* y, double* F) {
    double e3[3], e2[3];
    cross3(x, y, e3);
    const double n = sqrt(dot3(e3, e3));
    if (n < 1e-12) return false;
    e3[0] /= n; e3[1] /= n; e3[2] /= n;
    cross3(e3, x, e2);
    for (int i = 0; i < 3; ++i) {
      F[i * 3] = x[i];
      F[i * 3 + 1] = e2[i];
      F[i * 3 + 2] = e3[i];
    }
    return true;
  };
  double B1[9], B2[9], R[9];
  if (!triad(a1, b1, B1) || !triad(a2, b2, B2)) return false;
  mul_abt(B1, B2, R);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) M[i * 4 + j] = R[i * 3 + j];
    M[i * 4 + 3] = 0.0;
  }
  return true;
}

// cyclic Jacobi on a symmetric N x N matrix (row-major, destroyed), eigenvectors in the columns of V
template <int N>
__device__ void jacobi_sym(double* A, double* V) {
  for (int i = 0; i < N * N; ++i) V[i] = (i / N == i % N) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) (i == j ? dg : off) += A[i * N + j] * A[i * N + j];
    if (off <= 1e-30 * dg || off == 0.0) break;
    for (int p = 0; p < N - 1; ++p)
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p * N + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * N + q] - A[p * N + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < N; ++k) {
          const double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - s * akq;
          A[k * N + q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; ++k) {
          const double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - s * aqk;
          A[q * N + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < N; ++k) {
          const double vkp = V[k * N + p], vkq = V[k * N + q];
          V[k * N + p] = c * vkp - s * vkq;
          V[k * N + q] = s * vkp + c * vkq;
        }
      }
  }
}

// Eight-point algorithm (f1^T E f2 = 0) on the 8 samples, essential-matrix projection, the decomposition with the most
// sample points in front of both cameras.
__device__ bool solve_rel(const RData& d, const RProb& P, const int* smp, double* M) {
  for (int i = 0; i < 8; ++i)
    for (int j = i + 1; j < 8; ++j)
      if (smp[i] == smp[j]) return false;
  double G[81], V9[81];
  for (int i = 0; i < 81; ++i) G[i] = 0.0;
  for (int s = 0; s < 8; ++s) {
    const double* a = d.A3 + 3 * (size_t)(P.corr_off + smp[s]);
    const double* b = d.B3 + 3 * (size_t)(P.corr_off + smp[s]);
    double row[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) row[i * 3 + j] = a[i] * b[j];
    for (int i = 0; i < 9; ++i)
      for (int j = 0; j < 9; ++j) G[i * 9 + j] += row[i] * row[j];
  }
  jacobi_sym<9>(G, V9);
  int kmin = 0;
  for (int k = 1; k < 9; ++k)
    if (G[k * 9 + k] < G[kmin * 9 + kmin]) kmin = k;
  double E[9];
  for (int i = 0; i < 9; ++i) E[i] = V9[i * 9 + kmin];
  // SVD of E through the eigen-decomposition of E^T E
  double T3[9], V3[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T3[i * 3 + j] = E[i] * E[j] + E[3 + i] * E[3 + j] + E[6 + i] * E[6 + j];
  jacobi_sym<3>(T3, V3);
  int o0 = 0, o1 = 1, o2 = 2;   // descending eigenvalues
  if (T3[o0 * 4] < T3[o1 * 4]) { const int t = o0; o0 = o1; o1 = t; }
  if (T3[o1 * 4] < T3[o2 * 4]) { const int t = o1; o1 = o2; o2 = t; }
  if (T3[o0 * 4] < T3[o1 * 4]) { const int t = o0; o0 = o1; o1 = t; }
  double v0[3] = {V3[o0], V3[3 + o0], V3[6 + o0]}, v1[3] = {V3[o1], V3[3 + o1], V3[6 + o1]}, v2[3];
  cross3(v0, v1, v2);                                              // det V = +1
  double u0[3], u1[3], u2[3];
  for (int i = 0; i < 3; ++i) {
    u0[i] = E[i * 3] * v0[0] + E[i * 3 + 1] * v0[1] + E[i * 3 + 2] * v0[2];
    u1[i] = E[i * 3] * v1[0] + E[i * 3 + 1] * v1[1] + E[i * 3 + 2] * v1[2];
  }
  if (!normalize3(u0)) return false;
  // Gram-Schmidt keeps U orthonormal when the two leading singular values are close
  const double pr = dot3(u0, u1);
  for (int i = 0; i < 3; ++i) u1[i] -= pr * u0[i];
  if (!normalize3(u1)) return false;
  cross3(u0, u1, u2);                                              // det U = +1
  const double U[9] = {u0[0], u1[0], u2[0], u0[1], u1[1], u2[1], u0[2], u1[2], u2[2]};
  const double Vt[9] = {v0[0], v0[1], v0[2], v1[0], v1[1], v1[2], v2[0], v2[1], v2[2]};
  const double W[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1}, Wt[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
  int best_n = -1;
  for (int r = 0; r < 2; ++r) {
    double UW[9], R[9];
    mul_ab(U, r == 0 ? W : Wt, UW);
    mul_ab(UW, Vt, R);
    for (int sgn = 0; sgn < 2; ++sgn) {
      double Mc[12];
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Mc[i * 4 + j] = R[i * 3 + j];
        Mc[i * 4 + 3] = sgn == 0 ? u2[i] : -u2[i];
      }
      int n = 0;
      for (int s = 0; s < 8; ++s) {
        const double* a = d.A3 + 3 * (size_t)(P.corr_off + smp[s]);
        const double* b = d.B3 + 3 * (size_t)(P.corr_off + smp[s]);
        double p[3];
        triangulate2(Mc, a, b, p);
        const double dd[3] = {p[0] - Mc[3], p[1] - Mc[7], p[2] - Mc[11]};
        double q[3];
        for (int j = 0; j < 3; ++j) q[j] = Mc[j] * dd[0] + Mc[4 + j] * dd[1] + Mc[8 + j] * dd[2];
        if (dot3(p, a) > 0.0 && dot3(q, b) > 0.0) ++n;
      }
      if (n > best_n) {
        best_n = n;
        for (int i = 0; i < 12; ++i) M[i] = Mc[i];
      }
    }
  }
  return best_n >= 0;
}

__global__ void __launch_bounds__(64) k_hypotheses(RData d, int H) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const RProb& P = d.prob[d.hyp_prob[h]];
  const int* smp = d.samples + P.samp_off + (size_t)(h - P.hyp_off) * P.sample_size;
  for (int k = 0; k < P.sample_size; ++k)
    if (smp[k] < 0 || smp[k] >= P.n) {
      d.valid[h] = 0;
      return;
    }
  double M[12];
  bool ok;
  if (P.kind == KIND_ABS) ok = solve_abs(d, P, smp, M);
  else if (P.kind == KIND_REL) ok = solve_rel(d, P, smp, M);
  else ok = solve_rot(d, P, smp, M);
  for (int i = 0; i < 12; ++i) ok = ok && isfinite(M[i]);
  d.valid[h] = ok ? 1 : 0;
  if (ok)
    for (int i = 0; i < 12; ++i) d.models[12 * (size_t)h + i] = M[i];
}

__device__ __forceinline__ bool is_inlier(const RData& d, const RProb& P, const double* M, int i) {
  const int g = P.corr_off + i;
  double s;
  if (P.kind == KIND_ABS) {
    const int c = P.cam_off + d.cam[g];
    s = score_abs(M, d.A3 + 3 * (size_t)g, d.B3 + 3 * (size_t)g, d.camR + 9 * (size_t)c, d.camT + 3 * (size_t)c, d.s1[g]);
  } else if (P.kind == KIND_REL) {
    s = score_rel(M, d.A3 + 3 * (size_t)g, d.B3 + 3 * (size_t)g, d.s1[g], d.s2[g]);
  } else {
    s = score_rot(M, d.A3 + 3 * (size_t)g, d.B3 + 3 * (size_t)g, d.s1[g], d.s2[g]);
  }
  return s < P.threshold;   // countWithinDistance: NaN scores are not inliers
}

__global__ void __launch_bounds__(128) k_consensus(RData d) {
  const int h = blockIdx.x;
  if (!d.valid[h]) {
    if (threadIdx.x == 0) d.counts[h] = 0;
    return;
  }
  const RProb& P = d.prob[d.hyp_prob[h]];
  __shared__ double Ms[12];
  __shared__ int cnt;
  if (threadIdx.x < 12) Ms[threadIdx.x] = d.models[12 * (size_t)h + threadIdx.x];
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  int mine = 0;
  for (int i = threadIdx.x; i < P.n; i += blockDim.x) mine += is_inlier(d, P, Ms, i) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&cnt, mine);
  __syncthreads();
  if (threadIdx.x == 0) d.counts[h] = cnt;
}

__global__ void __launch_bounds__(128) k_select(RData d) {
  const int p = blockIdx.x;
  const RProb& P = d.prob[p];
  __shared__ int best_s;
  __shared__ double Ms[12];
  if (threadIdx.x == 0) {
    // opengv::sac::Ransac::computeModel, replayed over the pre-scored hypotheses
    int it = 0, skipped = 0, best = -2147483647, best_j = -1;
    double k = 1.0;
    const int max_skip = P.max_iterations * 10;
    for (int j = 0; j < P.ns; ++j) {
      if (!((double)it < k && skipped < max_skip)) break;
      const int h = P.hyp_off + j;
      if (!d.valid[h]) {
        ++skipped;
        continue;
      }
      if (d.counts[h] > best) {
        best = d.counts[h];
        best_j = j;
        const double w = (double)best / (double)P.n;
        double p_no = 1.0 - pow(w, (double)P.sample_size);
        p_no = fmin(fmax(2.220446049250313e-16, p_no), 1.0 - 2.220446049250313e-16);
        k = log(1.0 - 0.99) / log(p_no);
      }
      ++it;
      if (it > P.max_iterations) break;
    }
    best_s = best_j;
    d.best[p] = best_j;
    d.iters[p] = it;
    d.ninl[p] = best_j >= 0 ? d.counts[P.hyp_off + best_j] : 0;
  }
  __syncthreads();
  const int bj = best_s;
  if (bj >= 0 && threadIdx.x < 12) Ms[threadIdx.x] = d.models[12 * (size_t)(P.hyp_off + bj) + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.x; i < P.n; i += blockDim.x)
    d.mask[P.corr_off + i] = (bj >= 0 && is_inlier(d, P, Ms, i)) ? 1 : 0;
}

struct Arena {
  void* d = nullptr;
  void* h = nullptr;
  size_t cap = 0;
};

}  // namespace

struct svin_ransac_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  Arena in, out;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  double last_device_ms = 0.0;
  long long launches = 0;
};

namespace {

int ensure(Arena& a, size_t need) {
  if (need <= a.cap) return SVIN_OK;
  if (a.d) cudaFree(a.d);
  if (a.h) cudaFreeHost(a.h);
  a.d = a.h = nullptr;
  a.cap = 0;
  const size_t cap = need + need / 2 + 4096;
  SVIN_CUDA(cudaMalloc(&a.d, cap));
  SVIN_CUDA(cudaMallocHost(&a.h, cap));
  a.cap = cap;
  return SVIN_OK;
}

struct Layout {
  size_t bytes = 0;
  size_t add(size_t b) {
    const size_t o = bytes;
    bytes += (b + 255) & ~(size_t)255;
    return o;
  }
};

// One generic problem view used by both entry points.
struct HostProb {
  int kind, n, ns, sample_size, num_cams, max_iterations;
  double threshold;
  const double *A3, *B3, *s1, *s2, *camR, *camT;
  const int *cam, *samples;
};

int run(svin_ransac_ctx* c, const std::vector<HostProb>& hp, std::vector<SvinRansacResult*>& res) {
  SVIN_CUDA(cudaSetDevice(c->device));
  const int NP = (int)hp.size();
  size_t N = 0, H = 0, NS = 0, NC = 0;
  for (const HostProb& p : hp) {
    N += p.n;
    H += p.ns;
    NS += (size_t)p.ns * p.sample_size;
    NC += p.num_cams;
  }
  if (NP == 0) return SVIN_OK;
  Layout in;
  const size_t o_prob = in.add(sizeof(RProb) * NP), o_A = in.add(24 * N), o_B = in.add(24 * N), o_s1 = in.add(8 * N),
               o_s2 = in.add(8 * N), o_cam = in.add(4 * N), o_cR = in.add(72 * (NC + 1)), o_cT = in.add(24 * (NC + 1)),
               o_smp = in.add(4 * NS), o_hp = in.add(4 * H);
  Layout out;
  const size_t o_best = out.add(4 * NP), o_ninl = out.add(4 * NP), o_it = out.add(4 * NP), o_mask = out.add(N),
               o_models = out.add(96 * H), o_valid = out.add(4 * H), o_counts = out.add(4 * H);
  int rc;
  if ((rc = ensure(c->in, in.bytes)) != SVIN_OK) return rc;
  if ((rc = ensure(c->out, out.bytes)) != SVIN_OK) return rc;
  char* Hh = (char*)c->in.h;
  RProb* pr = (RProb*)(Hh + o_prob);
  size_t cn = 0, ch = 0, cs = 0, cc = 0;
  for (int k = 0; k < NP; ++k) {
    const HostProb& p = hp[k];
    pr[k] = RProb{p.kind, p.n, p.ns, p.sample_size, (int)cn, (int)cs, (int)ch, (int)cc, p.num_cams, p.max_iterations,
                  p.threshold};
    std::memcpy(Hh + o_A + 24 * cn, p.A3, 24 * (size_t)p.n);
    std::memcpy(Hh + o_B + 24 * cn, p.B3, 24 * (size_t)p.n);
    std::memcpy(Hh + o_s1 + 8 * cn, p.s1, 8 * (size_t)p.n);
    if (p.s2) std::memcpy(Hh + o_s2 + 8 * cn, p.s2, 8 * (size_t)p.n);
    if (p.cam) std::memcpy(Hh + o_cam + 4 * cn, p.cam, 4 * (size_t)p.n);
    if (p.num_cams) {
      std::memcpy(Hh + o_cR + 72 * cc, p.camR, 72 * (size_t)p.num_cams);
      std::memcpy(Hh + o_cT + 24 * cc, p.camT, 24 * (size_t)p.num_cams);
    }
    std::memcpy(Hh + o_smp + 4 * cs, p.samples, 4 * (size_t)p.ns * p.sample_size);
    int* hpi = (int*)(Hh + o_hp) + ch;
    for (int j = 0; j < p.ns; ++j) hpi[j] = k;
    cn += p.n;
    ch += p.ns;
    cs += (size_t)p.ns * p.sample_size;
    cc += p.num_cams;
  }
  char* D = (char*)c->in.d;
  char* O = (char*)c->out.d;
  SVIN_CUDA(cudaMemcpyAsync(D, Hh, in.bytes, cudaMemcpyHostToDevice, c->stream));
  RData d{};
  d.prob = (const RProb*)(D + o_prob);
  d.A3 = (const double*)(D + o_A); d.B3 = (const double*)(D + o_B);
  d.s1 = (const double*)(D + o_s1); d.s2 = (const double*)(D + o_s2);
  d.cam = (const int*)(D + o_cam);
  d.camR = (const double*)(D + o_cR); d.camT = (const double*)(D + o_cT);
  d.samples = (const int*)(D + o_smp);
  d.hyp_prob = (const int*)(D + o_hp);
  d.models = (double*)(O + o_models);
  d.valid = (int*)(O + o_valid); d.counts = (int*)(O + o_counts);
  d.best = (int*)(O + o_best); d.ninl = (int*)(O + o_ninl); d.iters = (int*)(O + o_it);
  d.mask = (unsigned char*)(O + o_mask);
  SVIN_CUDA(cudaEventRecord(c->ev[0], c->stream));
  if (H > 0) {
    k_hypotheses<<<(unsigned)((H + 63) / 64), 64, 0, c->stream>>>(d, (int)H);
    k_consensus<<<(unsigned)H, 128, 0, c->stream>>>(d);
  }
  k_select<<<NP, 128, 0, c->stream>>>(d);
  c->launches += 3;
  SVIN_CUDA(cudaEventRecord(c->ev[1], c->stream));
  SVIN_CUDA(cudaMemcpyAsync(c->out.h, O, out.bytes, cudaMemcpyDeviceToHost, c->stream));
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  SVIN_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->last_device_ms = ms;
  const char* R = (const char*)c->out.h;
  cn = 0;
  for (int k = 0; k < NP; ++k) {
    SvinRansacResult* r = res[k];
    const int bj = ((const int*)(R + o_best))[k];
    r->best_sample = bj;
    r->num_inliers = ((const int*)(R + o_ninl))[k];
    r->iterations = ((const int*)(R + o_it))[k];
    const double I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    std::memcpy(r->model, bj >= 0 ? (const double*)(R + o_models) + 12 * (size_t)(pr[k].hyp_off + bj) : I, 96);
    if (r->inliers) std::memcpy(r->inliers, R + o_mask + cn, (size_t)hp[k].n);
    if (r->hypothesis_inliers) std::memcpy(r->hypothesis_inliers, (const int*)(R + o_counts) + pr[k].hyp_off, 4 * (size_t)hp[k].ns);
    if (r->hypothesis_valid) {
      const int* v = (const int*)(R + o_valid) + pr[k].hyp_off;
      for (int j = 0; j < hp[k].ns; ++j) r->hypothesis_valid[j] = (uint8_t)v[j];
    }
    cn += hp[k].n;
  }
  return SVIN_OK;
}

}  // namespace

extern "C" {

int svin_ransac_create(int device, svin_ransac_ctx** out) {
  if (!out) {
    set_error("svin_ransac_create: out is NULL");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) {
    cudaGetLastError();
    set_error("svin_ransac_create: no CUDA device " + std::to_string(device) + " - this engine has no CPU fallback");
    return SVIN_ERR_NO_DEVICE;
  }
  SVIN_CUDA(cudaSetDevice(device));
  svin_ransac_ctx* c = new svin_ransac_ctx();
  c->device = device;
  SVIN_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  SVIN_CUDA(cudaEventCreate(&c->ev[0]));
  SVIN_CUDA(cudaEventCreate(&c->ev[1]));
  *out = c;
  return SVIN_OK;
}

void svin_ransac_destroy(svin_ransac_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  for (Arena* a : {&c->in, &c->out}) {
    if (a->d) cudaFree(a->d);
    if (a->h) cudaFreeHost(a->h);
  }
  if (c->ev[0]) cudaEventDestroy(c->ev[0]);
  if (c->ev[1]) cudaEventDestroy(c->ev[1]);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int svin_ransac_absolute(svin_ransac_ctx* c, int32_t num_problems, const SvinRansacAbsProblem* probs, SvinRansacResult* results) {
  if (!c || num_problems < 0 || (num_problems && (!probs || !results))) {
    set_error("svin_ransac_absolute: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  std::vector<HostProb> hp;
  std::vector<SvinRansacResult*> res;
  for (int k = 0; k < num_problems; ++k) {
    const SvinRansacAbsProblem& p = probs[k];
    if (p.num_correspondences < 0 || p.num_samples < 0 || p.num_cameras < 1 ||
        (p.num_correspondences && (!p.points || !p.bearings || !p.camera_index || !p.sigma_angle)) ||
        !p.camera_rotation || !p.camera_offset || (p.num_samples && !p.samples)) {
      set_error("svin_ransac_absolute: problem " + std::to_string(k) + " has NULL arrays or negative counts");
      return SVIN_ERR_INVALID_ARGUMENT;
    }
    for (int i = 0; i < p.num_correspondences; ++i)
      if (p.camera_index[i] < 0 || p.camera_index[i] >= p.num_cameras) {
        set_error("svin_ransac_absolute: camera_index out of range");
        return SVIN_ERR_INVALID_ARGUMENT;
      }
    hp.push_back(HostProb{KIND_ABS, p.num_correspondences, p.num_samples, 4, p.num_cameras, p.max_iterations, p.threshold,
                          p.points, p.bearings, p.sigma_angle, nullptr, p.camera_rotation, p.camera_offset, p.camera_index,
                          p.samples});
    res.push_back(&results[k]);
  }
  return run(c, hp, res);
}

int svin_ransac_relative(svin_ransac_ctx* c, int32_t num_problems, const SvinRansacRelProblem* probs,
                         SvinRansacResult* rotation_only, SvinRansacResult* relative_pose) {
  if (!c || num_problems < 0 || (num_problems && (!probs || !rotation_only || !relative_pose))) {
    set_error("svin_ransac_relative: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  std::vector<HostProb> hp;
  std::vector<SvinRansacResult*> res;
  for (int k = 0; k < num_problems; ++k) {
    const SvinRansacRelProblem& p = probs[k];
    if (p.num_correspondences < 0 || p.num_samples < 0 ||
        (p.num_correspondences && (!p.bearings1 || !p.bearings2 || !p.sigma_angle1 || !p.sigma_angle2)) ||
        (p.num_samples && (!p.samples_rotation || !p.samples_relative))) {
      set_error("svin_ransac_relative: problem " + std::to_string(k) + " has NULL arrays or negative counts");
      return SVIN_ERR_INVALID_ARGUMENT;
    }
    hp.push_back(HostProb{KIND_ROT, p.num_correspondences, p.num_samples, 2, 0, p.max_iterations, p.threshold, p.bearings1,
                          p.bearings2, p.sigma_angle1, p.sigma_angle2, nullptr, nullptr, nullptr, p.samples_rotation});
    res.push_back(&rotation_only[k]);
    hp.push_back(HostProb{KIND_REL, p.num_correspondences, p.num_samples, 8, 0, p.max_iterations, p.threshold, p.bearings1,
                          p.bearings2, p.sigma_angle1, p.sigma_angle2, nullptr, nullptr, nullptr, p.samples_relative});
    res.push_back(&relative_pose[k]);
  }
  return run(c, hp, res);
}

int svin_ransac_timings(svin_ransac_ctx* c, double* device_ms, int64_t* kernel_launches) {
  if (!c) {
    set_error("svin_ransac_timings: ctx is NULL");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  if (device_ms) *device_ms = c->last_device_ms;
  if (kernel_launches) *kernel_launches = c->launches;
  return SVIN_OK;
}

}  // extern "C"
