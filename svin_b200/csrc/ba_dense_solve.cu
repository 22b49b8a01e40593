// Reduced camera system, fast path: one CTA per window with the whole system in shared memory.
//
//   A   = H~ (landmark-eliminated pose blocks, from k_schur) + Jd^T Jd   (dense-term rows)
//   solve (S A S + mu * diag^2) y = S g~   with Jacobi scaling S and the dogleg's LM diagonal
//
// which is what Ceres' SchurComplementSolver hands to its Cholesky after SchurEliminator::Eliminate
// (the per-iteration linear solve inside Map::solve(), okvis_ceres/include/okvis/ceres/Map.hpp:347).
// Storage is packed lower-triangular with the right-hand side carried as an extra row, so the
// factorisation performs the forward substitution for free.  Threads are laid out 16x16 so the
// trailing update needs no integer division.
#include <cuda_runtime.h>
#include <math.h>

#include <cstdlib>

#include "ba_device_utils.cuh"
#include "ba_kernels.cuh"

namespace svin {

namespace {
constexpr int kTS = 16;              // thread layout kTS x kTS
constexpr int kNT = 7;               // accumulators per thread per dimension in one panel
constexpr int kPanel = kTS * kNT;    // 112 columns per panel

__device__ __forceinline__ int tri(int i, int j) { return (i * (i + 1) >> 1) + j; }  // j <= i
}  // namespace

// G = Jd^T Jd (lower triangle of a row-major n x n block at H_off) and gg = Jd^T rd of the dense-term rows of one
// state buffer.  Runs right after k_dense_eval - for the candidate on the side stream, hidden behind k_linearize -
// so that k_dense_solve_smem only adds it to the landmark-eliminated blocks.
constexpr int kGR = 16;  // rows staged per tile
size_t dense_gram_smem_bytes(int n) { return sizeof(double) * ((size_t)kGR * n + kGR); }

__global__ void __launch_bounds__(kTS* kTS, 2) k_dense_gram(Batch b, int which) {
  extern __shared__ double smem[];
  const int w = blockIdx.x;
  WinState& ws = b.ws[w];
  if (ws.done) return;
  if (which == 1 && (ws.skip_slot || ws.gn_failed || !(-ws.acc_mc > 0.0))) return;  // as k_dense_eval
  // which == 2: after k_decide, for the (possibly new) current buffer; a rejected step keeps the old system
  if (which == 2 && ws.reuse) return;
  const WinDesc& wd = b.win[w];
  const int buf = (which == 1) ? 1 - ws.cur : ws.cur;
  const int n = wd.n_dense, M = wd.n_rows;
  const int tid = threadIdx.x, tx = tid & (kTS - 1), ty = tid >> 4;
  double* Js = smem;
  double* rds = Js + (size_t)kGR * n;
  const double* Jd = b.Jd[buf] + wd.Jd_off;
  const double* rd = b.rd[buf] + wd.rd_off;
  double* G = b.gram[buf] + wd.H_off;
  double* gg = b.gram_g[buf] + wd.d_off;
  double gr_acc = 0.0;
  const int n_panels = (n + kPanel - 1) / kPanel;
  for (int ip = 0; ip < n_panels; ++ip)
    for (int jp = 0; jp <= ip; ++jp) {
      double acc[kNT][kNT];
#pragma unroll
      for (int x = 0; x < kNT; ++x)
#pragma unroll
        for (int y = 0; y < kNT; ++y) acc[x][y] = 0.0;
      const bool first = (ip == 0 && jp == 0);
      for (int r0 = 0; r0 < M; r0 += kGR) {
        const int rows = min(kGR, M - r0);
        __syncthreads();
        for (int e = tid; e < rows * n; e += kTS * kTS) Js[e] = Jd[(size_t)r0 * n + e];
        if (tid < rows) rds[tid] = rd[r0 + tid];
        __syncthreads();
        for (int r = 0; r < rows; ++r) {
          const double* row = Js + r * n;
          double a[kNT], c[kNT];
#pragma unroll
          for (int x = 0; x < kNT; ++x) {
            const int i = ip * kPanel + ty + kTS * x;
            a[x] = (i < n) ? row[i] : 0.0;
            const int j = jp * kPanel + tx + kTS * x;
            c[x] = (j < n) ? row[j] : 0.0;
          }
#pragma unroll
          for (int x = 0; x < kNT; ++x)
#pragma unroll
            for (int y = 0; y < kNT; ++y) acc[x][y] += a[x] * c[y];
          if (first && tid < n) gr_acc += row[tid] * rds[r];
        }
      }
#pragma unroll
      for (int x = 0; x < kNT; ++x)
#pragma unroll
        for (int y = 0; y < kNT; ++y) {
          const int i = ip * kPanel + ty + kTS * x, j = jp * kPanel + tx + kTS * y;
          if (i < n && j <= i) G[(size_t)i * n + j] = acc[x][y];
        }
    }
  if (tid < n) gg[tid] = gr_acc;
}
// Tensor-core version (fp64 mma.sync.m8n8k4, SASS DMMA): G = P^T P with P = [Jd | rd] staged kGP rows at a time;
// each warp owns every 8th 8x8 tile pair of the lower triangle and keeps it in registers over all row panels, so the
// Gram matrix costs ~1.5 k warp instructions per warp instead of the ~22 k of the register-tiled FMA version
// (ncu r1t: 89 us -> see profiles/).  The extra column rd yields gg = Jd^T rd in the same product.
constexpr int kGP = 32;         // rows per staged panel (8 k-steps)
constexpr int kGPairTable = 288; // tile pairs of the lower triangle: T (T + 1) / 2 <= 288  ->  n + 1 <= 184
__host__ __device__ inline int gram_ld(int n) {  // >= n + 1, = 4 (mod 16): conflict-free fragment loads
  int ld = n + 1;
  while ((ld & 15) != 4) ++ld;
  return ld;
}
size_t dense_gram_mma_smem_bytes(int n) {
  return sizeof(double) * (size_t)kGP * gram_ld(n) + sizeof(int) * kGPairTable;
}
static int gram_pairs(int n) {
  const int T = (n + 1 + 7) / 8;
  return T * (T + 1) / 2;
}

__device__ __forceinline__ void dmma_884(double& d0, double& d1, double a, double bq) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(bq));
}

// NP = tile pairs per warp (8 warps)
template <int NP>
__global__ void __launch_bounds__(256, NP <= 16 ? 2 : 1) k_dense_gram_mma(Batch b, int which) {
  extern __shared__ double smem[];
  const int w = blockIdx.x;
  WinState& ws = b.ws[w];
  if (ws.done) return;
  if (which == 1 && (ws.skip_slot || ws.gn_failed || !(-ws.acc_mc > 0.0))) return;  // as k_dense_eval
  // which == 2: after k_decide, for the (possibly new) current buffer; a rejected step keeps the old system
  if (which == 2 && ws.reuse) return;
  const WinDesc& wd = b.win[w];
  const int buf = (which == 1) ? 1 - ws.cur : ws.cur;
  const int n = wd.n_dense, M = wd.n_rows;
  const int ld = gram_ld(n);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const double* Jd = b.Jd[buf] + wd.Jd_off;
  const double* rd = b.rd[buf] + wd.rd_off;
  double* G = b.gram[buf] + wd.H_off;
  double* gg = b.gram_g[buf] + wd.d_off;
  const int T = (n + 1 + 7) >> 3;
  const int npairs = T * (T + 1) / 2;
  // pair p -> (tm >= tn), row-major over the lower triangle of the T x T tile grid
  int* pair_tab = reinterpret_cast<int*>(smem + (size_t)kGP * ld);
  for (int p = tid; p < npairs; p += 256) {
    int tm = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
    while (tm * (tm + 1) / 2 > p) --tm;
    while ((tm + 1) * (tm + 2) / 2 <= p) ++tm;
    pair_tab[p] = (tm << 8) | (p - tm * (tm + 1) / 2);
  }
  double acc[NP][2];
#pragma unroll
  for (int q = 0; q < NP; ++q) acc[q][0] = acc[q][1] = 0.0;
  for (int r0 = 0; r0 < M; r0 += kGP) {
    const int rows = min(kGP, M - r0);
    __syncthreads();
    for (int e = tid; e < kGP * ld; e += 256) {
      const int r = e / ld, c = e - r * ld;
      double v = 0.0;
      if (r < rows) {
        if (c < n)
          v = Jd[(size_t)(r0 + r) * n + c];
        else if (c == n)
          v = rd[r0 + r];
      }
      smem[e] = v;
    }
    __syncthreads();
    const int ksteps = (rows + 3) >> 2;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      if (wid + 8 * q < npairs) {  // warp-uniform
        const int tt = pair_tab[wid + 8 * q];
        const double* pa = smem + t * ld + 8 * (tt >> 8) + g;
        const double* pb = smem + t * ld + 8 * (tt & 255) + g;
        for (int ks = 0; ks < ksteps; ++ks) dmma_884(acc[q][0], acc[q][1], pa[4 * ks * ld], pb[4 * ks * ld]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    if (wid + 8 * q < npairs) {
      const int tt = pair_tab[wid + 8 * q];
      const int i = 8 * (tt >> 8) + g;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 8 * (tt & 255) + 2 * t + e;
        const double v = acc[q][e];
        if (i == n && j < n)
          gg[j] = v;
        else if (i < n && j <= i)
          G[(size_t)i * n + j] = v;
      }
    }
  }
}
cudaError_t configure_dense_gram(int smem_bytes) {
  return cudaFuncSetAttribute(k_dense_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
}
void launch_dense_gram(const Batch& b, int which, int n_max, cudaStream_t st) {
  static const bool no_mma = std::getenv("SVIN_GRAM_FMA") != nullptr;  // A/B knob
  const int pairs = gram_pairs(n_max);
  const size_t sm = dense_gram_mma_smem_bytes(n_max);
  if (!no_mma && pairs <= 8 * 16 && sm <= 48 * 1024)
    k_dense_gram_mma<16><<<b.B, 256, sm, st>>>(b, which);
  else if (!no_mma && pairs <= kGPairTable && sm <= 48 * 1024)
    k_dense_gram_mma<36><<<b.B, 256, sm, st>>>(b, which);
  else
    k_dense_gram<<<b.B, kTS * kTS, dense_gram_smem_bytes(n_max), st>>>(b, which);
}

size_t dense_solve_smem_bytes(int n) {
  const size_t a = (size_t)(n + 1) * (n + 2) / 2;
  return sizeof(double) * (a + 8 * (size_t)n + 16);
}

__global__ void __launch_bounds__(kTS* kTS, 2) k_dense_solve_smem(Batch b, SvinBaOptions opt) {
  extern __shared__ double smem[];
  const int w = blockIdx.x;
  WinState& ws = b.ws[w];
  if (ws.done || ws.reuse) return;
  const WinDesc& wd = b.win[w];
  const int n = wd.n_dense, buf = ws.cur;
  const int tid = threadIdx.x, tx = tid & (kTS - 1), ty = tid >> 4;
  const int lane = tid & 31, wid = tid >> 5;
  double* A = smem;
  double* v_hd = A + (size_t)(n + 1) * (n + 2) / 2;  // column square norms
  double* v_graw = v_hd + n;    // unreduced gradient
  double* v_gred = v_graw + n;  // reduced rhs
  double* v_sc = v_gred + n;    // Jacobi scale
  double* v_dg = v_sc + n;      // dogleg diagonal
  double* v_c = v_dg + n;       // scale * gradient_ / diag (Cauchy direction, unscaled space)
  double* v_tmp = v_c + n;
  double* colk = v_tmp + n;     // current Cholesky column, dense copy (n + 1 entries)
  const double* Hg = b.H + wd.H_off;
  __shared__ unsigned long long gmax_s;

  // ---- 1. A(lower) = H~ + Jd^T Jd; the Gram matrix of the dense rows comes from k_dense_gram
  const double* G = b.gram[buf] + wd.H_off;
  for (int i = ty; i < n; i += kTS) {
    const double* Gi = G + (size_t)i * n;
    double* Ai = A + tri(i, 0);
    for (int j = tx; j <= i; j += kTS) Ai[j] = Gi[j] + Hg[(size_t)j * n + i];  // H~ keeps the upper triangle
  }
  const double hd_acc = (tid < n) ? G[(size_t)tid * n + tid] : 0.0;
  const double gr_acc = (tid < n) ? b.gram_g[buf][wd.d_off + tid] : 0.0;
  // ---- 2. vectors
  double gmax_l = 0.0;
  if (tid < n) {
    const double hd = b.Hdiag[wd.d_off + tid] + hd_acc;
    const double gr = b.g_raw[wd.d_off + tid] + gr_acc;
    v_hd[tid] = hd;
    v_graw[tid] = gr;
    v_gred[tid] = b.g_red[wd.d_off + tid] + gr_acc;
    gmax_l = fabs(gr);
  }

  atomic_max_nonneg(&ws.gmax_bits, gmax_l);
  __syncthreads();
  if (tid == 0) gmax_s = atomicMax(&ws.gmax_bits, 0ull);  // atomic read (L1 may hold a stale WinState line)
  __syncthreads();
  // ---- 3. gradient tolerance (checked after the max-iteration test of the previous slot)
  if (ws.last_successful && __longlong_as_double((long long)gmax_s) <= opt.gradient_tolerance) {
    if (tid == 0) {
      ws.done = 1;
      ws.termination = SVIN_TERM_CONVERGENCE;
    }
    return;
  }
  // ---- 4. Jacobi scaling (first Jacobian only), dogleg diagonal, scaled gradient
  const bool first_jac = (ws.iter == 0 && ws.num_successful == 0 && !ws.invalid);
  const double mu = ws.mu;
  if (tid < n) {
    double s;
    if (first_jac) {
      s = opt.jacobi_scaling ? 1.0 / (1.0 + sqrt(v_hd[tid])) : 1.0;
      b.scale_d[wd.d_off + tid] = s;
    } else {
      s = b.scale_d[wd.d_off + tid];
    }
    const double d = sqrt(fmin(fmax(v_hd[tid] * s * s, opt.min_lm_diagonal), opt.max_lm_diagonal));
    const double g = s * v_graw[tid] / d;
    v_sc[tid] = s;
    v_dg[tid] = d;
    v_tmp[tid] = g;  // gradient_
    b.diag_d[wd.d_off + tid] = d;
    b.grad_d[wd.d_off + tid] = g;
  }
  __syncthreads();
  // ---- 5. scaled system + LM diagonal; rhs row n
  for (int i = ty; i <= n; i += kTS) {
    if (i < n) {
      const double si = v_sc[i];
      for (int j = tx; j <= i; j += kTS) {
        double v = A[tri(i, j)] * si * v_sc[j];
        if (i == j) v += mu * v_dg[i] * v_dg[i];
        A[tri(i, j)] = v;
      }
    } else {
      for (int j = tx; j < n; j += kTS) A[tri(n, j)] = v_sc[j] * v_gred[j];
    }
  }
  // ---- 6. right-looking Cholesky, rhs row included (forward substitution)
  bool fail = false;
  for (int k = 0; k < n; ++k) {
    __syncthreads();
    const double akk = A[tri(k, k)];
    if (!(akk > 0.0) || !isfinite(akk)) {
      fail = true;
      break;
    }
    const double d = sqrt(akk);
    for (int i = k + 1 + tid; i <= n; i += kTS * kTS) {
      const double v = A[tri(i, k)] / d;
      A[tri(i, k)] = v;
      colk[i] = v;  // unit-stride copy: the trailing update below needs no triangular index arithmetic
    }
    __syncthreads();
    if (tid == 0) A[tri(k, k)] = d;
    for (int i = k + 1 + ty; i <= n; i += kTS) {
      const double aik = colk[i];
      const int jmax = min(i, n - 1);
      double* Ai = A + tri(i, 0);
      for (int j = k + 1 + tx; j <= jmax; j += kTS) Ai[j] -= aik * colk[j];
    }
  }
  __syncthreads();
  if (fail) {
    if (tid == 0) {
      // DoglegStrategy::ComputeGaussNewtonStep: mu *= 10 and retry while mu < max_mu (1.0)
      ws.mu *= 10.0;
      if (ws.mu < 1.0)
        ws.skip_slot = 1;  // re-eliminate next slot, no iteration consumed
      else
        ws.gn_failed = 1;  // linear solver FAILURE -> invalid step
    }
    return;
  }
  // ---- 7. backward solve L^T y = z (row n) by warp 0
  double* z = A + tri(n, 0);
  if (wid == 0) {
    for (int k = n - 1; k >= 0; --k) {
      const double yk = z[k] / A[tri(k, k)];
      __syncwarp();
      if (lane == 0) z[k] = yk;
      const double* Lk = A + tri(k, 0);
      for (int i = lane; i < k; i += 32) z[i] -= Lk[i] * yk;
      __syncwarp();
    }
  }
  __syncthreads();
  // ---- 8. outputs
  bool bad = false;
  double g2 = 0, n2 = 0, gd = 0;
  if (tid < n) {
    const double y = z[tid];
    if (!isfinite(y)) bad = true;
    const double g = v_tmp[tid];
    const double gni = -y * v_dg[tid];
    b.gn_d[wd.d_off + tid] = gni;
    b.u_d[wd.d_off + tid] = v_sc[tid] * y;
    const double cv = v_sc[tid] * g / v_dg[tid];
    b.c_d[wd.d_off + tid] = cv;
    v_c[tid] = cv;
    g2 = g * g;
    n2 = gni * gni;
    gd = g * gni;
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
  // Cauchy point over the dense rows: sum_r (Jd[r,:] . c)^2 = c^T (Jd^T Jd) c with the Gram matrix (lower triangle)
  __syncthreads();
  double jg2 = 0.0;
  for (int i = wid; i < n; i += kTS * kTS / 32) {
    const double* Gi = G + (size_t)i * n;
    double sdot = 0.0;
    for (int j = lane; j < i; j += 32) sdot += Gi[j] * v_c[j];
    sdot = warp_sum(sdot);
    if (lane == 0) jg2 += v_c[i] * (2.0 * sdot + Gi[i] * v_c[i]);
  }
  double v[4] = {g2, n2, gd, jg2};
  double* const dst[4] = {&ws.acc_g2, &ws.acc_n2, &ws.acc_gdot, &ws.acc_Jg2};
  block_atomic_add<4, kTS * kTS>(v, dst);
}

// Register-resident variant for n + 1 <= 16 NB: thread (ty, tx) of the 16 x 16 layout owns the elements (i, j) with
// i = ty, j = tx (mod 16) of the lower triangle (NB (NB + 1) / 2 blocks -> as many registers), the right-hand side is
// row n.  Per column only the scaled column travels through shared memory (one vector + the next pivot), the rank-1
// update is pure register arithmetic: 2 barriers and <= NB (NB + 1) / 2 DFMA per thread per column instead of
// 4 shared-memory accesses per updated element (ncu r1t: the update loop was 54 % of the smem kernel's instructions).
template <int NB, int KB>
__device__ __forceinline__ void chol_column(double (&a)[NB * (NB + 1) / 2], int k, int n, int tx, int ty, double akk,
                                            double* colk, double* akk_s) {
  constexpr int DI = KB * (KB + 1) / 2 + KB;  // block (KB, KB)
  const int kt = k & 15;
  // owners of column k (one half-warp: tx is the slow thread index) scale it and publish it; one reciprocal per
  // column instead of a division per element (1 ulp from Eigen's A21 /= x, far inside the 1e-6 parity gate)
  if (tx == kt) {
    const double d = sqrt(akk);
    const double inv = 1.0 / d;
#pragma unroll
    for (int x = KB; x < NB; ++x) {
      const int i = 16 * x + ty;
      double& e = a[x * (x + 1) / 2 + KB];
      if (i > k && i <= n) {
        e = e * inv;
        colk[i] = e;
      } else if (i == k) {
        e = d;
      }
    }
  }
  __syncthreads();
  double ri[NB], cj[NB];
#pragma unroll
  for (int x = KB; x < NB; ++x) {
    const int i = 16 * x + ty, j = 16 * x + tx;
    ri[x] = (i > k && i <= n) ? colk[i] : 0.0;
    cj[x] = (j > k && j < n) ? colk[j] : 0.0;
  }
#pragma unroll
  for (int x = KB; x < NB; ++x)
#pragma unroll
    for (int y = KB; y <= x; ++y) a[x * (x + 1) / 2 + y] -= ri[x] * cj[y];
  // the next pivot
  const int nt = (k + 1) & 15;
  if (ty == nt && tx == nt) {
    if (kt == 15) {
      if (KB + 1 < NB) *akk_s = a[(KB + 1 < NB ? KB + 1 : KB) * ((KB + 1 < NB ? KB + 1 : KB) + 1) / 2 + (KB + 1 < NB ? KB + 1 : KB)];
    } else {
      *akk_s = a[DI];
    }
  }
  __syncthreads();
}

template <int NB>
__global__ void __launch_bounds__(kTS* kTS, 2) k_dense_solve_reg(Batch b, SvinBaOptions opt) {
  extern __shared__ double smem[];
  const int w = blockIdx.x;
  WinState& ws = b.ws[w];
  if (ws.done || ws.reuse) return;
  const WinDesc& wd = b.win[w];
  const int n = wd.n_dense, buf = ws.cur;
  const int tid = threadIdx.x, ty = tid & (kTS - 1), tx = tid >> 4;  // a column's owners share a half-warp
  const int lane = tid & 31, wid = tid >> 5;
  double* A = smem;
  double* v_hd = A + (size_t)(n + 1) * (n + 2) / 2;  // column square norms
  double* v_graw = v_hd + n;    // unreduced gradient
  double* v_gred = v_graw + n;  // reduced rhs
  double* v_sc = v_gred + n;    // Jacobi scale
  double* v_dg = v_sc + n;      // dogleg diagonal
  double* v_c = v_dg + n;       // scale * gradient_ / diag (Cauchy direction, unscaled space)
  double* v_tmp = v_c + n;
  double* colk = v_tmp + n;     // current Cholesky column (n + 1 entries)
  __shared__ double akk_s;
  const double* Hg = b.H + wd.H_off;
  __shared__ unsigned long long gmax_s;

  // ---- 1. A(lower) = H~ + Jd^T Jd; the Gram matrix of the dense rows comes from k_dense_gram
  const double* G = b.gram[buf] + wd.H_off;
  double a[NB * (NB + 1) / 2];
#pragma unroll
  for (int x = 0; x < NB; ++x)
#pragma unroll
    for (int y = 0; y <= x; ++y) {
      const int i = 16 * x + ty, j = 16 * y + tx;
      a[x * (x + 1) / 2 + y] = (i < n && j <= i) ? G[(size_t)i * n + j] + Hg[(size_t)j * n + i] : 0.0;  // H~: upper
    }
  const double hd_acc = (tid < n) ? G[(size_t)tid * n + tid] : 0.0;
  const double gr_acc = (tid < n) ? b.gram_g[buf][wd.d_off + tid] : 0.0;
  // ---- 2. vectors
  double gmax_l = 0.0;
  if (tid < n) {
    const double hd = b.Hdiag[wd.d_off + tid] + hd_acc;
    const double gr = b.g_raw[wd.d_off + tid] + gr_acc;
    v_hd[tid] = hd;
    v_graw[tid] = gr;
    v_gred[tid] = b.g_red[wd.d_off + tid] + gr_acc;
    gmax_l = fabs(gr);
  }

  atomic_max_nonneg(&ws.gmax_bits, gmax_l);
  __syncthreads();
  if (tid == 0) gmax_s = atomicMax(&ws.gmax_bits, 0ull);  // atomic read (L1 may hold a stale WinState line)
  __syncthreads();
  // ---- 3. gradient tolerance (checked after the max-iteration test of the previous slot)
  if (ws.last_successful && __longlong_as_double((long long)gmax_s) <= opt.gradient_tolerance) {
    if (tid == 0) {
      ws.done = 1;
      ws.termination = SVIN_TERM_CONVERGENCE;
    }
    return;
  }
  // ---- 4. Jacobi scaling (first Jacobian only), dogleg diagonal, scaled gradient
  const bool first_jac = (ws.iter == 0 && ws.num_successful == 0 && !ws.invalid);
  const double mu = ws.mu;
  if (tid < n) {
    double s;
    if (first_jac) {
      s = opt.jacobi_scaling ? 1.0 / (1.0 + sqrt(v_hd[tid])) : 1.0;
      b.scale_d[wd.d_off + tid] = s;
    } else {
      s = b.scale_d[wd.d_off + tid];
    }
    const double d = sqrt(fmin(fmax(v_hd[tid] * s * s, opt.min_lm_diagonal), opt.max_lm_diagonal));
    const double g = s * v_graw[tid] / d;
    v_sc[tid] = s;
    v_dg[tid] = d;
    v_tmp[tid] = g;  // gradient_
    b.diag_d[wd.d_off + tid] = d;
    b.grad_d[wd.d_off + tid] = g;
  }
  __syncthreads();
  // ---- 5. scaled system + LM diagonal; rhs row n
#pragma unroll
  for (int x = 0; x < NB; ++x)
#pragma unroll
    for (int y = 0; y <= x; ++y) {
      const int i = 16 * x + ty, j = 16 * y + tx;
      double& e = a[x * (x + 1) / 2 + y];
      if (i < n && j <= i) {
        double v = e * v_sc[i] * v_sc[j];
        if (i == j) v += mu * v_dg[i] * v_dg[i];
        e = v;
      } else if (i == n && j < n) {
        e = v_sc[j] * v_gred[j];
      }
    }
  // ---- 6. right-looking Cholesky in registers, rhs row included (forward substitution)
  if (tid == 0) akk_s = a[0];
  __syncthreads();
  bool fail = false;
  for (int k = 0; k < n; ++k) {
    const double akk = akk_s;
    if (!(akk > 0.0) || !isfinite(akk)) {
      fail = true;
      break;
    }
    switch (k >> 4) {
      case 0: chol_column<NB, 0>(a, k, n, tx, ty, akk, colk, &akk_s); break;
      case 1: if (NB > 1) chol_column<NB, (NB > 1 ? 1 : 0)>(a, k, n, tx, ty, akk, colk, &akk_s); break;
      case 2: if (NB > 2) chol_column<NB, (NB > 2 ? 2 : 0)>(a, k, n, tx, ty, akk, colk, &akk_s); break;
      case 3: if (NB > 3) chol_column<NB, (NB > 3 ? 3 : 0)>(a, k, n, tx, ty, akk, colk, &akk_s); break;
      case 4: if (NB > 4) chol_column<NB, (NB > 4 ? 4 : 0)>(a, k, n, tx, ty, akk, colk, &akk_s); break;
      case 5: if (NB > 5) chol_column<NB, (NB > 5 ? 5 : 0)>(a, k, n, tx, ty, akk, colk, &akk_s); break;
      default: if (NB > 6) chol_column<NB, (NB > 6 ? 6 : 0)>(a, k, n, tx, ty, akk, colk, &akk_s); break;
    }
  }
  // the factor (and the forward-substituted rhs row) go to the packed layout the back-substitution reads
  if (!fail) {
#pragma unroll
    for (int x = 0; x < NB; ++x)
#pragma unroll
      for (int y = 0; y <= x; ++y) {
        const int i = 16 * x + ty, j = 16 * y + tx;
        if (i <= n && j <= i && j < n) A[tri(i, j)] = a[x * (x + 1) / 2 + y];
      }
  }
  __syncthreads();
  if (fail) {
    if (tid == 0) {
      // DoglegStrategy::ComputeGaussNewtonStep: mu *= 10 and retry while mu < max_mu (1.0)
      ws.mu *= 10.0;
      if (ws.mu < 1.0)
        ws.skip_slot = 1;  // re-eliminate next slot, no iteration consumed
      else
        ws.gn_failed = 1;  // linear solver FAILURE -> invalid step
    }
    return;
  }
  // ---- 7. backward solve L^T y = z (row n) by warp 0
  double* z = A + tri(n, 0);
  if (wid == 0) {
    for (int k = n - 1; k >= 0; --k) {
      const double yk = z[k] / A[tri(k, k)];
      __syncwarp();
      if (lane == 0) z[k] = yk;
      const double* Lk = A + tri(k, 0);
      for (int i = lane; i < k; i += 32) z[i] -= Lk[i] * yk;
      __syncwarp();
    }
  }
  __syncthreads();
  // ---- 8. outputs
  bool bad = false;
  double g2 = 0, n2 = 0, gd = 0;
  if (tid < n) {
    const double y = z[tid];
    if (!isfinite(y)) bad = true;
    const double g = v_tmp[tid];
    const double gni = -y * v_dg[tid];
    b.gn_d[wd.d_off + tid] = gni;
    b.u_d[wd.d_off + tid] = v_sc[tid] * y;
    const double cv = v_sc[tid] * g / v_dg[tid];
    b.c_d[wd.d_off + tid] = cv;
    v_c[tid] = cv;
    g2 = g * g;
    n2 = gni * gni;
    gd = g * gni;
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
  // Cauchy point over the dense rows: sum_r (Jd[r,:] . c)^2 = c^T (Jd^T Jd) c with the Gram matrix (lower triangle)
  __syncthreads();
  double jg2 = 0.0;
  for (int i = wid; i < n; i += kTS * kTS / 32) {
    const double* Gi = G + (size_t)i * n;
    double sdot = 0.0;
    for (int j = lane; j < i; j += 32) sdot += Gi[j] * v_c[j];
    sdot = warp_sum(sdot);
    if (lane == 0) jg2 += v_c[i] * (2.0 * sdot + Gi[i] * v_c[i]);
  }
  double v[4] = {g2, n2, gd, jg2};
  double* const dst[4] = {&ws.acc_g2, &ws.acc_n2, &ws.acc_gdot, &ws.acc_Jg2};
  block_atomic_add<4, kTS * kTS>(v, dst);
}

cudaError_t configure_dense_solve(int smem_bytes) {
  cudaError_t e = cudaFuncSetAttribute(k_dense_solve_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_dense_solve_reg<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_dense_solve_reg<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  return e;
}
void launch_dense_solve_generic(const Batch& b, const SvinBaOptions& opt, cudaStream_t st);
void launch_dense_solve(const Batch& b, const SvinBaOptions& opt, int smem_bytes, int n_max, cudaStream_t st) {
  static const bool no_reg = std::getenv("SVIN_SOLVE_SMEM") != nullptr;  // A/B knob
  if (smem_bytes > 0 && !no_reg && n_max + 1 <= 16 * 5)
    k_dense_solve_reg<5><<<b.B, kTS * kTS, smem_bytes, st>>>(b, opt);
  else if (smem_bytes > 0 && !no_reg && n_max + 1 <= 16 * 7)
    k_dense_solve_reg<7><<<b.B, kTS * kTS, smem_bytes, st>>>(b, opt);
  else if (smem_bytes > 0)
    k_dense_solve_smem<<<b.B, kTS * kTS, smem_bytes, st>>>(b, opt);
  else
    launch_dense_solve_generic(b, opt, st);
}

}  // namespace svin
