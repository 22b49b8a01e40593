// Device data model of a batch of sliding windows (see DESIGN.md "Data layout in HBM").
#pragma once
#include <stdint.h>

#include "../../include/svin_b200.h"

namespace svin {

constexpr int kObsTile = 128;  // observations per CTA (linearise)
constexpr int kLmTile = 128;   // landmarks per CTA (Schur / back-substitution / step)
constexpr int kDenseThreads = 256;
constexpr int kSchurClasses = 9;  // lane mappings of the Schur chunk kernels

struct ImuP {
  double sigma_g_c, sigma_a_c, sigma_gw_c, sigma_aw_c, g, g_max, a_max;
};

// Static description of one window inside the batch (all indices global into the batch arrays).
struct WinDesc {
  int pose_begin, pose_end;
  int sb_begin, sb_end;
  int lm_begin, lm_end;
  int obs_begin, obs_end;
  int cam_begin;     // first camera (intrinsics) of the window
  int n_dense;       // reduced-system dimension (6 per free pose block + 9 per free speed/bias block)
  int n_rows;        // rows of the dense-term Jacobian Jd
  int d_off;         // offset of this window in the concatenated dense vectors
  int rd_off;        // offset in the dense residual vectors
  long long H_off;   // offset in the reduced-Hessian storage (n_dense^2 doubles)
  long long Jd_off;  // offset in Jd storage (n_rows * n_dense doubles), per buffer
  int imu_begin, imu_end;
  int pp_begin, pp_end;  // pose priors
  int sp_begin, sp_end;  // speed/bias priors
  int rp_begin, rp_end;  // relative pose
  int so_begin, so_end;  // sonar
  int de_begin, de_end;  // depth
  int marg_blk_begin, marg_blk_end;
  int marg_dim, marg_row0;
  long long margJ_off;
  int marg_e0_off;
  int marg_lin_off;
  int loss_type;
  int info_uniform;  // every observation of the window carries the same 2x2 information: info3 below, nothing uploaded per observation
  double loss_scale;
  double info3[3];   // a00, a10, a11 of that information
  ImuP imu;
  double T_SSo[7];
};

// Dynamic trust-region state of one window.
struct WinState {
  int cur;  // which state/linearisation buffer holds x
  int done, termination;
  int iter, num_successful, invalid;
  int reuse, skip_slot, gn_failed, last_successful;
  int imu_redo;
  int pad0;
  double cost_x, cost_cand;
  double radius, mu;
  double alpha, dogleg_step_norm, cg, cn;
  double x_norm;
  double initial_cost;
  // accumulators (cleared by k_decide / k_init)
  double acc_g2, acc_n2, acc_gdot, acc_Jg2, acc_mc, acc_step2, acc_xnorm2;
  // reprojection rows of the model, split so that k_step_dense can evaluate them for any dogleg (cg, cn)
  // without another pass over the Jacobians: sum mg.r, sum an.r, sum |mg|^2, sum mg.an, sum |an|^2
  // (mg = J * Cauchy direction, an = J * (-Gauss-Newton step), both per observation)
  double acc_A[5];
  unsigned long long gmax_bits;
  unsigned long long t_start_ns, t_iter_start_ns, t_last_iter_ns;
};

struct ImuTerm {
  int pose0, sb0, pose1, sb1;  // global block indices
  int meas_begin, meas_end;
  int row0;
  int win;
  long long t0, t1;
};
// mutable pre-integration cache of an ImuError (ImuError.hpp:239-270), device resident
struct ImuCache {
  double Delta_q[4];
  double C_integral[9], C_doubleintegral[9];
  double acc_integral[3], acc_doubleintegral[3];
  double dalpha_db_g[9], dv_db_g[9], dp_db_g[9];
  double sb_ref[9];
  double sqrt_info[225];
  int redo;
  int redo_counter;
};
struct PosePrior {
  int block, row0, win, pad;
  double meas[7];
  double U[36];
};
struct SbPrior {
  int block, row0, win, pad;
  double meas[9];
  double U[81];
};
struct RelPose {
  int block0, block1, row0, win;
  double U[36];
};
struct SonarTerm {
  int pose, row0, win, pad;
  double range, heading, sqrt_info;
  double mean[3];
};
struct DepthTerm {
  int pose, row0, win, pad;
  double depth, first, sqrt_info;
};
struct MargBlock {
  int kind, index;  // global block index
  int col0;         // first column in the marginalisation prior (or -1 if the block is fixed)
  int lin_off;      // offset of its linearisation point
};

// All device pointers of a batch.
constexpr int kShardAcc = 12;  // [0..3] g2 n2 gdot Jg2, [4..8] acc_A, [9] step2, [10] xnorm2, [11] candidate cost

struct Batch {
  int B, NPB, NSB, NL, NC, NOBS, NIMU, NMEAS;
  int n_obs_tiles, n_lm_tiles;
  int has_ext;  // any observation whose extrinsics block is estimated
  int fused;    // compact linearisation: r and Jl planes only, Jp rebuilt from Jl in the Schur / back-substitution kernels
  size_t obs_stride;  // plane stride of the per-observation SoA arrays
  WinDesc* win;
  WinState* ws;
  // parameter blocks, double-buffered (x / candidate), plus the uploaded initial values
  double* pose[2];
  double* sb[2];
  double* lm[2];
  double *pose_init, *sb_init, *lm_init;
  int* pose_off;  // [NPB] offset in the window's dense vector or -1
  int* sb_off;    // [NSB]
  uint8_t* lm_fixed;
  int* lm_win;  // [NL]
  double* intr;  // [NC*8]
  // observations (sorted by landmark, pose, camera)
  int *obs_pose, *obs_lm, *obs_ext, *obs_cam;
  int* obs_poff;  // dense offset of the observation's pose block (pose_off[obs_pose]), filled on the device at upload
  double *obs_zx, *obs_zy, *obs_u00, *obs_u01, *obs_u11;
  // observations of landmark l are  lm_obs_first[l] + k * lm_obs_stride[l],  k < lm_obs_cnt[l].
  // Inside a Schur chunk (landmarks sharing one observation pattern) they are stored pattern-major
  // (stride = chunk size) so lane-consecutive landmarks read consecutive addresses; otherwise stride = 1.
  int *lm_obs_first, *lm_obs_stride, *lm_obs_cnt;  // [NL]
  int *obs_tile_win, *obs_tile_begin;
  int *lm_tile_win, *lm_tile_begin;
  // Schur warp chunks: <= 32 consecutive landmarks with an identical (pose, camera) observation pattern
  int n_schur_warps;
  int *sw_win, *sw_lm_begin, *sw_count;
  int* sw_list;           // chunk ids ordered by lane-mapping class (schur_chunk_class), sw_class_count each
  int sw_class_count[kSchurClasses];  // 0..2 k_schur_mma<1|2|4>, 3/7/8 k_schur_lr<1|2|4>, 4..6 k_schur_wr<2..4>
  int *sw_nruns, *sw_run_first;  // pose runs of the chunk's pattern: count, first entry in run_off / run_k0m
  int *run_off, *run_k0m;        // per run: dense offset of its pose block (-1 fixed), (first obs k << 8) | obs count
  // linearisation, two buffers: planes [k][obs_stride]
  double* lin_r[2];   // 2 planes
  double* lin_Jp[2];  // 12 planes
  double* lin_Jl[2];  // 6 planes
  double* lin_Je[2];  // 12 planes (only if has_ext)
  // per-landmark solver state
  double *lm_scale, *lm_Vinv, *lm_bs, *lm_diag, *lm_grad, *lm_gn;
  // per-window dense solver state (concatenated by d_off)
  double *H;  // reduced system accumulators (upper triangle), by H_off
  double *g_red, *g_raw, *Hdiag;  // reduced rhs, unreduced gradient, column square norms  (cleared each slot)
  double *scale_d, *diag_d, *grad_d, *gn_d, *u_d, *c_d, *delta_d;
  double* Jd[2];  // dense-term Jacobians, by Jd_off
  double* gram[2];    // Jd^T Jd per buffer (lower triangle, n x n at H_off), k_dense_gram
  double* gram_g[2];  // Jd^T rd per buffer (at d_off)
  double* rd[2];  // dense-term residuals, by rd_off
  // terms
  ImuTerm* imu;
  ImuCache* imu_cache;
  long long* imu_meas_t;
  double *imu_meas_gyro, *imu_meas_accel;
  PosePrior* pp;
  SbPrior* sp;
  RelPose* rp;
  SonarTerm* so;
  DepthTerm* de;
  MargBlock* marg_blk;
  double *marg_J, *marg_e0, *marg_lin;
  double* lm_quality;
  // Sharded single-window mode (landmarks split over ranks): landmark-side partial sums go here instead of
  // WinState so that they can be all-reduced before k_fold adds them to the replicated dense-side sums.
  // [B][8]: g2, n2, gdot, Jg2 (k_backsub) | mc, step2, xnorm2 (k_step_lm) | cost of the reprojection terms.
  double* shard_acc;  // nullptr when not sharded; kShardAcc doubles per window
  double* gmax_buf;   // [B][comm_world] landmark gradient max, one slot per rank (inside the all-reduced clear region)
  int comm_rank, comm_world;
};

// The caller's observation arrays (window by window, caller order) as uploaded; k_pack_obs gathers them.
struct RawObs {
  const int* pec;                    // packed pose | ext << 10 | cam << 20 (window-local), or nullptr -> the three arrays below
  int one_word;                      // pec holds landmark | pose << 18 | ext << 24 | cam << 30 instead, `lm` is not uploaded
  const int *pose, *lm, *ext, *cam;  // window-local indices
  const double* meas;                // [2] per observation
  const float* meas32;               // the same as float when every coordinate is float-exact (BRISK keypoints are), else nullptr
  const double* info3;               // a00, a10, a11 of the 2x2 information
  const int* order;                  // internal observation -> caller observation (window-local)
  const int* lm_inv;                 // [NL] caller landmark -> internal landmark (window-local)
};

struct SolveParams {
  SvinBaOptions opt;
};

}  // namespace svin
