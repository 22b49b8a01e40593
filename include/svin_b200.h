/*
 * svin_b200.h — C ABI of the B200-native SVIn hot-path engine.
 *
 * Plain pointers and sizes only; no C++/torch types cross this boundary.  All
 * functions return SVIN_OK (0) or a negative status; the message is available
 * from svin_last_error().  Nothing throws across the ABI.  A context is not
 * re-entrant: serialise calls per context (the reference serialises the
 * back-end with estimator_mutex_, ThreadedKFVio.cpp:737,1083, and the front-end
 * per camera with featureDetectorMutexes_[cam], Frontend.cpp:96).
 *
 * Reference seams each entry point replaces (paths relative to
 * okvis_ros/okvis/ in AutonomousFieldRoboticsLab/SVIn):
 *   svin_ba_*   : okvis::Estimator::optimize -> ceres::Map::solve()
 *                 (okvis_ceres/src/Estimator.cpp:876-929,
 *                  okvis_ceres/include/okvis/ceres/Map.hpp:347) and the
 *                 ErrorInterface::EvaluateWithMinimalJacobians seam
 *                 (okvis_ceres/include/okvis/ceres/ErrorInterface.hpp:93-96).
 *   svin_fe_*   : okvis::Frame::detect / describe
 *                 (okvis_cv/include/okvis/implementation/Frame.hpp:93-135),
 *                 i.e. cv::FeatureDetector::detect + cv::DescriptorExtractor::compute
 *                 as constructed at okvis_frontend/src/Frontend.cpp:997-1007.
 *   svin_match_*: okvis::DenseMatcher::match<VioKeyframeWindowMatchingAlgorithm>
 *                 (okvis_matcher/include/okvis/implementation/DenseMatcher.hpp:195-203,
 *                  okvis_frontend/src/VioKeyframeWindowMatchingAlgorithm.cpp:124-323).
 */
#ifndef SVIN_B200_H_
#define SVIN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status */
#define SVIN_OK 0
#define SVIN_ERR_INVALID_ARGUMENT (-1)
#define SVIN_ERR_CUDA (-2)
#define SVIN_ERR_NO_DEVICE (-3)
#define SVIN_ERR_OUT_OF_MEMORY (-4)
#define SVIN_ERR_STATE (-5)

/* Thread-local message for the last failing call on this thread. */
const char* svin_last_error(void);
/* Library version string, e.g. "svin_b200 0.1 (sm_100a)". */
const char* svin_version(void);

/* =====================================================================
 *  (B) sliding-window bundle adjustment
 * ===================================================================== */

/* Robust loss on reprojection terms. Estimator.cpp:61 uses CauchyLoss(1). */
enum { SVIN_LOSS_NONE = 0, SVIN_LOSS_CAUCHY = 1, SVIN_LOSS_HUBER = 2 };

/* Kinds of parameter block referenced by the marginalisation prior. */
enum { SVIN_BLOCK_POSE = 0, SVIN_BLOCK_SPEEDBIAS = 1, SVIN_BLOCK_LANDMARK = 2 };

/* Termination, mirroring ceres::TerminationType as used by Map::solve. */
enum {
  SVIN_TERM_NO_CONVERGENCE = 0, /* max_num_iterations reached */
  SVIN_TERM_CONVERGENCE = 1,    /* function / gradient / parameter tolerance */
  SVIN_TERM_FAILURE = 2,        /* too many invalid steps / radius underflow */
  SVIN_TERM_USER_SUCCESS = 3    /* time limit hit after min_iterations (CeresIterationCallback.hpp:55-80) */
};

/* okvis::ImuParameters fields used by ImuError (Parameters.hpp; ImuError.cpp:76-263). */
typedef struct SvinImuParams {
  double sigma_g_c;  /* gyro noise density */
  double sigma_a_c;  /* accelerometer noise density */
  double sigma_gw_c; /* gyro drift noise density */
  double sigma_aw_c; /* accelerometer drift noise density */
  double g;          /* earth acceleration magnitude */
  double g_max;      /* gyro saturation */
  double a_max;      /* accelerometer saturation */
} SvinImuParams;

/*
 * One sliding window = the contents of okvis::ceres::Map at the moment
 * Estimator::optimize is called, flattened to SoA.  All pointers are HOST
 * pointers owned by the caller.  pose_blocks/speedbias/landmarks are in/out
 * (svin_ba_download / svin_ba_optimize write the solution back); everything
 * else is read-only.
 *
 * Pose blocks hold every okvis::ceres::PoseParameterBlock of the window: the
 * T_WS states AND the camera extrinsics T_SCi (Estimator.cpp:179-262), layout
 * [x y z qx qy qz qw] (PoseParameterBlock.hpp:53).  Speed-and-bias blocks are
 * [v(3) b_g(3) b_a(3)] (SpeedAndBiasParameterBlock.hpp:56).  Landmarks are
 * homogeneous [x y z w] with Euclidean 3-dof update (HomogeneousPointManifold.cpp:57-66).
 */
typedef struct SvinBaWindow {
  /* ---- parameter blocks ---- */
  int32_t num_pose_blocks;
  int32_t num_speedbias;
  int32_t num_landmarks;
  int32_t num_cameras;
  double* pose_blocks;            /* [num_pose_blocks][7]  in/out */
  double* speedbias;              /* [num_speedbias][9]    in/out */
  double* landmarks;              /* [num_landmarks][4]    in/out */
  const uint8_t* pose_fixed;      /* [num_pose_blocks]  Map::setParameterBlockConstant */
  const uint8_t* speedbias_fixed; /* [num_speedbias] */
  const uint8_t* landmark_fixed;  /* [num_landmarks] or NULL (none fixed) */
  const double* intrinsics;       /* [num_cameras][8] fu fv cu cv k1 k2 p1 p2 (PinholeCamera<RadialTangentialDistortion>) */

  /* ---- ReprojectionError terms (ReprojectionError.hpp impl:85-229) ---- */
  int32_t num_obs;
  int32_t loss_type;            /* SVIN_LOSS_* applied to every reprojection term */
  double loss_scale;            /* 'a' of CauchyLoss(a)/HuberLoss(a) */
  const int32_t* obs_pose;      /* [num_obs] pose block index of T_WS */
  const int32_t* obs_landmark;  /* [num_obs] */
  const int32_t* obs_extrinsics;/* [num_obs] pose block index of T_SC */
  const int32_t* obs_camera;    /* [num_obs] intrinsics index */
  const double* obs_measurement;/* [num_obs][2] keypoint */
  const double* obs_information;/* [num_obs][4] row-major 2x2 (Estimator.hpp impl:64-67: 64/size^2 * I) */

  /* ---- ImuError terms (ImuError.cpp) ---- */
  int32_t num_imu;
  SvinImuParams imu_params;
  const int32_t* imu_pose0;       /* [num_imu] pose block of T_WS_0 */
  const int32_t* imu_speedbias0;  /* [num_imu] */
  const int32_t* imu_pose1;       /* [num_imu] */
  const int32_t* imu_speedbias1;  /* [num_imu] */
  const int64_t* imu_t0_ns;       /* [num_imu] okvis::Time t0_ as nanoseconds */
  const int64_t* imu_t1_ns;       /* [num_imu] */
  const int32_t* imu_meas_offset; /* [num_imu+1] into the measurement arrays */
  const int64_t* imu_meas_t_ns;   /* [M] measurement stamps */
  const double* imu_meas_gyro;    /* [M][3] */
  const double* imu_meas_accel;   /* [M][3] */

  /* ---- PoseError terms (PoseError.cpp:85-132) on any pose block ---- */
  int32_t num_pose_priors;
  const int32_t* pose_prior_block;      /* [n] */
  const double* pose_prior_measurement; /* [n][7] */
  const double* pose_prior_information; /* [n][36] row-major 6x6 (may be singular, Estimator.cpp:321-326) */

  /* ---- SpeedAndBiasError terms (SpeedAndBiasError.cpp:84-111) ---- */
  int32_t num_speedbias_priors;
  const int32_t* speedbias_prior_block;      /* [n] */
  const double* speedbias_prior_measurement; /* [n][9] */
  const double* speedbias_prior_information; /* [n][81] */

  /* ---- RelativePoseError terms (RelativePoseError.cpp:76-147) ---- */
  int32_t num_relative_pose;
  const int32_t* relative_pose_block0;      /* [n] */
  const int32_t* relative_pose_block1;      /* [n] */
  const double* relative_pose_information;  /* [n][36] */

  /* ---- SonarError terms (SonarError.cpp:113-183) ---- */
  int32_t num_sonar;
  const int32_t* sonar_pose;         /* [n] */
  const double* sonar_range;         /* [n] */
  const double* sonar_heading;       /* [n] */
  const double* sonar_information;   /* [n] scalar */
  const double* sonar_landmark_mean; /* [n][3] mean of landmarkSubset_ (constant after construction) */
  const double* sonar_T_SSo;         /* [7] sonar extrinsics (config sonar_params.T_SSo) */

  /* ---- DepthError terms (DepthError.cpp:70-139) ---- */
  int32_t num_depth;
  const int32_t* depth_pose;        /* [n] */
  const double* depth_measurement;  /* [n] */
  const double* depth_first;        /* [n] first_depth_ */
  const double* depth_information;  /* [n] scalar */

  /* ---- MarginalizationError (MarginalizationError.cpp:798-844) ---- */
  int32_t marg_num_blocks;                 /* 0 = no prior */
  int32_t marg_dim;                        /* rows of e0 = cols of J = sum of minimal dims */
  const int32_t* marg_block_kind;          /* [marg_num_blocks] SVIN_BLOCK_* */
  const int32_t* marg_block_index;         /* [marg_num_blocks] index into the matching block array */
  const double* marg_linearization_points; /* concatenated 7/9/4 doubles per block, in block order */
  const double* marg_J;                    /* [marg_dim][marg_dim] row-major  (J_) */
  const double* marg_e0;                   /* [marg_dim] */
} SvinBaWindow;

/* Solver options: the subset of ceres::Solver::Options Estimator::optimize sets
 * (Estimator.cpp:878-899) plus the Ceres 2.2 defaults it leaves untouched. */
typedef struct SvinBaOptions {
  int32_t max_num_iterations;   /* numIter */
  int32_t min_num_iterations;   /* CeresIterationCallback min iterations; only with time_limit >= 0 */
  double time_limit_seconds;    /* < 0: none (Estimator.cpp:934-938) */
  double initial_trust_region_radius; /* 1e4 */
  double max_trust_region_radius;     /* 1e16 */
  double min_trust_region_radius;     /* 1e-32 */
  double min_relative_decrease;       /* 1e-3 */
  double min_lm_diagonal;             /* 1e-6 */
  double max_lm_diagonal;             /* 1e32 */
  double function_tolerance;          /* 1e-6 */
  double gradient_tolerance;          /* 1e-10 */
  double parameter_tolerance;         /* 1e-8 */
  int32_t max_num_consecutive_invalid_steps; /* 5 */
  int32_t jacobi_scaling;             /* 1 */
  int32_t compute_landmark_quality;   /* 1: Estimator.cpp:903-922 post-pass */
} SvinBaOptions;

void svin_ba_default_options(SvinBaOptions* opt);

typedef struct SvinBaSummary {
  int32_t iterations;           /* trust-region iterations performed (successful + unsuccessful) */
  int32_t num_successful_steps;
  int32_t termination;          /* SVIN_TERM_* */
  int32_t imu_repropagations;   /* total redoPreintegration calls (ImuError.cpp:739-747) */
  double initial_cost;
  double final_cost;
  double final_trust_region_radius;
} SvinBaSummary;

/* Residual/Jacobian dump of one window at its current estimate (the
 * EvaluateWithMinimalJacobians seam).  Any pointer may be NULL to skip.
 * Reprojection outputs are the *raw* (not loss-corrected) weighted residuals
 * and minimal Jacobians exactly as ReprojectionError returns them. */
typedef struct SvinBaEvaluation {
  double* reproj_residuals;   /* [num_obs][2] */
  double* reproj_J_pose;      /* [num_obs][2][6]  J0_minimal */
  double* reproj_J_landmark;  /* [num_obs][2][3]  J1_minimal */
  double* reproj_J_extrinsics;/* [num_obs][2][6]  J2_minimal */
  double* imu_residuals;      /* [num_imu][15] */
  double* imu_J_pose0;        /* [num_imu][15][6] */
  double* imu_J_speedbias0;   /* [num_imu][15][9] */
  double* imu_J_pose1;        /* [num_imu][15][6] */
  double* imu_J_speedbias1;   /* [num_imu][15][9] */
  double* cost;               /* [1] 0.5*sum rho(|r|^2) over all terms */
} SvinBaEvaluation;

typedef struct svin_ba_ctx svin_ba_ctx;

int svin_ba_create(int device, svin_ba_ctx** out);
void svin_ba_destroy(svin_ba_ctx* ctx);

/* Copy `num_windows` independent windows to the device (replaces any previous
 * batch).  Windows in one batch are solved concurrently by the same kernels. */
int svin_ba_upload(svin_ba_ctx* ctx, const SvinBaWindow* windows, int32_t num_windows);

/* Evaluate all terms of window `window_index` at the uploaded estimate. */
int svin_ba_evaluate(svin_ba_ctx* ctx, int32_t window_index, SvinBaEvaluation* out);

/* Run the trust-region (dogleg + Schur) solve on every uploaded window.
 * `summaries` may be NULL, else [num_windows].  Synchronous on return. */
int svin_ba_solve(svin_ba_ctx* ctx, const SvinBaOptions* opt, SvinBaSummary* summaries);

/* Copy the solution of window `window_index` back into the caller's window
 * (pose_blocks, speedbias, landmarks).  landmark_quality ([num_landmarks], may
 * be NULL) receives sqrt(lambda_min/lambda_max) of the landmark Hessian block
 * (Estimator.cpp:903-922). */
int svin_ba_download(svin_ba_ctx* ctx, int32_t window_index, SvinBaWindow* window, double* landmark_quality);
/* All windows of the uploaded batch with one device-to-host copy (windows[i] must have the shape of the uploaded
 * window i; landmark_quality may be NULL or hold one pointer per window).  upload / solve / download_all is
 * svin_ba_optimize in three stages, for callers that overlap the upload of one batch (on one context) with the
 * solve of another (on a second context). */
int svin_ba_download_all(svin_ba_ctx* ctx, SvinBaWindow* windows, int32_t num_windows, double* const* landmark_quality);

/* Reset the estimate on the device to the uploaded initial values (and the IMU
 * pre-integration state) without a new host->device copy: lets a benchmark
 * repeat a solve on HBM-resident inputs. */
int svin_ba_reset(svin_ba_ctx* ctx);

/* upload + solve + download in one call: the drop-in for Estimator::optimize.
 * landmark_quality: NULL or array of num_windows pointers (each NULL or [num_landmarks]). */
int svin_ba_optimize(svin_ba_ctx* ctx, SvinBaWindow* windows, int32_t num_windows, const SvinBaOptions* opt,
                     SvinBaSummary* summaries, double* const* landmark_quality);

/* Device time (ms, CUDA events on the engine's stream) of the last svin_ba_solve,
 * and per-kernel-family launch counts / accumulated ms since the last upload. */
typedef struct SvinBaTimings {
  double solve_ms;
  double h2d_ms;
  double d2h_ms;
  int64_t kernel_launches;
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  /* host wall-clock parts of the last svin_ba_upload / svin_ba_optimize (the end-to-end path) */
  double host_order_ms;   /* validation + landmark pattern ordering */
  double host_fill_ms;    /* packing the pinned staging arena */
  double host_upload_ms;  /* whole svin_ba_upload, incl. the H2D copy and work-buffer initialisation */
  double host_scatter_ms; /* writing results back into the caller's arrays (svin_ba_optimize) */
} SvinBaTimings;
int svin_ba_timings(svin_ba_ctx* ctx, SvinBaTimings* out);

/* Per-kernel-family device time of the last svin_ba_solve, measured with CUDA events on the engine's
 * stream around every launch.  Only filled when profiling is enabled (it adds an event pair per launch,
 * so throughput numbers are taken with it off). */
enum {
  SVIN_BA_K_LINEARIZE = 0, /* reprojection residual + Jacobian, one thread per observation */
  SVIN_BA_K_DENSE_EVAL,    /* IMU / prior / sonar / depth / marginalisation terms */
  SVIN_BA_K_SCHUR,         /* landmark-block elimination into the reduced system */
  SVIN_BA_K_DENSE_SOLVE,   /* reduced-system Cholesky + solve */
  SVIN_BA_K_BACKSUB,       /* landmark back-substitution + Cauchy point */
  SVIN_BA_K_STEP_DENSE,    /* dogleg step, candidate poses */
  SVIN_BA_K_STEP_LM,       /* candidate landmarks + model cost */
  SVIN_BA_K_DECIDE,        /* accept / reject */
  SVIN_BA_K_CLEAR,         /* per-slot memset of the reduced-system accumulators */
  SVIN_BA_K_COUNT
};
typedef struct SvinBaKernelTimes {
  double ms[SVIN_BA_K_COUNT];
  int64_t launches[SVIN_BA_K_COUNT];
} SvinBaKernelTimes;
/* ---- marginalisation (MarginalizationError.cpp:126-397 addResidualBlock, :463-721 marginalizeOut,
 * :725-758 updateErrorComputation).  The host keeps the graph book-keeping of
 * Estimator::applyMarginalizationStrategy (Estimator.cpp:495-814) and hands the numeric core one window that
 * contains exactly the residual blocks to be linearised into the prior, with every parameter block set to its
 * linearisation point (first-estimate point where the block is already part of the prior).  All landmarks of
 * that window are marginalised; dense blocks are marginalised where the flags say so.  The dense ordering is
 * the engine's: non-fixed pose blocks in index order (6 each), then non-fixed speed/bias blocks (9 each). */
typedef struct SvinMargSpec {
  int32_t prior_num_blocks;           /* blocks of the existing prior (0 = none), in the order of prior_H rows */
  const int32_t* prior_block_kind;    /* SVIN_BLOCK_POSE / SVIN_BLOCK_SPEEDBIAS */
  const int32_t* prior_block_index;   /* index into the window's pose / speedbias arrays */
  int32_t prior_dim;
  const double* prior_H;              /* [prior_dim][prior_dim] H_ */
  const double* prior_b0;             /* [prior_dim] b0_ */
  const uint8_t* marginalize_pose;      /* [num_pose_blocks] */
  const uint8_t* marginalize_speedbias; /* [num_speedbias] */
} SvinMargSpec;
typedef struct SvinMargResult {     /* caller-allocated, capacity = the window's dense dimension */
  int32_t dim;                      /* dimension after marginalisation */
  int32_t num_blocks;
  int32_t* block_kind;              /* [<= num_pose_blocks + num_speedbias] kept blocks, dense order */
  int32_t* block_index;
  double* H;                        /* [dim][dim] */
  double* b0;                       /* [dim] */
  double* J;                        /* [dim][dim]  J_ of updateErrorComputation */
  double* e0;                       /* [dim] */
} SvinMargResult;
/* Runs on window `window_index` of the uploaded batch (svin_ba_upload). */
int svin_ba_marginalize(svin_ba_ctx* ctx, int32_t window_index, const SvinMargSpec* spec, SvinMargResult* out);

/* ---- sharded single-window mode (BASELINE configs[3]): every rank uploads the SAME windows with the same pose /
 * speed-bias blocks and dense terms but a disjoint subset of the landmarks and their observations.  Each
 * trust-region iteration has three exchange steps, each ONE sum all-reduce of one packed buffer on the engine's
 * stream: (1) after the landmark elimination the reduced system [H | g_red | g_raw | Hdiag] of every window plus one
 * gradient-max slot per rank, (2) after the landmark back-substitution the nine landmark-side sums the dogleg
 * coefficients need, (3) before accept/reject the landmark step / state norms and the candidate's reprojection cost.
 * (2) and (3) cannot be folded into (1): the Gauss-Newton step norm needs the solved reduced system, the candidate cost
 * needs the step.  The Cholesky, the dogleg logic and accept/reject run replicated and bit-identically on every rank.
 * time_limit_seconds must be < 0 in this mode (ranks must not decide on their own clocks).
 * NCCL is dlopen'ed (libnccl.so.2); the communicator is created from a 128-byte unique id that rank 0 obtains
 * with svin_nccl_unique_id() and the host distributes (e.g. torch.distributed / MPI / a file).
 * svin_ba_comm_init_local joins `world_size` contexts of THIS process (one device) into an in-process group instead:
 * rank = position in `ctxs`; their svin_ba_solve calls must then run concurrently on `world_size` host threads. */
int svin_nccl_unique_id(uint8_t out[128]);
int svin_ba_comm_init(svin_ba_ctx* ctx, const uint8_t unique_id[128], int32_t rank, int32_t world_size);
int svin_ba_comm_init_local(svin_ba_ctx* const* ctxs, int32_t world_size);

/* Host-side planning only (no device needed): how svin_ba_upload would order one window's landmarks and cut them
 * into Schur chunks.  landmark_order [num_landmarks]: internal position -> caller landmark; per chunk (in internal
 * order) its kernel kind (0..2 lane = landmark, 3/7/8 run-parallel with 1/2/4 warps, 4..6 warp per run for 2..4 runs),
 * landmark count and pose-run count.  Returns SVIN_ERR_INVALID_ARGUMENT if `capacity` chunks are not enough
 * (*num_chunks is still set).  Test / diagnostics hook for the ordering logic of the upload path. */
int svin_ba_plan(const SvinBaWindow* window, int32_t* landmark_order, int32_t capacity, int32_t* chunk_kind,
                 int32_t* chunk_landmarks, int32_t* chunk_runs, int32_t* num_chunks);

/* The same planner's observation order: internal position -> caller observation (window-local), [num_obs]. */
int svin_ba_plan_observations(const SvinBaWindow* window, int32_t* observation_order);
/* What svin_ba_upload actually put on the device for window `window_index` (read back from HBM), in the format of
 * svin_ba_plan + svin_ba_plan_observations.  *planned_on_device = 1 when the device planner (csrc/ba_plan.cu) produced it:
 * svin_ba_upload orders the landmarks, cuts the chunks and builds the observation order ON THE GPU when every window of
 * the batch has its observations sorted by (landmark, pose, camera), fixed extrinsics, <= 64 pose blocks and <= 8192
 * landmarks, and on the host threads otherwise (SVIN_BA_DEVICE_PLAN=0 forces the host).  Both must agree entry by entry:
 * tests/test_plan_gpu.py. */
int svin_ba_uploaded_plan(svin_ba_ctx* ctx, int32_t window_index, int32_t* landmark_order, int32_t capacity,
                          int32_t* chunk_kind, int32_t* chunk_landmarks, int32_t* chunk_runs, int32_t* num_chunks,
                          int32_t* observation_order, int32_t* planned_on_device);

int svin_ba_set_profiling(svin_ba_ctx* ctx, int enable);
int svin_ba_kernel_times(svin_ba_ctx* ctx, SvinBaKernelTimes* out);

/* =====================================================================
 *  (A) BRISK-2 front-end: detect / describe / match
 * ===================================================================== */

/* cv::KeyPoint memory layout (28 bytes), as filled by Frame::detect / Frame::describe. */
typedef struct SvinKeypoint {
  float x, y;     /* pt */
  float size;     /* 12 at octave 0 */
  float angle;    /* degrees, gravity-aligned (Frame.hpp impl:113-129) */
  float response; /* Harris score */
  int32_t octave;
  int32_t class_id;
} SvinKeypoint;

#define SVIN_DESCRIPTOR_BYTES 48 /* 384 bit, Hamming over 3 x 128 bit (VioKeyframeWindowMatchingAlgorithm.hpp:263) */

/* The arguments of brisk::ScaleSpaceFeatureDetector<HarrisScoreCalculator>(threshold, octaves, absoluteThreshold,
 * maxNoKeypoints) and brisk::BriskDescriptorExtractor(rotationInvariance, scaleInvariance) as constructed at
 * okvis_frontend/src/Frontend.cpp:997-1007 from detection_options.* (config_fpga_p2_euroc.yaml:65-68). */
typedef struct SvinFeOptions {
  int32_t image_width, image_height;
  double detection_threshold; /* uniformity radius in pixels (40) */
  int32_t detection_octaves;  /* only 0 (single scale) is implemented, as in both shipped configs */
  double absolute_threshold;  /* 800 (Frontend.cpp:75) */
  int32_t max_keypoints;      /* 400 */
  int32_t rotation_invariance;/* 1: use the supplied keypoint angle */
  int32_t scale_invariance;   /* must be 0 */
  int32_t max_images;         /* capacity of one batched call */
} SvinFeOptions;

void svin_fe_default_options(SvinFeOptions* opt);

typedef struct svin_fe_ctx svin_fe_ctx;
int svin_fe_create(int device, const SvinFeOptions* opt, svin_fe_ctx** out);
void svin_fe_destroy(svin_fe_ctx* ctx);

/* Frame::detect + Frame::describe for `num_images` images of the configured size in one launch sequence.
 * images[i] points to 8-bit grey rows with `stride` bytes between rows; intrinsics [n][8] (fu fv cu cv k1 k2 p1 p2);
 * extraction_direction [n][3] is gravity in the camera frame (Frontend.cpp:107-108).  Outputs, all host buffers:
 * keypoints [n][max_keypoints], descriptors [n][max_keypoints][48], counts [n]. */
int svin_fe_detect_describe(svin_fe_ctx* ctx, int32_t num_images, const uint8_t* const* images, int32_t stride,
                            const double* intrinsics, const double* extraction_direction, SvinKeypoint* keypoints,
                            uint8_t* descriptors, int32_t* counts);
/* Split form for device-resident benchmarking: upload once, run many times, download. */
int svin_fe_upload(svin_fe_ctx* ctx, int32_t num_images, const uint8_t* const* images, int32_t stride,
                   const double* intrinsics, const double* extraction_direction);
int svin_fe_run(svin_fe_ctx* ctx);
int svin_fe_download(svin_fe_ctx* ctx, SvinKeypoint* keypoints, uint8_t* descriptors, int32_t* counts);
/* Harris score image of uploaded image `index` (int32 [H][W]); test hook. */
int svin_fe_scores(svin_fe_ctx* ctx, int32_t index, int32_t* out);

/* One DenseMatcher::match<VioKeyframeWindowMatchingAlgorithm> call (3.2 in SURVEY.md). */
enum { SVIN_MATCH_3D2D = 0, SVIN_MATCH_2D2D = 1 };
typedef struct SvinMatchProblem {
  int32_t type;                 /* SVIN_MATCH_* (matchingType_) */
  int32_t nA, nB;
  const uint8_t* descA;         /* [nA][48] */
  const uint8_t* descB;         /* [nB][48] */
  const uint8_t* skipA;         /* [nA] graph-state skips decided by the host (landmark not added/initialised...), or NULL */
  const uint8_t* skipB;         /* [nB] or NULL */
  const SvinKeypoint* kpA;      /* [nA] */
  const SvinKeypoint* kpB;      /* [nB] */
  float distance_threshold;     /* 60 (Frontend.cpp:79) */
  const double* landmarksA;     /* 3D-2D: [nA][4] hp_W of the landmark each A keypoint observes */
  const double* T_CbW;          /* 3D-2D: [7] */
  double pose_uncertainty;      /* 3D-2D: UOplus translation variance (VKWMA.cpp:132-144) */
  const double* intrA;          /* [8] */
  const double* intrB;          /* [8] */
  const double* T_CaCb;         /* 2D-2D: [7] */
  int32_t image_width, image_height;
} SvinMatchProblem;
typedef struct SvinMatchResult {
  int32_t* best_index;     /* [nA][4] best-4 list per A keypoint (index in B, -1 = empty) */
  float* best_distance;    /* [nA][4] */
  int32_t* match_of_B;     /* [nB] final pairing per B keypoint (index in A or -1), B-index order like matchBody */
  float* match_distance;   /* [nB] */
  uint8_t* skipA_effective;/* [nA] skipA after doSetup (projection failures added), or NULL */
} SvinMatchResult;
/* Solves `num_problems` independent match problems; assignment order is the declared deterministic one
 * (A ascending, single worker).  Host buffers in and out. */
int svin_match(svin_fe_ctx* ctx, int32_t num_problems, const SvinMatchProblem* problems, SvinMatchResult* results);

/* =====================================================================
 *  (A10) RANSAC between matching and bundle adjustment  
 *  Replaces the three opengv::sac::Ransac<...>::computeModel calls of
 *    Frontend::runRansac3d2d               (okvis_frontend/src/Frontend.cpp:617-676)  FrameAbsolutePoseSacProblem
 *    Frontend::runRansac2d2d               (:832-980)   FrameRotationOnlySacProblem + FrameRelativePoseSacProblem
 *    Frontend::runRansac2d2dToRefineScale  (:680-830)   the same pair between camera 0 and camera 1
 *  The adapter flattens its correspondences exactly as FrameNoncentralAbsoluteAdapter / FrameRelativeAdapter do
 *  (okvis_frontend/src/FrameNoncentralAbsoluteAdapter.cpp:64-131, FrameRelativeAdapter.cpp:60-190): unit bearing
 *  vectors from backProject, sigma_angle = sqrt(2) (0.8 size / 12)^2 / fu^2, camera extrinsics per correspondence.
 *  Consensus scores: the in-tree getSelectedDistancesToModel of the three problems, exactly.  The loop
 *  (most inliers wins, adaptive bound k = log(1 - 0.99) / log(1 - w^sampleSize), <= max_iterations + 1 iterations,
 *  failed models skipped) follows OpenGV's sac::Ransac.  Minimal solvers (OpenGV is not in the reference tree):
 *  absolute pose - central P3P (Grunert) on three points of ONE camera + 1 point to pick the solution, consensus over
 *  all cameras (the reference asks OpenGV for GP3P; samples mixing cameras give no model); relative pose - eight-point
 *  on 8 samples (OpenGV's STEWENIUS uses 5 + 3; EIGHTPT is another algorithm of the same problem class);
 *  rotation only - 2 points.  The sample index sets are an INPUT so that results are reproducible: sample j is
 *  hypothesis j, consumed in order.  All hypotheses are evaluated concurrently on the device and the sequential
 *  stop rule is replayed over their inlier counts, which selects the hypothesis the sequential loop would.
 *  Models are row-major 3x4 [R | t]: absolute = body pose in the world (T_WS), relative = T_C1C2 with |t| = 1,
 *  rotation only = [R12 | 0].
 * ===================================================================== */
typedef struct svin_ransac_ctx svin_ransac_ctx;
typedef struct SvinRansacAbsProblem {
  int32_t num_correspondences;
  const double* points;          /* [n][3] landmarks in the world (hp.head<3>() / hp[3]) */
  const double* bearings;        /* [n][3] unit bearing vectors in their camera frame */
  const int32_t* camera_index;   /* [n] */
  const double* sigma_angle;     /* [n] */
  int32_t num_cameras;
  const double* camera_rotation; /* [num_cameras][9] row-major C_SC (frame->T_SC(im)->C()) */
  const double* camera_offset;   /* [num_cameras][3] r_SC */
  int32_t num_samples;           /* hypotheses offered; OpenGV consumes at most max_iterations + 1 valid ones */
  const int32_t* samples;        /* [num_samples][4] correspondence indices, the first three in one camera */
  double threshold;              /* 9 (Frontend.cpp:643) */
  int32_t max_iterations;        /* 50 (Frontend.cpp:644) */
} SvinRansacAbsProblem;
typedef struct SvinRansacRelProblem {
  int32_t num_correspondences;
  const double* bearings1;       /* [n][3] unit, frame 1 (older frame / camera A) */
  const double* bearings2;       /* [n][3] unit, frame 2 */
  const double* sigma_angle1;    /* [n] */
  const double* sigma_angle2;    /* [n] */
  int32_t num_samples;
  const int32_t* samples_rotation; /* [num_samples][2] */
  const int32_t* samples_relative; /* [num_samples][8] */
  double threshold;
  int32_t max_iterations;
} SvinRansacRelProblem;
typedef struct SvinRansacResult {
  int32_t best_sample;           /* winning hypothesis (-1: no model, inliers all 0) */
  int32_t num_inliers;
  int32_t iterations;            /* valid hypotheses the sequential loop would have evaluated */
  double model[12];
  uint8_t* inliers;              /* [n] caller-allocated or NULL: selectWithinDistance of the winning model */
  int32_t* hypothesis_inliers;   /* [num_samples] caller-allocated or NULL: inlier count of every hypothesis */
  uint8_t* hypothesis_valid;     /* [num_samples] caller-allocated or NULL: the minimal solver produced a model */
} SvinRansacResult;
int svin_ransac_create(int device, svin_ransac_ctx** out);
void svin_ransac_destroy(svin_ransac_ctx* ctx);
int svin_ransac_absolute(svin_ransac_ctx* ctx, int32_t num_problems, const SvinRansacAbsProblem* problems,
                         SvinRansacResult* results);
/* Both RANSACs of runRansac2d2d on every problem; the ratio rule between them (Frontend.cpp:876-905) stays in the caller. */
int svin_ransac_relative(svin_ransac_ctx* ctx, int32_t num_problems, const SvinRansacRelProblem* problems,
                         SvinRansacResult* rotation_only, SvinRansacResult* relative_pose);
int svin_ransac_timings(svin_ransac_ctx* ctx, double* device_ms, int64_t* kernel_launches);

/* =====================================================================
 *  (L) loop-closure QUERY path of pose_graph  (BASELINE configs[4]) - partial row:
 *  the data-parallel part of LoopClosure / PoseGraph::detectLoop, i.e. DBoW2's vocabulary transform, the inverted-file L1
 *  query and the BRIEF-256 candidate search.  FAST + BRIEF extraction, PnPRANSAC and the pose-graph optimisation are
 *  NOT here.  Replaces
 *    voc->transform(brief_descriptors, bowVec)     pose_graph/src/pose_graph/PoseGraph.cpp:176,
 *                                                   pose_graph/ThirdParty/DBoW/TemplatedVocabulary.h:981-1028,1114-1153
 *    db.add(brief_descriptors)                      PoseGraph.cpp:197
 *    db.query(bowVec, ret, 4, frame_index - 50)     PoseGraph.cpp:196, ThirdParty/DBoW/TemplatedDatabase.h:587-646
 *    Keyframe::searchByBRIEFDes                     pose_graph/src/pose_graph/Keyframe.cpp:262-306
 *  Scoring: TF-IDF weights, L1 norm, L1 score (the defaults of TemplatedVocabulary.h:48).  Equal scores are returned in
 *  ascending entry id (std::sort leaves that unspecified in the reference).
 *  The vocabulary is the flattened DBoW2 tree: node 0 is the root, the children of a node are contiguous.
 *  Sharding over ranks: entry e lives on rank e % world_size; every rank adds the same image sequence and keeps its own
 *  entries; svin_loop_query returns the rank's best results with GLOBAL entry ids, the caller merges 4 x world_size pairs.
 * ===================================================================== */
typedef struct svin_loop_ctx svin_loop_ctx;
typedef struct SvinVocabulary {
  int32_t num_nodes;
  const int32_t* first_child;   /* [num_nodes] index of the first child (> node) */
  const int32_t* num_children;  /* [num_nodes] 0 = leaf (word) */
  const uint8_t* descriptor;    /* [num_nodes][32] BRIEF-256 */
  const double* weight;         /* [num_nodes] idf weight of a word (leaves) */
  const int32_t* word_id;       /* [num_nodes] word id of a leaf, -1 otherwise */
} SvinVocabulary;
int svin_loop_create(int device, const SvinVocabulary* vocabulary, int32_t rank, int32_t world_size, svin_loop_ctx** out);
void svin_loop_destroy(svin_loop_ctx* ctx);
/* descriptors: the images' BRIEF descriptors back to back ([sum counts][32]); word_ids / values: capacity sum counts,
 * image i's sparse vector starts at offset sum(counts[0..i)) and has num_words[i] entries (ids ascending). */
int svin_loop_transform(svin_loop_ctx* ctx, int32_t num_images, const uint8_t* descriptors, const int32_t* counts,
                        int32_t* word_ids, double* values, int32_t* num_words);
int svin_loop_add(svin_loop_ctx* ctx, int32_t num_images, const uint8_t* descriptors, const int32_t* counts);
/* max_id = -1: all entries, else only entries with id < max_id (PoseGraph.cpp:196: frame_index - 50). */
int svin_loop_query(svin_loop_ctx* ctx, const uint8_t* descriptors, int32_t count, int32_t max_results, int32_t max_id,
                    int32_t* entry_ids, double* scores, int32_t* num_results);
int svin_loop_brief_search(svin_loop_ctx* ctx, const uint8_t* window_descriptors, int32_t num_window,
                           const uint8_t* old_descriptors, int32_t num_old, int32_t* best_index, int32_t* best_distance,
                           uint8_t* status);
int svin_loop_stats(svin_loop_ctx* ctx, int32_t* entries_total, int32_t* entries_local, double* last_device_ms);

/* =====================================================================
 *  (A0) image pre-processing in front of the detector  
 *  Replaces the OpenCV chain of Subscriber::imageCallback (okvis_ros/src/Subscriber.cpp:123-147):
 *    cv::resize(raw, Size(), resizeFactor, resizeFactor)  [INTER_LINEAR; an exact 2x decimation takes OpenCV's
 *    INTER_AREA fast path]  ->  cv::medianBlur(3) if optimization.useMedianFilter  ->  CLAHE::apply / cv::equalizeHist
 *  per histogramParams (config_stereorig_v1.yaml:103-106).  8-bit single channel, bit-exact with OpenCV 4.x.
 * ===================================================================== */
enum { SVIN_HIST_NONE = 0, SVIN_HIST_EQUALIZE = 1, SVIN_HIST_CLAHE = 2 };
typedef struct SvinPreOptions {
  int32_t src_width, src_height;  /* raw image size */
  double resize_factor;           /* miscParams.resizeFactor (1.0 = no resize) */
  int32_t median_filter;          /* optimization.useMedianFilter */
  int32_t histogram_method;       /* SVIN_HIST_* (histogramParams.histogramMethod) */
  double clahe_clip_limit;        /* histogramParams.claheClipLimit */
  int32_t clahe_tiles;            /* histogramParams.claheTilesGridSize (tiles x tiles) */
  int32_t max_images;             /* capacity of one batched call */
} SvinPreOptions;
typedef struct SvinPreTimings {
  double run_ms, h2d_ms, d2h_ms;  /* CUDA events of the last call */
  int64_t h2d_bytes, d2h_bytes;
  int64_t kernel_launches;
  double kernel_ms[5];            /* resize, median, histogram, lut, apply */
} SvinPreTimings;
typedef struct svin_pre_ctx svin_pre_ctx;
int svin_pre_create(int device, const SvinPreOptions* opt, svin_pre_ctx** out);
void svin_pre_destroy(svin_pre_ctx* ctx);
int svin_pre_output_size(svin_pre_ctx* ctx, int32_t* width, int32_t* height);
/* Host buffers in and out: src[i] rows `src_stride` bytes apart, dst[i] dense [height][width]. */
int svin_pre_process(svin_pre_ctx* ctx, int32_t num_images, const uint8_t* const* src, int32_t src_stride,
                     uint8_t* const* dst);
/* Split form for device-resident timing and for chaining into svin_fe_upload_device. */
int svin_pre_upload(svin_pre_ctx* ctx, int32_t num_images, const uint8_t* const* src, int32_t src_stride);
int svin_pre_run(svin_pre_ctx* ctx);
int svin_pre_download(svin_pre_ctx* ctx, uint8_t* const* dst);
/* Device pointer of the processed images of the last run, dense [num_images][height][width] (valid until the next
 * call on this context); the stream-ordered hand-over to the detector without a host round trip. */
const uint8_t* svin_pre_device_output(svin_pre_ctx* ctx);
int svin_pre_timings(svin_pre_ctx* ctx, SvinPreTimings* out);
/* svin_fe_upload with the images already on the device (dense rows of image_width bytes), e.g. svin_pre_device_output. */
int svin_fe_upload_device(svin_fe_ctx* ctx, int32_t num_images, const uint8_t* device_images,
                          const double* intrinsics, const double* extraction_direction);

typedef struct SvinFeTimings {
  double run_ms;   /* device time of the last svin_fe_run / svin_match (CUDA events) */
  double h2d_ms, d2h_ms;
  int64_t kernel_launches;
  int64_t h2d_bytes, d2h_bytes;
  double kernel_ms[8]; /* harris, nms-compact, sort, uniformity, orient-describe, match, assign, (unused) */
} SvinFeTimings;
int svin_fe_timings(svin_fe_ctx* ctx, SvinFeTimings* out);

#ifdef __cplusplus
}
#endif
#endif /* SVIN_B200_H_ */
