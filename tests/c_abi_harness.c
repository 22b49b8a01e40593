/* C harness for include/svin_b200.h: proves that the header is valid C99, prints the layout of every
 * struct of the ABI (tests/test_c_abi.py compares it with the ctypes mirror in svin_b200/capi.py), and pushes one window
 * through svin_ba_optimize the way a C/C++ adapter would - no Python between the caller's buffers and the library.
 *
 *   c_abi_harness layout                       -> "Struct.field offset size" per line, "Struct sizeof" per struct
 *   c_abi_harness solve <in.bin> <out.bin>     -> reads a window dump, runs svin_ba_optimize on device 0, writes the
 *                                                 summary + solution (needs a GPU)
 * Window dump: int32 counts[16] then the arrays in SvinBaWindow order, each as int64 byte count + bytes.
 * Built by the test with:  gcc -std=c99 -Wall -Wextra -Werror -pedantic tests/c_abi_harness.c -Iinclude -Lsvin_b200/lib -lsvin_b200 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "svin_b200.h"

#define FIELD(S, f) printf(#S "." #f " %zu %zu\n", offsetof(S, f), sizeof(((S*)0)->f))
#define SIZE(S) printf(#S " %zu\n", sizeof(S))

static void layout(void) {
  SIZE(SvinImuParams);
  FIELD(SvinImuParams, sigma_g_c);
  FIELD(SvinImuParams, sigma_a_c);
  FIELD(SvinImuParams, sigma_gw_c);
  FIELD(SvinImuParams, sigma_aw_c);
  FIELD(SvinImuParams, g);
  FIELD(SvinImuParams, g_max);
  FIELD(SvinImuParams, a_max);
  SIZE(SvinBaWindow);
  FIELD(SvinBaWindow, num_pose_blocks);
  FIELD(SvinBaWindow, num_speedbias);
  FIELD(SvinBaWindow, num_landmarks);
  FIELD(SvinBaWindow, num_cameras);
  FIELD(SvinBaWindow, pose_blocks);
  FIELD(SvinBaWindow, speedbias);
  FIELD(SvinBaWindow, landmarks);
  FIELD(SvinBaWindow, pose_fixed);
  FIELD(SvinBaWindow, speedbias_fixed);
  FIELD(SvinBaWindow, landmark_fixed);
  FIELD(SvinBaWindow, intrinsics);
  FIELD(SvinBaWindow, num_obs);
  FIELD(SvinBaWindow, loss_type);
  FIELD(SvinBaWindow, loss_scale);
  FIELD(SvinBaWindow, obs_pose);
  FIELD(SvinBaWindow, obs_landmark);
  FIELD(SvinBaWindow, obs_extrinsics);
  FIELD(SvinBaWindow, obs_camera);
  FIELD(SvinBaWindow, obs_measurement);
  FIELD(SvinBaWindow, obs_information);
  FIELD(SvinBaWindow, num_imu);
  FIELD(SvinBaWindow, imu_params);
  FIELD(SvinBaWindow, imu_pose0);
  FIELD(SvinBaWindow, imu_speedbias0);
  FIELD(SvinBaWindow, imu_pose1);
  FIELD(SvinBaWindow, imu_speedbias1);
  FIELD(SvinBaWindow, imu_t0_ns);
  FIELD(SvinBaWindow, imu_t1_ns);
  FIELD(SvinBaWindow, imu_meas_offset);
  FIELD(SvinBaWindow, imu_meas_t_ns);
  FIELD(SvinBaWindow, imu_meas_gyro);
  FIELD(SvinBaWindow, imu_meas_accel);
  FIELD(SvinBaWindow, num_pose_priors);
  FIELD(SvinBaWindow, pose_prior_block);
  FIELD(SvinBaWindow, pose_prior_measurement);
  FIELD(SvinBaWindow, pose_prior_information);
  FIELD(SvinBaWindow, num_speedbias_priors);
  FIELD(SvinBaWindow, speedbias_prior_block);
  FIELD(SvinBaWindow, speedbias_prior_measurement);
  FIELD(SvinBaWindow, speedbias_prior_information);
  FIELD(SvinBaWindow, num_relative_pose);
  FIELD(SvinBaWindow, relative_pose_block0);
  FIELD(SvinBaWindow, relative_pose_block1);
  FIELD(SvinBaWindow, relative_pose_information);
  FIELD(SvinBaWindow, num_sonar);
  FIELD(SvinBaWindow, sonar_pose);
  FIELD(SvinBaWindow, sonar_range);
  FIELD(SvinBaWindow, sonar_heading);
  FIELD(SvinBaWindow, sonar_information);
  FIELD(SvinBaWindow, sonar_landmark_mean);
  FIELD(SvinBaWindow, sonar_T_SSo);
  FIELD(SvinBaWindow, num_depth);
  FIELD(SvinBaWindow, depth_pose);
  FIELD(SvinBaWindow, depth_measurement);
  FIELD(SvinBaWindow, depth_first);
  FIELD(SvinBaWindow, depth_information);
  FIELD(SvinBaWindow, marg_num_blocks);
  FIELD(SvinBaWindow, marg_dim);
  FIELD(SvinBaWindow, marg_block_kind);
  FIELD(SvinBaWindow, marg_block_index);
  FIELD(SvinBaWindow, marg_linearization_points);
  FIELD(SvinBaWindow, marg_J);
  FIELD(SvinBaWindow, marg_e0);
  SIZE(SvinBaOptions);
  FIELD(SvinBaOptions, max_num_iterations);
  FIELD(SvinBaOptions, min_num_iterations);
  FIELD(SvinBaOptions, time_limit_seconds);
  FIELD(SvinBaOptions, initial_trust_region_radius);
  FIELD(SvinBaOptions, max_trust_region_radius);
  FIELD(SvinBaOptions, min_trust_region_radius);
  FIELD(SvinBaOptions, min_relative_decrease);
  FIELD(SvinBaOptions, min_lm_diagonal);
  FIELD(SvinBaOptions, max_lm_diagonal);
  FIELD(SvinBaOptions, function_tolerance);
  FIELD(SvinBaOptions, gradient_tolerance);
  FIELD(SvinBaOptions, parameter_tolerance);
  FIELD(SvinBaOptions, max_num_consecutive_invalid_steps);
  FIELD(SvinBaOptions, jacobi_scaling);
  FIELD(SvinBaOptions, compute_landmark_quality);
  SIZE(SvinBaSummary);
  FIELD(SvinBaSummary, iterations);
  FIELD(SvinBaSummary, num_successful_steps);
  FIELD(SvinBaSummary, termination);
  FIELD(SvinBaSummary, imu_repropagations);
  FIELD(SvinBaSummary, initial_cost);
  FIELD(SvinBaSummary, final_cost);
  FIELD(SvinBaSummary, final_trust_region_radius);
  SIZE(SvinBaEvaluation);
  FIELD(SvinBaEvaluation, reproj_residuals);
  FIELD(SvinBaEvaluation, reproj_J_pose);
  FIELD(SvinBaEvaluation, reproj_J_landmark);
  FIELD(SvinBaEvaluation, reproj_J_extrinsics);
  FIELD(SvinBaEvaluation, imu_residuals);
  FIELD(SvinBaEvaluation, imu_J_pose0);
  FIELD(SvinBaEvaluation, imu_J_speedbias0);
  FIELD(SvinBaEvaluation, imu_J_pose1);
  FIELD(SvinBaEvaluation, imu_J_speedbias1);
  FIELD(SvinBaEvaluation, cost);
  SIZE(SvinBaTimings);
  FIELD(SvinBaTimings, solve_ms);
  FIELD(SvinBaTimings, h2d_ms);
  FIELD(SvinBaTimings, d2h_ms);
  FIELD(SvinBaTimings, kernel_launches);
  FIELD(SvinBaTimings, h2d_bytes);
  FIELD(SvinBaTimings, d2h_bytes);
  FIELD(SvinBaTimings, host_order_ms);
  FIELD(SvinBaTimings, host_fill_ms);
  FIELD(SvinBaTimings, host_upload_ms);
  FIELD(SvinBaTimings, host_scatter_ms);
  SIZE(SvinBaKernelTimes);
  FIELD(SvinBaKernelTimes, ms);
  FIELD(SvinBaKernelTimes, launches);
  SIZE(SvinMargSpec);
  FIELD(SvinMargSpec, prior_num_blocks);
  FIELD(SvinMargSpec, prior_block_kind);
  FIELD(SvinMargSpec, prior_block_index);
  FIELD(SvinMargSpec, prior_dim);
  FIELD(SvinMargSpec, prior_H);
  FIELD(SvinMargSpec, prior_b0);
  FIELD(SvinMargSpec, marginalize_pose);
  FIELD(SvinMargSpec, marginalize_speedbias);
  SIZE(SvinMargResult);
  FIELD(SvinMargResult, dim);
  FIELD(SvinMargResult, num_blocks);
  FIELD(SvinMargResult, block_kind);
  FIELD(SvinMargResult, block_index);
  FIELD(SvinMargResult, H);
  FIELD(SvinMargResult, b0);
  FIELD(SvinMargResult, J);
  FIELD(SvinMargResult, e0);
  SIZE(SvinKeypoint);
  FIELD(SvinKeypoint, x);
  FIELD(SvinKeypoint, y);
  FIELD(SvinKeypoint, size);
  FIELD(SvinKeypoint, angle);
  FIELD(SvinKeypoint, response);
  FIELD(SvinKeypoint, octave);
  FIELD(SvinKeypoint, class_id);
  SIZE(SvinFeOptions);
  FIELD(SvinFeOptions, image_width);
  FIELD(SvinFeOptions, image_height);
  FIELD(SvinFeOptions, detection_threshold);
  FIELD(SvinFeOptions, detection_octaves);
  FIELD(SvinFeOptions, absolute_threshold);
  FIELD(SvinFeOptions, max_keypoints);
  FIELD(SvinFeOptions, rotation_invariance);
  FIELD(SvinFeOptions, scale_invariance);
  FIELD(SvinFeOptions, max_images);
  SIZE(SvinMatchProblem);
  FIELD(SvinMatchProblem, type);
  FIELD(SvinMatchProblem, nA);
  FIELD(SvinMatchProblem, nB);
  FIELD(SvinMatchProblem, descA);
  FIELD(SvinMatchProblem, descB);
  FIELD(SvinMatchProblem, skipA);
  FIELD(SvinMatchProblem, skipB);
  FIELD(SvinMatchProblem, kpA);
  FIELD(SvinMatchProblem, kpB);
  FIELD(SvinMatchProblem, distance_threshold);
  FIELD(SvinMatchProblem, landmarksA);
  FIELD(SvinMatchProblem, T_CbW);
  FIELD(SvinMatchProblem, pose_uncertainty);
  FIELD(SvinMatchProblem, intrA);
  FIELD(SvinMatchProblem, intrB);
  FIELD(SvinMatchProblem, T_CaCb);
  FIELD(SvinMatchProblem, image_width);
  FIELD(SvinMatchProblem, image_height);
  SIZE(SvinMatchResult);
  FIELD(SvinMatchResult, best_index);
  FIELD(SvinMatchResult, best_distance);
  FIELD(SvinMatchResult, match_of_B);
  FIELD(SvinMatchResult, match_distance);
  FIELD(SvinMatchResult, skipA_effective);
  SIZE(SvinRansacAbsProblem);
  FIELD(SvinRansacAbsProblem, num_correspondences);
  FIELD(SvinRansacAbsProblem, points);
  FIELD(SvinRansacAbsProblem, bearings);
  FIELD(SvinRansacAbsProblem, camera_index);
  FIELD(SvinRansacAbsProblem, sigma_angle);
  FIELD(SvinRansacAbsProblem, num_cameras);
  FIELD(SvinRansacAbsProblem, camera_rotation);
  FIELD(SvinRansacAbsProblem, camera_offset);
  FIELD(SvinRansacAbsProblem, num_samples);
  FIELD(SvinRansacAbsProblem, samples);
  FIELD(SvinRansacAbsProblem, threshold);
  FIELD(SvinRansacAbsProblem, max_iterations);
  SIZE(SvinRansacRelProblem);
  FIELD(SvinRansacRelProblem, num_correspondences);
  FIELD(SvinRansacRelProblem, bearings1);
  FIELD(SvinRansacRelProblem, bearings2);
  FIELD(SvinRansacRelProblem, sigma_angle1);
  FIELD(SvinRansacRelProblem, sigma_angle2);
  FIELD(SvinRansacRelProblem, num_samples);
  FIELD(SvinRansacRelProblem, samples_rotation);
  FIELD(SvinRansacRelProblem, samples_relative);
  FIELD(SvinRansacRelProblem, threshold);
  FIELD(SvinRansacRelProblem, max_iterations);
  SIZE(SvinRansacResult);
  FIELD(SvinRansacResult, best_sample);
  FIELD(SvinRansacResult, num_inliers);
  FIELD(SvinRansacResult, iterations);
  FIELD(SvinRansacResult, model);
  FIELD(SvinRansacResult, inliers);
  FIELD(SvinRansacResult, hypothesis_inliers);
  FIELD(SvinRansacResult, hypothesis_valid);
  SIZE(SvinVocabulary);
  FIELD(SvinVocabulary, num_nodes);
  FIELD(SvinVocabulary, first_child);
  FIELD(SvinVocabulary, num_children);
  FIELD(SvinVocabulary, descriptor);
  FIELD(SvinVocabulary, weight);
  FIELD(SvinVocabulary, word_id);
  SIZE(SvinPreOptions);
  FIELD(SvinPreOptions, src_width);
  FIELD(SvinPreOptions, src_height);
  FIELD(SvinPreOptions, resize_factor);
  FIELD(SvinPreOptions, median_filter);
  FIELD(SvinPreOptions, histogram_method);
  FIELD(SvinPreOptions, clahe_clip_limit);
  FIELD(SvinPreOptions, clahe_tiles);
  FIELD(SvinPreOptions, max_images);
  SIZE(SvinPreTimings);
  FIELD(SvinPreTimings, run_ms);
  FIELD(SvinPreTimings, h2d_ms);
  FIELD(SvinPreTimings, d2h_ms);
  FIELD(SvinPreTimings, h2d_bytes);
  FIELD(SvinPreTimings, d2h_bytes);
  FIELD(SvinPreTimings, kernel_launches);
  FIELD(SvinPreTimings, kernel_ms);
  SIZE(SvinFeTimings);
  FIELD(SvinFeTimings, run_ms);
  FIELD(SvinFeTimings, h2d_ms);
  FIELD(SvinFeTimings, d2h_ms);
  FIELD(SvinFeTimings, kernel_launches);
  FIELD(SvinFeTimings, h2d_bytes);
  FIELD(SvinFeTimings, d2h_bytes);
  FIELD(SvinFeTimings, kernel_ms);
}

static void* read_array(FILE* f) {
  long long n = 0;
  void* p;
  if (fread(&n, sizeof n, 1, f) != 1) return NULL;
  if (n <= 0) return NULL;
  p = malloc((size_t)n);
  if (!p || fread(p, 1, (size_t)n, f) != (size_t)n) {
    fprintf(stderr, "short read\n");
    exit(2);
  }
  return p;
}

static int solve(const char* in_path, const char* out_path) {
  FILE* f = fopen(in_path, "rb");
  int32_t cnt[16];
  SvinBaWindow w;
  SvinBaOptions opt;
  SvinBaSummary summary;
  svin_ba_ctx* ctx = NULL;
  double* quality;
  double* qptr[1];
  int rc;
  FILE* o;
  if (!f || fread(cnt, sizeof(int32_t), 16, f) != 16) {
    fprintf(stderr, "cannot read %s\n", in_path);
    return 2;
  }
  memset(&w, 0, sizeof w);
  w.num_pose_blocks = cnt[0]; w.num_speedbias = cnt[1]; w.num_landmarks = cnt[2]; w.num_cameras = cnt[3];
  w.num_obs = cnt[4]; w.loss_type = cnt[5]; w.num_imu = cnt[6]; w.num_pose_priors = cnt[7];
  w.num_speedbias_priors = cnt[8]; w.num_relative_pose = cnt[9]; w.num_sonar = cnt[10]; w.num_depth = cnt[11];
  w.marg_num_blocks = cnt[12]; w.marg_dim = cnt[13];
  if (w.num_relative_pose || w.num_sonar || w.num_depth) {
    fprintf(stderr, "the dump format carries no relative-pose / sonar / depth terms\n");
    return 2;
  }
  if (fread(&w.loss_scale, sizeof(double), 1, f) != 1 || fread(&w.imu_params, sizeof(SvinImuParams), 1, f) != 1) return 2;
  w.pose_blocks = (double*)read_array(f); w.speedbias = (double*)read_array(f); w.landmarks = (double*)read_array(f);
  w.pose_fixed = (const uint8_t*)read_array(f); w.speedbias_fixed = (const uint8_t*)read_array(f);
  w.landmark_fixed = (const uint8_t*)read_array(f); w.intrinsics = (const double*)read_array(f);
  w.obs_pose = (const int32_t*)read_array(f); w.obs_landmark = (const int32_t*)read_array(f);
  w.obs_extrinsics = (const int32_t*)read_array(f); w.obs_camera = (const int32_t*)read_array(f);
  w.obs_measurement = (const double*)read_array(f); w.obs_information = (const double*)read_array(f);
  w.imu_pose0 = (const int32_t*)read_array(f); w.imu_speedbias0 = (const int32_t*)read_array(f);
  w.imu_pose1 = (const int32_t*)read_array(f); w.imu_speedbias1 = (const int32_t*)read_array(f);
  w.imu_t0_ns = (const int64_t*)read_array(f); w.imu_t1_ns = (const int64_t*)read_array(f);
  w.imu_meas_offset = (const int32_t*)read_array(f); w.imu_meas_t_ns = (const int64_t*)read_array(f);
  w.imu_meas_gyro = (const double*)read_array(f); w.imu_meas_accel = (const double*)read_array(f);
  w.pose_prior_block = (const int32_t*)read_array(f); w.pose_prior_measurement = (const double*)read_array(f);
  w.pose_prior_information = (const double*)read_array(f);
  w.speedbias_prior_block = (const int32_t*)read_array(f); w.speedbias_prior_measurement = (const double*)read_array(f);
  w.speedbias_prior_information = (const double*)read_array(f);
  w.marg_block_kind = (const int32_t*)read_array(f); w.marg_block_index = (const int32_t*)read_array(f);
  w.marg_linearization_points = (const double*)read_array(f); w.marg_J = (const double*)read_array(f);
  w.marg_e0 = (const double*)read_array(f);
  fclose(f);
  svin_ba_default_options(&opt);
  rc = svin_ba_create(0, &ctx);
  if (rc != SVIN_OK) {
    fprintf(stderr, "svin_ba_create: %d %s\n", rc, svin_last_error());
    return rc == SVIN_ERR_NO_DEVICE ? 77 : 3;
  }
  quality = (double*)calloc((size_t)(w.num_landmarks > 0 ? w.num_landmarks : 1), sizeof(double));
  qptr[0] = quality;
  rc = svin_ba_optimize(ctx, &w, 1, &opt, &summary, qptr);
  if (rc != SVIN_OK) {
    fprintf(stderr, "svin_ba_optimize: %d %s\n", rc, svin_last_error());
    return 3;
  }
  o = fopen(out_path, "wb");
  if (!o) return 2;
  fwrite(&summary, sizeof summary, 1, o);
  fwrite(w.pose_blocks, sizeof(double), (size_t)7 * (size_t)w.num_pose_blocks, o);
  fwrite(w.speedbias, sizeof(double), (size_t)9 * (size_t)w.num_speedbias, o);
  fwrite(w.landmarks, sizeof(double), (size_t)4 * (size_t)w.num_landmarks, o);
  fwrite(quality, sizeof(double), (size_t)w.num_landmarks, o);
  fclose(o);
  svin_ba_destroy(ctx);
  printf("iterations %d cost %.6f -> %.6f\n", summary.iterations, summary.initial_cost, summary.final_cost);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && strcmp(argv[1], "layout") == 0) {
    layout();
    return 0;
  }
  if (argc >= 4 && strcmp(argv[1], "solve") == 0) return solve(argv[2], argv[3]);
  fprintf(stderr, "usage: %s layout | solve <in.bin> <out.bin>\n", argv[0]);
  return 64;
}
