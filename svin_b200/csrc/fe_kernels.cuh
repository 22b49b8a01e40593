// Device data model + launchers of the front-end kernels (definitions in fe_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/svin_b200.h"

namespace svin {

struct FeBatch {
  int W, H, pitch;
  int max_kp, border, cand_cap;
  int abs_threshold;
  double uniformity_radius;
  const uint8_t* images;     // [n][H][pitch]
  const double* intrinsics;  // [n][8]
  const double* extraction_dir;  // [n][3]
  int* scores;               // [n][H][W]
  unsigned* cand_count;      // [n]
  unsigned long long* cand_keys;  // [n][cand_cap]
  int* kept_xy;              // [n][max_kp][2]
  int* kept_score;           // [n][max_kp]
  int* kept_count;           // [n]
  SvinKeypoint* keypoints;   // [n][max_kp]
  uint8_t* descriptors;      // [n][max_kp][48]
  const int8_t *pat_dx, *pat_dy;  // [1024][60]
  const int* pat_half;       // [60]
  const int *pair_i, *pair_j;  // [384]
};

struct MatchDesc {
  int type, nA, nB;
  int a0, b0;  // offsets of this problem's keypoints in the concatenated A / B arrays
  int W, H;
  float threshold;
  double pose_uncertainty;
  double intrA[8], intrB[8], T_CbW[7], T_CaCb[7];
};

struct MatchBatch {
  const MatchDesc* desc;
  const uint8_t *descA, *descB;      // [sum nA][48], [sum nB][48]
  const uint8_t *skipA_in, *skipB;   // [sum nA], [sum nB]
  const SvinKeypoint *kpA, *kpB;
  const double* landmarksA;          // [sum nA][4]
  uint8_t* skipA;                    // effective
  double *proj, *cov, *sigA, *sigB, *rayA, *rayB;
  int* best_idx;
  float* best_dist;
  int* match_of_B;
  float* match_dist;
};

void fe_launch_detect(const FeBatch& f, int n_images, bool use_tma, size_t occ_bytes, cudaStream_t st,
                      cudaEvent_t* ev);
cudaError_t fe_configure(size_t occ_bytes);
void fe_launch_match(const MatchBatch& mb, int n_problems, int max_nA, int max_nAB, cudaStream_t st, cudaEvent_t* ev);

}  // namespace svin
