// Upload-path planning shared by the host checker (order_window / svin_ba_plan) and the device planner (ba_plan.cu):
// how the landmarks of one observation pattern are cut into Schur chunks.  One definition, so that the two cannot drift.
#pragma once
#include <cuda_runtime.h>

#include "ba_types.cuh"

namespace svin {

constexpr int kSchurLrBelow = 16;  // fewer landmarks of one pattern than this -> run-parallel chunks
constexpr int kPlanMaxRuns = 64;   // pattern grouping needs <= 64 pose blocks per window
constexpr int kPlanMaxLandmarks = 8192;  // device planner: (key, index) pairs of one window sorted in shared memory

struct ChunkCaps {  // limits of the chunk kernels for a pattern of `runs` pose runs (ba_kernels.cu)
  int cap;          // k_schur_mma: lane = landmark chunk
  int lr1, lr2, lr4;  // k_schur_lr<1|2|4>: landmarks per run-parallel chunk, 0 = not eligible
};

// Next chunk of a pattern with `rem` landmarks left: its size and lane-mapping class
// (0..2 k_schur_mma<1|2|4>, 3/7/8 k_schur_lr<1|2|4>, 4..6 k_schur_wr<2..4>).
// use_lr: SVIN_SCHUR_LR A/B knob (0 lane = landmark only, 1 + k_schur_lr, 2 + k_schur_wr, 3 + multi-warp k_schur_lr).
__host__ __device__ inline void plan_next_chunk(int rem, int runs, const ChunkCaps& cc, int use_lr, int& c, int& kind) {
  if (use_lr >= 2 && runs >= 2 && runs <= 4 && rem >= kSchurLrBelow) {
    c = rem < 32 ? rem : 32;
    kind = 2 + runs;  // warp per run (k_schur_wr<runs>)
  } else if (cc.lr1 >= 1 && (use_lr >= 2 || rem < kSchurLrBelow || cc.cap < kSchurLrBelow)) {
    // run-parallel: the narrowest CTA that takes what is left, else equal parts of the widest
    if (rem <= cc.lr1 || cc.lr2 < 1) {
      c = cc.lr1 < rem ? cc.lr1 : rem;
      kind = 3;
    } else if (rem <= cc.lr2) {
      c = rem;
      kind = 7;
    } else {
      const int parts = (rem + cc.lr4 - 1) / cc.lr4;
      c = (rem + parts - 1) / parts;
      kind = c <= cc.lr1 ? 3 : (c <= cc.lr2 ? 7 : 8);
    }
  } else {
    c = cc.cap < rem ? cc.cap : rem;
    kind = c <= 8 ? 2 : (c <= 16 ? 1 : 0);  // schur_chunk_class
  }
  if (c < 1) c = 1;  // never reached with the caps of ba_kernels.cu; keeps the callers' loops finite
}

struct PlanCapsTable {
  unsigned char cap[kPlanMaxRuns + 1], lr1[kPlanMaxRuns + 1], lr2[kPlanMaxRuns + 1], lr4[kPlanMaxRuns + 1];
  int use_lr;
};
int schur_use_lr();                      // SVIN_SCHUR_LR (default 3)
ChunkCaps chunk_caps_for(int runs);      // host: from schur_mma_max_chunk / schur_lr_max_chunk
const PlanCapsTable& plan_caps_table();  // the same for runs 0..64, passed to the device planner by value

// Device planner (ba_plan.cu): one CTA per window orders the landmarks by observation pattern, cuts the patterns into
// chunks and writes every table order_window + fill_window used to produce on the host.
struct PlanArgs {
  int B;
  int packed;  // 0: rpose / rcam / rlm, 1: rpec (pose | ext << 10 | cam << 20) + rlm, 2: rpec = landmark | pose << 18 | ext << 24 | cam << 30
  const WinDesc* win;
  const int *rlm, *rpec, *rpose, *rcam;  // caller's observation arrays (window by window), already on the device
  const double* lm_raw;                  // [NL][4] landmarks in caller order
  const unsigned char* lmfix_raw;        // [NL] in caller order
  const int* poff;                       // [NPB] reduced-system offset of a pose block, -1 = fixed
  // outputs (batch-global arrays)
  double* lm_init;
  unsigned char* lm_fixed;
  int *lm_win, *linv, *lm_perm;
  int *lmof, *lmos, *lmoc, *rord;
  int *sw_win, *sw_lm_begin, *sw_count, *sw_nruns, *sw_run_first, *sw_kind;  // chunk id = lm_begin + local chunk
  int *run_off, *run_k0m;                                                    // run id = obs_begin + local run
  int* sw_list;
  int* win_class_count;  // [B][kSchurClasses]
  int* win_class_base;   // [B][kSchurClasses] position in sw_list
  int* win_nchunks;      // [B]
  int* class_total;      // [kSchurClasses] (+ [kSchurClasses] = total chunks) -> host
  int* scratch[5];       // NL + 2 B + 8 ints each
  PlanCapsTable caps;
};
cudaError_t launch_plan(const PlanArgs& a, int max_landmarks, cudaStream_t st);

}  // namespace svin
