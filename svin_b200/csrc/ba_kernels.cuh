// Launchers of the BA kernels (definitions in ba_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "ba_types.cuh"

namespace svin {
// compact: Batch::fused - r and Jl only (64 B per observation); the pose Jacobian is rebuilt where it is consumed
void launch_linearize(const Batch& b, int which, int raw, cudaStream_t st, bool compact = false);
void launch_dense_eval(const Batch& b, int which, int raw, const double* const* dump, cudaStream_t st);
struct SchurStreams {  // auxiliary streams + events for the concurrent chunk kernels of one slot
  int n;
  cudaStream_t aux[5];
  cudaEvent_t fork, join[5];
};
int launch_schur(const Batch& b, const SvinBaOptions& opt, cudaStream_t st, const SchurStreams* par);  // -> kernels launched
size_t dense_solve_smem_bytes(int n_max);
cudaError_t configure_dense_solve(int smem_bytes);
size_t dense_gram_smem_bytes(int n_max);
cudaError_t configure_dense_gram(int smem_bytes);
void launch_dense_gram(const Batch& b, int which, int n_max, cudaStream_t st);  // after launch_dense_eval(which)
void launch_dense_solve(const Batch& b, const SvinBaOptions& opt, int smem_bytes, int n_max, cudaStream_t st);
void launch_backsub(const Batch& b, cudaStream_t st);
void launch_step_dense(const Batch& b, const SvinBaOptions& opt, cudaStream_t st);
void launch_step_lm(const Batch& b, cudaStream_t st);
void launch_decide(const Batch& b, const SvinBaOptions& opt, cudaStream_t st);
void launch_init(const Batch& b, const SvinBaOptions& opt, cudaStream_t st);
void launch_count_active(const Batch& b, int* out, cudaStream_t st);
void launch_quality(const Batch& b, cudaStream_t st);
size_t schur_mma_smem_bytes();
int schur_mma_max_chunk(int runs);
int schur_lr_max_chunk(int runs, int wpc);  // landmarks per run-parallel (k_schur_lr<wpc>) chunk, 0 = not eligible
cudaError_t configure_schur();
int schur_chunk_class(int count);  // which k_schur_mma<G> handles a chunk of `count` landmarks
void launch_fold(const Batch& b, int stage, cudaStream_t st);
void launch_pack_obs(const Batch& b, const RawObs& raw, cudaStream_t st);  // raw caller arrays -> internal SoA planes
void launch_obs_poff(const Batch& b, cudaStream_t st);  // obs_poff[o] = pose_off[obs_pose[o]], once per upload
void launch_gmax_pack(const Batch& b, int unpack, cudaStream_t st);
struct MargArgs {
  int w;                // window
  int n, nk, nm;        // dense dim, kept, marginalised
  double* H;            // [n][n] full symmetric accumulation
  double* b;            // [n]
  const int* keep_idx;  // [nk]
  const int* marg_idx;  // [nm]
  const int* prior_map; // [prior_dim] -> dense index
  int prior_dim;
  const double *prior_H, *prior_b;
  // scratch (global)
  double *A, *U, *ev;   // eigen-solver work: max(n, nm)^2 each
  double *Vp, *Wm, *WV, *pvec;
  int* players;
  // outputs
  double *Hk, *bk, *J, *e0;
};

void launch_marg(const Batch& b, const MargArgs& m, int num_landmarks, cudaStream_t st);
}  // namespace svin
