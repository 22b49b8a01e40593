// Host side of the BA engine: flattens SvinBaWindow batches into one device arena, drives the
// trust-region slots on a CUDA stream and implements the svin_ba_* C ABI (include/svin_b200.h).
//
// Reference seam: okvis::Estimator::optimize -> okvis::ceres::Map::solve()
// (okvis_ros/okvis/okvis_ceres/src/Estimator.cpp:876-929, include/okvis/ceres/Map.hpp:347).
#include <cuda_runtime.h>

#include <chrono>
#include <cstdlib>
#include <mutex>
#include <functional>
#include <condition_variable>

#include <algorithm>
#include <atomic>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include <dlfcn.h>

#include "ba_kernels.cuh"
#include "ba_plan.cuh"
#include "common.hpp"
#include "host_pool.hpp"

namespace svin {
void launch_gather_state(const Batch& b, const int* pose_win, const int* sb_win, double* pose_out, double* sb_out,
                         double* lm_out, cudaStream_t st);
// x <- uploaded initial values (both state buffers)
static void launch_reset_state(const Batch& b, cudaStream_t st) {
  for (int k = 0; k < 2; ++k) {
    if (b.NPB) cudaMemcpyAsync(b.pose[k], b.pose_init, 56 * (size_t)b.NPB, cudaMemcpyDeviceToDevice, st);
    if (b.NSB) cudaMemcpyAsync(b.sb[k], b.sb_init, 72 * (size_t)b.NSB, cudaMemcpyDeviceToDevice, st);
    if (b.NL) cudaMemcpyAsync(b.lm[k], b.lm_init, 32 * (size_t)b.NL, cudaMemcpyDeviceToDevice, st);
  }
}
}  // namespace svin

using namespace svin;

namespace {

// Eigen::LLT (unblocked, lower) incl. its early exit on a non-positive pivot, then U = L^T.
// (ReprojectionError.hpp impl:66-72, PoseError.cpp:70-76, ...)
void llt_sqrt_information(const double* info, double* U, int n) {
  std::vector<double> L(info, info + (size_t)n * n);
  for (int k = 0; k < n; ++k) {
    double x = L[(size_t)k * n + k];
    for (int j = 0; j < k; ++j) x -= L[(size_t)k * n + j] * L[(size_t)k * n + j];
    if (x <= 0.0) break;
    x = std::sqrt(x);
    L[(size_t)k * n + k] = x;
    for (int i = k + 1; i < n; ++i) {
      double s = L[(size_t)i * n + k];
      for (int j = 0; j < k; ++j) s -= L[(size_t)i * n + j] * L[(size_t)k * n + j];
      L[(size_t)i * n + k] = s / x;
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) U[(size_t)i * n + j] = (j >= i) ? L[(size_t)j * n + i] : 0.0;
}

// (the 2x2 case for the observations' information matrices runs on the device: k_pack_obs)

// Internal ordering of one window's landmarks and observations.
//  * observations sorted by (landmark, pose block, camera)
//  * when `group` is set, landmarks are additionally grouped by their exact (pose, camera) observation
//    pattern so that k_schur_warp can hand each warp <= 32 landmarks with identical pose runs.
struct WindowOrder {
  std::vector<int> lm_perm;   // internal landmark k -> caller landmark
  std::vector<int> obs_order; // internal observation k -> caller observation
  std::vector<int> lm_count;  // observations of internal landmark k
  std::vector<int> lm_first, lm_stride;  // window-local position of its first observation, stride between them
  std::vector<int> chunk_begin, chunk_count;  // Schur warp chunks (window-local internal landmark indices)
  std::vector<int> chunk_kind;                // lane mapping: schur_chunk_class() or 3 = run-parallel (k_schur_lr)
  std::vector<int> chunk_run_first, chunk_nruns;  // pose runs of the chunk's pattern -> run_pose / run_k0m
  std::vector<int> run_pose, run_k0m;             // window-local pose block, (first obs k << 8) | obs count
};


void order_window(const SvinBaWindow& w, bool group, WindowOrder& out) {
  const int N = w.num_obs, L = w.num_landmarks;
  std::vector<int> ord(N);
  std::iota(ord.begin(), ord.end(), 0);
  // already in (landmark, pose, camera) order?  one packed key per observation, no data-dependent branches
  bool sorted = true;
  {
    uint64_t prev = 0;
    unsigned bad = 0;
    for (int o = 0; o < N; ++o) {
      const uint64_t k = ((uint64_t)(uint32_t)w.obs_landmark[o] << 32) | ((uint64_t)(uint32_t)w.obs_pose[o] << 8) |
                         (uint64_t)(w.obs_camera[o] & 255);
      bad |= (unsigned)(k < prev);
      prev = k;
    }
    sorted = bad == 0 && w.num_cameras <= 256 && w.num_pose_blocks < (1 << 24);
  }
  if (!sorted)
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b2) {
      if (w.obs_landmark[a] != w.obs_landmark[b2]) return w.obs_landmark[a] < w.obs_landmark[b2];
      if (w.obs_pose[a] != w.obs_pose[b2]) return w.obs_pose[a] < w.obs_pose[b2];
      return w.obs_camera[a] < w.obs_camera[b2];
    });
  std::vector<int> start(L + 1, 0);
  for (int o = 0; o < N; ++o) start[w.obs_landmark[o] + 1]++;
  for (int k = 0; k < L; ++k) start[k + 1] += start[k];
  out.lm_perm.resize(L);
  std::iota(out.lm_perm.begin(), out.lm_perm.end(), 0);
  auto fixed = [&](int l) { return (w.landmark_fixed && w.landmark_fixed[l]) ? 1 : 0; };
  auto same_pattern = [&](int la, int lb) {
    const int na = start[la + 1] - start[la];
    if (na != start[lb + 1] - start[lb] || fixed(la) != fixed(lb)) return false;
    for (int k = 0; k < na; ++k) {
      const int oa = ord[start[la] + k], ob = ord[start[lb] + k];
      if (w.obs_pose[oa] != w.obs_pose[ob] || w.obs_camera[oa] != w.obs_camera[ob]) return false;
    }
    return true;
  };
  if (group) {
    std::vector<uint64_t> key(L);
    for (int l = 0; l < L; ++l) {
      uint64_t h = 1469598103934665603ull ^ (uint64_t)fixed(l);
      for (int k = start[l]; k < start[l + 1]; ++k) {
        const int o = ord[k];
        h = (h ^ (uint64_t)(w.obs_pose[o] * 4 + w.obs_camera[o] + 1)) * 1099511628211ull;
      }
      // primary: number of observations, then first pose (locality), then the pattern hash
      const uint64_t nobs = (uint64_t)(start[l + 1] - start[l]);
      const uint64_t first = nobs ? (uint64_t)w.obs_pose[ord[start[l]]] : 0;
      key[l] = (nobs << 54) | ((first & 0x3ff) << 44) | (h >> 20);
    }
    // Landmarks ordered by key, ties in caller order (= a stable sort by key).  The distinct keys are few (one per
    // observation pattern, ~70 for BASELINE configs[1]): bucket the landmarks through a small open-addressing table,
    // sort the buckets, and emit them - O(L + P log P) instead of an O(L log L) sort with random key lookups.
    int tsize = 64;
    while (tsize < 4 * std::min(L, 4096) && tsize < (1 << 15)) tsize <<= 1;
    std::vector<int> slot_of(tsize, -1);        // table position -> bucket id
    std::vector<uint64_t> bkey;                 // bucket id -> key
    std::vector<int> bcount, lm_bucket(L);
    bool overflow = false;
    for (int l = 0; l < L && !overflow; ++l) {
      const uint64_t k = key[l];
      unsigned pos = (unsigned)((k * 0x9E3779B97F4A7C15ull) >> 40) & (unsigned)(tsize - 1);
      int probes = 0;
      while (slot_of[pos] >= 0 && bkey[slot_of[pos]] != k && probes < tsize) {
        pos = (pos + 1) & (unsigned)(tsize - 1);
        ++probes;
      }
      if (slot_of[pos] < 0) {
        if ((int)bkey.size() * 2 > tsize) {
          overflow = true;   // more distinct patterns than the table was sized for: plain sort below
          break;
        }
        slot_of[pos] = (int)bkey.size();
        bkey.push_back(k);
        bcount.push_back(0);
      }
      lm_bucket[l] = slot_of[pos];
      bcount[slot_of[pos]]++;
    }
    if (overflow) {
      std::stable_sort(out.lm_perm.begin(), out.lm_perm.end(), [&](int a, int b2) { return key[a] < key[b2]; });
    } else {
      const int P = (int)bkey.size();
      std::vector<int> border(P);
      std::iota(border.begin(), border.end(), 0);
      std::sort(border.begin(), border.end(), [&](int a, int b2) { return bkey[a] < bkey[b2]; });
      std::vector<int> bstart(P);
      int acc = 0;
      for (int r = 0; r < P; ++r) {
        bstart[border[r]] = acc;
        acc += bcount[border[r]];
      }
      for (int l = 0; l < L; ++l) out.lm_perm[bstart[lm_bucket[l]]++] = l;
    }
  }
  out.lm_count.resize(L);
  for (int k = 0; k < L; ++k) out.lm_count[k] = start[out.lm_perm[k] + 1] - start[out.lm_perm[k]];
  out.chunk_begin.clear();
  out.chunk_count.clear();
  out.chunk_run_first.clear();
  out.chunk_nruns.clear();
  out.run_pose.clear();
  out.run_k0m.clear();
  out.chunk_kind.clear();
  if (group) {
    // A/B knob: SVIN_SCHUR_LR=0 keeps the lane = landmark mappings for every chunk, 1 adds k_schur_lr only
    // 2 adds k_schur_wr, 3 (default) also the 2- and 4-warp k_schur_lr
    const int use_lr = schur_use_lr();
    int k = 0;
    while (k < L) {
      // pose runs of this pattern (shared by all of its chunks)
      const int lk = out.lm_perm[k];
      int runs = 0, prev = -1;
      const int run_first = (int)out.run_pose.size();
      for (int q = start[lk]; q < start[lk + 1]; ++q) {
        const int pz = w.obs_pose[ord[q]];
        if (pz != prev) {
          ++runs;
          out.run_pose.push_back(pz);
          out.run_k0m.push_back(((q - start[lk]) << 8) | 1);
        } else {
          out.run_k0m.back() += 1;
        }
        prev = pz;
      }
      int e_all = k + 1;
      while (e_all < L && same_pattern(out.lm_perm[k], out.lm_perm[e_all])) ++e_all;
      // The pattern's landmarks go to lane = landmark chunks (bounded by the operand tiles of k_schur_mma) while at
      // least kSchurLrBelow of them are left (one warp per run for 2..4 runs, k_schur_wr), the rest to run-parallel
      // chunks of <= 32 / runs landmarks (k_schur_lr); patterns neither kernel takes keep the lane = landmark mappings.
      const ChunkCaps cc = chunk_caps_for(runs);
      while (k < e_all) {
        int c, kind;
        plan_next_chunk(e_all - k, runs, cc, use_lr, c, kind);   // shared with the device planner (ba_plan.cuh)
        out.chunk_run_first.push_back(run_first);
        out.chunk_nruns.push_back(runs);
        out.chunk_begin.push_back(k);
        out.chunk_count.push_back(c);
        out.chunk_kind.push_back(kind);
        k += c;
      }
    }
  }
  out.obs_order.resize(N);
  out.lm_first.resize(L);
  out.lm_stride.resize(L);
  int pos = 0;
  if (group) {
    // pattern-major inside a chunk: position (k, lane) -> pos + k * count + lane
    for (size_t ch = 0; ch < out.chunk_begin.size(); ++ch) {
      const int k0 = out.chunk_begin[ch], cntc = out.chunk_count[ch], m = out.lm_count[k0];
      for (int lane = 0; lane < cntc; ++lane) {
        const int l = out.lm_perm[k0 + lane];
        out.lm_first[k0 + lane] = pos + lane;
        out.lm_stride[k0 + lane] = cntc;
        for (int q = 0; q < m; ++q) out.obs_order[pos + q * cntc + lane] = ord[start[l] + q];
      }
      pos += m * cntc;
    }
  } else {
    for (int k = 0; k < L; ++k) {
      const int l = out.lm_perm[k];
      out.lm_first[k] = pos;
      out.lm_stride[k] = 1;
      for (int q = start[l]; q < start[l + 1]; ++q) out.obs_order[pos++] = ord[q];
    }
  }
}

// ---- NCCL, resolved at run time so that the library has no link-time dependency on it -------------------
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ struct NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
struct NcclId {
  char internal[128];
};
constexpr int kNcclFloat64 = 8, kNcclSum = 0;
NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return nullptr;
  api.GetUniqueId = (int (*)(void*))dlsym(api.lib, "ncclGetUniqueId");
  api.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(api.lib, "ncclCommInitRank");
  api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
  api.CommDestroy = (int (*)(void*))dlsym(api.lib, "ncclCommDestroy");
  api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce) {
    api.lib = nullptr;
    return nullptr;
  }
  return &api;
}

struct Region {
  size_t bytes = 0;
  size_t add(size_t b) {
    const size_t off = bytes;
    bytes += (b + 255) & ~(size_t)255;
    return off;
  }
};


// ---- in-process communicator: `world` contexts of ONE process on one device (svin_ba_comm_init_local) -----------------
// The exchange steps of the sharded mode with every rank's context in this process: each rank's host thread enqueues,
// on its own stream, "sum the peers' buffers in rank order into my scratch, then copy it back" with two host barriers
// and cross-stream events in between.  All ranks sum in the same order, so the result is bit-identical on every rank
// (as NCCL guarantees for its all-reduce).  Used by the single-device parity test of BASELINE configs[3]; the
// multi-process path is NCCL (svin_ba_comm_init).
struct LocalGroup {
  int world = 0;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  long long generation = 0;
  bool failed = false;
  std::vector<double*> bufs;
  std::vector<cudaEvent_t> ready, read;
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    const long long g = generation;
    if (++arrived == world) {
      arrived = 0;
      ++generation;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return generation != g; });
    }
  }
};
__global__ void k_sum_peers(double* dst, const double* const* src, int world, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = src[0][i];
  for (int r = 1; r < world; ++r) s += src[r][i];
  dst[i] = s;
}

}  // namespace

struct svin_ba_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;  // dense-term evaluation runs here, concurrently with k_linearize
  cudaStream_t up = nullptr;    // svin_ba_upload's copies and kernels (high priority)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_gram = nullptr;
  SchurStreams schur_par{};               // concurrent Schur chunk kernels (SVIN_SCHUR_STREAMS=0 disables)
  cudaGraphExec_t graph_exec = nullptr;  // first solve pass of the current upload (SVIN_BA_GRAPH)
  bool graph_valid = false, graph_opt_known = false;
  SvinBaOptions graph_opt{};
  long long graph_launches = 0;
  bool gram_pending = false;  // a k_dense_gram launch on the side stream the next reduced-system solve must wait for
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  Batch b{};
  bool uploaded = false;
  // arenas
  void* d_in = nullptr;       // uploaded inputs
  size_t d_in_cap = 0;
  void* h_in = nullptr;       // pinned staging mirror of d_in
  size_t h_in_cap = 0;
  void* d_work = nullptr;     // scratch
  size_t d_work_cap = 0;
  void* d_out = nullptr;      // results gathered for download
  size_t d_out_cap = 0;
  void* h_out = nullptr;      // pinned
  size_t h_out_cap = 0;
  size_t in_bytes = 0, out_bytes = 0;
  // host metadata
  std::vector<WinDesc> h_win;
  std::vector<int> obs_perm;  // sorted obs position -> caller's obs index (window-local)
  std::vector<int> lm_perm;   // internal landmark position (global) -> caller's landmark index (window-local)
  // device planner: the landmark permutation is read back into pinned memory (after 16 ints of class totals), the
  // observation permutation only when svin_ba_evaluate asks for it
  bool device_plan = false, obs_perm_valid = true;
  int* h_lm_perm = nullptr;
  size_t h_lm_perm_cap = 0;
  char* h_obs = nullptr;       // pinned: compact observation records written by the validation pass
  size_t h_obs_cap = 0;
  const int* lm_perm_p = nullptr;
  const int* d_rord = nullptr;
  size_t n_obs_total = 0, n_sw_cap = 0;
  int n_max = 0, smem_bytes = 0;
  int* d_active = nullptr;
  int* h_active = nullptr;
  // pointers into arenas kept for reset/download
  WinState* d_ws_init = nullptr;
  ImuCache* d_imu_cache_init = nullptr;
  int *d_pose_win = nullptr, *d_sb_win = nullptr;
  double *d_pose_out = nullptr, *d_sb_out = nullptr, *d_lm_out = nullptr;
  size_t out_off_pose = 0, out_off_sb = 0, out_off_lm = 0, out_off_q = 0, out_off_ws = 0;
  size_t clear_bytes = 0;
  void* d_clear = nullptr;
  bool quality_valid = false;
  bool solved = false;
  int solves_since_upload = 0;
  SvinBaTimings tm{};
  HostPool* pool = nullptr;
  // sharded mode
  void* nccl_comm = nullptr;
  std::shared_ptr<LocalGroup> local_comm;  // in-process communicator (svin_ba_comm_init_local)
  double* local_tmp = nullptr;             // its reduction scratch
  const double** local_srcs = nullptr;     // device array of the peers' buffer pointers
  size_t local_tmp_cap = 0;
  int comm_rank = 0, comm_world = 1;
  bool sharded() const { return nccl_comm != nullptr || local_comm != nullptr; }
  // optional per-kernel profiling
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
  std::vector<int> prof_family;  // family of the launch bracketed by events (2i, 2i+1)
  SvinBaKernelTimes ktimes{};
};

namespace {

double wall_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int ensure(void** p, size_t* cap, size_t need, bool pinned) {
  if (*cap >= need && *p) return SVIN_OK;
  if (*p) {
    if (pinned)
      cudaFreeHost(*p);
    else
      cudaFree(*p);
    *p = nullptr;
    *cap = 0;
  }
  const size_t want = need + need / 4 + 4096;
  cudaError_t e = pinned ? cudaMallocHost(p, want) : cudaMalloc(p, want);
  if (e != cudaSuccess) {
    set_error(std::string("allocation of ") + std::to_string(want) + " bytes failed: " + cudaGetErrorString(e));
    cudaGetLastError();
    return SVIN_ERR_OUT_OF_MEMORY;
  }
  *cap = want;
  return SVIN_OK;
}

// flags (optional): bit 0 = observations not in (landmark, pose, camera) order, bit 1 = an observation's extrinsics block is
// free, bit 2 = a measurement coordinate is not exactly a float
// sink (optional): the same pass writes the compact upload records - one index word and two floats per observation - so
// that the caller's index and measurement arrays are read once, not once to validate and once to pack
struct ObsSink {
  int* word = nullptr;   // landmark | pose << 18 | ext << 24 | cam << 30
  float* z = nullptr;    // [2] per observation
};
int validate(const SvinBaWindow& w, int idx, int* flags = nullptr, const ObsSink* sink = nullptr) {
  auto bad = [&](const char* what) {
    set_error("window " + std::to_string(idx) + ": " + what);
    return SVIN_ERR_INVALID_ARGUMENT;
  };
  if (w.num_pose_blocks < 0 || w.num_speedbias < 0 || w.num_landmarks < 0 || w.num_obs < 0 || w.num_cameras < 0 ||
      w.num_imu < 0 || w.num_pose_priors < 0 || w.num_speedbias_priors < 0 || w.num_relative_pose < 0 ||
      w.num_sonar < 0 || w.num_depth < 0 || w.marg_num_blocks < 0 || w.marg_dim < 0)
    return bad("negative count");
  // every array must be there whenever its count says it is read (a C caller's malformed window must come back as
  // SVIN_ERR_INVALID_ARGUMENT, not as a segfault in the packers)
  if (w.num_imu && (!w.imu_pose0 || !w.imu_pose1 || !w.imu_speedbias0 || !w.imu_speedbias1 || !w.imu_t0_ns ||
                    !w.imu_t1_ns || !w.imu_meas_offset || !w.imu_meas_t_ns || !w.imu_meas_gyro || !w.imu_meas_accel))
    return bad("IMU term arrays are NULL");
  if (w.num_pose_priors && (!w.pose_prior_block || !w.pose_prior_measurement || !w.pose_prior_information))
    return bad("pose prior arrays are NULL");
  if (w.num_speedbias_priors &&
      (!w.speedbias_prior_block || !w.speedbias_prior_measurement || !w.speedbias_prior_information))
    return bad("speedbias prior arrays are NULL");
  if (w.num_relative_pose && (!w.relative_pose_block0 || !w.relative_pose_block1 || !w.relative_pose_information))
    return bad("relative pose arrays are NULL");
  if (w.num_sonar && (!w.sonar_pose || !w.sonar_range || !w.sonar_heading || !w.sonar_information ||
                      !w.sonar_landmark_mean || !w.sonar_T_SSo))
    return bad("sonar arrays are NULL");
  if (w.num_depth && (!w.depth_pose || !w.depth_measurement || !w.depth_first || !w.depth_information))
    return bad("depth arrays are NULL");
  if (w.marg_num_blocks && (!w.marg_block_kind || !w.marg_block_index || !w.marg_linearization_points))
    return bad("marginalisation prior block arrays are NULL");
  if (w.marg_dim && (!w.marg_J || !w.marg_e0)) return bad("marg_J / marg_e0 is NULL");
  if (w.marg_dim && !w.marg_num_blocks) return bad("marg_dim > 0 without prior blocks");
  if (w.num_imu && w.imu_meas_offset[0] != 0) return bad("imu_meas_offset[0] must be 0");
  if (w.num_pose_blocks && (!w.pose_blocks || !w.pose_fixed)) return bad("pose_blocks/pose_fixed is NULL");
  if (w.num_speedbias && (!w.speedbias || !w.speedbias_fixed)) return bad("speedbias/speedbias_fixed is NULL");
  if (w.num_landmarks && !w.landmarks) return bad("landmarks is NULL");
  if (w.num_obs && (!w.obs_pose || !w.obs_landmark || !w.obs_extrinsics || !w.obs_camera || !w.obs_measurement ||
                    !w.obs_information || !w.intrinsics))
    return bad("observation arrays / intrinsics are NULL");
  {
    // one pass: ranges (any failure is diagnosed by the slow loop below), sortedness, free extrinsics
    const unsigned np = (unsigned)w.num_pose_blocks, nl = (unsigned)w.num_landmarks, nc = (unsigned)w.num_cameras;
    unsigned oob = 0, unsorted = 0, ext_free = 0, inexact = 0;
    uint64_t prev = 0;
    int* const sw = sink ? sink->word : nullptr;
    float* const sz = sink ? sink->z : nullptr;
    for (int i = 0; i < w.num_obs; ++i) {
      const unsigned p = (unsigned)w.obs_pose[i], e = (unsigned)w.obs_extrinsics[i], l = (unsigned)w.obs_landmark[i],
                     c = (unsigned)w.obs_camera[i];
      const unsigned bad_i = (unsigned)(p >= np) | (unsigned)(e >= np) | (unsigned)(l >= nl) | (unsigned)(c >= nc);
      oob |= bad_i;
      const uint64_t k = ((uint64_t)l << 32) | ((uint64_t)(p & 0xffffffu) << 8) | (uint64_t)(c & 255u);
      unsorted |= (unsigned)(k < prev);
      prev = k;
      if (!bad_i) ext_free |= (unsigned)(w.pose_fixed[e] == 0);
      const double zx = w.obs_measurement[2 * (size_t)i], zy = w.obs_measurement[2 * (size_t)i + 1];
      inexact |= (unsigned)((double)(float)zx != zx) | (unsigned)((double)(float)zy != zy);
      if (sw) sw[i] = (int)(l | (p << 18) | (e << 24) | (c << 30));
      if (sz) {
        sz[2 * (size_t)i] = (float)zx;
        sz[2 * (size_t)i + 1] = (float)zy;
      }
    }
    if (oob)
      for (int i = 0; i < w.num_obs; ++i) {
        if (w.obs_pose[i] < 0 || w.obs_pose[i] >= w.num_pose_blocks) return bad("obs_pose out of range");
        if (w.obs_extrinsics[i] < 0 || w.obs_extrinsics[i] >= w.num_pose_blocks) return bad("obs_extrinsics out of range");
        if (w.obs_landmark[i] < 0 || w.obs_landmark[i] >= w.num_landmarks) return bad("obs_landmark out of range");
        if (w.obs_camera[i] < 0 || w.obs_camera[i] >= w.num_cameras) return bad("obs_camera out of range");
      }
    if (flags)
      *flags = ((unsorted || w.num_cameras > 256 || w.num_pose_blocks >= (1 << 24)) ? 1 : 0) | (ext_free ? 2 : 0) |
               (inexact ? 4 : 0);
  }
  for (int i = 0; i < w.num_imu; ++i) {
    if (w.imu_pose0[i] < 0 || w.imu_pose0[i] >= w.num_pose_blocks || w.imu_pose1[i] < 0 ||
        w.imu_pose1[i] >= w.num_pose_blocks || w.imu_speedbias0[i] < 0 || w.imu_speedbias0[i] >= w.num_speedbias ||
        w.imu_speedbias1[i] < 0 || w.imu_speedbias1[i] >= w.num_speedbias)
      return bad("IMU term block index out of range");
    const int a = w.imu_meas_offset[i], e = w.imu_meas_offset[i + 1];
    if (a < 0 || e < a) return bad("imu_meas_offset must be non-decreasing");
    if (e - a < 2) return bad("IMU term needs at least two measurements");
    if (w.imu_t1_ns[i] < w.imu_t0_ns[i]) return bad("IMU term with t1 < t0");
    for (int k = a + 1; k < e; ++k)
      if (w.imu_meas_t_ns[k] < w.imu_meas_t_ns[k - 1]) return bad("IMU measurement stamps must be ordered");
    if (w.imu_meas_t_ns[a] > w.imu_t0_ns[i] || w.imu_meas_t_ns[e - 1] < w.imu_t1_ns[i])
      return bad("IMU measurements do not cover [t0, t1] (ImuError.cpp:66-73)");
  }
  for (int i = 0; i < w.num_pose_priors; ++i)
    if (w.pose_prior_block[i] < 0 || w.pose_prior_block[i] >= w.num_pose_blocks) return bad("pose prior block");
  for (int i = 0; i < w.num_speedbias_priors; ++i)
    if (w.speedbias_prior_block[i] < 0 || w.speedbias_prior_block[i] >= w.num_speedbias) return bad("speedbias prior block");
  for (int i = 0; i < w.num_relative_pose; ++i)
    if (w.relative_pose_block0[i] < 0 || w.relative_pose_block0[i] >= w.num_pose_blocks ||
        w.relative_pose_block1[i] < 0 || w.relative_pose_block1[i] >= w.num_pose_blocks)
      return bad("relative pose block");
  for (int i = 0; i < w.num_sonar; ++i)
    if (w.sonar_pose[i] < 0 || w.sonar_pose[i] >= w.num_pose_blocks) return bad("sonar pose");
  for (int i = 0; i < w.num_depth; ++i)
    if (w.depth_pose[i] < 0 || w.depth_pose[i] >= w.num_pose_blocks) return bad("depth pose");
  if (w.marg_num_blocks > 64) return bad("marginalisation prior with more than 64 blocks is not supported");
  if (w.marg_dim > 512) return bad("marginalisation prior dimension > 512 is not supported");
  int md = 0;
  for (int i = 0; i < w.marg_num_blocks; ++i) {
    const int k = w.marg_block_kind[i], x = w.marg_block_index[i];
    if (k == SVIN_BLOCK_LANDMARK)
      return bad("landmark blocks inside the marginalisation prior are not supported (OKVIS marginalises them out, "
                 "Estimator.cpp:731-741)");
    if (k == SVIN_BLOCK_POSE) {
      if (x < 0 || x >= w.num_pose_blocks) return bad("marg pose block index");
      if (!w.pose_fixed[x]) md += 6;
    } else if (k == SVIN_BLOCK_SPEEDBIAS) {
      if (x < 0 || x >= w.num_speedbias) return bad("marg speedbias block index");
      if (!w.speedbias_fixed[x]) md += 9;
    } else {
      return bad("unknown marg block kind");
    }
  }
  if (md != w.marg_dim) return bad("marg_dim does not equal the sum of the minimal dimensions of its free blocks");
  return SVIN_OK;
}

}  // namespace

extern "C" {

void svin_ba_default_options(SvinBaOptions* o) {
  if (!o) return;
  o->max_num_iterations = 10;
  o->min_num_iterations = 3;
  o->time_limit_seconds = -1.0;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
  o->compute_landmark_quality = 1;
}

int svin_ba_create(int device, svin_ba_ctx** out) {
  if (!out) {
    set_error("svin_ba_create: out is NULL");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    set_error("no CUDA device available: the svin_b200 engine has no CPU fallback");
    return SVIN_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) {
    set_error("device index out of range");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(cudaSetDevice(device));
  svin_ba_ctx* c = new svin_ba_ctx();
  c->device = device;
  SVIN_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    // higher priority: its small grid must be dispatched ahead of the remaining k_linearize CTAs, otherwise the
    // work distributor only starts it in k_linearize's tail
    int lo = 0, hi = 0;
    SVIN_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    SVIN_CUDA(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi));
    SVIN_CUDA(cudaStreamCreateWithPriority(&c->up, cudaStreamNonBlocking, hi));
  }
  SVIN_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  SVIN_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  SVIN_CUDA(cudaEventCreateWithFlags(&c->ev_gram, cudaEventDisableTiming));
  c->schur_par.n = 5;
  SVIN_CUDA(cudaEventCreateWithFlags(&c->schur_par.fork, cudaEventDisableTiming));
  for (int k = 0; k < 5; ++k) {
    SVIN_CUDA(cudaStreamCreateWithFlags(&c->schur_par.aux[k], cudaStreamNonBlocking));
    SVIN_CUDA(cudaEventCreateWithFlags(&c->schur_par.join[k], cudaEventDisableTiming));
  }
  for (auto& ev : c->ev) SVIN_CUDA(cudaEventCreate(&ev));
  SVIN_CUDA(cudaMalloc(&c->d_active, sizeof(int)));
  SVIN_CUDA(cudaMallocHost(&c->h_active, sizeof(int)));
  // worker threads of the upload / scatter path; SVIN_HOST_THREADS lets several contexts (BaPipeline) share the cores
  int host_threads = std::min(32, (int)std::thread::hardware_concurrency());
  // one process per GPU (torchrun sets LOCAL_WORLD_SIZE): the ranks of a node share its cores - a full-size pool per
  // rank oversubscribed the host 8x at 8 ranks and made the end-to-end path host-bound (SCALE_r01: 0.25 efficiency)
  if (const char* lw = std::getenv("LOCAL_WORLD_SIZE"))
    host_threads = std::max(2, host_threads / std::max(1, std::atoi(lw)));
  if (const char* e = std::getenv("SVIN_HOST_THREADS")) host_threads = std::max(1, std::min(64, std::atoi(e)));
  c->pool = new HostPool(std::max(0, host_threads - 1));
  *out = c;
  return SVIN_OK;
}

void svin_ba_destroy(svin_ba_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->d_in);
  cudaFree(c->d_work);
  cudaFree(c->d_out);
  cudaFreeHost(c->h_in);
  cudaFreeHost(c->h_out);
  cudaFreeHost(c->h_lm_perm);
  cudaFreeHost(c->h_obs);
  cudaFree(c->d_active);
  cudaFreeHost(c->h_active);
  for (auto& ev : c->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->prof_events) cudaEventDestroy(ev);
  if (c->nccl_comm && nccl_api() && nccl_api()->CommDestroy) nccl_api()->CommDestroy(c->nccl_comm);
  if (c->local_tmp) cudaFree(c->local_tmp);
  if (c->local_srcs) cudaFree(c->local_srcs);
  c->local_comm.reset();
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->ev_gram) cudaEventDestroy(c->ev_gram);
  if (c->schur_par.fork) cudaEventDestroy(c->schur_par.fork);
  for (int k = 0; k < 5; ++k) {
    if (c->schur_par.join[k]) cudaEventDestroy(c->schur_par.join[k]);
    if (c->schur_par.aux[k]) cudaStreamDestroy(c->schur_par.aux[k]);
  }
  if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->up) cudaStreamDestroy(c->up);
  delete c->pool;
  delete c;
}

static bool graph_wanted();
static int graph_min_windows();
static int build_graph(svin_ba_ctx* c, const SvinBaOptions& opt);

int svin_ba_upload(svin_ba_ctx* c, const SvinBaWindow* wins, int32_t B) {
  if (!c || !wins || B <= 0) {
    set_error("svin_ba_upload: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  const double t_begin = wall_ms();
  std::vector<int> vflags;
  bool prepacked = false;
  {
    // per window on the pool; the message of the first failing window is re-raised on the calling thread
    std::vector<int> vrc(B, SVIN_OK);
    std::vector<std::string> vmsg(B);
    vflags.assign(B, 0);
    // the compact records are written by the validation pass itself when the batch's COUNTS allow the one-word format
    // (whether the measurements are float-exact is only known afterwards: if not, the doubles are copied by the fill pass)
    const bool compact_wanted = !(std::getenv("SVIN_BA_COMPACT_OBS") && std::atoi(std::getenv("SVIN_BA_COMPACT_OBS")) == 0);
    std::vector<size_t> obs0(B + 1, 0);
    prepacked = compact_wanted;
    for (int i = 0; i < B; ++i) {
      obs0[i + 1] = obs0[i] + (size_t)std::max(0, wins[i].num_obs);
      prepacked = prepacked && wins[i].num_pose_blocks <= 64 && wins[i].num_cameras <= 4 && wins[i].num_landmarks < (1 << 18);
    }
    if (prepacked && ensure((void**)&c->h_obs, &c->h_obs_cap, 12 * obs0[B] + 64, true) != SVIN_OK) prepacked = false;
    c->pool->run(B, [&](int i) {
      ObsSink sink;
      if (prepacked) {
        sink.word = reinterpret_cast<int*>(c->h_obs) + obs0[i];
        sink.z = reinterpret_cast<float*>(c->h_obs + 4 * obs0[B]) + 2 * obs0[i];
      }
      vrc[i] = validate(wins[i], i, &vflags[i], prepacked ? &sink : nullptr);
      if (vrc[i] != SVIN_OK) vmsg[i] = svin_last_error();
    });
    for (int i = 0; i < B; ++i)
      if (vrc[i] != SVIN_OK) {
        set_error(vmsg[i]);
        return vrc[i];
      }
  }
  c->uploaded = false;
  c->solved = false;
  // ---------------- totals
  long long NPB = 0, NSB = 0, NL = 0, NC = 0, NOBS = 0, NIMU = 0, NMEAS = 0, NPP = 0, NSP = 0, NRP = 0, NSO = 0, NDE = 0;
  long long NMB = 0, NMJ = 0, NME = 0, NMLIN = 0, ND = 0, NROWS = 0, NH = 0, NJD = 0;
  int n_obs_tiles = 0, n_lm_tiles = 0;
  c->h_win.assign(B, WinDesc{});
  c->n_max = 0;
  int has_ext = 0;
  for (int i = 0; i < B; ++i) has_ext |= (vflags[i] >> 1) & 1;
  // pattern grouping (k_schur_mma) needs fixed extrinsics and at most 64 pose runs per landmark
  bool group = !has_ext;
  for (int i = 0; i < B; ++i)
    if (wins[i].num_pose_blocks > 64) group = false;
  // Device planner (ba_plan.cu): pattern grouping, chunking and the observation order are computed on the GPU from the raw
  // arrays when every window qualifies (observations sorted by (landmark, pose, camera) - what the adapter's Map walk
  // delivers - and <= kPlanMaxLandmarks landmarks); otherwise the host threads do it (order_window).  SVIN_BA_DEVICE_PLAN=0
  // forces the host planner (A/B, and the tests of that path).
  const bool dev_plan_wanted = !(std::getenv("SVIN_BA_DEVICE_PLAN") && std::atoi(std::getenv("SVIN_BA_DEVICE_PLAN")) == 0);
  bool device_plan = group && dev_plan_wanted;
  int max_landmarks = 0;
  for (int i = 0; i < B; ++i) {
    device_plan = device_plan && !(vflags[i] & 1) && wins[i].num_landmarks <= kPlanMaxLandmarks;
    max_landmarks = std::max(max_landmarks, wins[i].num_landmarks);
  }
  c->device_plan = device_plan;
  std::vector<WindowOrder> orders(B);
  long long NSW = 0;
  // independent per window: spread over the context's host threads (this is on the end-to-end path)
  if (!device_plan) c->pool->run(B, [&](int i) { order_window(wins[i], group, orders[i]); });
  const double t_ordered = wall_ms();
  long long NRUN = 0;
  for (int i = 0; i < B; ++i) NSW += (long long)orders[i].chunk_begin.size();
  for (int i = 0; i < B; ++i) NRUN += (long long)orders[i].run_pose.size();
  bool any_lmfix = false;
  for (int i = 0; i < B; ++i) any_lmfix = any_lmfix || wins[i].landmark_fixed != nullptr;
  for (int i = 0; i < B; ++i) {
    const SvinBaWindow& w = wins[i];
    WinDesc& d = c->h_win[i];
    d.pose_begin = (int)NPB; NPB += w.num_pose_blocks; d.pose_end = (int)NPB;
    d.sb_begin = (int)NSB; NSB += w.num_speedbias; d.sb_end = (int)NSB;
    d.lm_begin = (int)NL; NL += w.num_landmarks; d.lm_end = (int)NL;
    d.obs_begin = (int)NOBS; NOBS += w.num_obs; d.obs_end = (int)NOBS;
    NC += w.num_cameras;
    int n = 0;
    for (int k = 0; k < w.num_pose_blocks; ++k) n += w.pose_fixed[k] ? 0 : 6;
    for (int k = 0; k < w.num_speedbias; ++k) n += w.speedbias_fixed[k] ? 0 : 9;
    d.n_dense = n;
    c->n_max = std::max(c->n_max, n);
    d.imu_begin = (int)NIMU; NIMU += w.num_imu; d.imu_end = (int)NIMU;
    if (w.num_imu) NMEAS += w.imu_meas_offset[w.num_imu];
    d.pp_begin = (int)NPP; NPP += w.num_pose_priors; d.pp_end = (int)NPP;
    d.sp_begin = (int)NSP; NSP += w.num_speedbias_priors; d.sp_end = (int)NSP;
    d.rp_begin = (int)NRP; NRP += w.num_relative_pose; d.rp_end = (int)NRP;
    d.so_begin = (int)NSO; NSO += w.num_sonar; d.so_end = (int)NSO;
    d.de_begin = (int)NDE; NDE += w.num_depth; d.de_end = (int)NDE;
    d.marg_blk_begin = (int)NMB; NMB += w.marg_num_blocks; d.marg_blk_end = (int)NMB;
    d.marg_dim = w.marg_dim;
    d.margJ_off = NMJ; NMJ += (long long)w.marg_dim * w.marg_dim;
    d.marg_e0_off = (int)NME; NME += w.marg_dim;
    d.marg_lin_off = (int)NMLIN;
    for (int k = 0; k < w.marg_num_blocks; ++k) NMLIN += (w.marg_block_kind[k] == SVIN_BLOCK_POSE) ? 7 : 9;
    const int rows = 15 * w.num_imu + 6 * w.num_pose_priors + 9 * w.num_speedbias_priors + 6 * w.num_relative_pose +
                     w.num_sonar + w.num_depth + w.marg_dim;
    d.n_rows = rows;
    d.d_off = (int)ND; ND += n;
    d.rd_off = (int)NROWS; NROWS += rows;
    d.H_off = NH; NH += (long long)n * n;
    d.Jd_off = NJD; NJD += (long long)rows * n;
    d.loss_type = w.loss_type;
    d.loss_scale = w.loss_scale;
    d.imu = ImuP{w.imu_params.sigma_g_c, w.imu_params.sigma_a_c, w.imu_params.sigma_gw_c, w.imu_params.sigma_aw_c,
                 w.imu_params.g, w.imu_params.g_max, w.imu_params.a_max};
    for (int k = 0; k < 7; ++k) d.T_SSo[k] = (w.num_sonar && w.sonar_T_SSo) ? w.sonar_T_SSo[k] : (k == 6 ? 1.0 : 0.0);
    n_obs_tiles += (w.num_obs + kObsTile - 1) / kObsTile;
    n_lm_tiles += (w.num_landmarks + kLmTile - 1) / kLmTile;
  }
  if (NOBS >= (1ll << 31) || NH >= (1ll << 40)) {
    set_error("batch too large");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  // ---------------- input arena layout
  Region in;
  const size_t o_win = in.add(sizeof(WinDesc) * B), o_ws = in.add(sizeof(WinState) * B);
  const size_t o_pose = in.add(8 * 7 * NPB), o_sb = in.add(8 * 9 * NSB), o_lm = in.add(8 * 4 * NL);
  const size_t o_poff = in.add(4 * NPB), o_sboff = in.add(4 * NSB), o_lmfix = in.add(NL);
  const size_t o_pwin = in.add(4 * NPB), o_sbwin = in.add(4 * NSB);
  const size_t o_intr = in.add(8 * 8 * NC);
  const size_t o_opose = in.add(4 * NOBS), o_olm = in.add(4 * NOBS), o_oext = in.add(4 * NOBS), o_ocam = in.add(4 * NOBS);
  const size_t o_zx = in.add(8 * NOBS), o_zy = in.add(8 * NOBS), o_u00 = in.add(8 * NOBS), o_u01 = in.add(8 * NOBS),
               o_u11 = in.add(8 * NOBS);
  // raw observation arrays in the caller's order (plain copies on the host); k_pack_obs gathers them into the SoA
  // planes above on the device - the host no longer touches every observation with a sqrt and a scatter
  const size_t o_rpose = in.add(4 * NOBS), o_rlm = in.add(4 * NOBS), o_rext = in.add(4 * NOBS), o_rcam = in.add(4 * NOBS);
  const size_t o_rmeas = in.add(16 * NOBS), o_rinfo = in.add(24 * NOBS), o_rord = in.add(4 * NOBS);
  // compact upload: pose | ext << 10 | cam << 20 in one word per observation when every window has <= 1024 pose blocks and
  // cameras (else the three separate arrays travel), and no per-observation information when a window's is uniform
  bool packed_idx = true;
  for (int i = 0; i < B; ++i) packed_idx = packed_idx && wins[i].num_pose_blocks <= 1024 && wins[i].num_cameras <= 1024;
  const size_t o_rpec = in.add(4 * NOBS);
  // more compact still (SVIN_BA_COMPACT_OBS=0 turns both off): landmark | pose << 18 | ext << 24 | cam << 30 in ONE word
  // when every window has <= 64 pose blocks, <= 4 cameras and < 2^18 landmarks (the landmark array then stays home), and
  // float measurements when every coordinate is float-exact - BRISK keypoints are (cv::KeyPoint::pt is float, the
  // reference widens it in Frame::getKeypoint), so nothing is rounded.  12 B instead of 24 B per observation.
  const bool compact_wanted = !(std::getenv("SVIN_BA_COMPACT_OBS") && std::atoi(std::getenv("SVIN_BA_COMPACT_OBS")) == 0);
  bool one_word = packed_idx && compact_wanted, meas_f32 = compact_wanted;
  for (int i = 0; i < B; ++i) {
    one_word = one_word && wins[i].num_pose_blocks <= 64 && wins[i].num_cameras <= 4 && wins[i].num_landmarks < (1 << 18);
    meas_f32 = meas_f32 && !(vflags[i] & 4);
  }
  // the planner's tables: filled by the host threads and copied (host planner), or written by ba_plan.cu on the device,
  // where chunk ids are lm_begin + local chunk and run ids obs_begin + local run (capacities NL / NOBS, nothing travels)
  if (device_plan) {
    NSW = NL;
    NRUN = NOBS;
  }
  const size_t o_plan = in.bytes;
  const size_t o_lmwin = in.add(4 * NL);
  const size_t o_lmof = in.add(4 * NL), o_lmos = in.add(4 * NL), o_lmoc = in.add(4 * NL), o_linv = in.add(4 * NL);
  const size_t o_sww = in.add(4 * (size_t)NSW), o_swb = in.add(4 * (size_t)NSW), o_swc = in.add(4 * (size_t)NSW);
  const size_t o_swnr = in.add(4 * (size_t)NSW), o_swrf = in.add(4 * (size_t)NSW), o_swlist = in.add(4 * (size_t)NSW);
  const size_t o_runoff = in.add(4 * (size_t)NRUN), o_runkm = in.add(4 * (size_t)NRUN);
  const size_t o_tail = in.bytes;
  const size_t o_otw = in.add(4 * (size_t)n_obs_tiles), o_otb = in.add(4 * (size_t)n_obs_tiles);
  const size_t o_ltw = in.add(4 * (size_t)n_lm_tiles), o_ltb = in.add(4 * (size_t)n_lm_tiles);
  const size_t o_imu = in.add(sizeof(ImuTerm) * NIMU), o_imuc = in.add(sizeof(ImuCache) * NIMU);
  const size_t o_mt = in.add(8 * NMEAS), o_mg = in.add(24 * NMEAS), o_ma = in.add(24 * NMEAS);
  const size_t o_pp = in.add(sizeof(PosePrior) * NPP), o_sp = in.add(sizeof(SbPrior) * NSP);
  const size_t o_rp = in.add(sizeof(RelPose) * NRP), o_so = in.add(sizeof(SonarTerm) * NSO);
  const size_t o_de = in.add(sizeof(DepthTerm) * NDE), o_mb = in.add(sizeof(MargBlock) * NMB);
  const size_t o_mJ = in.add(8 * NMJ), o_me = in.add(8 * NME), o_ml = in.add(8 * NMLIN);
  // device only (device planner): permuted landmark state, landmark permutation for the download, chunk kinds, class
  // counts and scratch.  With the device planner o_lm / o_lmfix hold the caller-order arrays.
  const size_t host_bytes = in.bytes;
  size_t o_lmi = 0, o_lmfixi = 0, o_lmperm = 0, o_swkind = 0, o_wcc = 0, o_wcb = 0, o_wnc = 0, o_ctot = 0, o_scr[5] = {};
  if (device_plan) {
    o_lmi = in.add(32 * NL);
    o_lmfixi = in.add(NL);
    o_lmperm = in.add(4 * NL);
    o_swkind = in.add(4 * NL);
    o_wcc = in.add(4 * (size_t)kSchurClasses * B);
    o_wcb = in.add(4 * (size_t)kSchurClasses * B);
    o_wnc = in.add(4 * (size_t)B);
    o_ctot = in.add(4 * (kSchurClasses + 1));
    for (int k = 0; k < 5; ++k) o_scr[k] = in.add(4 * ((size_t)NL + 2 * (size_t)B + 8));
  }
  int rc;
  if ((rc = ensure(&c->h_in, &c->h_in_cap, host_bytes, true)) != SVIN_OK) return rc;
  if ((rc = ensure(&c->d_in, &c->d_in_cap, in.bytes, false)) != SVIN_OK) return rc;
  char* H = (char*)c->h_in;
  auto hp = [&](size_t off) { return H + off; };
  // ---------------- fill staging
  WinDesc* hw = (WinDesc*)hp(o_win);
  WinState* hs = (WinState*)hp(o_ws);
  double *h_pose = (double*)hp(o_pose), *h_sb = (double*)hp(o_sb), *h_lm = (double*)hp(o_lm);
  int *h_poff = (int*)hp(o_poff), *h_sboff = (int*)hp(o_sboff), *h_lmwin = (int*)hp(o_lmwin);
  int *h_pwin = (int*)hp(o_pwin), *h_sbwin = (int*)hp(o_sbwin);
  uint8_t* h_lmfix = (uint8_t*)hp(o_lmfix);
  double* h_intr = (double*)hp(o_intr);
  int *h_lmof = (int*)hp(o_lmof), *h_lmos = (int*)hp(o_lmos), *h_lmoc = (int*)hp(o_lmoc), *h_linv = (int*)hp(o_linv);
  int *h_rpose = (int*)hp(o_rpose), *h_rlm = (int*)hp(o_rlm), *h_rext = (int*)hp(o_rext), *h_rcam = (int*)hp(o_rcam);
  int* h_rord = (int*)hp(o_rord);
  int* h_rpec = (int*)hp(o_rpec);
  double *h_rmeas = (double*)hp(o_rmeas), *h_rinfo = (double*)hp(o_rinfo);
  int *h_otw = (int*)hp(o_otw), *h_otb = (int*)hp(o_otb), *h_ltw = (int*)hp(o_ltw), *h_ltb = (int*)hp(o_ltb);
  int *h_sww = (int*)hp(o_sww), *h_swb = (int*)hp(o_swb), *h_swc = (int*)hp(o_swc);
  int *h_swnr = (int*)hp(o_swnr), *h_swrf = (int*)hp(o_swrf), *h_runoff = (int*)hp(o_runoff), *h_runkm = (int*)hp(o_runkm);
  ImuTerm* h_imu = (ImuTerm*)hp(o_imu);
  ImuCache* h_imuc = (ImuCache*)hp(o_imuc);
  long long* h_mt = (long long*)hp(o_mt);
  double *h_mg = (double*)hp(o_mg), *h_ma = (double*)hp(o_ma);
  PosePrior* h_pp = (PosePrior*)hp(o_pp);
  SbPrior* h_sp = (SbPrior*)hp(o_sp);
  RelPose* h_rp = (RelPose*)hp(o_rp);
  SonarTerm* h_so = (SonarTerm*)hp(o_so);
  DepthTerm* h_de = (DepthTerm*)hp(o_de);
  MargBlock* h_mb = (MargBlock*)hp(o_mb);
  double *h_mJ = (double*)hp(o_mJ), *h_me = (double*)hp(o_me), *h_ml = (double*)hp(o_ml);

  c->obs_perm_valid = !device_plan;
  if (!device_plan) {
    c->obs_perm.resize((size_t)NOBS);
    c->lm_perm.resize((size_t)NL);
    c->lm_perm_p = c->lm_perm.data();
  } else {
    if ((rc = ensure((void**)&c->h_lm_perm, &c->h_lm_perm_cap, 4 * (size_t)NL + 64, true)) != SVIN_OK) return rc;
    c->lm_perm_p = c->h_lm_perm + 16;   // [0..15]: class totals read back from the planner
  }
  // per-window bases of the running counters, so that windows can be packed by independent host threads
  std::vector<int> cam_base_v(B), ot_v(B), lt_v(B), sw_v(B), run_v(B);
  std::vector<long long> meas_base_v(B);
  {
    int cb = 0, ot0 = 0, lt0 = 0, sw0 = 0, run0 = 0;
    long long mb0 = 0;
    for (int i = 0; i < B; ++i) {
      cam_base_v[i] = cb; ot_v[i] = ot0; lt_v[i] = lt0; sw_v[i] = sw0; meas_base_v[i] = mb0; run_v[i] = run0;
      run0 += (int)orders[i].run_pose.size();
      cb += wins[i].num_cameras;
      ot0 += (wins[i].num_obs + kObsTile - 1) / kObsTile;
      lt0 += (wins[i].num_landmarks + kLmTile - 1) / kLmTile;
      sw0 += (int)orders[i].chunk_begin.size();
      if (wins[i].num_imu) mb0 += wins[i].imu_meas_offset[wins[i].num_imu];
    }
  }
  const int G = std::max(1, std::min(8, B / 16));   // window groups of the overlapped H2D copy below
  auto group_of = [&](int i) { return (int)((long long)i * G / B); };
  std::vector<std::atomic<int>> group_info(G);   // 1: a window of the group has per-observation information
  for (int g = 0; g < G; ++g) group_info[g].store(0);
  auto fill_window = [&](int i) {
    int cam_base = cam_base_v[i], ot = ot_v[i], lt = lt_v[i], sw = sw_v[i];
    long long meas_base = meas_base_v[i];
    std::vector<int> cnt;
    const SvinBaWindow& w = wins[i];
    WinDesc& d = c->h_win[i];
    std::memcpy(h_pose + 7 * (size_t)d.pose_begin, w.pose_blocks, 56 * (size_t)w.num_pose_blocks);
    std::memcpy(h_sb + 9 * (size_t)d.sb_begin, w.speedbias, 72 * (size_t)w.num_speedbias);
    int off = 0;
    for (int k = 0; k < w.num_pose_blocks; ++k) {
      h_poff[d.pose_begin + k] = w.pose_fixed[k] ? -1 : off;
      if (!w.pose_fixed[k]) off += 6;
      h_pwin[d.pose_begin + k] = i;
    }
    for (int k = 0; k < w.num_speedbias; ++k) {
      h_sboff[d.sb_begin + k] = w.speedbias_fixed[k] ? -1 : off;
      if (!w.speedbias_fixed[k]) off += 9;
      h_sbwin[d.sb_begin + k] = i;
    }
    const WindowOrder& wo = orders[i];
    std::vector<int>& inv = cnt;  // caller landmark -> internal landmark (reuses the scratch vector)
    if (device_plan) {
      // caller order; the planner permutes on the device
      if (w.num_landmarks) std::memcpy(h_lm + 4 * (size_t)d.lm_begin, w.landmarks, 32 * (size_t)w.num_landmarks);
      if (any_lmfix) {
        if (w.landmark_fixed)
          for (int k = 0; k < w.num_landmarks; ++k) h_lmfix[d.lm_begin + k] = w.landmark_fixed[k] ? 1 : 0;
        else
          std::memset(h_lmfix + d.lm_begin, 0, (size_t)w.num_landmarks);
      }
    } else {
      inv.assign(w.num_landmarks, 0);
      for (int k = 0; k < w.num_landmarks; ++k) {
        const int lc = wo.lm_perm[k];
        inv[lc] = k;
        c->lm_perm[(size_t)d.lm_begin + k] = lc;
        std::memcpy(h_lm + 4 * ((size_t)d.lm_begin + k), w.landmarks + 4 * (size_t)lc, 32);
        h_lmfix[d.lm_begin + k] = (w.landmark_fixed && w.landmark_fixed[lc]) ? 1 : 0;
        h_lmwin[d.lm_begin + k] = i;
      }
    }
    std::memcpy(h_intr + 8 * (size_t)cam_base, w.intrinsics, 64 * (size_t)w.num_cameras);
    // observations: the caller's arrays as they are + the internal order; the gather into the internal
    // (landmark-major, pattern-major inside a chunk) SoA planes and U = chol(information)^T run on the device
    const int N = w.num_obs;
    const size_t g0 = (size_t)d.obs_begin;
    if (one_word) {
      if (!prepacked)
        for (int o = 0; o < N; ++o)
          h_rpec[g0 + o] = (int)((unsigned)w.obs_landmark[o] | ((unsigned)w.obs_pose[o] << 18) |
                                 ((unsigned)w.obs_extrinsics[o] << 24) | ((unsigned)w.obs_camera[o] << 30));
    } else if (packed_idx) {
      std::memcpy(h_rlm + g0, w.obs_landmark, 4 * (size_t)N);
      for (int o = 0; o < N; ++o) h_rpec[g0 + o] = w.obs_pose[o] | (w.obs_extrinsics[o] << 10) | (w.obs_camera[o] << 20);
    } else {
      std::memcpy(h_rlm + g0, w.obs_landmark, 4 * (size_t)N);
      std::memcpy(h_rpose + g0, w.obs_pose, 4 * (size_t)N);
      std::memcpy(h_rext + g0, w.obs_extrinsics, 4 * (size_t)N);
      std::memcpy(h_rcam + g0, w.obs_camera, 4 * (size_t)N);
    }
    if (meas_f32) {
      if (!prepacked) {
        float* zf = reinterpret_cast<float*>(h_rmeas) + 2 * g0;
        for (int o = 0; o < 2 * N; ++o) zf[o] = (float)w.obs_measurement[o];
      }
    } else {
      std::memcpy(h_rmeas + 2 * g0, w.obs_measurement, 16 * (size_t)N);
    }
    // information: the three entries the Cholesky factor reads (row-major 2x2: a00, a10, a11); one copy per window when
    // they are all equal (OKVIS: 64 / size^2 * I with one keypoint size, Estimator.hpp impl:64-67)
    bool uniform = N > 0;
    for (int o = 1; o < N && uniform; ++o) {
      const double* a = w.obs_information + 4 * (size_t)o;
      uniform = a[0] == w.obs_information[0] && a[2] == w.obs_information[2] && a[3] == w.obs_information[3];
    }
    d.info_uniform = uniform ? 1 : 0;
    if (uniform) {
      d.info3[0] = w.obs_information[0];
      d.info3[1] = w.obs_information[2];
      d.info3[2] = w.obs_information[3];
    } else {
      d.info3[0] = d.info3[1] = d.info3[2] = 0.0;
      group_info[group_of(i)].store(1, std::memory_order_relaxed);   // this group's information slice must travel
      for (int o = 0; o < N; ++o) {
        const double* a = w.obs_information + 4 * (size_t)o;
        double* t = h_rinfo + 3 * (g0 + o);
        t[0] = a[0];
        t[1] = a[2];
        t[2] = a[3];
      }
    }
    d.cam_begin = cam_base;
    if (!device_plan) {
      std::memcpy(h_rord + g0, wo.obs_order.data(), 4 * (size_t)N);
      std::memcpy(c->obs_perm.data() + g0, wo.obs_order.data(), 4 * (size_t)N);
      std::memcpy(h_linv + d.lm_begin, inv.data(), 4 * (size_t)w.num_landmarks);
      for (int k = 0; k < w.num_landmarks; ++k) {
        h_lmof[d.lm_begin + k] = d.obs_begin + wo.lm_first[k];
        h_lmos[d.lm_begin + k] = wo.lm_stride[k];
        h_lmoc[d.lm_begin + k] = wo.lm_count[k];
      }
      for (size_t k = 0; k < wo.chunk_begin.size(); ++k) {
        h_sww[sw] = i;
        h_swb[sw] = d.lm_begin + wo.chunk_begin[k];
        h_swc[sw] = wo.chunk_count[k];
        h_swnr[sw] = wo.chunk_nruns[k];
        h_swrf[sw] = run_v[i] + wo.chunk_run_first[k];
        ++sw;
      }
      for (size_t k = 0; k < wo.run_pose.size(); ++k) {
        h_runoff[run_v[i] + k] = h_poff[d.pose_begin + wo.run_pose[k]];
        h_runkm[run_v[i] + k] = wo.run_k0m[k];
      }
    }
    for (int t0 = 0; t0 < N; t0 += kObsTile) {
      h_otw[ot] = i;
      h_otb[ot] = d.obs_begin + t0;
      ++ot;
    }
    for (int t0 = 0; t0 < w.num_landmarks; t0 += kLmTile) {
      h_ltw[lt] = i;
      h_ltb[lt] = d.lm_begin + t0;
      ++lt;
    }
    // dense terms, rows laid out IMU | pose priors | speedbias priors | relative pose | sonar | depth | marg
    int row = 0;
    for (int k = 0; k < w.num_imu; ++k) {
      ImuTerm& t = h_imu[d.imu_begin + k];
      t.pose0 = d.pose_begin + w.imu_pose0[k];
      t.pose1 = d.pose_begin + w.imu_pose1[k];
      t.sb0 = d.sb_begin + w.imu_speedbias0[k];
      t.sb1 = d.sb_begin + w.imu_speedbias1[k];
      t.meas_begin = (int)(meas_base + w.imu_meas_offset[k]);
      t.meas_end = (int)(meas_base + w.imu_meas_offset[k + 1]);
      t.row0 = row;
      t.win = i;
      t.t0 = w.imu_t0_ns[k];
      t.t1 = w.imu_t1_ns[k];
      row += 15;
      ImuCache& ic = h_imuc[d.imu_begin + k];
      std::memset(&ic, 0, sizeof ic);
      ic.Delta_q[3] = 1.0;
      ic.redo = 1;
    }
    if (w.num_imu) {
      const int nm = w.imu_meas_offset[w.num_imu];
      std::memcpy(h_mt + meas_base, w.imu_meas_t_ns, 8 * (size_t)nm);
      std::memcpy(h_mg + 3 * meas_base, w.imu_meas_gyro, 24 * (size_t)nm);
      std::memcpy(h_ma + 3 * meas_base, w.imu_meas_accel, 24 * (size_t)nm);
      meas_base += nm;
    }
    for (int k = 0; k < w.num_pose_priors; ++k) {
      PosePrior& t = h_pp[d.pp_begin + k];
      t.block = d.pose_begin + w.pose_prior_block[k];
      t.row0 = row;
      t.win = i;
      t.pad = 0;
      std::memcpy(t.meas, w.pose_prior_measurement + 7 * (size_t)k, 56);
      llt_sqrt_information(w.pose_prior_information + 36 * (size_t)k, t.U, 6);
      row += 6;
    }
    for (int k = 0; k < w.num_speedbias_priors; ++k) {
      SbPrior& t = h_sp[d.sp_begin + k];
      t.block = d.sb_begin + w.speedbias_prior_block[k];
      t.row0 = row;
      t.win = i;
      t.pad = 0;
      std::memcpy(t.meas, w.speedbias_prior_measurement + 9 * (size_t)k, 72);
      llt_sqrt_information(w.speedbias_prior_information + 81 * (size_t)k, t.U, 9);
      row += 9;
    }
    for (int k = 0; k < w.num_relative_pose; ++k) {
      RelPose& t = h_rp[d.rp_begin + k];
      t.block0 = d.pose_begin + w.relative_pose_block0[k];
      t.block1 = d.pose_begin + w.relative_pose_block1[k];
      t.row0 = row;
      t.win = i;
      llt_sqrt_information(w.relative_pose_information + 36 * (size_t)k, t.U, 6);
      row += 6;
    }
    for (int k = 0; k < w.num_sonar; ++k) {
      SonarTerm& t = h_so[d.so_begin + k];
      t.pose = d.pose_begin + w.sonar_pose[k];
      t.row0 = row;
      t.win = i;
      t.pad = 0;
      t.range = w.sonar_range[k];
      t.heading = w.sonar_heading[k];
      t.sqrt_info = std::sqrt(w.sonar_information[k]);  // SonarError.cpp:97-104
      for (int x = 0; x < 3; ++x) t.mean[x] = w.sonar_landmark_mean[3 * (size_t)k + x];
      row += 1;
    }
    for (int k = 0; k < w.num_depth; ++k) {
      DepthTerm& t = h_de[d.de_begin + k];
      t.pose = d.pose_begin + w.depth_pose[k];
      t.row0 = row;
      t.win = i;
      t.pad = 0;
      t.depth = w.depth_measurement[k];
      t.first = w.depth_first[k];
      t.sqrt_info = std::sqrt(w.depth_information[k]);  // DepthError.cpp:55-62
      row += 1;
    }
    d.marg_row0 = row;
    int col = 0, lin = 0;
    for (int k = 0; k < w.marg_num_blocks; ++k) {
      MargBlock& mb = h_mb[d.marg_blk_begin + k];
      mb.kind = w.marg_block_kind[k];
      const int x = w.marg_block_index[k];
      const bool is_pose = mb.kind == SVIN_BLOCK_POSE;
      const bool fixed = is_pose ? w.pose_fixed[x] : w.speedbias_fixed[x];
      mb.index = (is_pose ? d.pose_begin : d.sb_begin) + x;
      mb.col0 = fixed ? -1 : col;
      mb.lin_off = lin;
      if (!fixed) col += is_pose ? 6 : 9;
      lin += is_pose ? 7 : 9;
    }
    if (w.marg_dim) {
      std::memcpy(h_mJ + d.margJ_off, w.marg_J, 8 * (size_t)w.marg_dim * w.marg_dim);
      std::memcpy(h_me + d.marg_e0_off, w.marg_e0, 8 * (size_t)w.marg_dim);
      std::memcpy(h_ml + d.marg_lin_off, w.marg_linearization_points, 8 * (size_t)lin);
    }
    hw[i] = d;
    WinState& s = hs[i];
    std::memset(&s, 0, sizeof s);
    s.radius = 1e4;  // overwritten by svin_ba_solve with the options' initial radius
    s.mu = 1e-8;
    s.last_successful = 1;
    (void)cam_base; (void)ot; (void)lt; (void)sw; (void)meas_base;
  };
  // chunk ids by lane-mapping class (one k_schur_mma<G> launch per class)
  int sw_class_count[kSchurClasses] = {};
  {
    int* h_swlist = (int*)hp(o_swlist);
    for (int i = 0; i < B && !device_plan; ++i)
      for (int kind : orders[i].chunk_kind) sw_class_count[kind]++;
    int pos[kSchurClasses] = {};
    for (int k = 1; k < kSchurClasses; ++k) pos[k] = pos[k - 1] + sw_class_count[k - 1];
    for (int i = 0; i < B && !device_plan; ++i)
      for (size_t k = 0; k < orders[i].chunk_count.size(); ++k)
        h_swlist[pos[orders[i].chunk_kind[k]]++] = sw_v[i] + (int)k;
  }
  // Packing runs on the pool while this thread lays out the work arena; the big observation arrays are copied
  // group by group as soon as their windows are packed, so that the H2D transfer overlaps the packing.
  std::vector<std::atomic<int>> group_left(G);
  for (int g = 0; g < G; ++g) group_left[g].store(0);
  for (int i = 0; i < B; ++i) group_left[group_of(i)].fetch_add(1);
  c->pool->start(B, [&](int i) {
    fill_window(i);
    group_left[group_of(i)].fetch_sub(1, std::memory_order_release);
  });
  // any early return below must first let the workers finish: they reference this frame
  struct PoolJoin {
    HostPool* p;
    ~PoolJoin() { p->help(); p->wait(); }
  } pool_join{c->pool};
  // ---------------- work arena
  const size_t S = ((size_t)NOBS + 31) & ~(size_t)31;
  Region wk;
  size_t o_poseb[2], o_sbb[2], o_lmb[2], o_r[2], o_Jp[2], o_Jl[2], o_Je[2], o_Jd[2], o_rd[2], o_gram[2], o_gramg[2];
  for (int k = 0; k < 2; ++k) {
    o_poseb[k] = wk.add(56 * NPB);
    o_sbb[k] = wk.add(72 * NSB);
    o_lmb[k] = wk.add(32 * NL);
    o_r[k] = wk.add(8 * 2 * S);
    o_Jp[k] = wk.add(8 * 12 * S);
    o_Jl[k] = wk.add(8 * 6 * S);
    o_Je[k] = has_ext ? wk.add(8 * 12 * S) : 0;
    o_Jd[k] = wk.add(8 * NJD);
    o_rd[k] = wk.add(8 * NROWS);
    o_gram[k] = wk.add(8 * NH);
    o_gramg[k] = wk.add(8 * ND);
  }
  const size_t o_wsw = wk.add(sizeof(WinState) * B), o_imucw = wk.add(sizeof(ImuCache) * NIMU);
  const size_t o_opoff = wk.add(4 * S);
  const size_t o_sacc = wk.add(8 * kShardAcc * (size_t)B);
  const size_t o_lms = wk.add(24 * NL), o_lmV = wk.add(48 * NL), o_lmb2 = wk.add(24 * NL), o_lmd = wk.add(24 * NL),
               o_lmg = wk.add(24 * NL), o_lmgn = wk.add(24 * NL);
  // cleared every slot: H, g_red, g_raw, Hdiag (kept contiguous)
  const size_t o_clear = wk.bytes;
  const size_t o_H = wk.add(8 * NH), o_gred = wk.add(8 * ND), o_graw = wk.add(8 * ND), o_hd = wk.add(8 * ND);
  // sharded mode: one slot per (window, rank) for the landmark gradient max - every rank fills its own slot, the SUM
  // all-reduce of the whole clear region then carries all of them and the max is taken locally (k_gmax_pack)
  const size_t o_gmaxb = wk.add(8 * (size_t)B * (size_t)(c->sharded() ? c->comm_world : 1));
  const size_t clear_bytes = wk.bytes - o_clear;
  const size_t o_sc = wk.add(8 * ND), o_dg = wk.add(8 * ND), o_gr = wk.add(8 * ND), o_gn = wk.add(8 * ND),
               o_u = wk.add(8 * ND), o_c = wk.add(8 * ND), o_dl = wk.add(8 * ND);
  if ((rc = ensure(&c->d_work, &c->d_work_cap, wk.bytes, false)) != SVIN_OK) return rc;
  // ---------------- output arena
  Region out;
  c->out_off_pose = out.add(56 * NPB);
  c->out_off_sb = out.add(72 * NSB);
  c->out_off_lm = out.add(32 * NL);
  c->out_off_q = out.add(8 * NL);
  c->out_off_ws = out.add(sizeof(WinState) * B);
  c->out_bytes = out.bytes;
  if ((rc = ensure(&c->d_out, &c->d_out_cap, out.bytes, false)) != SVIN_OK) return rc;
  if ((rc = ensure(&c->h_out, &c->h_out_cap, out.bytes, true)) != SVIN_OK) return rc;

  // ---------------- device pointers
  char* D = (char*)c->d_in;
  char* Wk = (char*)c->d_work;
  char* O = (char*)c->d_out;
  Batch& b = c->b;
  b = Batch{};
  b.B = B; b.NPB = (int)NPB; b.NSB = (int)NSB; b.NL = (int)NL; b.NC = (int)NC; b.NOBS = (int)NOBS;
  b.NIMU = (int)NIMU; b.NMEAS = (int)NMEAS;
  b.n_obs_tiles = n_obs_tiles; b.n_lm_tiles = n_lm_tiles; b.has_ext = has_ext; b.obs_stride = S;
  {
    // fused linearisation needs the pattern-grouped chunk kernels; SVIN_BA_FUSED=0 keeps the materialised Jacobians (A/B)
    static const bool fused_wanted = !(std::getenv("SVIN_BA_FUSED") && std::atoi(std::getenv("SVIN_BA_FUSED")) == 0);
    b.fused = (group && fused_wanted) ? 1 : 0;
  }
  b.win = (WinDesc*)(D + o_win);
  c->d_ws_init = (WinState*)(D + o_ws);
  b.ws = (WinState*)(Wk + o_wsw);
  b.pose_init = (double*)(D + o_pose); b.sb_init = (double*)(D + o_sb); b.lm_init = (double*)(D + o_lm);
  b.pose_off = (int*)(D + o_poff); b.sb_off = (int*)(D + o_sboff);
  b.lm_fixed = (uint8_t*)(D + (device_plan ? o_lmfixi : o_lmfix)); b.lm_win = (int*)(D + o_lmwin);
  if (device_plan) b.lm_init = (double*)(D + o_lmi);
  c->d_pose_win = (int*)(D + o_pwin); c->d_sb_win = (int*)(D + o_sbwin);
  b.intr = (double*)(D + o_intr);
  b.obs_pose = (int*)(D + o_opose); b.obs_lm = (int*)(D + o_olm); b.obs_ext = (int*)(D + o_oext);
  b.obs_cam = (int*)(D + o_ocam);
  b.obs_poff = (int*)(Wk + o_opoff);
  b.obs_zx = (double*)(D + o_zx); b.obs_zy = (double*)(D + o_zy);
  b.obs_u00 = (double*)(D + o_u00); b.obs_u01 = (double*)(D + o_u01); b.obs_u11 = (double*)(D + o_u11);
  b.lm_obs_first = (int*)(D + o_lmof); b.lm_obs_stride = (int*)(D + o_lmos); b.lm_obs_cnt = (int*)(D + o_lmoc);
  b.obs_tile_win = (int*)(D + o_otw); b.obs_tile_begin = (int*)(D + o_otb);
  b.lm_tile_win = (int*)(D + o_ltw); b.lm_tile_begin = (int*)(D + o_ltb);
  b.n_schur_warps = (int)NSW;
  for (int k = 0; k < kSchurClasses; ++k) b.sw_class_count[k] = sw_class_count[k];
  b.sw_win = (int*)(D + o_sww); b.sw_lm_begin = (int*)(D + o_swb); b.sw_count = (int*)(D + o_swc);
  b.sw_nruns = (int*)(D + o_swnr); b.sw_run_first = (int*)(D + o_swrf); b.sw_list = (int*)(D + o_swlist);
  b.run_off = (int*)(D + o_runoff); b.run_k0m = (int*)(D + o_runkm);
  b.imu = (ImuTerm*)(D + o_imu);
  c->d_imu_cache_init = (ImuCache*)(D + o_imuc);
  b.imu_cache = (ImuCache*)(Wk + o_imucw);
  b.imu_meas_t = (long long*)(D + o_mt); b.imu_meas_gyro = (double*)(D + o_mg); b.imu_meas_accel = (double*)(D + o_ma);
  b.pp = (PosePrior*)(D + o_pp); b.sp = (SbPrior*)(D + o_sp); b.rp = (RelPose*)(D + o_rp);
  b.so = (SonarTerm*)(D + o_so); b.de = (DepthTerm*)(D + o_de); b.marg_blk = (MargBlock*)(D + o_mb);
  b.marg_J = (double*)(D + o_mJ); b.marg_e0 = (double*)(D + o_me); b.marg_lin = (double*)(D + o_ml);
  for (int k = 0; k < 2; ++k) {
    b.pose[k] = (double*)(Wk + o_poseb[k]); b.sb[k] = (double*)(Wk + o_sbb[k]); b.lm[k] = (double*)(Wk + o_lmb[k]);
    b.lin_r[k] = (double*)(Wk + o_r[k]); b.lin_Jp[k] = (double*)(Wk + o_Jp[k]); b.lin_Jl[k] = (double*)(Wk + o_Jl[k]);
    b.lin_Je[k] = has_ext ? (double*)(Wk + o_Je[k]) : nullptr;
    b.Jd[k] = (double*)(Wk + o_Jd[k]); b.rd[k] = (double*)(Wk + o_rd[k]);
    b.gram[k] = (double*)(Wk + o_gram[k]); b.gram_g[k] = (double*)(Wk + o_gramg[k]);
  }
  b.lm_scale = (double*)(Wk + o_lms); b.lm_Vinv = (double*)(Wk + o_lmV); b.lm_bs = (double*)(Wk + o_lmb2);
  b.lm_diag = (double*)(Wk + o_lmd); b.lm_grad = (double*)(Wk + o_lmg); b.lm_gn = (double*)(Wk + o_lmgn);
  b.H = (double*)(Wk + o_H); b.g_red = (double*)(Wk + o_gred); b.g_raw = (double*)(Wk + o_graw);
  b.Hdiag = (double*)(Wk + o_hd);
  b.scale_d = (double*)(Wk + o_sc); b.diag_d = (double*)(Wk + o_dg); b.grad_d = (double*)(Wk + o_gr);
  b.gn_d = (double*)(Wk + o_gn); b.u_d = (double*)(Wk + o_u); b.c_d = (double*)(Wk + o_c);
  b.delta_d = (double*)(Wk + o_dl);
  c->d_pose_out = (double*)(O + c->out_off_pose); c->d_sb_out = (double*)(O + c->out_off_sb);
  c->d_lm_out = (double*)(O + c->out_off_lm);
  b.lm_quality = (double*)(O + c->out_off_q);
  b.shard_acc = c->sharded() ? (double*)(Wk + o_sacc) : nullptr;
  b.comm_rank = c->comm_rank;
  b.comm_world = c->sharded() ? c->comm_world : 1;
  b.gmax_buf = (double*)(Wk + o_gmaxb);
  c->d_clear = Wk + o_clear;
  c->clear_bytes = clear_bytes;
  if (c->local_comm) {
    const size_t need = std::max(clear_bytes / 8, kShardAcc * (size_t)B);
    if (need > c->local_tmp_cap) {
      if (c->local_tmp) cudaFree(c->local_tmp);
      c->local_tmp = nullptr;
      SVIN_CUDA(cudaMalloc(&c->local_tmp, 8 * need));
      c->local_tmp_cap = need;
    }
  }
  c->in_bytes = in.bytes;

  // ---------------- copy + initialise
  // The upload's copies and its small kernels (planner, pack) go to a high-priority stream: in a pipeline another
  // context's solve is usually running, and on an equal-priority stream these grids would only be dispatched in the
  // tails of its kernels - the upload would take as long as that solve.
  cudaStream_t up = c->up;
  SVIN_CUDA(cudaEventRecord(c->ev[0], up));
  size_t h2d_skipped = 0;   // slices of the observation arrays that did not have to travel
  if (c->pool->workers() == 0) c->pool->help();
  {
    // the seven raw per-observation arrays, sliced by window group
    const size_t obs_arr[9] = {o_rpose, o_rlm, o_rext, o_rcam, o_rmeas, o_rinfo, o_rord, o_rpec, 0};
    const size_t obs_elt[9] = {4, 4, 4, 4, 16, 24, 4, 4, 0};
    const size_t meas_elt = meas_f32 ? 8 : 16;
    int w0 = 0;
    for (int g = 0; g < G; ++g) {
      int w1 = w0;
      while (w1 < B && group_of(w1) == g) ++w1;
      while (group_left[g].load(std::memory_order_acquire) > 0) std::this_thread::yield();
      if (w1 > w0) {
        const size_t e0 = (size_t)c->h_win[w0].obs_begin, e1 = (size_t)c->h_win[w1 - 1].obs_end;
        for (int a = 0; a < 8 && e1 > e0; ++a) {
          const bool skip = (packed_idx && (a == 0 || a == 2 || a == 3)) || (!packed_idx && a == 7) ||
                            (a == 5 && group_info[g].load(std::memory_order_acquire) == 0) ||
                            (a == 6 && device_plan) ||   // the observation order is computed on the device
                            (a == 1 && one_word);
          if (skip) {
            h2d_skipped += obs_elt[a] * (e1 - e0);
            continue;
          }
          const size_t elt = a == 4 ? meas_elt : obs_elt[a];
          h2d_skipped += (obs_elt[a] - elt) * (e1 - e0);
          // the records the validation pass wrote live in their own staging buffer (words, then floats)
          const char* src = H + obs_arr[a] + elt * e0;
          if (prepacked && a == 7 && one_word) src = c->h_obs + 4 * e0;
          if (prepacked && a == 4 && meas_f32) src = c->h_obs + 4 * (size_t)NOBS + 8 * e0;
          SVIN_CUDA(cudaMemcpyAsync(D + obs_arr[a] + elt * e0, src, elt * (e1 - e0), cudaMemcpyHostToDevice, up));
        }
      }
      w0 = w1;
    }
    c->pool->wait();
    // everything else: the regions before and after the observation arrays (the planner's tables only when the host
    // filled them)
    const size_t obs_lo = o_opose, obs_hi = device_plan ? o_tail : o_plan;
    SVIN_CUDA(cudaMemcpyAsync(D, H, obs_lo, cudaMemcpyHostToDevice, up));
    SVIN_CUDA(cudaMemcpyAsync(D + obs_hi, H + obs_hi, host_bytes - obs_hi, cudaMemcpyHostToDevice, up));
    h2d_skipped += obs_hi - o_plan;
  }
  const double t_filled = wall_ms();
  SVIN_CUDA(cudaEventRecord(c->ev[1], up));
  SVIN_CUDA(cudaMemsetAsync(Wk + o_sacc, 0, 8 * kShardAcc * (size_t)B, up));
  // Jd must be zero outside the blocks the terms write (structure is static)
  for (int k = 0; k < 2; ++k) SVIN_CUDA(cudaMemsetAsync(b.Jd[k], 0, 8 * (size_t)NJD + 8, up));
  SVIN_CUDA(cudaMemsetAsync(b.lm_quality, 0, 8 * (size_t)NL + 8, up));
  if (device_plan) {
    PlanArgs pa{};
    pa.B = B;
    pa.packed = one_word ? 2 : (packed_idx ? 1 : 0);
    pa.win = b.win;
    pa.rlm = (const int*)(D + o_rlm); pa.rpec = (const int*)(D + o_rpec);
    pa.rpose = (const int*)(D + o_rpose); pa.rcam = (const int*)(D + o_rcam);
    pa.lm_raw = (const double*)(D + o_lm); pa.lmfix_raw = (const unsigned char*)(D + o_lmfix);
    pa.poff = b.pose_off;
    pa.lm_init = b.lm_init; pa.lm_fixed = b.lm_fixed; pa.lm_win = b.lm_win;
    pa.linv = (int*)(D + o_linv); pa.lm_perm = (int*)(D + o_lmperm);
    pa.lmof = b.lm_obs_first; pa.lmos = b.lm_obs_stride; pa.lmoc = b.lm_obs_cnt; pa.rord = (int*)(D + o_rord);
    pa.sw_win = b.sw_win; pa.sw_lm_begin = b.sw_lm_begin; pa.sw_count = b.sw_count; pa.sw_nruns = b.sw_nruns;
    pa.sw_run_first = b.sw_run_first; pa.sw_kind = (int*)(D + o_swkind);
    pa.run_off = b.run_off; pa.run_k0m = b.run_k0m; pa.sw_list = b.sw_list;
    pa.win_class_count = (int*)(D + o_wcc); pa.win_class_base = (int*)(D + o_wcb); pa.win_nchunks = (int*)(D + o_wnc);
    pa.class_total = (int*)(D + o_ctot);
    for (int k = 0; k < 5; ++k) pa.scratch[k] = (int*)(D + o_scr[k]);
    pa.caps = plan_caps_table();
    if (!any_lmfix) SVIN_CUDA(cudaMemsetAsync(D + o_lmfix, 0, (size_t)NL + 1, up));
    SVIN_CUDA(launch_plan(pa, max_landmarks, up));
    SVIN_CUDA(cudaMemcpyAsync(c->h_lm_perm, D + o_ctot, 4 * (kSchurClasses + 1), cudaMemcpyDeviceToHost, up));
    if (NL)
      SVIN_CUDA(cudaMemcpyAsync(c->h_lm_perm + 16, D + o_lmperm, 4 * (size_t)NL, cudaMemcpyDeviceToHost, up));
  }
  c->d_rord = (const int*)(D + o_rord);
  c->n_obs_total = (size_t)NOBS;
  c->n_sw_cap = (size_t)NSW;
  launch_reset_state(b, up);
  {
    RawObs raw{packed_idx ? (const int*)(D + o_rpec) : nullptr, one_word ? 1 : 0,
               (const int*)(D + o_rpose), (const int*)(D + o_rlm), (const int*)(D + o_rext), (const int*)(D + o_rcam),
               (const double*)(D + o_rmeas), meas_f32 ? (const float*)(D + o_rmeas) : nullptr,
               (const double*)(D + o_rinfo), (const int*)(D + o_rord), (const int*)(D + o_linv)};
    launch_pack_obs(b, raw, up);
  }
  launch_obs_poff(b, up);
  SVIN_CUDA(cudaMemcpyAsync(b.ws, c->d_ws_init, sizeof(WinState) * B, cudaMemcpyDeviceToDevice, up));
  if (NIMU)
    SVIN_CUDA(cudaMemcpyAsync(b.imu_cache, c->d_imu_cache_init, sizeof(ImuCache) * NIMU, cudaMemcpyDeviceToDevice,
                              up));
  SVIN_CUDA(cudaGetLastError());
  // the reduced system lives in shared memory when it fits
  {
    const size_t need = dense_solve_smem_bytes(c->n_max);
    c->smem_bytes = (need <= 227 * 1024 && c->n_max <= 256 && c->n_max > 0) ? (int)need : 0;
    if (c->smem_bytes > 0) {
      SVIN_CUDA(configure_dense_solve(c->smem_bytes));
      SVIN_CUDA(configure_dense_gram((int)dense_gram_smem_bytes(c->n_max)));
    }
    SVIN_CUDA(configure_schur());
  }
  SVIN_CUDA(cudaStreamSynchronize(up));
  if (device_plan) {
    // the chunk counts size the Schur launches
    for (int k = 0; k < kSchurClasses; ++k) b.sw_class_count[k] = c->h_lm_perm[k];
    b.n_schur_warps = c->h_lm_perm[kSchurClasses];
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->tm = SvinBaTimings{};
  c->tm.h2d_ms = ms;
  // the SoA observation planes are produced on the device; packed indices / uniform information skip their slices
  c->tm.h2d_bytes = (int64_t)(host_bytes - (o_rpose - o_opose)) - (int64_t)h2d_skipped;
  c->tm.host_order_ms = t_ordered - t_begin;
  c->tm.host_fill_ms = t_filled - t_ordered;
  c->tm.host_upload_ms = wall_ms() - t_begin;
  c->uploaded = true;
  c->solves_since_upload = 0;
  c->graph_valid = false;
  c->solved = false;
  if (graph_wanted() && c->graph_opt_known && !c->profiling && !c->sharded() && B >= graph_min_windows()) {
    const int grc = build_graph(c, c->graph_opt);
    if (grc != SVIN_OK) return grc;
  }
  c->quality_valid = false;
  return SVIN_OK;
}

int svin_ba_plan(const SvinBaWindow* w, int32_t* lm_order, int32_t cap, int32_t* kind, int32_t* count, int32_t* runs,
                 int32_t* num_chunks) {
  if (!w || !num_chunks) {
    set_error("svin_ba_plan: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  const int rc = validate(*w, 0);
  if (rc != SVIN_OK) return rc;
  bool group = w->num_pose_blocks <= 64;
  for (int o = 0; o < w->num_obs && group; ++o)
    if (!w->pose_fixed[w->obs_extrinsics[o]]) group = false;  // estimated extrinsics: no pattern grouping
  WindowOrder wo;
  order_window(*w, group, wo);
  if (lm_order)
    for (int k = 0; k < w->num_landmarks; ++k) lm_order[k] = wo.lm_perm[k];
  const int n = (int)wo.chunk_begin.size();
  *num_chunks = n;
  if (n > cap) {
    set_error("svin_ba_plan: capacity too small");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  for (int k = 0; k < n; ++k) {
    if (kind) kind[k] = wo.chunk_kind[k];
    if (count) count[k] = wo.chunk_count[k];
    if (runs) runs[k] = wo.chunk_nruns[k];
  }
  return SVIN_OK;
}

int svin_ba_plan_observations(const SvinBaWindow* w, int32_t* observation_order) {
  if (!w || !observation_order) {
    set_error("svin_ba_plan_observations: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  const int rc = validate(*w, 0);
  if (rc != SVIN_OK) return rc;
  bool group = w->num_pose_blocks <= 64;
  for (int o = 0; o < w->num_obs && group; ++o)
    if (!w->pose_fixed[w->obs_extrinsics[o]]) group = false;
  WindowOrder wo;
  order_window(*w, group, wo);
  for (int o = 0; o < w->num_obs; ++o) observation_order[o] = wo.obs_order[o];
  return SVIN_OK;
}

int svin_ba_uploaded_plan(svin_ba_ctx* c, int32_t wi, int32_t* lm_order, int32_t cap, int32_t* kind, int32_t* count,
                          int32_t* runs, int32_t* num_chunks, int32_t* observation_order, int32_t* planned_on_device) {
  if (!c || !c->uploaded || wi < 0 || wi >= c->b.B || !num_chunks) {
    set_error("svin_ba_uploaded_plan: nothing uploaded / window index out of range");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  const Batch& b = c->b;
  const WinDesc& d = c->h_win[wi];
  const int L = d.lm_end - d.lm_begin, N = d.obs_end - d.obs_begin;
  if (planned_on_device) *planned_on_device = c->device_plan ? 1 : 0;
  if (lm_order)
    for (int k = 0; k < L; ++k) lm_order[k] = c->lm_perm_p[d.lm_begin + k];
  if (observation_order && N)
    SVIN_CUDA(cudaMemcpy(observation_order, c->d_rord + d.obs_begin, 4 * (size_t)N, cudaMemcpyDeviceToHost));
  int total = 0;
  for (int k = 0; k < kSchurClasses; ++k) total += b.sw_class_count[k];
  std::vector<int> list(total > 0 ? total : 1), win(c->n_sw_cap + 1), beg(c->n_sw_cap + 1), cnt(c->n_sw_cap + 1),
      nr(c->n_sw_cap + 1);
  if (total) SVIN_CUDA(cudaMemcpy(list.data(), b.sw_list, 4 * (size_t)total, cudaMemcpyDeviceToHost));
  if (c->n_sw_cap) {
    SVIN_CUDA(cudaMemcpy(win.data(), b.sw_win, 4 * c->n_sw_cap, cudaMemcpyDeviceToHost));
    SVIN_CUDA(cudaMemcpy(beg.data(), b.sw_lm_begin, 4 * c->n_sw_cap, cudaMemcpyDeviceToHost));
    SVIN_CUDA(cudaMemcpy(cnt.data(), b.sw_count, 4 * c->n_sw_cap, cudaMemcpyDeviceToHost));
    SVIN_CUDA(cudaMemcpy(nr.data(), b.sw_nruns, 4 * c->n_sw_cap, cudaMemcpyDeviceToHost));
  }
  // this window's chunks in internal landmark order, their class from the position in the class-sorted list
  std::vector<std::pair<int, int>> mine;   // (first landmark, list position)
  for (int p = 0; p < total; ++p)
    if (list[p] >= 0 && (size_t)list[p] < c->n_sw_cap && win[list[p]] == wi) mine.push_back({beg[list[p]], p});
  std::sort(mine.begin(), mine.end());
  *num_chunks = (int)mine.size();
  if ((int)mine.size() > cap) {
    set_error("svin_ba_uploaded_plan: capacity too small");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  for (size_t k = 0; k < mine.size(); ++k) {
    const int p = mine[k].second, id = list[p];
    int cls = 0, acc = 0;
    while (cls < kSchurClasses && p >= acc + b.sw_class_count[cls]) acc += b.sw_class_count[cls++];
    if (kind) kind[k] = cls;
    if (count) count[k] = cnt[id];
    if (runs) runs[k] = nr[id];
  }
  return SVIN_OK;
}

int svin_ba_reset(svin_ba_ctx* c) {
  if (!c || !c->uploaded) {
    set_error("svin_ba_reset: nothing uploaded");
    return SVIN_ERR_STATE;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  Batch& b = c->b;
  launch_reset_state(b, c->stream);
  SVIN_CUDA(cudaMemcpyAsync(b.ws, c->d_ws_init, sizeof(WinState) * b.B, cudaMemcpyDeviceToDevice, c->stream));
  if (b.NIMU)
    SVIN_CUDA(cudaMemcpyAsync(b.imu_cache, c->d_imu_cache_init, sizeof(ImuCache) * b.NIMU, cudaMemcpyDeviceToDevice,
                              c->stream));
  if (b.shard_acc) SVIN_CUDA(cudaMemsetAsync(b.shard_acc, 0, 8 * kShardAcc * (size_t)b.B, c->stream));
  SVIN_CUDA(cudaGetLastError());
  c->solved = false;
  c->quality_valid = false;
  return SVIN_OK;
}

// SUM all-reduce of `count` doubles in place on the context's stream: NCCL between processes, or the in-process group.
static int comm_allreduce(svin_ba_ctx* c, double* buf, size_t count) {
  if (c->nccl_comm) {
    NcclApi* api = nccl_api();
    const int r = api->AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, c->nccl_comm, c->stream);
    if (r != 0) {
      set_error(std::string("ncclAllReduce failed: ") + (api->GetErrorString ? api->GetErrorString(r) : "?"));
      return SVIN_ERR_CUDA;
    }
    return SVIN_OK;
  }
  LocalGroup& g = *c->local_comm;
  const int me = c->comm_rank;
  if (count > c->local_tmp_cap) {
    set_error("svin_ba: in-process all-reduce scratch too small");
    return SVIN_ERR_STATE;
  }
  g.bufs[me] = buf;
  cudaEventRecord(g.ready[me], c->stream);
  g.barrier();  // every rank has published its buffer and recorded `ready`
  std::vector<const double*> srcs(g.bufs.begin(), g.bufs.end());
  for (int r = 0; r < g.world; ++r)
    if (r != me) cudaStreamWaitEvent(c->stream, g.ready[r], 0);
  cudaMemcpyAsync(c->local_srcs, srcs.data(), sizeof(double*) * g.world, cudaMemcpyHostToDevice, c->stream);
  cudaStreamSynchronize(c->stream);  // srcs is a stack array; the test-only path can afford the sync
  k_sum_peers<<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(c->local_tmp, c->local_srcs, g.world, count);
  cudaEventRecord(g.read[me], c->stream);
  g.barrier();  // every rank has enqueued its read of all buffers
  for (int r = 0; r < g.world; ++r)
    if (r != me) cudaStreamWaitEvent(c->stream, g.read[r], 0);
  SVIN_CUDA(cudaMemcpyAsync(buf, c->local_tmp, 8 * count, cudaMemcpyDeviceToDevice, c->stream));
  return SVIN_OK;
}

// bracket one launch with an event pair when profiling
struct ProfScope {
  svin_ba_ctx* c;
  ProfScope(svin_ba_ctx* c_, int family) : c(c_) {
    if (!c->profiling) return;
    const size_t i = c->prof_family.size();
    while (c->prof_events.size() < 2 * (i + 1)) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      c->prof_events.push_back(e);
    }
    c->prof_family.push_back(family);
    cudaEventRecord(c->prof_events[2 * i], c->stream);
  }
  ~ProfScope() {
    if (!c->profiling) return;
    cudaEventRecord(c->prof_events[2 * (c->prof_family.size() - 1) + 1], c->stream);
  }
};

static int enqueue_slot(svin_ba_ctx* c, const SvinBaOptions& opt) {
  Batch& b = c->b;
  {
    ProfScope p(c, SVIN_BA_K_CLEAR);
    SVIN_CUDA(cudaMemsetAsync(c->d_clear, 0, c->clear_bytes, c->stream));
  }
  const bool sharded = c->sharded();
  int rc;
  static const bool schur_par = !(std::getenv("SVIN_SCHUR_STREAMS") && std::atoi(std::getenv("SVIN_SCHUR_STREAMS")) == 0);
  int n_schur = 0;
  { ProfScope p(c, SVIN_BA_K_SCHUR); n_schur = launch_schur(b, opt, c->stream, schur_par ? &c->schur_par : nullptr); }
  if (sharded) {
    // exchange 1 of 3 per iteration: the reduced system [H | g_red | g_raw | Hdiag] of every window and the per-rank
    // landmark gradient maxima, one packed buffer, one all-reduce
    launch_gmax_pack(b, 0, c->stream);
    if ((rc = comm_allreduce(c, (double*)c->d_clear, c->clear_bytes / 8)) != SVIN_OK) return rc;
    launch_gmax_pack(b, 1, c->stream);
  }
  if (c->gram_pending) {
    SVIN_CUDA(cudaStreamWaitEvent(c->stream, c->ev_gram, 0));
    c->gram_pending = false;
  }
  { ProfScope p(c, SVIN_BA_K_DENSE_SOLVE); launch_dense_solve(b, opt, c->smem_bytes, c->n_max, c->stream); }
  { ProfScope p(c, SVIN_BA_K_BACKSUB); launch_backsub(b, c->stream); }
  if (sharded) {
    // exchange 2: the landmark-side sums the dogleg coefficients need (step norms, Cauchy point, model rows)
    if ((rc = comm_allreduce(c, b.shard_acc, kShardAcc * (size_t)b.B)) != SVIN_OK) return rc;
    launch_fold(b, 2, c->stream);
  }
  { ProfScope p(c, SVIN_BA_K_STEP_DENSE); launch_step_dense(b, opt, c->stream); }
  static const bool no_fork = std::getenv("SVIN_BA_NO_FORK") != nullptr;  // debugging knob
  const bool fork = !c->profiling && !no_fork;
  if (fork) {
    SVIN_CUDA(cudaEventRecord(c->ev_fork, c->stream));
    SVIN_CUDA(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
    launch_dense_eval(b, 1, 0, nullptr, c->side);
    SVIN_CUDA(cudaEventRecord(c->ev_join, c->side));
  }
  { ProfScope p(c, SVIN_BA_K_STEP_LM); launch_step_lm(b, c->stream); }
  // candidate evaluation: the reprojection terms (k_linearize) and the dense terms (k_dense_eval, a small
  // latency-bound grid) are independent -> run them concurrently on two streams.  The dense terms only need the
  // candidate poses / speed-biases, so the side stream forks right after k_step_dense; k_decide needs their cost
  // (join), the Gram matrix is only needed by the next slot's reduced-system solve and is computed after k_decide
  // (for the buffer that is current by then) beside the next slot's Schur kernels.  Serial when profiling
  // (per-kernel events).
  { ProfScope p(c, SVIN_BA_K_LINEARIZE); launch_linearize(b, 1, 0, c->stream, b.fused != 0); }
  if (sharded) {
    // exchange 3: landmark step / state norms (k_step_lm) and the candidate's reprojection cost, both only needed by k_decide
    if ((rc = comm_allreduce(c, b.shard_acc, kShardAcc * (size_t)b.B)) != SVIN_OK) return rc;
    launch_fold(b, 3, c->stream);
  }
  if (fork) {
    SVIN_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  } else {
    ProfScope p(c, SVIN_BA_K_DENSE_EVAL);
    launch_dense_eval(b, 1, 0, nullptr, c->stream);
    if (c->smem_bytes > 0) launch_dense_gram(b, 1, c->n_max, c->stream);
  }
  { ProfScope p(c, SVIN_BA_K_DECIDE); launch_decide(b, opt, c->stream); }
  if (fork && c->smem_bytes > 0) {
    SVIN_CUDA(cudaEventRecord(c->ev_fork, c->stream));
    SVIN_CUDA(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
    launch_dense_gram(b, 2, c->n_max, c->side);
    SVIN_CUDA(cudaEventRecord(c->ev_gram, c->side));
    c->gram_pending = true;
  }
  // dense solve, backsub, step (2), linearise, dense eval, Gram, decide + the Schur chunk kernels
  c->tm.kernel_launches += 8 + n_schur;
  return SVIN_OK;
}

// One pass of the solver's launch sequence: [initial evaluation] + max_num_iterations slots.
static int enqueue_pass(svin_ba_ctx* c, const SvinBaOptions& opt, bool with_init) {
  Batch& b = c->b;
  const int chunk = std::max(1, opt.max_num_iterations);
  if (with_init) {
    { ProfScope p(c, SVIN_BA_K_LINEARIZE); launch_linearize(b, 0, 0, c->stream, b.fused != 0); }
    if (c->sharded()) {
      const int rc0 = comm_allreduce(c, b.shard_acc, kShardAcc * (size_t)b.B);
      if (rc0 != SVIN_OK) return rc0;
      launch_fold(b, 3, c->stream);
    }
    {
      ProfScope p(c, SVIN_BA_K_DENSE_EVAL);
      launch_dense_eval(b, 0, 0, nullptr, c->stream);
      if (c->smem_bytes > 0) launch_dense_gram(b, 0, c->n_max, c->stream);
    }
    launch_init(b, opt, c->stream);
    c->tm.kernel_launches += 4;  // linearise, dense eval, Gram, init
  }
  for (int s = 0; s < chunk; ++s) {
    const int rc = enqueue_slot(c, opt);
    if (rc != SVIN_OK) return rc;
  }
  if (c->gram_pending) {
    SVIN_CUDA(cudaStreamWaitEvent(c->stream, c->ev_gram, 0));
    c->gram_pending = false;
  }
  return SVIN_OK;
}

static bool graph_wanted() {
  static const bool w = !(std::getenv("SVIN_BA_GRAPH") && std::atoi(std::getenv("SVIN_BA_GRAPH")) == 0);
  return w;
}
// Capturing + instantiating the ~190-node graph costs the host about as much as a small batch takes to solve: a graph
// pays when the batch is solved more than once per upload or is big enough that launch gaps matter.  Below this many
// windows a fresh upload is solved with plain stream launches (SVIN_BA_GRAPH_MIN_WINDOWS, measured in DESIGN.md §3.5).
static int graph_min_windows() {
  static const int v = std::getenv("SVIN_BA_GRAPH_MIN_WINDOWS") ? std::atoi(std::getenv("SVIN_BA_GRAPH_MIN_WINDOWS")) : 8;
  return v;
}
// Capture the first pass for the current upload (kernel arguments hold the batch by value) and instantiate it.
static int build_graph(svin_ba_ctx* c, const SvinBaOptions& opt) {
  if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
  c->graph_exec = nullptr;
  c->graph_valid = false;
  const long long l0 = c->tm.kernel_launches;
  cudaGraph_t g = nullptr;
  SVIN_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  const int rc = enqueue_pass(c, opt, true);
  const cudaError_t ce = cudaStreamEndCapture(c->stream, &g);
  c->graph_launches = c->tm.kernel_launches - l0;
  c->tm.kernel_launches = l0;
  if (rc != SVIN_OK) return rc;
  SVIN_CUDA(ce);
  const cudaError_t ie = cudaGraphInstantiate(&c->graph_exec, g, 0);
  cudaGraphDestroy(g);
  SVIN_CUDA(ie);
  c->graph_opt = opt;
  c->graph_valid = true;
  c->graph_opt_known = true;
  return SVIN_OK;
}

int svin_ba_solve(svin_ba_ctx* c, const SvinBaOptions* opt_in, SvinBaSummary* summaries) {
  if (!c || !c->uploaded) {
    set_error("svin_ba_solve: nothing uploaded");
    return SVIN_ERR_STATE;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  SvinBaOptions opt;
  if (opt_in)
    opt = *opt_in;
  else
    svin_ba_default_options(&opt);
  Batch& b = c->b;
  if (c->sharded() && opt.time_limit_seconds >= 0.0) {
    // k_decide stops a window on each rank's own clock: ranks could leave the iteration loop in different slots and
    // stop issuing the collectives their peers wait for.  The replicated decisions must not depend on local time.
    set_error("svin_ba_solve: time_limit_seconds is not supported in sharded mode (set it < 0)");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  if (c->solved) {
    const int rc = svin_ba_reset(c);
    if (rc != SVIN_OK) return rc;
  }
  SVIN_CUDA(cudaEventRecord(c->ev[2], c->stream));
  // initial evaluation (IterationZero); k_init also installs the options' initial trust-region radius
  c->prof_family.clear();
  const int chunk = std::max(1, opt.max_num_iterations);
  const int hard_cap = opt.max_num_iterations * 10 + 16;
  // The first pass (initial evaluation + max_num_iterations slots) is a fixed launch sequence - all control flow is
  // on the device - so it is captured once per upload into a CUDA graph and replayed (SVIN_BA_GRAPH=0 disables).
  // svin_ba_upload pre-builds it with the options of the previous solve, outside the caller's GPU critical section.
  // A small batch only gets a graph from its second solve on.  Sharded mode runs with plain stream launches: a captured
  // pass with its ncclAllReduce nodes deadlocked at replay on this image (NCCL 2.28.9, 2 ranks, profiles/r2u_*), the
  // same launch sequence without capture runs fine.
  const bool use_graph = graph_wanted() && !c->profiling && !c->sharded() &&
                         (b.B >= graph_min_windows() || c->solves_since_upload > 0);
  int slots_done = 0;
  bool first = true;
  while (true) {
    if (first && use_graph) {
      if (!c->graph_exec || !c->graph_valid || std::memcmp(&c->graph_opt, &opt, sizeof opt) != 0) {
        const int rc = build_graph(c, opt);
        if (rc != SVIN_OK) return rc;
      }
      SVIN_CUDA(cudaGraphLaunch(c->graph_exec, c->stream));
      c->tm.kernel_launches += c->graph_launches;
    } else {
      const int rc = enqueue_pass(c, opt, first);
      if (rc != SVIN_OK) return rc;
    }
    first = false;
    slots_done += chunk;
    SVIN_CUDA(cudaMemsetAsync(c->d_active, 0, sizeof(int), c->stream));
    launch_count_active(b, c->d_active, c->stream);
    c->tm.kernel_launches += 1;
    SVIN_CUDA(cudaMemcpyAsync(c->h_active, c->d_active, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SVIN_CUDA(cudaStreamSynchronize(c->stream));
    SVIN_CUDA(cudaGetLastError());
    if (*c->h_active == 0) break;
    if (slots_done >= hard_cap) {
      set_error("svin_ba_solve: windows still active after the slot cap (solver did not terminate)");
      return SVIN_ERR_STATE;
    }
  }
  if (opt.compute_landmark_quality) {
    launch_quality(b, c->stream);
    c->tm.kernel_launches += 1;
    c->quality_valid = true;
  }
  SVIN_CUDA(cudaEventRecord(c->ev[3], c->stream));
  SVIN_CUDA(cudaMemcpyAsync((char*)c->h_out + c->out_off_ws, b.ws, sizeof(WinState) * b.B, cudaMemcpyDeviceToHost,
                            c->stream));
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  SVIN_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]);
  c->tm.solve_ms = ms;
  c->solved = true;
  c->solves_since_upload += 1;
  if (c->profiling) {
    c->ktimes = SvinBaKernelTimes{};
    for (size_t i = 0; i < c->prof_family.size(); ++i) {
      float t = 0;
      cudaEventElapsedTime(&t, c->prof_events[2 * i], c->prof_events[2 * i + 1]);
      c->ktimes.ms[c->prof_family[i]] += t;
      c->ktimes.launches[c->prof_family[i]] += 1;
    }
  }
  if (summaries) {
    const WinState* ws = (const WinState*)((char*)c->h_out + c->out_off_ws);
    for (int i = 0; i < b.B; ++i) {
      SvinBaSummary& s = summaries[i];
      s.iterations = ws[i].iter;
      s.num_successful_steps = ws[i].num_successful;
      s.termination = ws[i].termination;
      s.imu_repropagations = ws[i].imu_redo;
      s.initial_cost = ws[i].initial_cost;
      s.final_cost = ws[i].cost_x;
      s.final_trust_region_radius = ws[i].radius;
    }
  }
  return SVIN_OK;
}

static int fetch_results(svin_ba_ctx* c) {
  Batch& b = c->b;
  launch_gather_state(b, c->d_pose_win, c->d_sb_win, c->d_pose_out, c->d_sb_out, c->d_lm_out, c->stream);
  SVIN_CUDA(cudaEventRecord(c->ev[4], c->stream));
  SVIN_CUDA(cudaMemcpyAsync(c->h_out, c->d_out, c->out_off_ws, cudaMemcpyDeviceToHost, c->stream));
  SVIN_CUDA(cudaEventRecord(c->ev[5], c->stream));
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  SVIN_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]);
  c->tm.d2h_ms = ms;
  c->tm.d2h_bytes = (int64_t)c->out_off_ws;
  return SVIN_OK;
}

static void scatter_window(svin_ba_ctx* c, int i, SvinBaWindow* w, double* quality) {
  const WinDesc& d = c->h_win[i];
  const char* Hh = (const char*)c->h_out;
  std::memcpy(w->pose_blocks, (const double*)(Hh + c->out_off_pose) + 7 * (size_t)d.pose_begin,
              56 * (size_t)(d.pose_end - d.pose_begin));
  std::memcpy(w->speedbias, (const double*)(Hh + c->out_off_sb) + 9 * (size_t)d.sb_begin,
              72 * (size_t)(d.sb_end - d.sb_begin));
  const double* lm_out = (const double*)(Hh + c->out_off_lm) + 4 * (size_t)d.lm_begin;
  const double* q_out = (const double*)(Hh + c->out_off_q) + d.lm_begin;
  const int* perm = c->lm_perm_p + d.lm_begin;
  for (int k = 0; k < d.lm_end - d.lm_begin; ++k) {
    std::memcpy(w->landmarks + 4 * (size_t)perm[k], lm_out + 4 * (size_t)k, 32);
    // without compute_landmark_quality the device buffer holds an earlier batch's values (or nothing): report "unknown"
    if (quality) quality[perm[k]] = c->quality_valid ? q_out[k] : std::nan("");
  }
}

int svin_ba_download(svin_ba_ctx* c, int32_t i, SvinBaWindow* w, double* quality) {
  if (!c || !c->uploaded || !w || i < 0 || i >= c->b.B) {
    set_error("svin_ba_download: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  const WinDesc& d = c->h_win[i];
  if (w->num_pose_blocks != d.pose_end - d.pose_begin || w->num_speedbias != d.sb_end - d.sb_begin ||
      w->num_landmarks != d.lm_end - d.lm_begin) {
    set_error("svin_ba_download: window shape differs from the uploaded one");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  const int rc = fetch_results(c);
  if (rc != SVIN_OK) return rc;
  scatter_window(c, i, w, quality);
  return SVIN_OK;
}

int svin_ba_download_all(svin_ba_ctx* c, SvinBaWindow* wins, int32_t B, double* const* quality) {
  if (!c || !c->uploaded || !wins || B != c->b.B) {
    set_error("svin_ba_download_all: invalid arguments (num_windows must equal the uploaded batch size)");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  for (int i = 0; i < B; ++i) {
    const WinDesc& d = c->h_win[i];
    if (wins[i].num_pose_blocks != d.pose_end - d.pose_begin || wins[i].num_speedbias != d.sb_end - d.sb_begin ||
        wins[i].num_landmarks != d.lm_end - d.lm_begin) {
      set_error("svin_ba_download_all: window shape differs from the uploaded one");
      return SVIN_ERR_INVALID_ARGUMENT;
    }
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  const int rc = fetch_results(c);
  if (rc != SVIN_OK) return rc;
  const double t0 = wall_ms();
  c->pool->run(B, [&](int i) { scatter_window(c, i, &wins[i], quality ? quality[i] : nullptr); });
  c->tm.host_scatter_ms = wall_ms() - t0;
  return SVIN_OK;
}

int svin_ba_optimize(svin_ba_ctx* c, SvinBaWindow* wins, int32_t B, const SvinBaOptions* opt, SvinBaSummary* summaries,
                     double* const* quality) {
  int rc = svin_ba_upload(c, wins, B);
  if (rc != SVIN_OK) return rc;
  rc = svin_ba_solve(c, opt, summaries);
  if (rc != SVIN_OK) return rc;
  return svin_ba_download_all(c, wins, B, quality);
}

int svin_ba_evaluate(svin_ba_ctx* c, int32_t wi, SvinBaEvaluation* out) {
  if (!c || !c->uploaded || !out || wi < 0 || wi >= c->b.B) {
    set_error("svin_ba_evaluate: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  Batch& b = c->b;
  const WinDesc& d = c->h_win[wi];
  const int N = d.obs_end - d.obs_begin, NI = d.imu_end - d.imu_begin;
  if (!c->obs_perm_valid) {   // device planner: the observation order lives on the device until someone asks
    c->obs_perm.resize(c->n_obs_total);
    if (c->n_obs_total)
      SVIN_CUDA(cudaMemcpy(c->obs_perm.data(), c->d_rord, 4 * c->n_obs_total, cudaMemcpyDeviceToHost));
    c->obs_perm_valid = true;
  }
  // total cost at the current estimate: regular evaluation, then restore the state
  int rc = svin_ba_reset(c);
  if (rc != SVIN_OK) return rc;
  launch_linearize(b, 0, 0, c->stream);
  launch_dense_eval(b, 0, 0, nullptr, c->stream);
  double cost = 0;
  SVIN_CUDA(cudaMemcpyAsync(&cost, &b.ws[wi].cost_cand, 8, cudaMemcpyDeviceToHost, c->stream));
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  if (out->cost) out->cost[0] = cost;
  rc = svin_ba_reset(c);
  if (rc != SVIN_OK) return rc;
  // raw dump into the inactive buffers
  double* imu_dump = nullptr;
  const size_t imu_per = 15 + 90 + 135 + 90 + 135;
  if (b.NIMU) SVIN_CUDA(cudaMalloc(&imu_dump, 8 * imu_per * (size_t)b.NIMU));
  const double* dump[5] = {imu_dump, imu_dump ? imu_dump + 15 * (size_t)b.NIMU : nullptr,
                           imu_dump ? imu_dump + (15 + 90) * (size_t)b.NIMU : nullptr,
                           imu_dump ? imu_dump + (15 + 90 + 135) * (size_t)b.NIMU : nullptr,
                           imu_dump ? imu_dump + (15 + 90 + 135 + 90) * (size_t)b.NIMU : nullptr};
  // the solver only materialises the extrinsics Jacobian when some extrinsics block is estimated;
  // for the dump it is always produced, into a temporary plane set
  const size_t S = b.obs_stride;
  double* je_tmp = nullptr;
  Batch braw = b;
  if (!b.has_ext) {
    SVIN_CUDA(cudaMalloc(&je_tmp, 8 * 12 * S + 8));
    braw.lin_Je[1] = je_tmp;
  }
  launch_linearize(braw, 1, 1, c->stream);
  launch_dense_eval(b, 1, 1, dump, c->stream);
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  SVIN_CUDA(cudaGetLastError());
  std::vector<double> plane(N ? (size_t)N : 1);
  auto pull = [&](const double* base, int planes, double* dst, int stride) -> int {
    if (!dst) return SVIN_OK;
    for (int k = 0; k < planes; ++k) {
      SVIN_CUDA(cudaMemcpy(plane.data(), base + k * S + d.obs_begin, 8 * (size_t)N, cudaMemcpyDeviceToHost));
      for (int o = 0; o < N; ++o) dst[(size_t)c->obs_perm[(size_t)d.obs_begin + o] * stride + k] = plane[o];
    }
    return SVIN_OK;
  };
  // inactive buffer index of this window: state was reset so cur == 0 -> inactive == 1
  if ((rc = pull(b.lin_r[1], 2, out->reproj_residuals, 2)) != SVIN_OK) return rc;
  if ((rc = pull(b.lin_Jp[1], 12, out->reproj_J_pose, 12)) != SVIN_OK) return rc;
  if ((rc = pull(b.lin_Jl[1], 6, out->reproj_J_landmark, 6)) != SVIN_OK) return rc;
  if ((rc = pull(braw.lin_Je[1], 12, out->reproj_J_extrinsics, 12)) != SVIN_OK) return rc;
  if (je_tmp) cudaFree(je_tmp);
  if (NI) {
    auto pull_imu = [&](const double* src, int per, double* dst) -> int {
      if (!dst) return SVIN_OK;
      SVIN_CUDA(cudaMemcpy(dst, src + (size_t)per * d.imu_begin, 8 * (size_t)per * NI, cudaMemcpyDeviceToHost));
      return SVIN_OK;
    };
    if ((rc = pull_imu(dump[0], 15, out->imu_residuals)) != SVIN_OK) return rc;
    if ((rc = pull_imu(dump[1], 90, out->imu_J_pose0)) != SVIN_OK) return rc;
    if ((rc = pull_imu(dump[2], 135, out->imu_J_speedbias0)) != SVIN_OK) return rc;
    if ((rc = pull_imu(dump[3], 90, out->imu_J_pose1)) != SVIN_OK) return rc;
    if ((rc = pull_imu(dump[4], 135, out->imu_J_speedbias1)) != SVIN_OK) return rc;
  }
  if (imu_dump) cudaFree(imu_dump);
  return svin_ba_reset(c);
}

int svin_ba_marginalize(svin_ba_ctx* c, int32_t wi, const SvinMargSpec* spec, SvinMargResult* out) {
  if (!c || !c->uploaded || !spec || !out || wi < 0 || wi >= c->b.B) {
    set_error("svin_ba_marginalize: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  if (!out->block_kind || !out->block_index || !out->H || !out->b0 || !out->J || !out->e0) {
    set_error("svin_ba_marginalize: result arrays are NULL");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  Batch& b = c->b;
  const WinDesc& d = c->h_win[wi];
  const int n = d.n_dense;
  const int npb = d.pose_end - d.pose_begin, nsb = d.sb_end - d.sb_begin;
  // dense offsets of this window's blocks (engine order: free pose blocks, then free speed/bias blocks)
  std::vector<int> pose_off(npb > 0 ? npb : 1), sb_off(nsb > 0 ? nsb : 1);
  if (npb) SVIN_CUDA(cudaMemcpy(pose_off.data(), b.pose_off + d.pose_begin, 4 * (size_t)npb, cudaMemcpyDeviceToHost));
  if (nsb) SVIN_CUDA(cudaMemcpy(sb_off.data(), b.sb_off + d.sb_begin, 4 * (size_t)nsb, cudaMemcpyDeviceToHost));
  std::vector<int> keep_idx, marg_idx, prior_map;
  out->num_blocks = 0;
  for (int i = 0; i < npb; ++i)
    if (pose_off[i] >= 0) {
      const bool mg = spec->marginalize_pose && spec->marginalize_pose[i];
      for (int k = 0; k < 6; ++k) (mg ? marg_idx : keep_idx).push_back(pose_off[i] + k);
      if (!mg) {
        out->block_kind[out->num_blocks] = SVIN_BLOCK_POSE;
        out->block_index[out->num_blocks++] = i;
      }
    }
  for (int i = 0; i < nsb; ++i)
    if (sb_off[i] >= 0) {
      const bool mg = spec->marginalize_speedbias && spec->marginalize_speedbias[i];
      for (int k = 0; k < 9; ++k) (mg ? marg_idx : keep_idx).push_back(sb_off[i] + k);
      if (!mg) {
        out->block_kind[out->num_blocks] = SVIN_BLOCK_SPEEDBIAS;
        out->block_index[out->num_blocks++] = i;
      }
    }
  for (int k = 0; k < spec->prior_num_blocks; ++k) {
    const int kind = spec->prior_block_kind[k], idx = spec->prior_block_index[k];
    const bool pose = kind == SVIN_BLOCK_POSE;
    if (idx < 0 || idx >= (pose ? npb : nsb) || (pose ? pose_off[idx] : sb_off[idx]) < 0) {
      set_error("svin_ba_marginalize: prior block is not an estimated block of the window");
      return SVIN_ERR_INVALID_ARGUMENT;
    }
    const int off = pose ? pose_off[idx] : sb_off[idx];
    for (int k2 = 0; k2 < (pose ? 6 : 9); ++k2) prior_map.push_back(off + k2);
  }
  if ((int)prior_map.size() != spec->prior_dim) {
    set_error("svin_ba_marginalize: prior_dim does not match the prior block list");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  const int nk = (int)keep_idx.size(), nm = (int)marg_idx.size(), pd = spec->prior_dim;
  out->dim = nk;
  if (nk == 0) return SVIN_OK;
  // linearise at the uploaded (first-estimate) points
  int rc = svin_ba_reset(c);
  if (rc != SVIN_OK) return rc;
  launch_linearize(b, 0, 0, c->stream);
  launch_dense_eval(b, 0, 0, nullptr, c->stream);
  // scratch arena: doubles then ints
  const size_t nmax = (size_t)(nk > nm ? nk : nm);
  const size_t nd = (size_t)n * n + n                 // H, b
                    + 2 * nmax * nmax + nmax          // A, U, ev
                    + (size_t)nm * nm + 2 * (size_t)nk * nm + 2 * nmax + 4  // Vp, Wm, WV (also (c,s) scratch)
                    + n + 8                           // pvec (+ the cluster Jacobi convergence slots)
                    + (size_t)pd * pd + pd            // prior
                    + 2 * (size_t)nk * nk + 2 * nk;   // Hk, J, bk, e0
  const size_t ni = (size_t)nk + nm + pd + nmax + 2;
  void* scratch = nullptr;
  SVIN_CUDA(cudaMalloc(&scratch, 8 * nd + 4 * ni + 64));
  double* p = static_cast<double*>(scratch);
  MargArgs m{};
  m.w = wi; m.n = n; m.nk = nk; m.nm = nm; m.prior_dim = pd;
  m.H = p; p += (size_t)n * n;
  m.b = p; p += n;
  m.A = p; p += nmax * nmax;
  m.U = p; p += nmax * nmax;
  m.ev = p; p += nmax;
  m.Vp = p; p += (size_t)nm * nm;
  m.Wm = p; p += (size_t)nk * nm;
  m.WV = p; p += (size_t)nk * nm + 2 * nmax + 4;
  m.pvec = p; p += n + 8;
  double* d_prior_H = p; p += (size_t)pd * pd;
  double* d_prior_b = p; p += pd;
  m.Hk = p; p += (size_t)nk * nk;
  m.J = p; p += (size_t)nk * nk;
  m.bk = p; p += nk;
  m.e0 = p; p += nk;
  int* ip = reinterpret_cast<int*>(p);
  int* d_keep = ip; ip += nk;
  int* d_marg = ip; ip += nm;
  int* d_map = ip; ip += pd;
  m.players = ip;
  m.keep_idx = d_keep; m.marg_idx = d_marg; m.prior_map = d_map;
  m.prior_H = d_prior_H; m.prior_b = d_prior_b;
  auto fail = [&](cudaError_t e) {
    set_error(std::string("svin_ba_marginalize: ") + cudaGetErrorString(e));
    cudaFree(scratch);
    return SVIN_ERR_CUDA;
  };
  cudaError_t e;
  if ((e = cudaMemsetAsync(m.H, 0, 8 * ((size_t)n * n + n), c->stream)) != cudaSuccess) return fail(e);
  if ((e = cudaMemcpyAsync(d_keep, keep_idx.data(), 4 * (size_t)nk, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return fail(e);
  if (nm && (e = cudaMemcpyAsync(d_marg, marg_idx.data(), 4 * (size_t)nm, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return fail(e);
  if (pd) {
    if ((e = cudaMemcpyAsync(d_map, prior_map.data(), 4 * (size_t)pd, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return fail(e);
    if ((e = cudaMemcpyAsync(d_prior_H, spec->prior_H, 8 * (size_t)pd * pd, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return fail(e);
    if ((e = cudaMemcpyAsync(d_prior_b, spec->prior_b0, 8 * (size_t)pd, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return fail(e);
  }
  launch_marg(b, m, d.lm_end - d.lm_begin, c->stream);
  if ((e = cudaGetLastError()) != cudaSuccess) return fail(e);
  if ((e = cudaMemcpyAsync(out->H, m.Hk, 8 * (size_t)nk * nk, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) return fail(e);
  if ((e = cudaMemcpyAsync(out->J, m.J, 8 * (size_t)nk * nk, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) return fail(e);
  if ((e = cudaMemcpyAsync(out->b0, m.bk, 8 * (size_t)nk, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) return fail(e);
  if ((e = cudaMemcpyAsync(out->e0, m.e0, 8 * (size_t)nk, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) return fail(e);
  if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return fail(e);
  cudaFree(scratch);
  return svin_ba_reset(c);
}

int svin_nccl_unique_id(uint8_t out[128]) {
  NcclApi* api = nccl_api();
  if (!api) {
    set_error("libnccl.so.2 could not be loaded (dlopen)");
    return SVIN_ERR_STATE;
  }
  NcclId id;
  const int r = api->GetUniqueId(&id);
  if (r != 0) {
    set_error(std::string("ncclGetUniqueId failed: ") + (api->GetErrorString ? api->GetErrorString(r) : "?"));
    return SVIN_ERR_CUDA;
  }
  std::memcpy(out, id.internal, 128);
  return SVIN_OK;
}

int svin_ba_comm_init(svin_ba_ctx* c, const uint8_t unique_id[128], int32_t rank, int32_t world) {
  if (!c || !unique_id || world < 1 || rank < 0 || rank >= world) {
    set_error("svin_ba_comm_init: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  NcclApi* api = nccl_api();
  if (!api) {
    set_error("libnccl.so.2 could not be loaded (dlopen)");
    return SVIN_ERR_STATE;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  NcclId id;
  std::memcpy(id.internal, unique_id, 128);
  void* comm = nullptr;
  const int r = api->CommInitRank(&comm, world, id, rank);
  if (r != 0) {
    set_error(std::string("ncclCommInitRank failed: ") + (api->GetErrorString ? api->GetErrorString(r) : "?"));
    return SVIN_ERR_CUDA;
  }
  c->nccl_comm = comm;
  c->comm_rank = rank;
  c->comm_world = world;
  c->uploaded = false;  // the arena layout depends on the mode
  return SVIN_OK;
}

int svin_ba_comm_init_local(svin_ba_ctx* const* ctxs, int32_t world) {
  if (!ctxs || world < 1 || world > 64) {
    set_error("svin_ba_comm_init_local: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  for (int r = 0; r < world; ++r)
    if (!ctxs[r] || ctxs[r]->device != ctxs[0]->device || ctxs[r]->sharded()) {
      set_error("svin_ba_comm_init_local: contexts must be distinct, on one device and without a communicator");
      return SVIN_ERR_INVALID_ARGUMENT;
    }
  SVIN_CUDA(cudaSetDevice(ctxs[0]->device));
  auto g = std::make_shared<LocalGroup>();
  g->world = world;
  g->bufs.assign(world, nullptr);
  g->ready.resize(world);
  g->read.resize(world);
  for (int r = 0; r < world; ++r) {
    SVIN_CUDA(cudaEventCreateWithFlags(&g->ready[r], cudaEventDisableTiming));
    SVIN_CUDA(cudaEventCreateWithFlags(&g->read[r], cudaEventDisableTiming));
  }
  for (int r = 0; r < world; ++r) {
    svin_ba_ctx* c = ctxs[r];
    c->local_comm = g;
    c->comm_rank = r;
    c->comm_world = world;
    c->uploaded = false;  // the arena layout depends on the mode
    SVIN_CUDA(cudaMalloc(&c->local_srcs, sizeof(double*) * world));
  }
  return SVIN_OK;
}

int svin_ba_set_profiling(svin_ba_ctx* c, int enable) {
  if (!c) {
    set_error("svin_ba_set_profiling: ctx is NULL");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  c->profiling = enable != 0;
  return SVIN_OK;
}

int svin_ba_kernel_times(svin_ba_ctx* c, SvinBaKernelTimes* out) {
  if (!c || !out) {
    set_error("svin_ba_kernel_times: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  *out = c->ktimes;
  return SVIN_OK;
}

int svin_ba_timings(svin_ba_ctx* c, SvinBaTimings* out) {
  if (!c || !out) {
    set_error("svin_ba_timings: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  *out = c->tm;
  return SVIN_OK;
}

}  // extern "C"
