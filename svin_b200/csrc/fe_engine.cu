// Host side of the front-end engine: svin_fe_* / svin_match C ABI (include/svin_b200.h).
//
// Reference seams: Frame::detect / Frame::describe (okvis_cv/include/okvis/implementation/Frame.hpp:93-135) behind
// Frontend::detectAndDescribe (okvis_frontend/src/Frontend.cpp:91-113), and DenseMatcher::match over a
// VioKeyframeWindowMatchingAlgorithm (okvis_matcher/include/okvis/implementation/DenseMatcher.hpp:195-203).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.hpp"
#include "host_pool.hpp"
#include "fe_kernels.cuh"

using namespace svin;

namespace {

constexpr int kNumPoints = 60, kNumRot = 1024, kNumPairs = 384, kBorder = 16;

// The declared BRISK-2 sampling pattern: ring radii {0,2.9,4.9,7.4,10.8}*0.85 with
// {1,10,14,15,20} points, box half-widths {1,1,2,3,3}, 1024 rotations rounded to integer pixel offsets,
// and the 384 shortest point pairs (ties by (i,j)).
struct HostPattern {
  std::vector<int8_t> dx, dy;
  std::vector<int> half, pi, pj;
};
HostPattern make_pattern() {
  HostPattern P;
  const double radii[5] = {0.0, 2.9, 4.9, 7.4, 10.8};
  const int counts[5] = {1, 10, 14, 15, 20};
  const int halves[5] = {1, 1, 2, 3, 3};
  std::vector<double> px, py;
  for (int ring = 0; ring < 5; ++ring)
    for (int k = 0; k < counts[ring]; ++k) {
      const double a = 2.0 * M_PI * (double)k / (double)counts[ring];
      px.push_back(0.85 * radii[ring] * std::cos(a));
      py.push_back(0.85 * radii[ring] * std::sin(a));
      P.half.push_back(halves[ring]);
    }
  P.dx.resize((size_t)kNumRot * kNumPoints);
  P.dy.resize((size_t)kNumRot * kNumPoints);
  for (int r = 0; r < kNumRot; ++r) {
    const double th = 2.0 * M_PI * (double)r / (double)kNumRot;
    const double c = std::cos(th), s = std::sin(th);
    for (int i = 0; i < kNumPoints; ++i) {
      P.dx[(size_t)r * kNumPoints + i] = (int8_t)std::lround(c * px[i] - s * py[i]);
      P.dy[(size_t)r * kNumPoints + i] = (int8_t)std::lround(s * px[i] + c * py[i]);
    }
  }
  struct PairD {
    long long d;
    int i, j;
  };
  std::vector<PairD> all;
  for (int i = 1; i < kNumPoints; ++i)
    for (int j = 0; j < i; ++j) {
      const double ddx = px[i] - px[j], ddy = py[i] - py[j];
      all.push_back(PairD{(long long)std::llround(std::sqrt(ddx * ddx + ddy * ddy) * 1e9), i, j});
    }
  std::sort(all.begin(), all.end(), [](const PairD& a, const PairD& b) {
    if (a.d != b.d) return a.d < b.d;
    if (a.i != b.i) return a.i < b.i;
    return a.j < b.j;
  });
  for (int k = 0; k < kNumPairs; ++k) {
    P.pi.push_back(all[k].i);
    P.pj.push_back(all[k].j);
  }
  return P;
}

}  // namespace

struct svin_fe_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  SvinFeOptions opt{};
  FeBatch f{};
  int n_images = 0;
  int pitch = 0;
  size_t occ_bytes = 0;
  bool use_tma = true;
  svin::HostPool* pool = nullptr;
  // device buffers
  uint8_t* d_images = nullptr;
  double *d_intr = nullptr, *d_edir = nullptr;
  int* d_scores = nullptr;
  unsigned* d_cand_count = nullptr;
  unsigned long long* d_cand_keys = nullptr;
  int *d_kept_xy = nullptr, *d_kept_score = nullptr, *d_kept_count = nullptr;
  SvinKeypoint* d_kps = nullptr;
  uint8_t* d_desc = nullptr;
  int8_t *d_pdx = nullptr, *d_pdy = nullptr;
  int *d_half = nullptr, *d_pi = nullptr, *d_pj = nullptr;
  // pinned staging
  uint8_t* h_images = nullptr;
  double* h_small = nullptr;  // intrinsics + extraction dirs
  SvinKeypoint* h_kps = nullptr;
  uint8_t* h_desc = nullptr;
  int* h_counts = nullptr;
  // matching arena (grown on demand)
  void *d_match = nullptr, *h_match = nullptr;
  size_t match_cap = 0;
  SvinFeTimings tm{};
};

extern "C" {

void svin_fe_default_options(SvinFeOptions* o) {
  if (!o) return;
  o->image_width = 752;
  o->image_height = 480;
  o->detection_threshold = 40.0;  // config_fpga_p2_euroc.yaml:66
  o->detection_octaves = 0;       // :67
  o->absolute_threshold = 800.0;  // Frontend.cpp:75
  o->max_keypoints = 400;         // config :68
  o->rotation_invariance = 1;     // Frontend.cpp:77
  o->scale_invariance = 0;        // Frontend.cpp:78
  o->max_images = 2;
}

int svin_fe_create(int device, const SvinFeOptions* opt_in, svin_fe_ctx** out) {
  if (!out) {
    set_error("svin_fe_create: out is NULL");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  *out = nullptr;
  SvinFeOptions opt;
  if (opt_in)
    opt = *opt_in;
  else
    svin_fe_default_options(&opt);
  if (opt.detection_octaves != 0 || opt.scale_invariance != 0) {
    set_error("svin_fe_create: only octaves = 0 and scale_invariance = 0 are implemented (both shipped configs use them)");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  if (!opt.rotation_invariance) {
    set_error("svin_fe_create: rotation_invariance = 0 is not implemented (Frontend.cpp:77 sets it to true)");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  if (opt.image_width < 64 || opt.image_height < 64 || opt.max_keypoints < 1 || opt.max_images < 1) {
    set_error("svin_fe_create: invalid image size / capacities");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    set_error("no CUDA device available: the svin_b200 engine has no CPU fallback");
    return SVIN_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) {
    set_error("device index out of range");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(cudaSetDevice(device));
  svin_fe_ctx* c = new svin_fe_ctx();
  c->pool = new svin::HostPool(std::max(0, std::min(16, (int)std::thread::hardware_concurrency()) - 1));
  c->device = device;
  c->opt = opt;
  const int W = opt.image_width, H = opt.image_height, M = opt.max_images, K = opt.max_keypoints;
  c->pitch = (W + 15) & ~15;
  SVIN_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto& e : c->ev) SVIN_CUDA(cudaEventCreate(&e));
  int cap = 1;
  while (cap < (W * H) / 4 + 1024) cap <<= 1;
  SVIN_CUDA(cudaMalloc(&c->d_images, (size_t)M * H * c->pitch + 256));  // +pad: row segments may overrun by < 16 B
  SVIN_CUDA(cudaMalloc(&c->d_intr, sizeof(double) * 8 * M));
  SVIN_CUDA(cudaMalloc(&c->d_edir, sizeof(double) * 3 * M));
  SVIN_CUDA(cudaMalloc(&c->d_scores, sizeof(int) * (size_t)M * H * W));
  SVIN_CUDA(cudaMalloc(&c->d_cand_count, sizeof(unsigned) * M));
  SVIN_CUDA(cudaMalloc(&c->d_cand_keys, sizeof(unsigned long long) * (size_t)M * cap));
  SVIN_CUDA(cudaMalloc(&c->d_kept_xy, sizeof(int) * 2 * (size_t)M * K));
  SVIN_CUDA(cudaMalloc(&c->d_kept_score, sizeof(int) * (size_t)M * K));
  SVIN_CUDA(cudaMalloc(&c->d_kept_count, sizeof(int) * M));
  SVIN_CUDA(cudaMalloc(&c->d_kps, sizeof(SvinKeypoint) * (size_t)M * K));
  SVIN_CUDA(cudaMalloc(&c->d_desc, (size_t)M * K * 48));
  SVIN_CUDA(cudaMemset(c->d_kept_count, 0, sizeof(int) * M));
  const HostPattern P = make_pattern();
  SVIN_CUDA(cudaMalloc(&c->d_pdx, P.dx.size()));
  SVIN_CUDA(cudaMalloc(&c->d_pdy, P.dy.size()));
  SVIN_CUDA(cudaMalloc(&c->d_half, sizeof(int) * kNumPoints));
  SVIN_CUDA(cudaMalloc(&c->d_pi, sizeof(int) * kNumPairs));
  SVIN_CUDA(cudaMalloc(&c->d_pj, sizeof(int) * kNumPairs));
  SVIN_CUDA(cudaMemcpy(c->d_pdx, P.dx.data(), P.dx.size(), cudaMemcpyHostToDevice));
  SVIN_CUDA(cudaMemcpy(c->d_pdy, P.dy.data(), P.dy.size(), cudaMemcpyHostToDevice));
  SVIN_CUDA(cudaMemcpy(c->d_half, P.half.data(), sizeof(int) * kNumPoints, cudaMemcpyHostToDevice));
  SVIN_CUDA(cudaMemcpy(c->d_pi, P.pi.data(), sizeof(int) * kNumPairs, cudaMemcpyHostToDevice));
  SVIN_CUDA(cudaMemcpy(c->d_pj, P.pj.data(), sizeof(int) * kNumPairs, cudaMemcpyHostToDevice));
  SVIN_CUDA(cudaMallocHost(&c->h_images, (size_t)M * H * c->pitch));
  SVIN_CUDA(cudaMallocHost(&c->h_small, sizeof(double) * 11 * M));
  SVIN_CUDA(cudaMallocHost(&c->h_kps, sizeof(SvinKeypoint) * (size_t)M * K));
  SVIN_CUDA(cudaMallocHost(&c->h_desc, (size_t)M * K * 48));
  SVIN_CUDA(cudaMallocHost(&c->h_counts, sizeof(int) * M));
  FeBatch& f = c->f;
  f.W = W; f.H = H; f.pitch = c->pitch; f.max_kp = K; f.border = kBorder; f.cand_cap = cap;
  f.abs_threshold = (int)opt.absolute_threshold;
  f.uniformity_radius = opt.detection_threshold;
  f.images = c->d_images; f.intrinsics = c->d_intr; f.extraction_dir = c->d_edir; f.scores = c->d_scores;
  f.cand_count = c->d_cand_count; f.cand_keys = c->d_cand_keys; f.kept_xy = c->d_kept_xy;
  f.kept_score = c->d_kept_score; f.kept_count = c->d_kept_count; f.keypoints = c->d_kps; f.descriptors = c->d_desc;
  f.pat_dx = c->d_pdx; f.pat_dy = c->d_pdy; f.pat_half = c->d_half; f.pair_i = c->d_pi; f.pair_j = c->d_pj;
  c->occ_bytes = (size_t)(W / 2 + 32) * (H / 2 + 32);
  if (c->occ_bytes > 184 * 1024) {  // + 40 KB of static shared memory (cone table, staged candidate tile) <= 227 KB
    set_error("svin_fe_create: image too large for the shared-memory occupancy grid");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(fe_configure(c->occ_bytes));
  // descriptor tiles are staged with TMA bulk copies (cp.async.bulk); SVIN_FE_NO_TMA=1 selects plain loads (debug)
  c->use_tma = std::getenv("SVIN_FE_NO_TMA") == nullptr;
  *out = c;
  return SVIN_OK;
}

void svin_fe_destroy(svin_fe_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->d_images); cudaFree(c->d_intr); cudaFree(c->d_edir); cudaFree(c->d_scores); cudaFree(c->d_cand_count);
  cudaFree(c->d_cand_keys); cudaFree(c->d_kept_xy); cudaFree(c->d_kept_score); cudaFree(c->d_kept_count);
  cudaFree(c->d_kps); cudaFree(c->d_desc); cudaFree(c->d_pdx); cudaFree(c->d_pdy); cudaFree(c->d_half);
  cudaFree(c->d_pi); cudaFree(c->d_pj); cudaFree(c->d_match);
  cudaFreeHost(c->h_images); cudaFreeHost(c->h_small); cudaFreeHost(c->h_kps); cudaFreeHost(c->h_desc);
  cudaFreeHost(c->h_counts); cudaFreeHost(c->h_match);
  for (auto& e : c->ev)
    if (e) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c->pool;
  delete c;
}

int svin_fe_upload(svin_fe_ctx* c, int32_t n, const uint8_t* const* images, int32_t stride, const double* intrinsics,
                   const double* edir) {
  if (!c || !images || !intrinsics || !edir || n < 1 || n > c->opt.max_images || stride < c->opt.image_width) {
    set_error("svin_fe_upload: invalid arguments (num_images must be in [1, max_images], stride >= width)");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  const int W = c->opt.image_width, H = c->opt.image_height;
  for (int i = 0; i < n; ++i)
    if (!images[i]) {
      set_error("svin_fe_upload: NULL image");
      return SVIN_ERR_INVALID_ARGUMENT;
    }
  c->pool->run(n, [&](int i) {  // row-wise copy into the pitched pinned staging buffer
    for (int y = 0; y < H; ++y)
      std::memcpy(c->h_images + ((size_t)i * H + y) * c->pitch, images[i] + (size_t)y * stride, W);
  });
  std::memcpy(c->h_small, intrinsics, sizeof(double) * 8 * n);
  std::memcpy(c->h_small + 8 * (size_t)c->opt.max_images, edir, sizeof(double) * 3 * n);
  SVIN_CUDA(cudaEventRecord(c->ev[5], c->stream));
  SVIN_CUDA(cudaMemcpyAsync(c->d_images, c->h_images, (size_t)n * H * c->pitch, cudaMemcpyHostToDevice, c->stream));
  SVIN_CUDA(cudaMemcpyAsync(c->d_intr, c->h_small, sizeof(double) * 8 * n, cudaMemcpyHostToDevice, c->stream));
  SVIN_CUDA(cudaMemcpyAsync(c->d_edir, c->h_small + 8 * (size_t)c->opt.max_images, sizeof(double) * 3 * n,
                            cudaMemcpyHostToDevice, c->stream));
  SVIN_CUDA(cudaEventRecord(c->ev[6], c->stream));
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]);
  c->tm.h2d_ms = ms;
  c->tm.h2d_bytes = (int64_t)n * H * c->pitch + (int64_t)sizeof(double) * 11 * n;
  c->n_images = n;
  return SVIN_OK;
}

int svin_fe_upload_device(svin_fe_ctx* c, int32_t n, const uint8_t* d_images, const double* intrinsics,
                          const double* edir) {
  if (!c || !d_images || !intrinsics || !edir || n < 1 || n > c->opt.max_images) {
    set_error("svin_fe_upload_device: invalid arguments (num_images must be in [1, max_images])");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  const int W = c->opt.image_width, H = c->opt.image_height;
  std::memcpy(c->h_small, intrinsics, sizeof(double) * 8 * n);
  std::memcpy(c->h_small + 8 * (size_t)c->opt.max_images, edir, sizeof(double) * 3 * n);
  SVIN_CUDA(cudaEventRecord(c->ev[5], c->stream));
  // dense device rows -> the pitched image buffer (same device; the producer's stream was synchronised by its API)
  SVIN_CUDA(cudaMemcpy2DAsync(c->d_images, c->pitch, d_images, W, W, (size_t)n * H, cudaMemcpyDeviceToDevice,
                              c->stream));
  SVIN_CUDA(cudaMemcpyAsync(c->d_intr, c->h_small, sizeof(double) * 8 * n, cudaMemcpyHostToDevice, c->stream));
  SVIN_CUDA(cudaMemcpyAsync(c->d_edir, c->h_small + 8 * (size_t)c->opt.max_images, sizeof(double) * 3 * n,
                            cudaMemcpyHostToDevice, c->stream));
  SVIN_CUDA(cudaEventRecord(c->ev[6], c->stream));
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]);
  c->tm.h2d_ms = ms;
  c->tm.h2d_bytes = (int64_t)sizeof(double) * 11 * n;
  c->n_images = n;
  return SVIN_OK;
}

int svin_fe_run(svin_fe_ctx* c) {
  if (!c || c->n_images < 1) {
    set_error("svin_fe_run: nothing uploaded");
    return SVIN_ERR_STATE;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  SVIN_CUDA(cudaMemsetAsync(c->d_cand_count, 0, sizeof(unsigned) * c->n_images, c->stream));
  fe_launch_detect(c->f, c->n_images, c->use_tma, c->occ_bytes, c->stream, c->ev);
  SVIN_CUDA(cudaGetLastError());
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[4]);
  c->tm.run_ms = ms;
  for (int k = 0; k < 4; ++k) {
    cudaEventElapsedTime(&ms, c->ev[k], c->ev[k + 1]);
    c->tm.kernel_ms[k + (k >= 1 ? 1 : 0)] = ms;  // harris(+nms) -> slot 0, sort -> 2, uniformity -> 3, describe -> 4
  }
  c->tm.kernel_ms[1] = 0.0;
  c->tm.kernel_launches += 4;
  return SVIN_OK;
}

int svin_fe_download(svin_fe_ctx* c, SvinKeypoint* kps, uint8_t* desc, int32_t* counts) {
  if (!c || c->n_images < 1 || !kps || !desc || !counts) {
    set_error("svin_fe_download: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  const int n = c->n_images, K = c->opt.max_keypoints;
  SVIN_CUDA(cudaEventRecord(c->ev[5], c->stream));
  SVIN_CUDA(cudaMemcpyAsync(c->h_kps, c->d_kps, sizeof(SvinKeypoint) * (size_t)n * K, cudaMemcpyDeviceToHost, c->stream));
  SVIN_CUDA(cudaMemcpyAsync(c->h_desc, c->d_desc, (size_t)n * K * 48, cudaMemcpyDeviceToHost, c->stream));
  SVIN_CUDA(cudaMemcpyAsync(c->h_counts, c->d_kept_count, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
  SVIN_CUDA(cudaEventRecord(c->ev[6], c->stream));
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]);
  c->tm.d2h_ms = ms;
  c->tm.d2h_bytes = (int64_t)n * K * (sizeof(SvinKeypoint) + 48) + 4 * n;
  std::memcpy(kps, c->h_kps, sizeof(SvinKeypoint) * (size_t)n * K);
  std::memcpy(desc, c->h_desc, (size_t)n * K * 48);
  std::memcpy(counts, c->h_counts, sizeof(int) * n);
  return SVIN_OK;
}

int svin_fe_detect_describe(svin_fe_ctx* c, int32_t n, const uint8_t* const* images, int32_t stride,
                            const double* intrinsics, const double* edir, SvinKeypoint* kps, uint8_t* desc,
                            int32_t* counts) {
  int rc = svin_fe_upload(c, n, images, stride, intrinsics, edir);
  if (rc != SVIN_OK) return rc;
  rc = svin_fe_run(c);
  if (rc != SVIN_OK) return rc;
  return svin_fe_download(c, kps, desc, counts);
}

int svin_fe_scores(svin_fe_ctx* c, int32_t index, int32_t* out) {
  if (!c || !out || index < 0 || index >= c->n_images) {
    set_error("svin_fe_scores: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)c->opt.image_width * c->opt.image_height;
  SVIN_CUDA(cudaMemcpy(out, c->d_scores + n * index, sizeof(int) * n, cudaMemcpyDeviceToHost));
  return SVIN_OK;
}

int svin_match(svin_fe_ctx* c, int32_t np, const SvinMatchProblem* probs, SvinMatchResult* res) {
  if (!c || !probs || !res || np < 1) {
    set_error("svin_match: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  size_t NA = 0, NB = 0;
  int max_nA = 0, max_nAB = 0;
  for (int p = 0; p < np; ++p) {
    const SvinMatchProblem& q = probs[p];
    if (q.nA < 0 || q.nB < 0 || (q.nA && (!q.descA || !q.kpA)) || (q.nB && (!q.descB || !q.kpB)) || !q.intrB ||
        (q.type == SVIN_MATCH_3D2D && (!q.landmarksA || !q.T_CbW)) ||
        (q.type == SVIN_MATCH_2D2D && (!q.T_CaCb || !q.intrA)) || (q.type != SVIN_MATCH_3D2D && q.type != SVIN_MATCH_2D2D)) {
      set_error("svin_match: problem " + std::to_string(p) + " is malformed");
      return SVIN_ERR_INVALID_ARGUMENT;
    }
    NA += q.nA;
    NB += q.nB;
    max_nA = std::max(max_nA, q.nA);
    max_nAB = std::max(max_nAB, std::max(q.nA, q.nB));
  }
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  // input region
  size_t off = 0;
  const size_t o_desc = off; off += al(sizeof(MatchDesc) * np);
  const size_t o_dA = off; off += al(48 * NA);
  const size_t o_dB = off; off += al(48 * NB);
  const size_t o_sA = off; off += al(NA);
  const size_t o_sB = off; off += al(NB);
  const size_t o_kA = off; off += al(sizeof(SvinKeypoint) * NA);
  const size_t o_kB = off; off += al(sizeof(SvinKeypoint) * NB);
  const size_t o_lm = off; off += al(32 * NA);
  const size_t in_bytes = off;
  // output region
  const size_t o_bi = off; off += al(16 * NA);
  const size_t o_bd = off; off += al(16 * NA);
  const size_t o_mo = off; off += al(4 * NB);
  const size_t o_md = off; off += al(4 * NB);
  const size_t o_se = off; off += al(NA);
  const size_t out_bytes = off - in_bytes;
  // scratch
  const size_t o_proj = off; off += al(16 * NA);
  const size_t o_cov = off; off += al(32 * NA);
  const size_t o_sgA = off; off += al(8 * NA);
  const size_t o_sgB = off; off += al(8 * NB);
  const size_t o_rA = off; off += al(24 * NA);
  const size_t o_rB = off; off += al(24 * NB);
  const size_t total = off;
  if (total > c->match_cap) {
    cudaFree(c->d_match);
    cudaFreeHost(c->h_match);
    c->d_match = c->h_match = nullptr;
    c->match_cap = 0;
    const size_t want = total + total / 2 + 4096;
    SVIN_CUDA(cudaMalloc(&c->d_match, want));
    SVIN_CUDA(cudaMallocHost(&c->h_match, want));
    c->match_cap = want;
  }
  char* Hh = (char*)c->h_match;
  char* D = (char*)c->d_match;
  MatchDesc* hd = (MatchDesc*)(Hh + o_desc);
  std::vector<size_t> offA((size_t)np + 1, 0), offB((size_t)np + 1, 0);
  for (int p = 0; p < np; ++p) {
    offA[p + 1] = offA[p] + probs[p].nA;
    offB[p + 1] = offB[p] + probs[p].nB;
  }
  c->pool->run(np, [&](int p) {
    const SvinMatchProblem& q = probs[p];
    const size_t a0 = offA[p], b0 = offB[p];
    MatchDesc& d = hd[p];
    std::memset(&d, 0, sizeof d);
    d.type = q.type; d.nA = q.nA; d.nB = q.nB; d.a0 = (int)a0; d.b0 = (int)b0;
    d.W = q.image_width; d.H = q.image_height; d.threshold = q.distance_threshold;
    d.pose_uncertainty = q.pose_uncertainty;
    std::memcpy(d.intrB, q.intrB, 64);
    std::memcpy(d.intrA, q.intrA ? q.intrA : q.intrB, 64);
    if (q.T_CbW) std::memcpy(d.T_CbW, q.T_CbW, 56);
    if (q.T_CaCb) std::memcpy(d.T_CaCb, q.T_CaCb, 56);
    std::memcpy(Hh + o_dA + 48 * a0, q.descA, 48 * (size_t)q.nA);
    std::memcpy(Hh + o_dB + 48 * b0, q.descB, 48 * (size_t)q.nB);
    if (q.skipA) std::memcpy(Hh + o_sA + a0, q.skipA, q.nA); else std::memset(Hh + o_sA + a0, 0, q.nA);
    if (q.skipB) std::memcpy(Hh + o_sB + b0, q.skipB, q.nB); else std::memset(Hh + o_sB + b0, 0, q.nB);
    std::memcpy(Hh + o_kA + sizeof(SvinKeypoint) * a0, q.kpA, sizeof(SvinKeypoint) * (size_t)q.nA);
    std::memcpy(Hh + o_kB + sizeof(SvinKeypoint) * b0, q.kpB, sizeof(SvinKeypoint) * (size_t)q.nB);
    if (q.type == SVIN_MATCH_3D2D) std::memcpy(Hh + o_lm + 32 * a0, q.landmarksA, 32 * (size_t)q.nA);
    else std::memset(Hh + o_lm + 32 * a0, 0, 32 * (size_t)q.nA);
  });
  MatchBatch mb{};
  mb.desc = (const MatchDesc*)(D + o_desc);
  mb.descA = (const uint8_t*)(D + o_dA); mb.descB = (const uint8_t*)(D + o_dB);
  mb.skipA_in = (const uint8_t*)(D + o_sA); mb.skipB = (const uint8_t*)(D + o_sB);
  mb.kpA = (const SvinKeypoint*)(D + o_kA); mb.kpB = (const SvinKeypoint*)(D + o_kB);
  mb.landmarksA = (const double*)(D + o_lm);
  mb.skipA = (uint8_t*)(D + o_se);
  mb.proj = (double*)(D + o_proj); mb.cov = (double*)(D + o_cov); mb.sigA = (double*)(D + o_sgA);
  mb.sigB = (double*)(D + o_sgB); mb.rayA = (double*)(D + o_rA); mb.rayB = (double*)(D + o_rB);
  mb.best_idx = (int*)(D + o_bi); mb.best_dist = (float*)(D + o_bd);
  mb.match_of_B = (int*)(D + o_mo); mb.match_dist = (float*)(D + o_md);
  SVIN_CUDA(cudaEventRecord(c->ev[5], c->stream));
  SVIN_CUDA(cudaMemcpyAsync(D, Hh, in_bytes, cudaMemcpyHostToDevice, c->stream));
  SVIN_CUDA(cudaEventRecord(c->ev[6], c->stream));
  fe_launch_match(mb, np, max_nA, max_nAB, c->stream, c->ev);
  SVIN_CUDA(cudaGetLastError());
  SVIN_CUDA(cudaEventRecord(c->ev[3], c->stream));
  SVIN_CUDA(cudaMemcpyAsync(Hh + in_bytes, D + in_bytes, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  SVIN_CUDA(cudaEventRecord(c->ev[4], c->stream));
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[2]);
  c->tm.run_ms = ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->tm.kernel_ms[5] = ms;
  cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]);
  c->tm.kernel_ms[6] = ms;
  cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]);
  c->tm.h2d_ms = ms;
  cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]);
  c->tm.d2h_ms = ms;
  c->tm.h2d_bytes = (int64_t)in_bytes;
  c->tm.d2h_bytes = (int64_t)out_bytes;
  c->tm.kernel_launches += 3;
  c->pool->run(np, [&](int p) {
    const SvinMatchProblem& q = probs[p];
    const SvinMatchResult& r = res[p];
    const size_t a0 = offA[p], b0 = offB[p];
    if (r.best_index) std::memcpy(r.best_index, Hh + o_bi + 16 * a0, 16 * (size_t)q.nA);
    if (r.best_distance) std::memcpy(r.best_distance, Hh + o_bd + 16 * a0, 16 * (size_t)q.nA);
    if (r.match_of_B) std::memcpy(r.match_of_B, Hh + o_mo + 4 * b0, 4 * (size_t)q.nB);
    if (r.match_distance) std::memcpy(r.match_distance, Hh + o_md + 4 * b0, 4 * (size_t)q.nB);
    if (r.skipA_effective) std::memcpy(r.skipA_effective, Hh + o_se + a0, q.nA);
  });
  return SVIN_OK;
}

int svin_fe_timings(svin_fe_ctx* c, SvinFeTimings* out) {
  if (!c || !out) {
    set_error("svin_fe_timings: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  *out = c->tm;
  return SVIN_OK;
}

}  // extern "C"
