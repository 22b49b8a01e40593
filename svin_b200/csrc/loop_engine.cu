// Loop-closure QUERY path of pose_graph on the device (BASELINE configs[4]), behind svin_loop_*
// (include/svin_b200.h).  Reference arithmetic (vendored DBoW2 + pose_graph):
//   BRIEF-256 distance      FBrief::distance                           pose_graph/ThirdParty/DBoW/FBrief.cpp:44
//   word lookup             TemplatedVocabulary::transform(feature...)  pose_graph/ThirdParty/DBoW/TemplatedVocabulary.h:1114-1153
//   bag of words (TF-IDF, L1)  transform(features, BowVector)           …/TemplatedVocabulary.h:981-1028, BowVector.cpp:52-66
//   inverted-file query     TemplatedDatabase::queryL1                   …/TemplatedDatabase.h:587-646
//   candidate search        Keyframe::searchByBRIEFDes / searchInAera    pose_graph/src/pose_graph/Keyframe.cpp:262-306
// Not here: FAST/BRIEF extraction, PnPRANSAC, pose-graph optimisation.
//
// Kernels:
//   k_loop_words    thread per feature: k-ary descent, 256-bit Hamming per child (4 x popc64), first minimum on ties
//   k_loop_bow      CTA per image: (word, feature) keys sorted in shared memory, per-word sums in FEATURE order (the order
//                   BowVector::addWeight accumulates in), L1 norm summed in ascending word order, element-wise division
//   k_loop_query    thread per database entry: merge of the two ascending word lists, the reference's term order
//   k_loop_topk     one CTA: max_results rounds of (score, entry id) arg-min - ascending score, ties by entry id
//   k_brief_search  warp per window descriptor over all old descriptors, lowest index among equal distances
// The database is stored by ENTRY (CSR: word ids ascending + values); an entry-parallel merge visits exactly the terms the
// reference's word-major inverted file visits for that entry, in the same (ascending word) order, so scores are bit-equal.
// Sharding (configs[4]): rank r keeps the entries e with e % world == r; a query returns the rank's top results with global
// ids, the merge over ranks is 4 x world numbers (host / all-gather).
#include <cuda_runtime.h>
#include <math.h>

#include <cstring>
#include <string>
#include <vector>

#include "common.hpp"

using namespace svin;

namespace {

struct Voc {
  const int *first_child, *num_children, *word_id;
  const unsigned long long* desc;   // [nodes][4]
  const double* weight;
};

__device__ __forceinline__ int ham256(const unsigned long long* a, const unsigned long long* b) {
  return __popcll(a[0] ^ b[0]) + __popcll(a[1] ^ b[1]) + __popcll(a[2] ^ b[2]) + __popcll(a[3] ^ b[3]);
}

__global__ void k_loop_words(Voc v, const unsigned long long* feat, int n, int* word, double* weight) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  unsigned long long q[4];
  for (int k = 0; k < 4; ++k) q[k] = feat[4 * (size_t)f + k];
  int node = 0;
  while (v.num_children[node] > 0) {
    const int c0 = v.first_child[node], nc = v.num_children[node];
    int best = c0, best_d = ham256(q, v.desc + 4 * (size_t)c0);
    for (int c = c0 + 1; c < c0 + nc; ++c) {
      const int d = ham256(q, v.desc + 4 * (size_t)c);
      if (d < best_d) {
        best_d = d;
        best = c;
      }
    }
    node = best;
  }
  word[f] = v.word_id[node];
  weight[f] = v.weight[node];
}

constexpr int kMaxFeat = 2048;   // features per image (the reference extracts a few hundred FAST corners)
// one CTA per image: sparse L1-normalised TF-IDF vector, ids ascending
__global__ void __launch_bounds__(1024) k_loop_bow(const int* img_off, const int* word, const double* weight, int* out_ids,
                                                   double* out_vals, int* out_cnt) {
  __shared__ unsigned long long key[kMaxFeat];
  __shared__ int head[kMaxFeat];
  __shared__ int nwords_s;
  const int img = blockIdx.x, f0 = img_off[img], n = img_off[img + 1] - f0;
  int P = 1;
  while (P < n) P <<= 1;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    // stopped words (weight <= 0) are dropped (TemplatedVocabulary.h:1003); padding sorts last
    const bool ok = i < n && weight[f0 + i] > 0.0;
    key[i] = ok ? (((unsigned long long)(unsigned)word[f0 + i] << 32) | (unsigned)i) : ~0ull;
  }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = key[i], b = key[l];
          if (((i & k) == 0) ? (a > b) : (a < b)) {
            key[i] = b;
            key[l] = a;
          }
        }
      }
      __syncthreads();
    }
  // heads of the runs of equal words, compacted in order by thread 0 (a few hundred entries)
  if (threadIdx.x == 0) {
    int m = 0;
    for (int i = 0; i < P && key[i] != ~0ull; ++i)
      if (i == 0 || (key[i] >> 32) != (key[i - 1] >> 32)) head[m++] = i;
    nwords_s = m;
  }
  __syncthreads();
  const int m = nwords_s;
  int* ids = out_ids + f0;
  double* vals = out_vals + f0;
  for (int w = threadIdx.x; w < m; w += blockDim.x) {
    const int b = head[w];
    const unsigned wid = (unsigned)(key[b] >> 32);
    double s = 0.0;
    for (int i = b; i < P && key[i] != ~0ull && (unsigned)(key[i] >> 32) == wid; ++i)
      s += weight[f0 + (int)(key[i] & 0xffffffffu)];   // feature order within the word: the sort key's low half
    ids[w] = (int)wid;
    vals[w] = s;
  }
  __syncthreads();
  __shared__ double norm_s;
  if (threadIdx.x == 0) {
    double norm = 0.0;
    for (int w = 0; w < m; ++w) norm += fabs(vals[w]);   // BowVector::normalize(L1): ascending id order
    norm_s = norm;
    out_cnt[img] = m;
  }
  __syncthreads();
  if (norm_s > 0.0)
    for (int w = threadIdx.x; w < m; w += blockDim.x) vals[w] = vals[w] / norm_s;
}

// value = sum over common words of |q - d| - |q| - |d| (0 and `hit` = 0 when there is none)
__global__ void k_loop_query(const int* q_ids, const double* q_vals, int nq, const long long* e_off, const int* e_ids,
                             const double* e_vals, int n_entries, int entry_base, int entry_stride, int max_id,
                             double* score, int* hit) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entries) return;
  const int gid = entry_base + e * entry_stride;
  double s = 0.0;
  int common = 0;
  if (gid < max_id || max_id == -1) {
    long long j = e_off[e];
    const long long je = e_off[e + 1];
    int i = 0;
    while (i < nq && j < je) {
      const int a = q_ids[i], b = e_ids[j];
      if (a == b) {
        const double qv = q_vals[i], dv = e_vals[j];
        s += fabs(qv - dv) - fabs(qv) - fabs(dv);
        common = 1;
        ++i;
        ++j;
      } else if (a < b) {
        ++i;
      } else {
        ++j;
      }
    }
  }
  score[e] = s;
  hit[e] = common;
}

__global__ void __launch_bounds__(1024) k_loop_topk(const double* score, int* hit, int n_entries, int entry_base,
                                                    int entry_stride, int max_results, int* out_id, double* out_score,
                                                    int* out_n) {
  __shared__ double bs[32];
  __shared__ int bi[32];
  __shared__ int chosen;
  int found = 0;
  for (int r = 0; r < max_results; ++r) {
    double best = 1e300;
    int besti = 0x7fffffff;
    for (int e = threadIdx.x; e < n_entries; e += blockDim.x)
      if (hit[e] && (score[e] < best || (score[e] == best && e < besti))) {
        best = score[e];
        besti = e;
      }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob < best || (ob == best && oi < besti)) {
        best = ob;
        besti = oi;
      }
    }
    if ((threadIdx.x & 31) == 0) {
      bs[threadIdx.x >> 5] = best;
      bi[threadIdx.x >> 5] = besti;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (bs[w] < bs[0] || (bs[w] == bs[0] && bi[w] < bi[0])) {
          bs[0] = bs[w];
          bi[0] = bi[w];
        }
      chosen = bi[0];
      if (chosen != 0x7fffffff) {
        out_id[r] = entry_base + chosen * entry_stride;
        out_score[r] = -bs[0] / 2.0;
        hit[chosen] = 0;
      }
    }
    __syncthreads();
    if (chosen == 0x7fffffff) break;
    ++found;
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_n = found;
}

// Keyframe::searchByBRIEFDes: best old descriptor per window descriptor (strictly below 128), accepted below 80
__global__ void __launch_bounds__(128) k_brief_search(const unsigned long long* win, int nw, const unsigned long long* old,
                                                      int no, int* idx, int* dist, unsigned char* status) {
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= nw) return;
  unsigned long long q[4];
  for (int k = 0; k < 4; ++k) q[k] = win[4 * (size_t)w + k];
  int best = 128, besti = 0x7fffffff;
  for (int j = lane; j < no; j += 32) {
    const int d = ham256(q, old + 4 * (size_t)j);
    if (d < best) {   // ascending j per lane: keeps the lowest index of the lane's minimum
      best = d;
      besti = j;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const int ob = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob < best || (ob == best && oi < besti)) {
      best = ob;
      besti = oi;
    }
  }
  if (lane == 0) {
    const bool any = besti != 0x7fffffff;
    idx[w] = any ? besti : -1;
    dist[w] = any ? best : 128;
    status[w] = (any && best < 80) ? 1 : 0;
  }
}

template <class T>
struct DevVec {   // grow-only device array
  T* p = nullptr;
  size_t cap = 0;
  int reserve(size_t n) {
    if (n <= cap) return SVIN_OK;
    T* q = nullptr;
    const size_t ncap = n + n / 2 + 1024;
    SVIN_CUDA(cudaMalloc(&q, sizeof(T) * ncap));
    if (p) {
      SVIN_CUDA(cudaMemcpy(q, p, sizeof(T) * cap, cudaMemcpyDeviceToDevice));
      cudaFree(p);
    }
    p = q;
    cap = ncap;
    return SVIN_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

struct svin_loop_ctx {
  int device = 0, rank = 0, world = 1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int n_nodes = 0;
  int *d_first = nullptr, *d_num = nullptr, *d_word = nullptr;
  unsigned long long* d_desc = nullptr;
  double* d_weight = nullptr;
  // scratch of one call
  DevVec<unsigned long long> feat, feat2;
  DevVec<int> word, ids, cnt, off, hit, sidx, sdist;
  DevVec<double> wgt, vals, score;
  DevVec<unsigned char> sstat;
  // database (this rank's entries), CSR by entry
  DevVec<long long> e_off;
  DevVec<int> e_ids;
  DevVec<double> e_vals;
  std::vector<long long> h_off{0};
  int n_local = 0, n_total = 0;
  int* d_top = nullptr;      // [64] ids | n
  double* d_tops = nullptr;  // [64]
  double last_ms = 0.0;
};

namespace {

// features of `n_img` images -> sparse BoW vectors in c->ids / c->vals (slot = feature offset of the image), counts in c->cnt
int bow(svin_loop_ctx* c, int n_img, const uint8_t* desc, const int* counts, std::vector<int>& h_cnt, std::vector<int>& h_off) {
  h_off.assign(n_img + 1, 0);
  for (int i = 0; i < n_img; ++i) {
    if (counts[i] < 0 || counts[i] > kMaxFeat) {
      set_error("svin_loop: an image has more than " + std::to_string(kMaxFeat) + " features (or a negative count)");
      return SVIN_ERR_INVALID_ARGUMENT;
    }
    h_off[i + 1] = h_off[i] + counts[i];
  }
  const int n = h_off[n_img];
  int rc;
  if ((rc = c->feat.reserve((size_t)4 * n + 4)) != SVIN_OK || (rc = c->word.reserve(n + 1)) != SVIN_OK ||
      (rc = c->wgt.reserve(n + 1)) != SVIN_OK || (rc = c->ids.reserve(n + 1)) != SVIN_OK ||
      (rc = c->vals.reserve(n + 1)) != SVIN_OK || (rc = c->cnt.reserve(n_img + 1)) != SVIN_OK ||
      (rc = c->off.reserve(n_img + 2)) != SVIN_OK)
    return rc;
  SVIN_CUDA(cudaMemcpyAsync(c->feat.p, desc, 32 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  SVIN_CUDA(cudaMemcpyAsync(c->off.p, h_off.data(), 4 * (size_t)(n_img + 1), cudaMemcpyHostToDevice, c->stream));
  Voc v{c->d_first, c->d_num, c->d_word, c->d_desc, c->d_weight};
  if (n > 0) k_loop_words<<<(n + 127) / 128, 128, 0, c->stream>>>(v, c->feat.p, n, c->word.p, c->wgt.p);
  if (n_img > 0) k_loop_bow<<<n_img, 1024, 0, c->stream>>>(c->off.p, c->word.p, c->wgt.p, c->ids.p, c->vals.p, c->cnt.p);
  h_cnt.assign(n_img, 0);
  SVIN_CUDA(cudaMemcpyAsync(h_cnt.data(), c->cnt.p, 4 * (size_t)n_img, cudaMemcpyDeviceToHost, c->stream));
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  SVIN_CUDA(cudaGetLastError());
  return SVIN_OK;
}

}  // namespace

extern "C" {

int svin_loop_create(int device, const SvinVocabulary* voc, int32_t rank, int32_t world, svin_loop_ctx** out) {
  if (!out || !voc || voc->num_nodes < 1 || !voc->first_child || !voc->num_children || !voc->descriptor || !voc->weight ||
      !voc->word_id || world < 1 || rank < 0 || rank >= world) {
    set_error("svin_loop_create: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  for (int i = 0; i < voc->num_nodes; ++i) {
    const int nc = voc->num_children[i], c0 = voc->first_child[i];
    if (nc < 0 || (nc > 0 && (c0 <= i || c0 + nc > voc->num_nodes))) {
      set_error("svin_loop_create: vocabulary node " + std::to_string(i) + " has children outside (node, num_nodes)");
      return SVIN_ERR_INVALID_ARGUMENT;
    }
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) {
    cudaGetLastError();
    set_error("svin_loop_create: no CUDA device " + std::to_string(device) + " - this engine has no CPU fallback");
    return SVIN_ERR_NO_DEVICE;
  }
  SVIN_CUDA(cudaSetDevice(device));
  svin_loop_ctx* c = new svin_loop_ctx();
  c->device = device;
  c->rank = rank;
  c->world = world;
  c->n_nodes = voc->num_nodes;
  SVIN_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  SVIN_CUDA(cudaEventCreate(&c->ev[0]));
  SVIN_CUDA(cudaEventCreate(&c->ev[1]));
  const size_t n = (size_t)voc->num_nodes;
  SVIN_CUDA(cudaMalloc(&c->d_first, 4 * n));
  SVIN_CUDA(cudaMalloc(&c->d_num, 4 * n));
  SVIN_CUDA(cudaMalloc(&c->d_word, 4 * n));
  SVIN_CUDA(cudaMalloc(&c->d_desc, 32 * n));
  SVIN_CUDA(cudaMalloc(&c->d_weight, 8 * n));
  SVIN_CUDA(cudaMalloc(&c->d_top, 4 * 65));
  SVIN_CUDA(cudaMalloc(&c->d_tops, 8 * 64));
  SVIN_CUDA(cudaMemcpy(c->d_first, voc->first_child, 4 * n, cudaMemcpyHostToDevice));
  SVIN_CUDA(cudaMemcpy(c->d_num, voc->num_children, 4 * n, cudaMemcpyHostToDevice));
  SVIN_CUDA(cudaMemcpy(c->d_word, voc->word_id, 4 * n, cudaMemcpyHostToDevice));
  SVIN_CUDA(cudaMemcpy(c->d_desc, voc->descriptor, 32 * n, cudaMemcpyHostToDevice));
  SVIN_CUDA(cudaMemcpy(c->d_weight, voc->weight, 8 * n, cudaMemcpyHostToDevice));
  *out = c;
  return SVIN_OK;
}

void svin_loop_destroy(svin_loop_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  for (void* p : {(void*)c->d_first, (void*)c->d_num, (void*)c->d_word, (void*)c->d_desc, (void*)c->d_weight, (void*)c->d_top,
                  (void*)c->d_tops})
    if (p) cudaFree(p);
  c->feat.release(); c->feat2.release(); c->word.release(); c->ids.release(); c->cnt.release(); c->off.release();
  c->hit.release(); c->sidx.release(); c->sdist.release(); c->wgt.release(); c->vals.release(); c->score.release();
  c->sstat.release(); c->e_off.release(); c->e_ids.release(); c->e_vals.release();
  if (c->ev[0]) cudaEventDestroy(c->ev[0]);
  if (c->ev[1]) cudaEventDestroy(c->ev[1]);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int svin_loop_transform(svin_loop_ctx* c, int32_t n_img, const uint8_t* desc, const int32_t* counts, int32_t* word_ids,
                        double* values, int32_t* num_words) {
  if (!c || n_img < 0 || (n_img && (!desc || !counts || !num_words))) {
    set_error("svin_loop_transform: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  std::vector<int> h_cnt, h_off;
  const int rc = bow(c, n_img, desc, counts, h_cnt, h_off);
  if (rc != SVIN_OK) return rc;
  for (int i = 0; i < n_img; ++i) {
    num_words[i] = h_cnt[i];
    if (word_ids) SVIN_CUDA(cudaMemcpy(word_ids + h_off[i], c->ids.p + h_off[i], 4 * (size_t)h_cnt[i], cudaMemcpyDeviceToHost));
    if (values) SVIN_CUDA(cudaMemcpy(values + h_off[i], c->vals.p + h_off[i], 8 * (size_t)h_cnt[i], cudaMemcpyDeviceToHost));
  }
  return SVIN_OK;
}

int svin_loop_add(svin_loop_ctx* c, int32_t n_img, const uint8_t* desc, const int32_t* counts) {
  if (!c || n_img < 0 || (n_img && (!desc || !counts))) {
    set_error("svin_loop_add: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  std::vector<int> h_cnt, h_off;
  int rc = bow(c, n_img, desc, counts, h_cnt, h_off);
  if (rc != SVIN_OK) return rc;
  // entry ids continue the global sequence; this rank keeps those with id % world == rank
  for (int i = 0; i < n_img; ++i) {
    const int gid = c->n_total++;
    if (gid % c->world != c->rank) continue;
    const long long base = c->h_off.back();
    if ((rc = c->e_ids.reserve((size_t)(base + h_cnt[i]) + 1)) != SVIN_OK) return rc;
    if ((rc = c->e_vals.reserve((size_t)(base + h_cnt[i]) + 1)) != SVIN_OK) return rc;
    SVIN_CUDA(cudaMemcpyAsync(c->e_ids.p + base, c->ids.p + h_off[i], 4 * (size_t)h_cnt[i], cudaMemcpyDeviceToDevice, c->stream));
    SVIN_CUDA(cudaMemcpyAsync(c->e_vals.p + base, c->vals.p + h_off[i], 8 * (size_t)h_cnt[i], cudaMemcpyDeviceToDevice, c->stream));
    c->h_off.push_back(base + h_cnt[i]);
    c->n_local++;
  }
  if ((rc = c->e_off.reserve(c->h_off.size() + 1)) != SVIN_OK) return rc;
  SVIN_CUDA(cudaMemcpyAsync(c->e_off.p, c->h_off.data(), 8 * c->h_off.size(), cudaMemcpyHostToDevice, c->stream));
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  return SVIN_OK;
}

int svin_loop_query(svin_loop_ctx* c, const uint8_t* desc, int32_t count, int32_t max_results, int32_t max_id,
                    int32_t* entry_ids, double* scores, int32_t* num_results) {
  if (!c || !desc || count < 0 || max_results < 1 || max_results > 64 || !entry_ids || !scores || !num_results) {
    set_error("svin_loop_query: invalid arguments (1 <= max_results <= 64)");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  std::vector<int> h_cnt, h_off;
  int rc = bow(c, 1, desc, &count, h_cnt, h_off);
  if (rc != SVIN_OK) return rc;
  *num_results = 0;
  if (c->n_local == 0) return SVIN_OK;
  if ((rc = c->score.reserve(c->n_local + 1)) != SVIN_OK || (rc = c->hit.reserve(c->n_local + 1)) != SVIN_OK) return rc;
  SVIN_CUDA(cudaEventRecord(c->ev[0], c->stream));
  k_loop_query<<<(c->n_local + 127) / 128, 128, 0, c->stream>>>(c->ids.p, c->vals.p, h_cnt[0], c->e_off.p, c->e_ids.p,
                                                                  c->e_vals.p, c->n_local, c->rank, c->world, max_id,
                                                                  c->score.p, c->hit.p);
  k_loop_topk<<<1, 1024, 0, c->stream>>>(c->score.p, c->hit.p, c->n_local, c->rank, c->world, max_results, c->d_top,
                                         c->d_tops, c->d_top + 64);
  SVIN_CUDA(cudaEventRecord(c->ev[1], c->stream));
  int h_top[65];
  double h_tops[64];
  SVIN_CUDA(cudaMemcpyAsync(h_top, c->d_top, 4 * 65, cudaMemcpyDeviceToHost, c->stream));
  SVIN_CUDA(cudaMemcpyAsync(h_tops, c->d_tops, 8 * 64, cudaMemcpyDeviceToHost, c->stream));
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  SVIN_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->last_ms = ms;
  *num_results = h_top[64];
  for (int r = 0; r < h_top[64]; ++r) {
    entry_ids[r] = h_top[r];
    scores[r] = h_tops[r];
  }
  return SVIN_OK;
}

int svin_loop_brief_search(svin_loop_ctx* c, const uint8_t* window_desc, int32_t n_window, const uint8_t* old_desc,
                           int32_t n_old, int32_t* best_index, int32_t* best_distance, uint8_t* status) {
  if (!c || n_window < 0 || n_old < 0 || (n_window && (!window_desc || !best_index || !best_distance || !status)) ||
      (n_old && !old_desc)) {
    set_error("svin_loop_brief_search: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  if (n_window == 0) return SVIN_OK;
  SVIN_CUDA(cudaSetDevice(c->device));
  int rc;
  if ((rc = c->feat.reserve((size_t)4 * n_window + 4)) != SVIN_OK || (rc = c->feat2.reserve((size_t)4 * n_old + 4)) != SVIN_OK ||
      (rc = c->sidx.reserve(n_window)) != SVIN_OK || (rc = c->sdist.reserve(n_window)) != SVIN_OK ||
      (rc = c->sstat.reserve(n_window)) != SVIN_OK)
    return rc;
  SVIN_CUDA(cudaMemcpyAsync(c->feat.p, window_desc, 32 * (size_t)n_window, cudaMemcpyHostToDevice, c->stream));
  if (n_old) SVIN_CUDA(cudaMemcpyAsync(c->feat2.p, old_desc, 32 * (size_t)n_old, cudaMemcpyHostToDevice, c->stream));
  SVIN_CUDA(cudaEventRecord(c->ev[0], c->stream));
  k_brief_search<<<(n_window + 3) / 4, 128, 0, c->stream>>>(c->feat.p, n_window, c->feat2.p, n_old, c->sidx.p, c->sdist.p,
                                                            c->sstat.p);
  SVIN_CUDA(cudaEventRecord(c->ev[1], c->stream));
  SVIN_CUDA(cudaMemcpyAsync(best_index, c->sidx.p, 4 * (size_t)n_window, cudaMemcpyDeviceToHost, c->stream));
  SVIN_CUDA(cudaMemcpyAsync(best_distance, c->sdist.p, 4 * (size_t)n_window, cudaMemcpyDeviceToHost, c->stream));
  SVIN_CUDA(cudaMemcpyAsync(status, c->sstat.p, (size_t)n_window, cudaMemcpyDeviceToHost, c->stream));
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  SVIN_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->last_ms = ms;
  return SVIN_OK;
}

int svin_loop_stats(svin_loop_ctx* c, int32_t* entries_total, int32_t* entries_local, double* last_device_ms) {
  if (!c) {
    set_error("svin_loop_stats: ctx is NULL");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  if (entries_total) *entries_total = c->n_total;
  if (entries_local) *entries_local = c->n_local;
  if (last_device_ms) *last_device_ms = c->last_ms;
  return SVIN_OK;
}

}  // extern "C"
