// Device planner of the BA upload path: the ordering work that order_window + fill_window (ba_engine.cu) do on the host
// threads - landmarks grouped by exact (pose, camera) observation pattern, patterns cut into Schur chunks, the pattern-major
// observation order, the chunk / run tables and the class-sorted chunk list - done by one CTA per window from the caller's
// raw arrays once they are in HBM.  The host then only validates and copies.  svin_ba_plan (host) stays as the checker:
// tests/test_plan_gpu.py compares every table the two produce.
//
// Reference seam: none of this exists in the reference (Ceres orders its own parameter blocks, ceres::Problem /
// okvis::ceres::Map, okvis_ceres/src/Map.cpp:347); it is the layout step of the batched Schur kernels.
#include <algorithm>
#include <cstdlib>

#include "ba_kernels.cuh"
#include "ba_plan.cuh"

namespace svin {

int schur_use_lr() {
  static const int v = std::getenv("SVIN_SCHUR_LR") ? std::atoi(std::getenv("SVIN_SCHUR_LR")) : 3;
  return v;
}
ChunkCaps chunk_caps_for(int runs) {
  const int use_lr = schur_use_lr();
  ChunkCaps c;
  c.cap = std::max(1, std::min(32, schur_mma_max_chunk(std::max(runs, 1))));
  c.lr1 = use_lr ? schur_lr_max_chunk(runs, 1) : 0;
  c.lr2 = use_lr >= 3 ? schur_lr_max_chunk(runs, 2) : 0;
  c.lr4 = use_lr >= 3 ? schur_lr_max_chunk(runs, 4) : 0;
  return c;
}
const PlanCapsTable& plan_caps_table() {
  static const PlanCapsTable t = [] {
    PlanCapsTable x{};
    for (int r = 0; r <= kPlanMaxRuns; ++r) {
      const ChunkCaps c = chunk_caps_for(r);
      x.cap[r] = (unsigned char)c.cap;
      x.lr1[r] = (unsigned char)c.lr1;
      x.lr2[r] = (unsigned char)c.lr2;
      x.lr4[r] = (unsigned char)c.lr4;
    }
    x.use_lr = schur_use_lr();
    return x;
  }();
  return t;
}

namespace {

constexpr int kPlanThreads = 512;

// In-place exclusive scan of a[0..n) (global memory) by the whole CTA; returns the total.  The caller has synchronised
// the writes of a[]; the scan ends synchronised.
__device__ int block_excl_scan(int* a, int n, int* part) {
  const int T = blockDim.x, tid = threadIdx.x;
  const int ipt = (n + T - 1) / T;
  const int b0 = min(n, tid * ipt), b1 = min(n, b0 + ipt);
  int s = 0;
  for (int k = b0; k < b1; ++k) s += a[k];
  part[tid] = s;
  __syncthreads();
  for (int off = 1; off < T; off <<= 1) {
    const int v = tid >= off ? part[tid - off] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int run = part[tid] - s;
  const int total = part[T - 1];
  for (int k = b0; k < b1; ++k) {
    const int v = a[k];
    a[k] = run;
    run += v;
  }
  __syncthreads();
  return total;
}

// packed: 0 = separate arrays, 1 = pose | ext << 10 | cam << 20 (+ rlm), 2 = landmark | pose << 18 | ext << 24 | cam << 30
__device__ __forceinline__ int obs_pose_of(const PlanArgs& a, int g) {
  return a.packed == 2 ? (int)(((unsigned)a.rpec[g] >> 18) & 63u) : (a.packed ? (a.rpec[g] & 1023) : a.rpose[g]);
}
__device__ __forceinline__ int obs_cam_of(const PlanArgs& a, int g) {
  return a.packed == 2 ? (int)((unsigned)a.rpec[g] >> 30) : (a.packed ? ((a.rpec[g] >> 20) & 1023) : a.rcam[g]);
}
__device__ __forceinline__ int obs_lm_of(const PlanArgs& a, int g) {
  return a.packed == 2 ? (int)((unsigned)a.rpec[g] & 0x3ffffu) : a.rlm[g];
}

__global__ void __launch_bounds__(kPlanThreads) k_plan_window(PlanArgs a) {
  extern __shared__ unsigned long long sm_dyn[];
  __shared__ int part[kPlanThreads];
  __shared__ int cls[kSchurClasses];
  const int i = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const WinDesc& d = a.win[i];
  const int lm0 = d.lm_begin, ob0 = d.obs_begin, L = d.lm_end - d.lm_begin, N = d.obs_end - d.obs_begin;
  int Pw = 32;
  while (Pw < L) Pw <<= 1;
  unsigned long long* keys = sm_dyn;
  int* idx = reinterpret_cast<int*>(sm_dyn + Pw);
  const int sb = lm0 + 2 * i;
  int* start = a.scratch[0] + sb;        // [L + 1] first observation of caller landmark l (observations are sorted)
  int* e = a.scratch[1] + sb;            // pattern heads -> segment index; later chunk sizes -> first observation
  int* seg_begin = a.scratch[2] + sb;    // [segments + 1] internal landmark position
  int* seg_chunks = a.scratch[3] + sb;   // chunks of a segment -> first chunk
  int* seg_runs = a.scratch[4] + sb;     // pose runs of a segment -> first run
  if (tid < kSchurClasses) cls[tid] = 0;
  for (int l = tid; l <= L; l += T) start[l] = 0;
  __syncthreads();
  for (int o = tid; o < N; o += T) atomicAdd(&start[obs_lm_of(a, ob0 + o)], 1);
  __syncthreads();
  block_excl_scan(start, L + 1, part);

  // ---- pattern key per landmark (order_window: observation count, first pose, FNV-1a of the (pose, camera) list)
  for (int k = tid; k < Pw; k += T) {
    unsigned long long key = ~0ull;
    int id = 0x7fffffff;
    if (k < L) {
      const int s0 = start[k], s1 = start[k + 1];
      unsigned long long h = 1469598103934665603ull ^ (unsigned long long)(a.lmfix_raw[lm0 + k] ? 1 : 0);
      for (int q = s0; q < s1; ++q)
        h = (h ^ (unsigned long long)(obs_pose_of(a, ob0 + q) * 4 + obs_cam_of(a, ob0 + q) + 1)) * 1099511628211ull;
      const unsigned long long nobs = (unsigned long long)(s1 - s0);
      const unsigned long long first = nobs ? (unsigned long long)obs_pose_of(a, ob0 + s0) : 0ull;
      key = (nobs << 54) | ((first & 0x3ffull) << 44) | (h >> 20);
      id = k;
    }
    keys[k] = key;
    idx[k] = id;
  }
  __syncthreads();
  // ---- stable order by key = bitonic sort of (key, caller index)
  for (int k = 2; k <= Pw; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < Pw; t += T) {
        const int x = t ^ j;
        if (x > t) {
          const unsigned long long ka = keys[t], kb = keys[x];
          const int ia = idx[t], ib = idx[x];
          const bool a_after_b = ka > kb || (ka == kb && ia > ib);
          if (a_after_b == ((t & k) == 0)) {
            keys[t] = kb;
            keys[x] = ka;
            idx[t] = ib;
            idx[x] = ia;
          }
        }
      }
      __syncthreads();
    }
  // ---- landmark tables in internal order; pattern heads
  for (int k = tid; k < L; k += T) {
    const int l = idx[k];
    a.lm_perm[lm0 + k] = l;
    a.linv[lm0 + l] = k;
    const double4* src = reinterpret_cast<const double4*>(a.lm_raw) + (lm0 + l);
    reinterpret_cast<double4*>(a.lm_init)[lm0 + k] = *src;
    const int fx = a.lmfix_raw[lm0 + l] ? 1 : 0;
    a.lm_fixed[lm0 + k] = (unsigned char)fx;
    a.lm_win[lm0 + k] = i;
    int head = 1;
    if (k > 0) {
      const int lp = idx[k - 1];
      const int n0 = start[l + 1] - start[l];
      bool same = n0 == start[lp + 1] - start[lp] && fx == (a.lmfix_raw[lm0 + lp] ? 1 : 0);
      for (int q = 0; q < n0 && same; ++q) {
        const int ga = ob0 + start[l] + q, gb = ob0 + start[lp] + q;
        same = obs_pose_of(a, ga) == obs_pose_of(a, gb) && obs_cam_of(a, ga) == obs_cam_of(a, gb);
      }
      head = same ? 0 : 1;
    }
    e[k] = head;
  }
  if (tid == 0) e[L] = 0;
  __syncthreads();
  const int nseg = block_excl_scan(e, L + 1, part);
  for (int k = tid; k < L; k += T)
    if (e[k + 1] != e[k]) seg_begin[e[k]] = k;
  if (tid == 0) seg_begin[nseg] = L;
  __syncthreads();
  // ---- chunks and pose runs per pattern: count, scan, write
  for (int s = tid; s < nseg; s += T) {
    const int kb = seg_begin[s], ke = seg_begin[s + 1], lk = idx[kb];
    int runs = 0, prev = -1;
    for (int q = start[lk]; q < start[lk + 1]; ++q) {
      const int pz = obs_pose_of(a, ob0 + q);
      runs += pz != prev;
      prev = pz;
    }
    const int rr = runs <= kPlanMaxRuns ? runs : kPlanMaxRuns;
    const ChunkCaps cc{a.caps.cap[rr], a.caps.lr1[rr], a.caps.lr2[rr], a.caps.lr4[rr]};
    int nch = 0;
    for (int rem = ke - kb; rem > 0; ++nch) {
      int c, kind;
      plan_next_chunk(rem, runs, cc, a.caps.use_lr, c, kind);
      rem -= c;
    }
    seg_chunks[s] = nch;
    seg_runs[s] = runs;
  }
  if (tid == 0) seg_chunks[nseg] = seg_runs[nseg] = 0;
  __syncthreads();
  const int nchunks = block_excl_scan(seg_chunks, nseg + 1, part);
  block_excl_scan(seg_runs, nseg + 1, part);
  for (int s = tid; s < nseg; s += T) {
    const int kb = seg_begin[s], ke = seg_begin[s + 1], lk = idx[kb];
    const int r0 = seg_runs[s], runs = seg_runs[s + 1] - r0, ch0 = seg_chunks[s];
    const int m = start[lk + 1] - start[lk];
    int prev = -1, r = -1;
    for (int q = start[lk]; q < start[lk + 1]; ++q) {
      const int pz = obs_pose_of(a, ob0 + q);
      if (pz != prev) {
        ++r;
        a.run_off[ob0 + r0 + r] = a.poff[d.pose_begin + pz];
        a.run_k0m[ob0 + r0 + r] = ((q - start[lk]) << 8) | 1;
      } else {
        a.run_k0m[ob0 + r0 + r] += 1;
      }
      prev = pz;
    }
    const int rr = runs <= kPlanMaxRuns ? runs : kPlanMaxRuns;
    const ChunkCaps cc{a.caps.cap[rr], a.caps.lr1[rr], a.caps.lr2[rr], a.caps.lr4[rr]};
    int j = 0;
    for (int k = kb; k < ke; ++j) {
      int c, kind;
      plan_next_chunk(ke - k, runs, cc, a.caps.use_lr, c, kind);
      const int id = lm0 + ch0 + j;
      a.sw_win[id] = i;
      a.sw_lm_begin[id] = lm0 + k;
      a.sw_count[id] = c;
      a.sw_nruns[id] = runs;
      a.sw_run_first[id] = ob0 + r0;
      a.sw_kind[id] = kind;
      e[ch0 + j] = m * c;
      atomicAdd(&cls[kind], 1);
      k += c;
    }
  }
  if (tid == 0) e[nchunks] = 0;
  __syncthreads();
  block_excl_scan(e, nchunks + 1, part);
  // ---- pattern-major observation order inside a chunk: position (q, lane) -> first + q * count + lane
  {
    const int lane = tid & 31, wid = tid >> 5, nw = T >> 5;
    for (int j = wid; j < nchunks; j += nw) {
      const int id = lm0 + j;
      const int kb = a.sw_lm_begin[id] - lm0, c = a.sw_count[id], pos = e[j];
      if (lane < c) {
        const int k = kb + lane, l = idx[k];
        const int m = start[l + 1] - start[l];
        a.lmof[lm0 + k] = ob0 + pos + lane;
        a.lmos[lm0 + k] = c;
        a.lmoc[lm0 + k] = m;
        for (int q = 0; q < m; ++q) a.rord[ob0 + pos + q * c + lane] = start[l] + q;
      }
    }
  }
  if (tid < kSchurClasses) a.win_class_count[i * kSchurClasses + tid] = cls[tid];
  if (tid == 0) a.win_nchunks[i] = nchunks;
}

// Position of every window's chunks inside the class-sorted list (window-major inside a class, as the host builds it).
__global__ void k_plan_scan(PlanArgs a) {
  __shared__ int total[kSchurClasses + 1];
  const int c = threadIdx.x;
  if (c < kSchurClasses) {
    int s = 0;
    for (int i = 0; i < a.B; ++i) s += a.win_class_count[i * kSchurClasses + c];
    total[c] = s;
  }
  __syncthreads();
  if (c < kSchurClasses) {
    int pos = 0;
    for (int k = 0; k < c; ++k) pos += total[k];
    for (int i = 0; i < a.B; ++i) {
      a.win_class_base[i * kSchurClasses + c] = pos;
      pos += a.win_class_count[i * kSchurClasses + c];
    }
    a.class_total[c] = total[c];
  }
  if (c == 0) {
    int s = 0;
    for (int k = 0; k < kSchurClasses; ++k) s += total[k];
    a.class_total[kSchurClasses] = s;
  }
}

__global__ void k_plan_list(PlanArgs a) {
  const int i = blockIdx.x, c = threadIdx.x;
  if (c >= kSchurClasses) return;
  const int lm0 = a.win[i].lm_begin, n = a.win_nchunks[i];
  int pos = a.win_class_base[i * kSchurClasses + c];
  for (int j = 0; j < n; ++j)
    if (a.sw_kind[lm0 + j] == c) a.sw_list[pos++] = lm0 + j;
}

}  // namespace

cudaError_t launch_plan(const PlanArgs& a, int max_landmarks, cudaStream_t st) {
  if (a.B <= 0) return cudaSuccess;
  int P = 32;
  while (P < max_landmarks) P <<= 1;
  const size_t smem = 12 * (size_t)P;
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute((const void*)k_plan_window, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)(12 * (size_t)kPlanMaxLandmarks));
    if (e != cudaSuccess) return e;
  }
  k_plan_window<<<a.B, kPlanThreads, smem, st>>>(a);
  k_plan_scan<<<1, 32, 0, st>>>(a);
  k_plan_list<<<a.B, 32, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace svin
