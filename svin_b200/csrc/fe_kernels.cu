// CUDA kernels of the BRISK-2 style front-end (sm_100a).  Specification: DESIGN.md §A (declared, because
// brisk 2.0.8 is not vendored in the reference); in-tree arithmetic followed exactly is cited per kernel.
//
//   k_harris_nms       32x32 pixel tile per CTA   integer Harris score + 3x3 NMS + threshold -> candidate keys
//   k_sort_candidates  one CTA per image          bitonic sort of 64-bit keys (score desc, pixel index asc)
//   k_uniformity       one CTA per image          strongest-first uniformity enforcement, occupancy in smem
//   k_orient_describe  one warp per keypoint      sub-pixel refinement, Frame::describe orientation
//                                                 (okvis_cv/include/okvis/implementation/Frame.hpp:113-129),
//                                                 TMA bulk-copied 32x48 tile -> 60 box samples -> 384 bits via ballots
//   k_match_setup      one thread per keypoint    VKWMA::doSetup projections / rays (VKWMA.cpp:163-212)
//   k_match            one warp per A keypoint    Hamming (3 x 128 bit, VKWMA.hpp:258-264) + verifyMatch
//                                                 (VKWMA.cpp:292-323) + best-4 list (DenseMatcher.hpp impl:216-242)
//   k_assign           one thread per problem     assignbest (DenseMatcher.cpp:59-97), A ascending
#include <cuda_runtime.h>

#include <algorithm>
#include <math.h>
#include <stdint.h>

#include "ba_math.cuh"
#include "fe_kernels.cuh"

namespace svin {

// ------------------------------------------------------------------------------------------ Harris + NMS
constexpr int kTile = 32;
constexpr int kHalo = 3;  // Sobel 1 + box 1 + NMS 1

__global__ void __launch_bounds__(256) k_harris_nms(FeBatch f) {
  __shared__ uint8_t px[kTile + 2 * kHalo][kTile + 2 * kHalo + 2];
  __shared__ int gxx[kTile + 4][kTile + 4 + 1], gyy[kTile + 4][kTile + 4 + 1], gxy[kTile + 4][kTile + 4 + 1];
  __shared__ int sc[kTile + 2][kTile + 2 + 1];
  const int img = blockIdx.z;
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  const int W = f.W, H = f.H;
  const uint8_t* I = f.images + (size_t)img * f.H * f.pitch;
  const int tid = threadIdx.x;
  // pixels with halo 3 (zero outside the image; scores there are forced to 0 anyway)
  for (int e = tid; e < (kTile + 6) * (kTile + 6); e += 256) {
    const int ly = e / (kTile + 6), lx = e % (kTile + 6);
    const int gx = x0 + lx - kHalo, gy = y0 + ly - kHalo;
    px[ly][lx] = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? I[(size_t)gy * f.pitch + gx] : 0;
  }
  __syncthreads();
  // gradient products on (tile + 2*2); valid for image pixels 1..W-2, else 0 (matches the oracle's zero frame)
  for (int e = tid; e < (kTile + 4) * (kTile + 4); e += 256) {
    const int ly = e / (kTile + 4), lx = e % (kTile + 4);
    const int gx_ = x0 + lx - 2, gy_ = y0 + ly - 2;
    int a = 0, b = 0, c = 0;
    if (gx_ >= 1 && gx_ < W - 1 && gy_ >= 1 && gy_ < H - 1) {
      const int cy = ly + 1, cx = lx + 1;  // position in px
      const int dx = (px[cy - 1][cx + 1] + 2 * px[cy][cx + 1] + px[cy + 1][cx + 1]) -
                     (px[cy - 1][cx - 1] + 2 * px[cy][cx - 1] + px[cy + 1][cx - 1]);
      const int dy = (px[cy + 1][cx - 1] + 2 * px[cy + 1][cx] + px[cy + 1][cx + 1]) -
                     (px[cy - 1][cx - 1] + 2 * px[cy - 1][cx] + px[cy - 1][cx + 1]);
      a = dx * dx;
      b = dy * dy;
      c = dx * dy;
    }
    gxx[ly][lx] = a;
    gyy[ly][lx] = b;
    gxy[ly][lx] = c;
  }
  __syncthreads();
  // scores on (tile + 2*1)
  for (int e = tid; e < (kTile + 2) * (kTile + 2); e += 256) {
    const int ly = e / (kTile + 2), lx = e % (kTile + 2);
    const int gx_ = x0 + lx - 1, gy_ = y0 + ly - 1;
    int s = 0;
    if (gx_ >= 2 && gx_ < W - 2 && gy_ >= 2 && gy_ < H - 2) {
      long long a = 0, b = 0, c = 0;
#pragma unroll
      for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          a += gxx[ly + v][lx + u];
          b += gyy[ly + v][lx + u];
          c += gxy[ly + v][lx + u];
        }
      a >>= 10;
      b >>= 10;
      c >>= 10;
      long long v = (a * b - c * c) - (((a + b) * (a + b)) >> 4);
      v = v > 2147483647ll ? 2147483647ll : (v < -2147483647ll ? -2147483647ll : v);
      s = (int)v;
    }
    sc[ly][lx] = s;
    if (lx >= 1 && lx <= kTile && ly >= 1 && ly <= kTile && gx_ < W && gy_ < H)
      f.scores[((size_t)img * H + gy_) * W + gx_] = s;
  }
  __syncthreads();
  // 3x3 NMS on the tile, inside the descriptor border
  const int thr = f.abs_threshold;
  for (int e = tid; e < kTile * kTile; e += 256) {
    const int ly = e / kTile + 1, lx = e % kTile + 1;
    const int gx_ = x0 + lx - 1, gy_ = y0 + ly - 1;
    if (gx_ < f.border || gx_ >= W - f.border || gy_ < f.border || gy_ >= H - f.border) continue;
    const int s = sc[ly][lx];
    if (s < thr) continue;
    if (sc[ly - 1][lx - 1] > s || sc[ly - 1][lx] > s || sc[ly - 1][lx + 1] > s || sc[ly][lx - 1] > s) continue;
    if (sc[ly][lx + 1] >= s || sc[ly + 1][lx - 1] >= s || sc[ly + 1][lx] >= s || sc[ly + 1][lx + 1] >= s) continue;
    const unsigned slot = atomicAdd(&f.cand_count[img], 1u);
    if (slot < (unsigned)f.cand_cap) {
      const unsigned idx = (unsigned)(gy_ * W + gx_);
      f.cand_keys[(size_t)img * f.cand_cap + slot] = ((unsigned long long)(unsigned)s << 32) | (0xffffffffu - idx);
    }
  }
}

// ------------------------------------------------------------------------------------------ sort
// Descending bitonic sort of an image's candidate keys (unique), zero-padded to a power of two P >= n.  Stages whose
// partner distance j fits a 4096-key chunk run in shared memory (one CTA per chunk); the few stages with j >= 4096 are
// one compare-exchange per thread in global memory.  A launch is a no-op for images whose P is below its stage, so the
// fixed launch sequence (sized for cand_cap) costs a typical image (P <= 8192) two real kernels.  [r1: one CTA per image
// sorting in global memory, 0.90 ms for a 752x480 noise image.]
constexpr int kSortChunk = 4096;
__device__ __forceinline__ unsigned sort_size(const FeBatch& f, int img, unsigned& n) {
  n = f.cand_count[img];
  if (n > (unsigned)f.cand_cap) n = f.cand_cap;
  unsigned P = kSortChunk;
  while (P < n) P <<= 1;
  return P;
}
// K == 0: the complete network up to runs of kSortChunk.  K > 0: the tail (j < kSortChunk) of merge stage K.
__global__ void __launch_bounds__(1024) k_sort_local(FeBatch f, unsigned K) {
  __shared__ unsigned long long S[kSortChunk];
  const int img = blockIdx.y;
  unsigned n;
  const unsigned P = sort_size(f, img, n);
  const unsigned base = blockIdx.x * kSortChunk;
  if (base >= P || K > P) return;
  if (n == 0) return;
  unsigned long long* G = f.cand_keys + (size_t)img * f.cand_cap + base;
  for (unsigned i = threadIdx.x; i < kSortChunk; i += 1024) S[i] = (base + i < n || K != 0) ? G[i] : 0ull;
  __syncthreads();
  const unsigned k_first = K ? K : 2, k_last = K ? K : kSortChunk;
  for (unsigned k = k_first; k <= k_last; k <<= 1) {
    for (unsigned j = min(k >> 1, (unsigned)kSortChunk >> 1); j > 0; j >>= 1) {
      for (unsigned t = threadIdx.x; t < kSortChunk / 2; t += 1024) {
        const unsigned i = 2 * j * (t / j) + (t % j), l = i + j;
        const unsigned long long a = S[i], b = S[l];
        const bool desc = (((base + i) & k) == 0);
        if (desc ? (a < b) : (a > b)) {
          S[i] = b;
          S[l] = a;
        }
      }
      __syncthreads();
    }
    if (k == 0x80000000u) break;
  }
  for (unsigned i = threadIdx.x; i < kSortChunk; i += 1024) G[i] = S[i];
}
// one global compare-exchange step (K, j >= kSortChunk)
__global__ void __launch_bounds__(256) k_sort_global(FeBatch f, unsigned K, unsigned j) {
  const int img = blockIdx.y;
  unsigned n;
  const unsigned P = sort_size(f, img, n);
  if (K > P) return;
  const unsigned t = blockIdx.x * 256 + threadIdx.x;
  const unsigned i = 2 * j * (t / j) + (t % j), l = i + j;
  if (l >= P) return;
  unsigned long long* G = f.cand_keys + (size_t)img * f.cand_cap;
  const unsigned long long a = G[i], b = G[l];
  const bool desc = ((i & K) == 0);
  if (desc ? (a < b) : (a > b)) {
    G[i] = b;
    G[l] = a;
  }
}

// Batched path (many images per call): one CTA per image sorts in global memory - all SMs are busy across the images and it
// is ONE launch, where the chunked network above would issue ~20 mostly-empty grids per batch.
__global__ void __launch_bounds__(1024) k_sort_image(FeBatch f) {
  const int img = blockIdx.x;
  unsigned n = f.cand_count[img];
  if (n > (unsigned)f.cand_cap) n = f.cand_cap;
  unsigned long long* K = f.cand_keys + (size_t)img * f.cand_cap;
  unsigned P = 1;
  while (P < n) P <<= 1;
  for (unsigned i = n + threadIdx.x; i < P; i += blockDim.x) K[i] = 0ull;
  __syncthreads();
  for (unsigned k = 2; k <= P; k <<= 1)
    for (unsigned j = k >> 1; j > 0; j >>= 1) {
      for (unsigned i = threadIdx.x; i < P; i += blockDim.x) {
        const unsigned l = i ^ j;
        if (l > i) {
          const unsigned long long a = K[i], b = K[l];
          const bool desc = ((i & k) == 0);
          if (desc ? (a < b) : (a > b)) {
            K[i] = b;
            K[l] = a;
          }
        }
      }
      __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------ uniformity
// Strongest-first acceptance against the occupancy image.  The reference procedure is sequential in the accepted
// keypoints (each stamp changes the occupancy later tests read); this kernel produces exactly its result but accepts
// SEVERAL candidates per round:
//   * all candidates of a 512-wide window are tested against the current occupancy at once; those that fail now fail
//     for good (the occupancy only grows);
//   * the passing ones, in order, are accepted as long as each lies outside the 31x31 stamp footprint of every earlier
//     passing candidate of the round - its test cannot be changed by those stamps, so the sequential procedure would
//     accept it too.  The first one that is not (or the 33rd) starts the next round's window, which keeps the order.
//   * saturating adds commute, so stamps of one round only need a barrier between them where their footprints overlap.
// Candidates are staged in shared memory a tile at a time, decoded, with their stamp amplitude (two fp64 square roots)
// precomputed in parallel; the test reads a 256-entry table of (v / 255)^4 instead of dividing.  Bit-exact against the
// oracle's sequential loop (tests/test_fe_gpu.py).  [r1: one keypoint per round, 256 threads, 0.49 ms per image.]
constexpr int kUniThreads = 512;
constexpr int kUniTile = 2048;
constexpr int kUniList = 32;
__global__ void __launch_bounds__(kUniThreads) k_uniformity(FeBatch f) {
  extern __shared__ uint8_t occ[];  // (H/2+32) x (W/2+32)
  __shared__ double lut[31 * 31];
  __shared__ double p4[256];                     // ((s0 s0) s0) s0 with s0 = v / 255.0, the test's own operation order
  __shared__ double t_amp[kUniTile];             // 255.0 * sqrt(sqrt(score / maxScore))
  __shared__ int t_xy[kUniTile], t_s[kUniTile];
  __shared__ int wc[kUniThreads / 32];
  __shared__ int l_idx[kUniList + 1], l_cx[kUniList + 1], l_cy[kUniList + 1];
  __shared__ int m_s;
  __shared__ unsigned ovl_s;
  const int img = blockIdx.x;
  const int W = f.W, H = f.H;
  const int OW = W / 2 + 32, OH = H / 2 + 32;
  unsigned n = f.cand_count[img];
  if (n > (unsigned)f.cand_cap) n = f.cand_cap;
  const unsigned long long* K = f.cand_keys + (size_t)img * f.cand_cap;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int T = kUniThreads, NW = kUniThreads / 32;
  {
    uint32_t* o4 = reinterpret_cast<uint32_t*>(occ);   // the dynamic shared array is 16-byte aligned
    const int words = (OW * OH) / 4;
    for (int e = tid; e < words; e += T) o4[e] = 0u;
    for (int e = 4 * words + tid; e < OW * OH; e += T) occ[e] = 0;
  }
  const double half_radius = f.uniformity_radius / 2.0;
  for (int e = tid; e < 31 * 31; e += T) {
    const int dy = e / 31 - 15, dx = e % 31 - 15;
    const double d = sqrt((double)(dx * dx + dy * dy));
    const double v = 1.0 - d / half_radius;
    lut[e] = v > 0.0 ? v : 0.0;
  }
  for (int v = tid; v < 256; v += T) {
    const double s0 = (double)v / 255.0;
    p4[v] = s0 * s0 * s0 * s0;
  }
  __syncthreads();
  int* kept_xy = f.kept_xy + (size_t)img * f.max_kp * 2;
  int* kept_score = f.kept_score + (size_t)img * f.max_kp;
  if (n == 0) {
    if (tid == 0) f.kept_count[img] = 0;
    return;
  }
  const double maxScore = (double)(unsigned)(K[0] >> 32);
  const bool enforce = f.uniformity_radius > 0.0;
  unsigned i0 = 0, tile0 = 0, tile1 = 0;   // candidates [tile0, tile1) are staged
  int kept = 0;                            // identical in every thread
  while (i0 < n && kept < f.max_kp) {
    if (i0 + T > tile1 && tile1 < n) {     // (re)stage a tile starting at the scan position
      __syncthreads();
      tile0 = i0;
      tile1 = min(n, i0 + (unsigned)kUniTile);
      for (unsigned e = tile0 + tid; e < tile1; e += T) {
        const unsigned long long key = K[e];
        const int sc = (int)(unsigned)(key >> 32);
        const unsigned idx = 0xffffffffu - (unsigned)(key & 0xffffffffu);
        t_s[e - tile0] = sc;
        t_xy[e - tile0] = (int)idx;
        t_amp[e - tile0] = 255.0 * sqrt(sqrt((double)sc / maxScore));
      }
      __syncthreads();
    }
    // ---- test the window against the current occupancy
    const unsigned i = i0 + tid;
    bool pass = false;
    int cx = 0, cy = 0;
    if (i < tile1) {
      const int idx = t_xy[i - tile0];
      cy = (idx / W) / 2 + 16;
      cx = (idx % W) / 2 + 16;
      pass = true;
      if (enforce) {
        const double lim = p4[occ[cy * OW + cx]] * maxScore;
        pass = !((double)t_s[i - tile0] < lim);
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if (lane == 0) wc[wid] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const int c = wc[w];
      before += (w < wid) ? c : 0;
      total += c;
    }
    if (total == 0) {  // nothing in this window passes; occupancy only grows, so they are all rejected
      i0 = min(i0 + (unsigned)T, tile1);
      __syncthreads();   // wc is rewritten next round
      continue;
    }
    const int rank = before + __popc(bal & ((1u << lane) - 1u));
    if (pass && rank <= kUniList) {   // entry kUniList only marks where the next window starts
      l_idx[rank] = (int)i;
      l_cx[rank] = cx;
      l_cy[rank] = cy;
    }
    __syncthreads();
    // ---- warp 0: longest prefix of passing candidates that are outside each other's stamp footprints
    const int listed = min(total, kUniList);
    if (wid == 0) {
      bool unsafe = false, overlap = false;
      if (lane < listed && enforce) {
        const int mx = l_cx[lane], my = l_cy[lane];
        for (int q = 0; q < lane; ++q) {
          const int dx = abs(mx - l_cx[q]), dy = abs(my - l_cy[q]);
          unsafe = unsafe || (dx <= 15 && dy <= 15);
          overlap = overlap || (dx <= 30 && dy <= 30);
        }
      }
      const unsigned ub = __ballot_sync(0xffffffffu, unsafe), ob = __ballot_sync(0xffffffffu, overlap);
      if (lane == 0) {
        m_s = ub ? __ffs(ub) - 1 : listed;
        ovl_s = ob;
      }
    }
    __syncthreads();
    const int m = m_s;
    const unsigned ovl = ovl_s;
    const int accept = min(m, f.max_kp - kept);
    for (int r = 0; r < accept; ++r) {
      const int ci = l_idx[r] - (int)tile0;
      if (enforce) {
        if ((ovl >> r) & 1u) __syncthreads();   // this stamp overlaps an earlier one of the round
        const double amp = t_amp[ci];
        const int ccy = l_cy[r], ccx = l_cx[r];
        for (int e = tid; e < 31 * 31; e += T) {
          const int dy = e / 31 - 15, dx = e % 31 - 15;
          const int add = (int)floor(amp * lut[e]);
          const int o = (ccy + dy) * OW + ccx + dx;
          const int v = (int)occ[o] + add;
          occ[o] = (uint8_t)(v > 255 ? 255 : v);
        }
      }
      if (tid == 0) {
        const int idx = t_xy[ci];
        kept_xy[2 * (kept + r)] = idx % W;
        kept_xy[2 * (kept + r) + 1] = idx / W;
        kept_score[kept + r] = t_s[ci];
      }
    }
    kept += accept;
    // next window: the first passing candidate that was not decided, else past this window
    if (m < total)
      i0 = (unsigned)l_idx[m];
    else
      i0 = min(i0 + (unsigned)T, tile1);
    __syncthreads();   // stamps complete, lists free
  }
  if (tid == 0) f.kept_count[img] = kept;
}

// ------------------------------------------------------------------------------------------ camera helpers
// RadialTangentialDistortion::undistort (5 Gauss-Newton iterations) and PinholeCamera::backProject
__device__ __forceinline__ bool backproject(const double* intr, double ix, double iy, double& rx, double& ry) {
  const double y0 = (ix - intr[2]) * (1.0 / intr[0]), y1 = (iy - intr[3]) * (1.0 / intr[1]);
  double x0 = y0, x1 = y1;
  bool ok = false;
  for (int i = 0; i < 5; ++i) {
    double d0, d1, E00, E01, E10, E11;
    radtan_distort(intr, x0, x1, d0, d1, E00, E01, E10, E11);
    const double e0 = y0 - d0, e1 = y1 - d1;
    const double a = E00 * E00 + E10 * E10, b = E00 * E01 + E10 * E11, c = E01 * E01 + E11 * E11;
    const double det = a * c - b * b;
    const double i00 = c / det, i01 = -b / det, i11 = a / det;
    const double M00 = i00 * E00 + i01 * E01, M01 = i00 * E10 + i01 * E11;
    const double M10 = i01 * E00 + i11 * E01, M11 = i01 * E10 + i11 * E11;
    x0 += M00 * e0 + M01 * e1;
    x1 += M10 * e0 + M11 * e1;
    const double chi2 = e0 * e0 + e1 * e1;
    if (chi2 < 1e-4) ok = true;
    if (chi2 < 1e-15) {
      ok = true;
      break;
    }
  }
  rx = x0;
  ry = x1;
  return ok;
}
// PinholeCamera::project with point Jacobian; returns ProjectionStatus (0 Successful, 1 OutsideImage, 3 Behind, 4 Invalid)
__device__ __forceinline__ int project3(const double* intr, double X, double Y, double Z, int W, int H, double& u,
                                        double& v, double* J) {
  if (fabs(Z) < 1.0e-12) return 4;
  const double rz = 1.0 / Z, rz2 = rz * rz;
  double d0, d1, D00, D01, D10, D11;
  radtan_distort(intr, X * rz, Y * rz, d0, d1, D00, D01, D10, D11);
  if (J) {
    J[0] = intr[0] * D00 * rz;
    J[1] = intr[0] * D01 * rz;
    J[2] = -intr[0] * (X * D00 + Y * D01) * rz2;
    J[3] = intr[1] * D10 * rz;
    J[4] = intr[1] * D11 * rz;
    J[5] = -intr[1] * (X * D10 + Y * D11) * rz2;
  }
  u = intr[0] * d0 + intr[2];
  v = intr[1] * d1 + intr[3];
  if (W > 0) {
    if (u < 0.0 || v < 0.0) return 1;
    if (u >= W || v >= H) return 1;
  }
  return Z > 0.0 ? 0 : 3;
}
__device__ __forceinline__ int project_h(const double* intr, const double* hp, int W, int H, double& u, double& v,
                                         double* J) {
  double X = hp[0], Y = hp[1], Z = hp[2];
  if (hp[3] < 0) {
    X = -X;
    Y = -Y;
    Z = -Z;
  }
  return project3(intr, X, Y, Z, W, H, u, v, J);
}

// atan2 of the keypoint orientation (Frame.hpp impl:124) with +, -, *, / only: fdlibm's argument reduction and 11-term
// polynomial (< 1 ulp).  This translation unit is compiled with -fmad=false, so every fp64 operation here is a single
// IEEE operation in source order and the angle -> float -> 1024-step rotation bin chain is bit-reproducible against a CPU
// evaluation of the same expressions; CUDA's own atan2 is only specified to 2 ulp.
__device__ __forceinline__ double svin_atan(double x) {
  const double aT[11] = {3.33333333333329318027e-01,  -1.99999999998764832476e-01, 1.42857142725034663711e-01,
                                -1.11111104054623557880e-01, 9.09088713343650656196e-02,  -7.69187620504482999495e-02,
                                6.66107313738753120669e-02,  -5.83357013379057348645e-02, 4.97687799461593236017e-02,
                                -3.65315727442169155270e-02, 1.62858201153657823623e-02};
  const double hi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01,
                               1.57079632679489655800e+00};
  const double lo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17,
                               6.12323399573676603587e-17};
  const bool neg = x < 0.0;
  const double ax = neg ? -x : x;
  if (ax >= 73786976294838206464.0) {  // 2^66
    const double z = hi[3] + lo[3];
    return neg ? -z : z;
  }
  int id = -1;
  if (ax < 0.4375) {
    if (ax < 1.862645149230957e-09) return x;  // 2^-29
  } else {
    x = ax;
    if (x < 1.1875) {
      if (x < 0.6875) {
        id = 0;
        x = (2.0 * x - 1.0) / (2.0 + x);
      } else {
        id = 1;
        x = (x - 1.0) / (x + 1.0);
      }
    } else {
      if (x < 2.4375) {
        id = 2;
        x = (x - 1.5) / (1.0 + 1.5 * x);
      } else {
        id = 3;
        x = -1.0 / x;
      }
    }
  }
  const double z = x * x, w = z * z;
  const double s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
  const double s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
  if (id < 0) return x - x * (s1 + s2);
  const double r = hi[id] - ((x * (s1 + s2) - lo[id]) - x);
  return neg ? -r : r;
}
__device__ __forceinline__ double svin_atan2(double y, double x) {
  const double pi = 3.1415926535897931160e+00, pi_lo = 1.2246467991473531772e-16;
  if (y == 0.0) return x >= 0.0 ? 0.0 : pi;
  if (x == 0.0) return y > 0.0 ? 0.5 * pi : -0.5 * pi;
  const double q = y / x;
  const double z = svin_atan(q < 0.0 ? -q : q);
  if (x > 0.0) return y > 0.0 ? z : -z;
  return y > 0.0 ? pi - (z - pi_lo) : (z - pi_lo) - pi;
}

// ------------------------------------------------------------------------------------------ describe
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Tile staging: each warp pulls the 32 rows x 48 bytes around its keypoint into shared memory with the TMA
// bulk-copy engine (cp.async.bulk, SASS UBLKCP; one 48-byte row per lane, 16-byte aligned start, all
// completing on one mbarrier) while lane 0 does the fp64 orientation math.  The tensor-map form
// (cp.async.bulk.tensor / UTMALDG) raises "illegal instruction" on this pool's B200s even through libcu++'s
// reference wrapper (tools/tma_probe.cu), so the descriptor-free form is used.
constexpr int kTileStride = 48;
template <bool USE_TMA>
__global__ void __launch_bounds__(256) k_orient_describe(FeBatch f) {
  __shared__ __align__(128) uint8_t tiles[8][32 * kTileStride];
  __shared__ __align__(8) unsigned long long mbar[8];
  __shared__ int S[8][64];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int img = blockIdx.y;
  const int k = blockIdx.x * 8 + wid;
  const int count = f.kept_count[img];
  const bool valid = k < count;
  // all warps of the CTA take part in barrier setup
  if (USE_TMA && lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[wid])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (!valid) return;
  const int cx = f.kept_xy[((size_t)img * f.max_kp + k) * 2], cy = f.kept_xy[((size_t)img * f.max_kp + k) * 2 + 1];
  const int W = f.W, H = f.H;
  uint8_t* tile = tiles[wid];
  const int x_start = (cx - 16) & ~15;        // 16-byte aligned row segment [x_start, x_start + 48) covers cx-16 .. cx+15
  const int xoff = (cx - 16) - x_start;       // 0..15
  const uint8_t* I = f.images + (size_t)img * H * f.pitch;
  if (USE_TMA) {
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[wid])),
                   "r"(32 * kTileStride)
                   : "memory");
    __syncwarp();
    const uint8_t* src = I + (size_t)(cy - 16 + lane) * f.pitch + x_start;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(tile + lane * kTileStride)),
                 "l"(src), "r"(kTileStride), "r"(smem_u32(&mbar[wid]))
                 : "memory");
  } else {
    for (int e = lane; e < 32 * kTileStride; e += 32) {
      const int yy = cy - 16 + e / kTileStride, xx = x_start + e % kTileStride;
      tile[e] = (xx >= 0 && xx < f.pitch && yy >= 0 && yy < H) ? I[(size_t)yy * f.pitch + xx] : 0;
    }
  }
  // while the tile is in flight: sub-pixel refinement and orientation (lane 0), broadcast of the rotation bin
  float kx = 0, ky = 0, ang = 0, resp = 0;
  int rot = 0;
  if (lane == 0) {
    const int* sc = f.scores + ((size_t)img * H + cy) * W + cx;
    const double m = (double)sc[0];
    double ddx = 0.0, ddy = 0.0;
    {
      const double l = (double)sc[-1], r = (double)sc[1];
      const double den = 2.0 * (l - 2.0 * m + r);
      if (den < 0.0) {
        ddx = (l - r) / den;
        ddx = ddx > 0.5 ? 0.5 : (ddx < -0.5 ? -0.5 : ddx);
      }
    }
    {
      const double l = (double)sc[-W], r = (double)sc[W];
      const double den = 2.0 * (l - 2.0 * m + r);
      if (den < 0.0) {
        ddy = (l - r) / den;
        ddy = ddy > 0.5 ? 0.5 : (ddy < -0.5 ? -0.5 : ddy);
      }
    }
    kx = (float)((double)cx + ddx);
    ky = (float)((double)cy + ddy);
    resp = (float)f.kept_score[(size_t)img * f.max_kp + k];
    // Frame::describe (Frame.hpp impl:113-129)
    const double* intr = f.intrinsics + 8 * (size_t)img;
    const double* g = f.extraction_dir + 3 * (size_t)img;
    double rx, ry, u, v, J[6] = {0, 0, 0, 0, 0, 0};
    backproject(intr, (double)kx, (double)ky, rx, ry);
    project3(intr, rx, ry, 1.0, 0, 0, u, v, J);
    const double egx = J[0] * g[0] + J[1] * g[1] + J[2] * g[2];
    const double egy = J[3] * g[0] + J[4] * g[1] + J[5] * g[2];
    const double angle = svin_atan2(egy, egx);
    ang = (float)(angle / 3.14159265358979323846 * 180.0);
    double a = (double)ang;
    if (a < 0) a += 360.0;
    rot = (int)(a * 1024.0 / 360.0 + 0.5);
    rot &= 1023;
    SvinKeypoint kp;
    kp.x = kx;
    kp.y = ky;
    kp.size = 12.0f;
    kp.angle = ang;
    kp.response = resp;
    kp.octave = 0;
    kp.class_id = -1;
    f.keypoints[(size_t)img * f.max_kp + k] = kp;
  }
  rot = __shfl_sync(0xffffffffu, rot, 0);
  if (USE_TMA) {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done)
          : "r"(smem_u32(&mbar[wid])), "r"(0)
          : "memory");
    }
  }
  __syncwarp();
  // 60 box-smoothed samples
  const int8_t* pdx = f.pat_dx + rot * 60;
  const int8_t* pdy = f.pat_dy + rot * 60;
  for (int i = lane; i < 60; i += 32) {
    const int sx = 16 + xoff + pdx[i], sy = 16 + pdy[i], h = f.pat_half[i];
    int sum = 0;
    for (int v = -h; v <= h; ++v)
      for (int u = -h; u <= h; ++u) sum += tile[(sy + v) * kTileStride + sx + u];
    S[wid][i] = sum;
  }
  __syncwarp();
  // 384 comparisons -> 12 ballots
  uint8_t* out = f.descriptors + ((size_t)img * f.max_kp + k) * 48;
  uint32_t words[12];
#pragma unroll
  for (int w = 0; w < 12; ++w) {
    const int b = w * 32 + lane;
    const int i = f.pair_i[b], j = f.pair_j[b];
    const int hi = f.pat_half[i], hj = f.pat_half[j];
    const int Ai = (2 * hi + 1) * (2 * hi + 1), Aj = (2 * hj + 1) * (2 * hj + 1);
    const bool bit = S[wid][i] * Aj > S[wid][j] * Ai;
    words[w] = __ballot_sync(0xffffffffu, bit);
  }
  if (lane < 12) {
    uint32_t wv = words[0];
#pragma unroll
    for (int w = 1; w < 12; ++w) wv = (lane == w) ? words[w] : wv;
    reinterpret_cast<uint32_t*>(out)[lane] = wv;
  }
}

// ------------------------------------------------------------------------------------------ matching
// VKWMA::doSetup (VKWMA.cpp:124-265): projections of A's landmarks into B with 2x2 covariance, ray sigmas,
// back-projected rays for the 2D-2D triangulation.  One thread per keypoint (A then B).
__global__ void k_match_setup(MatchBatch mb) {
  const int p = blockIdx.y;
  const MatchDesc& md = mb.desc[p];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < md.nA) {
    const int a = md.a0 + t;
    const SvinKeypoint kp = mb.kpA[a];
    uint8_t skip = mb.skipA_in ? mb.skipA_in[a] : 0;
    const double fA = md.intrA[0];
    mb.sigA[a] = sqrt(sqrt(2.0)) * (0.8 * (double)kp.size / 12.0) / fA;
    if (md.type == SVIN_MATCH_3D2D) {
      if (!skip) {
        const Tf T = tf_load(md.T_CbW);
        const double* hw = mb.landmarksA + 4 * (size_t)a;
        const V3 tt = m3v(T.C, V3{hw[0], hw[1], hw[2]});
        const double s = hw[3];
        const double hc[4] = {tt.x + T.r.x * s, tt.y + T.r.y * s, tt.z + T.r.z * s, s};
        double u, v, J[6] = {0, 0, 0, 0, 0, 0};
        if (project_h(md.intrB, hc, md.W, md.H, u, v, J) != 0) {
          skip = 1;
        } else {
          const double pu = md.pose_uncertainty;
          mb.proj[2 * (size_t)a] = u;
          mb.proj[2 * (size_t)a + 1] = v;
          for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j)
              mb.cov[4 * (size_t)a + i * 2 + j] =
                  pu * (J[i * 3] * J[j * 3] + J[i * 3 + 1] * J[j * 3 + 1] + J[i * 3 + 2] * J[j * 3 + 2]);
        }
      }
    } else {
      double rx, ry;
      backproject(md.intrA, (double)kp.x, (double)kp.y, rx, ry);
      mb.rayA[3 * (size_t)a] = rx;
      mb.rayA[3 * (size_t)a + 1] = ry;
      mb.rayA[3 * (size_t)a + 2] = 1.0;
    }
    mb.skipA[a] = skip;
  }
  if (t < md.nB) {
    const int b = md.b0 + t;
    const SvinKeypoint kp = mb.kpB[b];
    mb.sigB[b] = sqrt(sqrt(2.0)) * (0.8 * (double)kp.size / 12.0) / md.intrB[0];
    if (md.type == SVIN_MATCH_2D2D) {
      double rx, ry;
      backproject(md.intrB, (double)kp.x, (double)kp.y, rx, ry);
      const Tf T = tf_load(md.T_CaCb);
      const V3 r = m3v(T.C, V3{rx, ry, 1.0});
      mb.rayB[3 * (size_t)b] = r.x;
      mb.rayB[3 * (size_t)b + 1] = r.y;
      mb.rayB[3 * (size_t)b + 2] = r.z;
    }
  }
}

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// triangulateFast (stereo_triangulation.cpp:51-125)
__device__ bool triangulate_fast(const double* p1, const double* e1, const double* p2, const double* e2, double sigma,
                                 bool& isValid, double* out) {
  isValid = false;
  const double t12[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  const double b0 = dot3(t12, e1), b1 = dot3(t12, e2);
  double A00 = dot3(e1, e1), A10 = dot3(e1, e2), A01 = -A10, A11 = -dot3(e2, e2);
  if (A10 < 0.0) {
    A10 = -A10;
    A01 = -A01;
  }
  const double det = A00 * A11 - A01 * A10;
  const bool invertible = fabs(det) > 1.0e-6;  // Eigen's fixed-size 2x2 computeInverseWithCheck: absolute threshold
  double x, y, z, w;
  if (!invertible) {
    const double cr[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    if (sqrt(dot3(cr, cr)) < 6 * sigma) isValid = true;
    x = (e1[0] + e2[0]) / 2.0;
    y = (e1[1] + e2[1]) / 2.0;
    z = (e1[2] + e2[2]) / 2.0;
    w = 1e-3;
  } else {
    const double invdet = 1.0 / det;
    const double i00 = A11 * invdet, i01 = -A01 * invdet, i10 = -A10 * invdet, i11 = A00 * invdet;
    const double l0 = i00 * b0 + i01 * b1, l1 = i10 * b0 + i11 * b1;
    double mid[3], err[3], diff[3];
    for (int k = 0; k < 3; ++k) {
      const double xm = l0 * e1[k] + p1[k], xn = l1 * e2[k] + p2[k];
      mid[k] = (xm + xn) / 2.0;
      err[k] = mid[k] - xm;
      diff[k] = mid[k] - (p1[k] + 0.5 * t12[k]);
    }
    const double diff_sq = dot3(diff, diff);
    const double chi2 = dot3(err, err) * (1.0 / (diff_sq * sigma * sigma));
    isValid = true;
    if (chi2 > 9) isValid = false;
    if (dot3(diff, e1) < 0)
      for (int k = 0; k < 3; ++k) mid[k] = (p1[k] + 0.5 * t12[k]) - diff[k];
    x = mid[0];
    y = mid[1];
    z = mid[2];
    w = 1.0;
  }
  const double nrm = sqrt(x * x + y * y + z * z + w * w);
  out[0] = x / nrm;
  out[1] = y / nrm;
  out[2] = z / nrm;
  out[3] = w / nrm;
  return invertible;
}

// computeReprojectionError4 (ProbabilisticStereoTriangulator.cpp:340-365)
__device__ __forceinline__ bool reproj_err4(const double* intr, int W, int H, const SvinKeypoint& kp, const double* hp,
                                            double& err) {
  double u, v;
  if (project_h(intr, hp, W, H, u, v, nullptr) != 0) return false;
  const double sd = 0.8 * (double)kp.size / 12.0;
  const double ic = 1.0 / (sd * sd);
  const double dx = u - (double)kp.x, dy = v - (double)kp.y;
  err = dx * (ic * dx) + dy * (ic * dy);
  return true;
}

// VKWMA::verifyMatch (VKWMA.cpp:292-323)
__device__ bool verify_match(const MatchBatch& mb, const MatchDesc& md, int a, int b) {
  if (md.type == SVIN_MATCH_2D2D) {
    const double sigmaR = fmax(mb.sigA[a], mb.sigB[b]);
    const double* ra = mb.rayA + 3 * (size_t)a;
    const double* rb = mb.rayB + 3 * (size_t)b;
    const double na = sqrt(dot3(ra, ra)), nb = sqrt(dot3(rb, rb));
    const double e1[3] = {ra[0] / na, ra[1] / na, ra[2] / na}, e2[3] = {rb[0] / nb, rb[1] / nb, rb[2] / nb};
    const double p1[3] = {0, 0, 0};
    const Tf T_AB = tf_load(md.T_CaCb);
    const double p2[3] = {T_AB.r.x, T_AB.r.y, T_AB.r.z};
    bool isValid;
    double hpA[4];
    triangulate_fast(p1, e1, p2, e2, sigmaR, isValid, hpA);
    if (!isValid) return false;
    double errA, errB;
    if (!reproj_err4(md.intrA, md.W, md.H, mb.kpA[a], hpA, errA)) return false;
    const Tf T_BA = tf_inverse(T_AB);
    const V3 t = m3v(T_BA.C, V3{hpA[0], hpA[1], hpA[2]});
    const double hpB[4] = {t.x + T_BA.r.x * hpA[3], t.y + T_BA.r.y * hpA[3], t.z + T_BA.r.z * hpA[3], hpA[3]};
    if (!reproj_err4(md.intrB, md.W, md.H, mb.kpB[b], hpB, errB)) return false;
    if (errA > 4.0 || errB > 4.0) return false;
    return true;
  }
  const SvinKeypoint kb = mb.kpB[b];
  const double sd = 0.8 * (double)kb.size / 12.0;
  const double* cv = mb.cov + 4 * (size_t)a;
  const double U00 = sd * sd + cv[0], U01 = cv[1], U10 = cv[2], U11 = sd * sd + cv[3];
  const double det = U00 * U11 - U01 * U10;
  const double i00 = U11 / det, i01 = -U01 / det, i10 = -U10 / det, i11 = U00 / det;
  const double ex = mb.proj[2 * (size_t)a] - (double)kb.x, ey = mb.proj[2 * (size_t)a + 1] - (double)kb.y;
  const int chi2 = (int)((ex * i00 + ey * i10) * ex + (ex * i01 + ey * i11) * ey);
  return chi2 < 4.0;
}

// One warp per A keypoint: distance() over all B, sorted best-4 list with the reference's insertion order.
__global__ void __launch_bounds__(128) k_match(MatchBatch mb) {
  const int p = blockIdx.y;
  const MatchDesc& md = mb.desc[p];
  const int lane = threadIdx.x & 31;
  const int ai = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (ai >= md.nA) return;
  const int a = md.a0 + ai;
  int bidx[4] = {-1, -1, -1, -1};
  float bdist[4] = {md.threshold, md.threshold, md.threshold, md.threshold};
  if (!mb.skipA[a]) {
    const unsigned long long* da = reinterpret_cast<const unsigned long long*>(mb.descA + 48 * (size_t)a);
    const unsigned long long a0 = da[0], a1 = da[1], a2 = da[2], a3 = da[3], a4 = da[4], a5 = da[5];
    for (int base = 0; base < md.nB; base += 32) {
      const int bi = base + lane;
      bool cand = false;
      float d = 0.f;
      if (bi < md.nB) {
        const int b = md.b0 + bi;
        if (!(mb.skipB && mb.skipB[b])) {
          const unsigned long long* db = reinterpret_cast<const unsigned long long*>(mb.descB + 48 * (size_t)b);
          const int h = __popcll(a0 ^ db[0]) + __popcll(a1 ^ db[1]) + __popcll(a2 ^ db[2]) + __popcll(a3 ^ db[3]) +
                        __popcll(a4 ^ db[4]) + __popcll(a5 ^ db[5]);
          d = (float)h;
          if (d < md.threshold) cand = verify_match(mb, md, a, b);
        }
      }
      unsigned mask = __ballot_sync(0xffffffffu, cand);
      while (mask) {
        const int l = __ffs(mask) - 1;
        mask &= mask - 1;
        const float dl = __shfl_sync(0xffffffffu, d, l);
        const int bl = base + l;
        // listBIteration: insert if better than the worst kept, before entries of equal distance
        if (dl < bdist[3]) {
          int pos = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) pos += (bdist[k] < dl) ? 1 : 0;
#pragma unroll
          for (int k = 3; k > 0; --k)
            if (k > pos) {
              bidx[k] = bidx[k - 1];
              bdist[k] = bdist[k - 1];
            }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k == pos) {
              bidx[k] = bl;
              bdist[k] = dl;
            }
        }
      }
    }
  }
  if (lane < 4) {
    int bi = bidx[0];
    float bd = bdist[0];
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      bi = (lane == k) ? bidx[k] : bi;
      bd = (lane == k) ? bdist[k] : bd;
    }
    mb.best_idx[4 * (size_t)a + lane] = bi;
    mb.best_dist[4 * (size_t)a + lane] = bd;
  }
}

// assignbest + matchBody tail, A ascending, single worker (one thread per problem)
__global__ void k_assign(MatchBatch mb, int n_problems) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_problems) return;
  const MatchDesc& md = mb.desc[p];
  const float FMAXV = 3.402823466e+38f;
  int* mo = mb.match_of_B + md.b0;
  float* mdist = mb.match_dist + md.b0;
  for (int b = 0; b < md.nB; ++b) {
    mo[b] = -1;
    mdist[b] = FMAXV;
  }
  for (int a0 = 0; a0 < md.nA; ++a0) {
    if (mb.skipA[md.a0 + a0]) continue;
    int a = a0, start = 0;
    while (true) {
      bool reassigned = false;
      const int* bi = mb.best_idx + 4 * (size_t)(md.a0 + a);
      const float* bd = mb.best_dist + 4 * (size_t)(md.a0 + a);
      for (int index = start; index < 4 && bi[index] != -1; ++index) {
        const int b = bi[index];
        if (mo[b] == -1) {
          mo[b] = a;
          mdist[b] = bd[index];
          break;
        }
        if (bd[index] < mdist[b]) {
          const int old = mo[b];
          mo[b] = a;
          mdist[b] = bd[index];
          a = old;
          start = 1;
          reassigned = true;
          break;
        }
      }
      if (!reassigned) break;
    }
  }
  for (int b = 0; b < md.nB; ++b)
    if (!(mdist[b] < md.threshold)) mo[b] = -1;
}

// ------------------------------------------------------------------------------------------ launchers
void fe_launch_detect(const FeBatch& f, int n_images, bool use_tma, size_t occ_bytes, cudaStream_t st,
                      cudaEvent_t* ev) {
  dim3 grid((f.W + kTile - 1) / kTile, (f.H + kTile - 1) / kTile, n_images);
  if (ev) cudaEventRecord(ev[0], st);
  k_harris_nms<<<grid, 256, 0, st>>>(f);
  if (ev) cudaEventRecord(ev[1], st);
  if (n_images >= 16) {
    k_sort_image<<<n_images, 1024, 0, st>>>(f);   // throughput path: one launch, one CTA per image
  } else {
    const unsigned cap = (unsigned)f.cand_cap;
    const dim3 gl(std::max(1u, cap / kSortChunk), n_images), gg(std::max(1u, cap / 2 / 256), n_images);
    k_sort_local<<<gl, 1024, 0, st>>>(f, 0u);
    for (unsigned K = 2 * kSortChunk; K <= cap; K <<= 1) {
      for (unsigned j = K >> 1; j >= (unsigned)kSortChunk; j >>= 1) k_sort_global<<<gg, 256, 0, st>>>(f, K, j);
      k_sort_local<<<gl, 1024, 0, st>>>(f, K);
    }
  }
  if (ev) cudaEventRecord(ev[2], st);
  k_uniformity<<<n_images, kUniThreads, occ_bytes, st>>>(f);
  if (ev) cudaEventRecord(ev[3], st);
  dim3 g2((f.max_kp + 7) / 8, n_images);
  if (use_tma)
    k_orient_describe<true><<<g2, 256, 0, st>>>(f);
  else
    k_orient_describe<false><<<g2, 256, 0, st>>>(f);
  if (ev) cudaEventRecord(ev[4], st);
}
cudaError_t fe_configure(size_t occ_bytes) {
  return cudaFuncSetAttribute(k_uniformity, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)occ_bytes);
}
void fe_launch_match(const MatchBatch& mb, int n_problems, int max_nA, int max_nAB, cudaStream_t st, cudaEvent_t* ev) {
  if (ev) cudaEventRecord(ev[0], st);
  dim3 gs((max_nAB + 127) / 128, n_problems);
  k_match_setup<<<gs, 128, 0, st>>>(mb);
  dim3 gm((max_nA + 3) / 4, n_problems);
  k_match<<<gm, 128, 0, st>>>(mb);
  if (ev) cudaEventRecord(ev[1], st);
  k_assign<<<(n_problems + 63) / 64, 64, 0, st>>>(mb, n_problems);
  if (ev) cudaEventRecord(ev[2], st);
}

}  // namespace svin
