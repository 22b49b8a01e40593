// TEST INFRASTRUCTURE — CPU oracle for hot path (A): BRISK-2 style detect / describe / match.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it.
//
// PARITY UNPINNED for detection and description: the reference delegates both to brisk 2.0.8
// (brisk::ScaleSpaceFeatureDetector<brisk::HarrisScoreCalculator>, brisk::BriskDescriptorExtractor;
// constructed at okvis_frontend/src/Frontend.cpp:997-1007, downloaded by
// okvis_ros/okvis/CMakeLists.txt:91-92), which is NOT vendored under /root/reference, and the
// reference's own tests at that boundary are smoke tests without assertions
// (okvis_cv/test/TestFrame.cpp:48-75).  What is restated here is therefore a DECLARED specification
// with the structure of brisk 2 (documented in DESIGN.md §A):
//   detect   : integer Harris score (3x3 Sobel, 3x3 box, k = 1/16), 3x3 non-maximum suppression,
//              absolute threshold, strongest-first uniformity enforcement on a half-resolution uint8
//              occupancy image with a 31x31 linear-cone kernel of radius `threshold`/2 (brisk's
//              EnforceKeyPointUniformity), cap maxNoKeypoints, 2-D parabolic sub-pixel refinement,
//              size 12 (octave 0), response = score.
//   orient   : exactly Frame::describe (okvis_cv/include/okvis/implementation/Frame.hpp:109-129):
//              backProject -> project with Jacobian -> atan2(J * extractionDirection).
//   describe : 60-point BRISK ring pattern (radii {0,2.9,4.9,7.4,10.8}*0.85, {1,10,14,15,20} points),
//              rotated by the keypoint angle quantised to 1024 steps, box-smoothed integer samples, the
//              384 shortest point pairs -> 48 bytes; keypoints closer than 16 px to the border are removed.
// What is IN-TREE and followed exactly: Hamming distance over 3 x 128 bit (VioKeyframeWindowMatchingAlgorithm.hpp:258-264),
// distance() with threshold + verifyMatch (ibid. :134-144, VKWMA.cpp:292-323), doSetup projections
// (VKWMA.cpp:163-212), stereoTriangulate / computeReprojectionError4 (ProbabilisticStereoTriangulator.cpp:154-212,340-365),
// triangulateFast (stereo_triangulation.cpp:51-125), DenseMatcher best-N lists and assignbest
// (DenseMatcher.hpp impl:216-340, DenseMatcher.cpp:59-97) executed in the declared deterministic
// order "A index ascending, single worker" (the reference's 4 pool threads race on assignbest).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/svin_b200.h"
#include "orc_terms.hpp"

namespace {
using namespace orc;

typedef SvinKeypoint Keypoint;  // cv::KeyPoint layout

constexpr int kBorder = 16;
constexpr int kNumPoints = 60;
constexpr int kNumRot = 1024;
constexpr int kNumPairs = 384;

struct Pattern {
  int8_t dx[kNumRot][kNumPoints], dy[kNumRot][kNumPoints];
  int half[kNumPoints];  // box half-width per point
  int pair_i[kNumPairs], pair_j[kNumPairs];
};

const Pattern& pattern() {
  static Pattern P;
  static bool init = false;
  if (init) return P;
  const double radii[5] = {0.0, 2.9, 4.9, 7.4, 10.8};
  const int counts[5] = {1, 10, 14, 15, 20};
  const int halves[5] = {1, 1, 2, 3, 3};
  double px[kNumPoints], py[kNumPoints];
  int idx = 0;
  for (int ring = 0; ring < 5; ++ring)
    for (int k = 0; k < counts[ring]; ++k) {
      const double a = 2.0 * M_PI * (double)k / (double)counts[ring];
      px[idx] = 0.85 * radii[ring] * std::cos(a);
      py[idx] = 0.85 * radii[ring] * std::sin(a);
      P.half[idx] = halves[ring];
      ++idx;
    }
  for (int r = 0; r < kNumRot; ++r) {
    const double th = 2.0 * M_PI * (double)r / (double)kNumRot;
    const double c = std::cos(th), s = std::sin(th);
    for (int i = 0; i < kNumPoints; ++i) {
      P.dx[r][i] = (int8_t)std::lround(c * px[i] - s * py[i]);
      P.dy[r][i] = (int8_t)std::lround(s * px[i] + c * py[i]);
    }
  }
  // the 384 shortest pairs (ties broken by (i, j)); distances compared on a 1e-9 grid
  struct PairD {
    long long d;
    int i, j;
  };
  std::vector<PairD> all;
  for (int i = 1; i < kNumPoints; ++i)
    for (int j = 0; j < i; ++j) {
      const double d = std::sqrt((px[i] - px[j]) * (px[i] - px[j]) + (py[i] - py[j]) * (py[i] - py[j]));
      all.push_back(PairD{(long long)std::llround(d * 1e9), i, j});
    }
  std::sort(all.begin(), all.end(), [](const PairD& a, const PairD& b) {
    if (a.d != b.d) return a.d < b.d;
    if (a.i != b.i) return a.i < b.i;
    return a.j < b.j;
  });
  for (int k = 0; k < kNumPairs; ++k) {
    P.pair_i[k] = all[k].i;
    P.pair_j[k] = all[k].j;
  }
  init = true;
  return P;
}

// integer Harris score image (int32), zero on a 2-pixel frame
void harris(const uint8_t* img, int stride, int W, int H, std::vector<int32_t>& score) {
  std::vector<int32_t> xx((size_t)W * H, 0), yy((size_t)W * H, 0), xy((size_t)W * H, 0);
  for (int y = 1; y < H - 1; ++y)
    for (int x = 1; x < W - 1; ++x) {
      const uint8_t* p = img + (size_t)y * stride + x;
      const int gx = (p[-stride + 1] + 2 * p[1] + p[stride + 1]) - (p[-stride - 1] + 2 * p[-1] + p[stride - 1]);
      const int gy = (p[stride - 1] + 2 * p[stride] + p[stride + 1]) - (p[-stride - 1] + 2 * p[-stride] + p[-stride + 1]);
      xx[(size_t)y * W + x] = gx * gx;
      yy[(size_t)y * W + x] = gy * gy;
      xy[(size_t)y * W + x] = gx * gy;
    }
  score.assign((size_t)W * H, 0);
  for (int y = 2; y < H - 2; ++y)
    for (int x = 2; x < W - 2; ++x) {
      long long a = 0, b = 0, c = 0;
      for (int v = -1; v <= 1; ++v)
        for (int u = -1; u <= 1; ++u) {
          a += xx[(size_t)(y + v) * W + x + u];
          b += yy[(size_t)(y + v) * W + x + u];
          c += xy[(size_t)(y + v) * W + x + u];
        }
      a >>= 10;
      b >>= 10;
      c >>= 10;  // arithmetic shift (floor), also for negative c
      const long long s = (a * b - c * c) - (((a + b) * (a + b)) >> 4);
      score[(size_t)y * W + x] = (int32_t)std::max<long long>(std::min<long long>(s, 2147483647ll), -2147483647ll);
    }
}

inline double cone(int dx, int dy, double half_radius) {
  const double d = std::sqrt((double)(dx * dx + dy * dy));
  const double v = 1.0 - d / half_radius;
  return v > 0.0 ? v : 0.0;
}

}  // namespace

extern "C" {

// atan2 of the keypoint orientation, restated with +, -, *, / only (fdlibm's s_atan.c / e_atan2.c argument reduction and
// 11-term polynomial, < 1 ulp) so that the CUDA kernel (compiled without FMA contraction) and this oracle
// (-ffp-contract=off) produce the same bits: libm's and CUDA's atan2 agree only to 1-2 ulp, and the angle feeds a
// float cast and the 1024-step rotation quantiser of the descriptor.  Against std::atan2 (what Frame.hpp impl:124 calls)
// the result differs by at most 1 ulp in fp64 (tests/test_oracle_fe.py).
inline double det_atan(double x) {
  static const double aT[11] = {3.33333333333329318027e-01,  -1.99999999998764832476e-01, 1.42857142725034663711e-01,
                                -1.11111104054623557880e-01, 9.09088713343650656196e-02,  -7.69187620504482999495e-02,
                                6.66107313738753120669e-02,  -5.83357013379057348645e-02, 4.97687799461593236017e-02,
                                -3.65315727442169155270e-02, 1.62858201153657823623e-02};
  static const double hi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01,
                               1.57079632679489655800e+00};
  static const double lo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17,
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
inline double det_atan2(double y, double x) {
  const double pi = 3.1415926535897931160e+00, pi_lo = 1.2246467991473531772e-16;
  if (y == 0.0) return x >= 0.0 ? 0.0 : pi;
  if (x == 0.0) return y > 0.0 ? 0.5 * pi : -0.5 * pi;
  const double q = y / x;
  const double z = det_atan(q < 0.0 ? -q : q);
  if (x > 0.0) return y > 0.0 ? z : -z;
  return y > 0.0 ? pi - (z - pi_lo) : (z - pi_lo) - pi;
}

// Detect + orient + describe one image.  kps/desc capacity = max_keypoints.  Returns the number of keypoints.
int svin_oracle_fe_detect_describe(const uint8_t* img, int stride, int W, int H, double uniformity_radius,
                                   double absolute_threshold, int max_keypoints, const double* intr,
                                   const double* extraction_dir, Keypoint* kps, uint8_t* desc) {
  std::vector<int32_t> score;
  harris(img, stride, W, H, score);
  // 3x3 non-maximum suppression: strictly greater than the 4 neighbours that come later in raster order,
  // greater-or-equal than the 4 earlier ones; inside the descriptor border
  struct Cand {
    int32_t s;
    int x, y;
  };
  std::vector<Cand> cand;
  const int32_t thr = (int32_t)absolute_threshold;
  for (int y = kBorder; y < H - kBorder; ++y)
    for (int x = kBorder; x < W - kBorder; ++x) {
      const int32_t s = score[(size_t)y * W + x];
      if (s < thr) continue;
      const int32_t* p = &score[(size_t)y * W + x];
      if (p[-W - 1] > s || p[-W] > s || p[-W + 1] > s || p[-1] > s) continue;
      if (p[1] >= s || p[W - 1] >= s || p[W] >= s || p[W + 1] >= s) continue;
      cand.push_back(Cand{s, x, y});
    }
  std::sort(cand.begin(), cand.end(), [&](const Cand& a, const Cand& b) {
    if (a.s != b.s) return a.s > b.s;
    return (a.y * W + a.x) < (b.y * W + b.x);
  });
  std::vector<Cand> kept;
  if (!cand.empty()) {
    const double maxScore = (double)cand.front().s;
    const int OW = W / 2 + 32, OH = H / 2 + 32;
    std::vector<uint8_t> occ((size_t)OW * OH, 0);
    const double half_radius = uniformity_radius / 2.0;
    for (const Cand& c : cand) {
      const int cy = c.y / 2 + 16, cx = c.x / 2 + 16;
      if (uniformity_radius > 0) {
        const double s0 = (double)occ[(size_t)cy * OW + cx] / 255.0;
        const double lim = s0 * s0 * s0 * s0 * maxScore;
        if ((double)c.s < lim) continue;
        const double nsc = std::sqrt(std::sqrt((double)c.s / maxScore));
        for (int dy = -15; dy <= 15; ++dy)
          for (int dx = -15; dx <= 15; ++dx) {
            const int add = (int)std::floor(255.0 * nsc * cone(dx, dy, half_radius));
            uint8_t& o = occ[(size_t)(cy + dy) * OW + cx + dx];
            o = (uint8_t)std::min(255, (int)o + add);
          }
      }
      kept.push_back(c);
      if ((int)kept.size() == max_keypoints) break;
    }
  }
  // refine, orient
  const Pattern& P = pattern();
  int n = 0;
  for (const Cand& c : kept) {
    const int32_t* p = &score[(size_t)c.y * W + c.x];
    auto parab = [](double l, double m, double r) {
      const double den = 2.0 * (l - 2.0 * m + r);
      if (den >= 0.0) return 0.0;  // not a strict maximum along this axis
      double d = (l - r) / den;
      if (d > 0.5) d = 0.5;
      if (d < -0.5) d = -0.5;
      return d;
    };
    const double ddx = parab((double)p[-1], (double)p[0], (double)p[1]);
    const double ddy = parab((double)p[-W], (double)p[0], (double)p[W]);
    Keypoint k;
    k.x = (float)((double)c.x + ddx);
    k.y = (float)((double)c.y + ddy);
    k.size = 12.0f;
    k.response = (float)c.s;
    k.octave = 0;
    k.class_id = -1;
    // Frame::describe orientation (Frame.hpp impl:113-129)
    const double ip[2] = {(double)k.x, (double)k.y};
    double ep[3], rp[2], J[6] = {0, 0, 0, 0, 0, 0};
    pinhole_backproject(intr, ip, ep);
    pinhole_project(intr, ep, rp, J);
    const double egx = J[0] * extraction_dir[0] + J[1] * extraction_dir[1] + J[2] * extraction_dir[2];
    const double egy = J[3] * extraction_dir[0] + J[4] * extraction_dir[1] + J[5] * extraction_dir[2];
    const double angle = det_atan2(egy, egx);  // std::atan2 in the reference, see det_atan2
    k.angle = (float)(angle / M_PI * 180.0);
    // descriptor: rotation bin from the float angle in degrees
    double a = (double)k.angle;
    if (a < 0) a += 360.0;
    int rot = (int)(a * (double)kNumRot / 360.0 + 0.5);
    rot &= (kNumRot - 1);
    int S[kNumPoints], A[kNumPoints];
    for (int i = 0; i < kNumPoints; ++i) {
      const int sx = c.x + P.dx[rot][i], sy = c.y + P.dy[rot][i], h = P.half[i];
      int sum = 0;
      for (int v = -h; v <= h; ++v)
        for (int u = -h; u <= h; ++u) {
          const int xx = sx + u, yy = sy + v;
          sum += (xx >= 0 && xx < W && yy >= 0 && yy < H) ? img[(size_t)yy * stride + xx] : 0;
        }
      S[i] = sum;
      A[i] = (2 * h + 1) * (2 * h + 1);
    }
    uint8_t* d = desc + (size_t)n * 48;
    std::memset(d, 0, 48);
    for (int b = 0; b < kNumPairs; ++b) {
      const int i = P.pair_i[b], j = P.pair_j[b];
      if (S[i] * A[j] > S[j] * A[i]) d[b >> 3] |= (uint8_t)(1u << (b & 7));
    }
    kps[n] = k;
    ++n;
  }
  return n;
}

// the orientation's deterministic atan2 (tests: <= 1 ulp from libm)
double svin_oracle_atan2(double y, double x) { return det_atan2(y, x); }

// raw Harris score dump (tests)
void svin_oracle_fe_harris(const uint8_t* img, int stride, int W, int H, int32_t* out) {
  std::vector<int32_t> s;
  harris(img, stride, W, H, s);
  std::memcpy(out, s.data(), sizeof(int32_t) * (size_t)W * H);
}

// brisk::Hamming::PopcntofXORed(a, b, 3): popcount over 3 x 128 bit
uint32_t svin_oracle_hamming48(const uint8_t* a, const uint8_t* b) {
  uint32_t d = 0;
  for (int k = 0; k < 48; ++k) d += (uint32_t)__builtin_popcount((unsigned)(a[k] ^ b[k]));
  return d;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ matching
namespace {

// stereo_triangulation.cpp:51-125
void triangulate_fast(const double p1[3], const double e1[3], const double p2[3], const double e2[3], double sigma,
                      bool& isValid, bool& isParallel, double out[4]) {
  isParallel = false;
  isValid = false;
  const double t12[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
  const double b0 = dot(t12, e1), b1 = dot(t12, e2);
  double A00 = dot(e1, e1), A10 = dot(e1, e2), A01 = -A10, A11 = -dot(e2, e2);
  if (A10 < 0.0) {
    A10 = -A10;
    A01 = -A01;
  }
  // Eigen computeInverseWithCheck(inverse, invertible, 1e-6) on a fixed-size 2x2 (stereo_triangulation.cpp:77-79):
  // compute_inverse_and_det_with_check<.., 2> tests the determinant against the threshold ABSOLUTELY,
  // invertible iff |det| > 1e-6 (no scaling by the largest coefficient), then inverse = adj * (1 / det).
  const double det = A00 * A11 - A01 * A10;
  const bool invertible = std::fabs(det) > 1.0e-6;
  auto normalize4 = [&](double x, double y, double z, double w) {
    const double n = std::sqrt(x * x + y * y + z * z + w * w);
    out[0] = x / n;
    out[1] = y / n;
    out[2] = z / n;
    out[3] = w / n;
  };
  if (!invertible) {
    isParallel = true;
    const double cr[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    if (std::sqrt(dot(cr, cr)) < 6 * sigma) isValid = true;
    normalize4((e1[0] + e2[0]) / 2.0, (e1[1] + e2[1]) / 2.0, (e1[2] + e2[2]) / 2.0, 1e-3);
    return;
  }
  const double invdet = 1.0 / det;
  const double i00 = A11 * invdet, i01 = -A01 * invdet, i10 = -A10 * invdet, i11 = A00 * invdet;
  const double l0 = i00 * b0 + i01 * b1, l1 = i10 * b0 + i11 * b1;
  double xm[3], xn[3], mid[3], err[3], diff[3];
  for (int k = 0; k < 3; ++k) {
    xm[k] = l0 * e1[k] + p1[k];
    xn[k] = l1 * e2[k] + p2[k];
    mid[k] = (xm[k] + xn[k]) / 2.0;
    err[k] = mid[k] - xm[k];
    diff[k] = mid[k] - (p1[k] + 0.5 * t12[k]);
  }
  const double diff_sq = dot(diff, diff);
  const double chi2 = dot(err, err) * (1.0 / (diff_sq * sigma * sigma));
  isValid = true;
  if (chi2 > 9) isValid = false;
  if (dot(diff, e1) < 0)
    for (int k = 0; k < 3; ++k) mid[k] = (p1[k] + 0.5 * t12[k]) - diff[k];
  normalize4(mid[0], mid[1], mid[2], 1.0);
}

// ProbabilisticStereoTriangulator::computeReprojectionError4 (…cpp:340-365)
bool reproj_err4(const double* intr, int W, int H, const Keypoint& kp, const double hp[4], double& err) {
  double y[2];
  if (pinhole_project_h(intr, hp, y, nullptr, W, H) != PROJ_SUCCESSFUL) return false;
  double sd = 0.8 * (double)kp.size / 12.0;
  const double ic = 1.0 / (sd * sd);
  const double dx = y[0] - (double)kp.x, dy = y[1] - (double)kp.y;
  err = dx * (ic * dx) + dy * (ic * dy);
  return true;
}

struct MatchProblem {  // view of SvinMatchProblem with the oracle's short names
  int type, nA, nB;
  const uint8_t *descA, *descB, *skipA, *skipB;
  const Keypoint *kpA, *kpB;
  float distance_threshold;
  const double *landmarksA, *T_CbW;
  double pose_uncertainty;
  const double *intrA, *intrB, *T_CaCb;
  int W, H;
};

}  // namespace

extern "C" {

// DenseMatcher::doWorkLinearMatching + listBIteration + assignbest + the matchBody tail on an explicit
// distance matrix D[nA][nB] (FLT_MAX = "absolutely no match"), worker order: A ascending, one worker.
void svin_oracle_match_matrix(int nA, int nB, const float* D, const uint8_t* skipA, const uint8_t* skipB,
                              float threshold, int32_t* best_idx, float* best_dist, int32_t* match_of_B,
                              float* match_dist) {
  const float FMAX = std::numeric_limits<float>::max();
  for (int a = 0; a < nA; ++a)
    for (int k = 0; k < 4; ++k) {
      best_idx[a * 4 + k] = -1;
      best_dist[a * 4 + k] = threshold;
    }
  for (int b = 0; b < nB; ++b) {
    match_of_B[b] = -1;
    match_dist[b] = FMAX;
  }
  // assignbest (DenseMatcher.cpp:59-97), iterative form of the tail recursion
  auto assignbest = [&](int a0, int start0) {
    int a = a0, start = start0;
    while (true) {
      bool reassigned = false;
      for (int index = start; index < 4 && best_idx[a * 4 + index] != -1; ++index) {
        const int b = best_idx[a * 4 + index];
        if (match_of_B[b] == -1) {
          match_of_B[b] = a;
          match_dist[b] = best_dist[a * 4 + index];
          return;
        }
        if (best_dist[a * 4 + index] < match_dist[b]) {
          const int old = match_of_B[b];
          match_of_B[b] = a;
          match_dist[b] = best_dist[a * 4 + index];
          a = old;
          start = 1;
          reassigned = true;
          break;
        }
      }
      if (!reassigned) return;
    }
  };
  for (int a = 0; a < nA; ++a) {
    if (skipA && skipA[a]) continue;
    for (int b = 0; b < nB; ++b) {
      if (skipB && skipB[b]) continue;
      const float d = D[(size_t)a * nB + b];
      // listBIteration (DenseMatcher.hpp impl:216-242): sorted insert, new entry before equal distances
      if (d < best_dist[a * 4 + 3]) {
        int pos = 0;
        while (pos < 4 && best_dist[a * 4 + pos] < d) ++pos;
        for (int k = 3; k > pos; --k) {
          best_idx[a * 4 + k] = best_idx[a * 4 + k - 1];
          best_dist[a * 4 + k] = best_dist[a * 4 + k - 1];
        }
        best_idx[a * 4 + pos] = b;
        best_dist[a * 4 + pos] = d;
      }
    }
    assignbest(a, 0);
  }
  // matchBody tail (DenseMatcher.hpp impl:95-119): setBestMatch only for pairings below the threshold
  for (int b = 0; b < nB; ++b)
    if (!(match_dist[b] < threshold)) match_of_B[b] = -1;
}

// Returns best[nA][4] (index in B, distance; -1 / FLT_MAX when empty), the final greedy assignment
// match_of_B[nB] (index in A or -1) + distance, and the effective skipA after doSetup.
void svin_oracle_match(const SvinMatchProblem* sp, int32_t* best_idx, float* best_dist, int32_t* match_of_B,
                       float* match_dist, uint8_t* skipA_out) {
  const MatchProblem mpv{sp->type, sp->nA, sp->nB, sp->descA, sp->descB, sp->skipA, sp->skipB, sp->kpA, sp->kpB,
                         sp->distance_threshold, sp->landmarksA, sp->T_CbW, sp->pose_uncertainty, sp->intrA, sp->intrB,
                         sp->T_CaCb, sp->image_width, sp->image_height};
  const MatchProblem* mp = &mpv;
  const int nA = mp->nA, nB = mp->nB;
  const float FMAX = std::numeric_limits<float>::max();
  std::vector<uint8_t> skipA(nA);
  std::vector<double> proj(2 * (size_t)nA, 0.0), cov(4 * (size_t)nA, 0.0), raysA(3 * (size_t)nA), raysB(3 * (size_t)nB);
  std::vector<double> sigA(nA), sigB(nB);
  const double fA = mp->intrA ? mp->intrA[0] : mp->intrB[0], fB = mp->intrB[0];
  Transform T_CbW, T_AB;
  if (mp->type == 0) T_CbW = Transform::from_params(mp->T_CbW);
  if (mp->type == 1) T_AB = Transform::from_params(mp->T_CaCb);
  for (int k = 0; k < nA; ++k) {
    skipA[k] = mp->skipA ? mp->skipA[k] : 0;
    sigA[k] = std::sqrt(std::sqrt(2.0)) * (0.8 * (double)mp->kpA[k].size / 12.0) / fA;
    if (mp->type == 0 && !skipA[k]) {
      // VKWMA.cpp:180-206
      const double* hw = mp->landmarksA + 4 * (size_t)k;
      double hc[4];
      const double s = hw[3];
      double t[3];
      mat3_vec(T_CbW.C, hw, t);
      hc[0] = t[0] + T_CbW.r[0] * s;
      hc[1] = t[1] + T_CbW.r[1] * s;
      hc[2] = t[2] + T_CbW.r[2] * s;
      hc[3] = s;
      double kpt[2], J[8];
      if (pinhole_project_h(mp->intrB, hc, kpt, J, mp->W, mp->H) != PROJ_SUCCESSFUL) {
        skipA[k] = 1;
        continue;
      }
      // J P_C J^T with P_C = diag(u,u,u,0)
      const double u = mp->pose_uncertainty;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
          cov[4 * (size_t)k + a * 2 + b] = u * (J[a * 4] * J[b * 4] + J[a * 4 + 1] * J[b * 4 + 1] + J[a * 4 + 2] * J[b * 4 + 2]);
      proj[2 * (size_t)k] = kpt[0];
      proj[2 * (size_t)k + 1] = kpt[1];
    }
    if (mp->type == 1) {
      const double ip[2] = {(double)mp->kpA[k].x, (double)mp->kpA[k].y};
      pinhole_backproject(mp->intrA, ip, &raysA[3 * (size_t)k]);
    }
  }
  for (int k = 0; k < nB; ++k) {
    sigB[k] = std::sqrt(std::sqrt(2.0)) * (0.8 * (double)mp->kpB[k].size / 12.0) / fB;
    if (mp->type == 1) {
      const double ip[2] = {(double)mp->kpB[k].x, (double)mp->kpB[k].y};
      double d[3];
      pinhole_backproject(mp->intrB, ip, d);
      mat3_vec(T_AB.C, d, &raysB[3 * (size_t)k]);
    }
  }
  if (skipA_out) std::memcpy(skipA_out, skipA.data(), nA);
  Transform T_BA;
  if (mp->type == 1) T_BA = T_AB.inverse();
  auto verify = [&](int a, int b) -> bool {
    if (mp->type == 1) {
      // stereoTriangulate (ProbabilisticStereoTriangulator.cpp:154-212)
      const double sigmaR = std::max(sigA[a], sigB[b]);
      double e1[3], e2[3];
      const double* ra = &raysA[3 * (size_t)a];
      const double* rb = &raysB[3 * (size_t)b];
      const double na = std::sqrt(ra[0] * ra[0] + ra[1] * ra[1] + ra[2] * ra[2]);
      const double nb = std::sqrt(rb[0] * rb[0] + rb[1] * rb[1] + rb[2] * rb[2]);
      for (int k = 0; k < 3; ++k) {
        e1[k] = ra[k] / na;
        e2[k] = rb[k] / nb;
      }
      const double p1[3] = {0, 0, 0};
      bool isValid, isParallel;
      double hpA[4];
      triangulate_fast(p1, e1, T_AB.r, e2, sigmaR, isValid, isParallel, hpA);
      if (!isValid) return false;
      double errA, errB;
      if (!reproj_err4(mp->intrA, mp->W, mp->H, mp->kpA[a], hpA, errA)) return false;
      double hpB[4], t[3];
      mat3_vec(T_BA.C, hpA, t);
      hpB[0] = t[0] + T_BA.r[0] * hpA[3];
      hpB[1] = t[1] + T_BA.r[1] * hpA[3];
      hpB[2] = t[2] + T_BA.r[2] * hpA[3];
      hpB[3] = hpA[3];
      if (!reproj_err4(mp->intrB, mp->W, mp->H, mp->kpB[b], hpB, errB)) return false;
      if (errA > 4.0 || errB > 4.0) return false;
      return true;
    }
    // 3D-2D chi2 gate (VKWMA.cpp:301-320; note the truncation to int)
    const double sd = 0.8 * (double)mp->kpB[b].size / 12.0;
    const double U00 = sd * sd + cov[4 * (size_t)a], U01 = cov[4 * (size_t)a + 1], U10 = cov[4 * (size_t)a + 2],
                 U11 = sd * sd + cov[4 * (size_t)a + 3];
    const double det = U00 * U11 - U01 * U10;
    const double i00 = U11 / det, i01 = -U01 / det, i10 = -U10 / det, i11 = U00 / det;
    const double ex = proj[2 * (size_t)a] - (double)mp->kpB[b].x, ey = proj[2 * (size_t)a + 1] - (double)mp->kpB[b].y;
    const int chi2 = (int)((ex * i00 + ey * i10) * ex + (ex * i01 + ey * i11) * ey);
    return chi2 < 4.0;
  };
  std::vector<float> D((size_t)nA * nB, FMAX);
  for (int a = 0; a < nA; ++a) {
    if (skipA[a]) continue;
    for (int b = 0; b < nB; ++b) {
      if (mp->skipB && mp->skipB[b]) continue;
      // distance() (VKWMA.hpp:134-144)
      float d = (float)svin_oracle_hamming48(mp->descA + 48 * (size_t)a, mp->descB + 48 * (size_t)b);
      if (!(d < mp->distance_threshold && verify(a, b))) d = FMAX;
      D[(size_t)a * nB + b] = d;
    }
  }
  svin_oracle_match_matrix(nA, nB, D.data(), skipA.data(), mp->skipB, mp->distance_threshold, best_idx, best_dist,
                           match_of_B, match_dist);
}

}  // extern "C"
