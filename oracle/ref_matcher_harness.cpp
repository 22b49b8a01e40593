// TEST INFRASTRUCTURE: drives the UNMODIFIED reference okvis::DenseMatcher (compiled from
// /root/reference/okvis_ros/okvis/okvis_matcher/src/*.cpp into oracle/_ref/) on a caller-supplied
// distance matrix, through the reference's own MatchingAlgorithm interface
// (okvis_matcher/include/okvis/MatchingAlgorithm.hpp:65-124).  Used to pin the oracle's restated
// best-N / assignbest logic against the real code (numMatcherThreads = 1 for a deterministic order).
#include <okvis/DenseMatcher.hpp>

#include <cstdint>
#include <limits>
#include <vector>

namespace {
class MatrixAlgorithm : public okvis::MatchingAlgorithm {
 public:
  int nA, nB;
  const float* dist;  // [nA][nB], FLT_MAX = no match
  const uint8_t *sA, *sB;
  float thr, ratio;
  std::vector<int> matchA;   // per B: index in A or -1
  std::vector<float> matchD;
  void doSetup() override {
    matchA.assign(nB, -1);
    matchD.assign(nB, std::numeric_limits<float>::max());
  }
  size_t sizeA() const override { return nA; }
  size_t sizeB() const override { return nB; }
  float distanceThreshold() const override { return thr; }
  float distanceRatioThreshold() const override { return ratio; }
  bool skipA(size_t i) const override { return sA && sA[i]; }
  bool skipB(size_t i) const override { return sB && sB[i]; }
  float distance(size_t a, size_t b) const override { return dist[a * (size_t)nB + b]; }
  void reserveMatches(size_t) override {}
  void setBestMatch(size_t a, size_t b, double d) override {
    matchA[b] = (int)a;
    matchD[b] = (float)d;
  }
};
}  // namespace

extern "C" int svin_ref_dense_match(int nA, int nB, const float* dist, const uint8_t* skipA, const uint8_t* skipB,
                                    float threshold, float ratio, int use_ratio, int threads, int32_t* match_of_B,
                                    float* match_dist) {
  MatrixAlgorithm alg;
  alg.nA = nA;
  alg.nB = nB;
  alg.dist = dist;
  alg.sA = skipA;
  alg.sB = skipB;
  alg.thr = threshold;
  alg.ratio = ratio;
  okvis::DenseMatcher matcher((unsigned char)threads, 4, use_ratio != 0);
  matcher.match<MatrixAlgorithm>(alg);
  for (int b = 0; b < nB; ++b) {
    match_of_B[b] = alg.matchA[b];
    match_dist[b] = alg.matchD[b];
  }
  return 0;
}
