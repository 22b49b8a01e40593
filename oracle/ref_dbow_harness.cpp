// TEST INFRASTRUCTURE - C entry points around the REFERENCE's own DBoW2 bag-of-words code, compiled from the sources where
// they lie (oracle/Makefile: pose_graph/ThirdParty/DBoW/BowVector.cpp + ScoringObject.cpp -> oracle/_ref/libref_dbow.so).
// Used to pin oracle/loop_oracle.py: TF-IDF accumulation (BowVector::addWeight in feature order), L1 normalisation
// (BowVector::normalize) and the L1 score (L1Scoring::score).  The vocabulary tree and the database are templates over
// OpenCV / boost types and cannot be built here; they stay restated-only.
#include <cstdint>

#include "BowVector.h"
#include "ScoringObject.h"

extern "C" {

// words / weights in FEATURE order, as TemplatedVocabulary::transform feeds them (weight <= 0 entries are skipped there,
// TemplatedVocabulary.h:1010); normalise != 0 applies L1Scoring's mustNormalize.  Returns the number of distinct words.
int ref_dbow_bow(const uint32_t* words, const double* weights, int n, int normalise, uint32_t* out_ids, double* out_vals) {
  DBoW2::BowVector v;
  for (int i = 0; i < n; ++i)
    if (weights[i] > 0) v.addWeight(words[i], weights[i]);
  if (normalise) {
    DBoW2::L1Scoring s;
    DBoW2::LNorm norm;
    if (s.mustNormalize(norm)) v.normalize(norm);
  }
  int k = 0;
  for (auto it = v.begin(); it != v.end(); ++it, ++k) {
    out_ids[k] = it->first;
    out_vals[k] = it->second;
  }
  return k;
}

double ref_dbow_l1_score(const uint32_t* ids_a, const double* val_a, int na, const uint32_t* ids_b, const double* val_b,
                         int nb) {
  DBoW2::BowVector a, b;
  for (int i = 0; i < na; ++i) a.addWeight(ids_a[i], val_a[i]);
  for (int i = 0; i < nb; ++i) b.addWeight(ids_b[i], val_b[i]);
  DBoW2::L1Scoring s;
  return s.score(a, b);
}

}  // extern "C"
