"""TEST INFRASTRUCTURE - CPU oracle (numpy / plain Python) of the loop-closure QUERY path of pose_graph
(BASELINE configs[4]).  Only tests/, smoke() and bench.py's CPU legs may import this file; the
product path is svin_b200/csrc/loop_engine.cu behind svin_loop_* (include/svin_b200.h).

Restated from the VENDORED sources under /root/reference/pose_graph.  PINNED against the reference's own code where that
compiles here: BowVector.cpp + ScoringObject.cpp -> oracle/_ref/libref_dbow.so (oracle/Makefile) - bag-of-words accumulation,
L1 normalisation and the L1 score are bit-identical to it on the committed golden vectors (tests/golden/dbow_golden.npz,
made by tests/golden/make_dbow_golden.py) and, when _ref is present, on fresh random vectors.  The vocabulary tree and the
database (TemplatedVocabulary.h / TemplatedDatabase.h) are templates over OpenCV + boost types, unbuildable in this image:
those parts follow the sources line by line and are pinned to hand-computed vectors only (tests/test_loop.py):
  * BRIEF-256 distance           FBrief::distance = popcount(a ^ b)            ThirdParty/DBoW/FBrief.cpp:44
  * word lookup                  TemplatedVocabulary::transform(feature, id, weight): from the root, at every level the
                                 child with the smallest distance, first one on ties (`d < best_d`)
                                                                              ThirdParty/DBoW/TemplatedVocabulary.h (transform)
  * bag of words                 transform(features, BowVector): TF-IDF weighting - BowVector::addWeight accumulates the
                                 word's idf weight per feature in feature order - then L1 normalisation (L1Scoring
                                 mustNormalize)                                TemplatedVocabulary.h:981-1028
  * L1 score                     sum over common words of |v - w| - |v| - |w|, score = -sum / 2   (Nister 2006)
  * database                     add(): one inverted-file row entry per word; queryL1(vec, max_results, max_id): entries
                                 with id < max_id, ascending score, cut, score = -score / 2
                                                                              ThirdParty/DBoW/TemplatedDatabase.h:587-646
  * candidate search             Keyframe::searchInAera / searchByBRIEFDes: best Hamming distance (< 128, strict), accepted
                                 if < 80                                       src/pose_graph/Keyframe.cpp:262-306
  * decision                     PoseGraph::detectLoop: query top 4 among entries older than frame_index - 50, a loop if a
                                 score exceeds 0.6 * the minimum score against the connected keyframes
                                                                              src/pose_graph/PoseGraph.cpp:170-224
Ties in std::sort(ret) are unspecified in the reference; declared here: equal scores keep ascending entry id.
Not restated (not built): FAST + BRIEF extraction, PnPRANSAC, the 4-/6-DoF pose-graph optimisation.
"""
from __future__ import annotations

import numpy as np

_POP8 = np.array([bin(i).count("1") for i in range(256)], dtype=np.int32)


def hamming(a, b):
    """a [..., 32] uint8, b [..., 32] uint8 -> popcount(a ^ b)."""
    return _POP8[np.bitwise_xor(a, b)].sum(axis=-1)


class Vocabulary:
    """Tree in the flattened layout of the C ABI: children of a node are contiguous."""

    def __init__(self, first_child, num_children, descriptor, weight, word_id):
        self.first_child = np.asarray(first_child, np.int32)
        self.num_children = np.asarray(num_children, np.int32)
        self.descriptor = np.asarray(descriptor, np.uint8).reshape(-1, 32)
        self.weight = np.asarray(weight, np.float64)
        self.word_id = np.asarray(word_id, np.int32)

    @staticmethod
    def random(k, L, seed=0):
        """Synthetic k-ary tree of depth L (svin_b200.synthetic_loop.random_vocabulary: the inputs' generator is shared with
        the bench, the arithmetic below is not)."""
        from svin_b200.synthetic_loop import random_vocabulary
        v = random_vocabulary(k, L, seed)
        return Vocabulary(v["first_child"], v["num_children"], v["descriptor"], v["weight"], v["word_id"])

    def transform_feature(self, f):
        node = 0
        while self.num_children[node] > 0:
            c0, nc = self.first_child[node], self.num_children[node]
            d = hamming(self.descriptor[c0:c0 + nc], f[None, :])
            node = c0 + int(np.argmin(d))          # first minimum = `d < best_d`
        return int(self.word_id[node]), float(self.weight[node])

    def words(self, features, chunk=65536):
        """Vectorised transform_feature over [n, 32] features -> (word ids, weights); same arithmetic (np.argmin returns
        the first minimum), used where the per-feature loop is too slow (BASELINE configs[4]: 5k keyframes)."""
        f = np.asarray(features, np.uint8).reshape(-1, 32)
        node = np.zeros(len(f), np.int64)
        for s in range(0, len(f), chunk):
            cur = node[s:s + chunk]
            ff = f[s:s + chunk]
            while True:
                nc = self.num_children[cur]
                act = np.nonzero(nc > 0)[0]
                if len(act) == 0:
                    break
                kmax = int(nc[act].max())
                c0 = self.first_child[cur[act]].astype(np.int64)
                cand = c0[:, None] + np.arange(kmax)[None, :]
                valid = np.arange(kmax)[None, :] < nc[act][:, None]
                d = hamming(self.descriptor[np.where(valid, cand, 0)], ff[act][:, None, :])
                d = np.where(valid, d, 1 << 20)
                cur[act] = c0 + np.argmin(d, axis=1)
            node[s:s + chunk] = cur
        return self.word_id[node], self.weight[node]

    def transform(self, features, fast=False):
        """-> (word ids ascending, L1-normalised values)."""
        if fast:
            w, v = (x.tolist() for x in self.words(features)) if len(features) else ([], [])
        else:
            pairs = [self.transform_feature(f) for f in features]
            w, v = [p[0] for p in pairs], [p[1] for p in pairs]
        return bow_from_words(w, v)


def bow_from_words(words, weights, normalise=True):
    """TemplatedVocabulary::transform's accumulation: words with weight > 0 are added in FEATURE order
    (BowVector::addWeight, BowVector.cpp:30-38), then BowVector::normalize(L1) sums |v| in ascending word order and divides
    (BowVector.cpp:52-66).  Pinned against the reference's own BowVector (oracle/_ref/libref_dbow.so, tests/golden/dbow_golden.npz)."""
    acc = {}
    for w, v in zip(words, weights):
        if v > 0:
            acc[w] = acc.get(w, 0.0) + v
    ids = sorted(acc)
    vals = np.array([acc[i] for i in ids], dtype=np.float64)
    if normalise:
        norm = 0.0
        for x in vals:
            norm += abs(x)
        if norm > 0.0:
            vals = vals / norm
    return np.array(ids, dtype=np.int32), vals


def l1_score(ids_a, val_a, ids_b, val_b):
    s, i, j = 0.0, 0, 0
    while i < len(ids_a) and j < len(ids_b):
        if ids_a[i] == ids_b[j]:
            s += abs(val_a[i] - val_b[j]) - abs(val_a[i]) - abs(val_b[j])
            i += 1
            j += 1
        elif ids_a[i] < ids_b[j]:
            i += 1
        else:
            j += 1
    return -s / 2.0


class Database:
    def __init__(self, voc: Vocabulary, fast=False):
        self.voc = voc
        self.fast = fast
        self.entries = []          # (ids, values) per entry id

    def add(self, features):
        self.entries.append(self.voc.transform(features, self.fast))
        return len(self.entries) - 1

    def query(self, features, max_results=4, max_id=-1):
        ids, vals = self.voc.transform(features, self.fast)
        out = []
        for e, (eid, ev) in enumerate(self.entries):
            if not (e < max_id or max_id == -1):
                continue
            s, i, j, common = 0.0, 0, 0, False
            while i < len(ids) and j < len(eid):
                if ids[i] == eid[j]:
                    s += abs(vals[i] - ev[j]) - abs(vals[i]) - abs(ev[j])
                    common = True
                    i += 1
                    j += 1
                elif ids[i] < eid[j]:
                    i += 1
                else:
                    j += 1
            if common:
                out.append((s, e))
        out.sort()                                   # ascending score, ties by entry id
        out = out[:max_results] if max_results > 0 else out
        return [(e, -s / 2.0) for s, e in out]


def search_by_brief(window_desc, old_desc):
    """Keyframe::searchByBRIEFDes -> (best index per window descriptor, its distance, status)."""
    n = len(window_desc)
    idx = np.full(n, -1, np.int32)
    dist = np.full(n, 128, np.int32)
    for i in range(n):
        if len(old_desc) == 0:
            continue
        d = hamming(old_desc, window_desc[i][None, :])
        j = int(np.argmin(d))
        if d[j] < 128:
            idx[i], dist[i] = j, d[j]
    status = (idx >= 0) & (dist < 80)
    return idx, dist, status.astype(np.uint8)


def detect_loop(results, min_score, frame_index):
    """PoseGraph.cpp:198-221 on the query results [(entry, score)] -> loop candidate index or -1."""
    find = any(s > 0.60 * min_score for _, s in results)
    if not (find and frame_index > 50):
        return -1
    best_index, best_score = -1, 0.0
    for e, s in results:
        if best_index == -1 or (s > best_score and s > 0.60 * min_score):
            best_index, best_score = e, s
    return best_index
