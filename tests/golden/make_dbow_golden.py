"""Generates tests/golden/dbow_golden.npz from the REFERENCE's own DBoW2 code (oracle/_ref/libref_dbow.so, built by
oracle/Makefile from /root/reference/pose_graph/ThirdParty/DBoW/{BowVector,ScoringObject}.cpp): bag-of-words vectors of random
(word, weight) sequences and their pairwise L1 scores.  Run in the build container (the reference tree is not on the GPU box)."""
import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]


def load_ref():
    lib = C.CDLL(str(ROOT / "oracle" / "_ref" / "libref_dbow.so"))
    u32p, f64p = C.POINTER(C.c_uint32), C.POINTER(C.c_double)
    lib.ref_dbow_bow.argtypes = [u32p, f64p, C.c_int, C.c_int, u32p, f64p]
    lib.ref_dbow_bow.restype = C.c_int
    lib.ref_dbow_l1_score.argtypes = [u32p, f64p, C.c_int, u32p, f64p, C.c_int]
    lib.ref_dbow_l1_score.restype = C.c_double
    return lib


def ref_bow(lib, words, weights, normalise=True):
    words = np.ascontiguousarray(words, np.uint32)
    weights = np.ascontiguousarray(weights, np.float64)
    ids, vals = np.zeros(max(len(words), 1), np.uint32), np.zeros(max(len(words), 1))
    u32p, f64p = C.POINTER(C.c_uint32), C.POINTER(C.c_double)
    n = lib.ref_dbow_bow(words.ctypes.data_as(u32p), weights.ctypes.data_as(f64p), len(words), int(normalise),
                         ids.ctypes.data_as(u32p), vals.ctypes.data_as(f64p))
    return ids[:n].astype(np.int32), vals[:n].copy()


def ref_score(lib, a, b):
    u32p, f64p = C.POINTER(C.c_uint32), C.POINTER(C.c_double)
    ia, va = np.ascontiguousarray(a[0], np.uint32), np.ascontiguousarray(a[1], np.float64)
    ib, vb = np.ascontiguousarray(b[0], np.uint32), np.ascontiguousarray(b[1], np.float64)
    return lib.ref_dbow_l1_score(ia.ctypes.data_as(u32p), va.ctypes.data_as(f64p), len(ia), ib.ctypes.data_as(u32p),
                                 vb.ctypes.data_as(f64p), len(ib))


def make_sequences(seed=7, n_seq=24):
    """(words, weights) per image: few words so that they repeat and collide between images; some zero weights."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n_seq):
        n = int(rng.integers(0, 400)) if k else 0           # the first one is empty
        words = rng.integers(0, 300, n).astype(np.uint32)
        weights = np.log(rng.uniform(1.0, 2000.0, n))
        weights[rng.random(n) < 0.05] = 0.0
        out.append((words, weights))
    return out


if __name__ == "__main__":
    lib = load_ref()
    seqs = make_sequences()
    data = {"n_seq": np.int32(len(seqs))}
    bows = []
    for k, (w, v) in enumerate(seqs):
        ids, vals = ref_bow(lib, w, v)
        data[f"words_{k}"], data[f"weights_{k}"], data[f"ids_{k}"], data[f"vals_{k}"] = w, v, ids, vals
        bows.append((ids, vals))
    data["scores"] = np.array([[ref_score(lib, a, b) for b in bows] for a in bows])
    np.savez_compressed(ROOT / "tests" / "golden" / "dbow_golden.npz", **data)
    print("wrote", len(seqs), "vectors,", data["scores"].shape, "scores")
