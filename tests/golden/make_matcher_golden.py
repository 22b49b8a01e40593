#!/usr/bin/env python
"""Generates tests/golden/matcher_golden.npz by running the UNMODIFIED reference okvis::DenseMatcher
(compiled from /root/reference into oracle/_ref by oracle/Makefile) on seeded distance matrices.
Run in the build container (needs /root/reference):  python tests/golden/make_matcher_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib  # noqa: E402

FMAX = np.finfo(np.float32).max


def cases():
    rng = np.random.default_rng(20260925)
    out = []
    for i in range(40):
        nA, nB = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        D = rng.integers(0, 90, (nA, nB)).astype(np.float32)  # integer Hamming-like distances -> many ties
        D[rng.uniform(size=D.shape) < 0.5] = FMAX
        skipA = (rng.uniform(size=nA) < 0.15).astype(np.uint8)
        skipB = (rng.uniform(size=nB) < 0.15).astype(np.uint8)
        out.append((D, skipA, skipB, 60.0))
    return out


def main():
    data = {}
    for k, (D, sA, sB, thr) in enumerate(cases()):
        mb, md = oracle_lib.ref_match_matrix(D, sA, sB, thr, threads=1)
        data[f"D{k}"], data[f"sA{k}"], data[f"sB{k}"] = D, sA, sB
        data[f"mb{k}"], data[f"md{k}"] = mb, md
    data["n"] = np.array(len(cases()))
    np.savez_compressed(os.path.join(HERE, "matcher_golden.npz"), **data)
    print("wrote", len(cases()), "cases")


if __name__ == "__main__":
    main()
