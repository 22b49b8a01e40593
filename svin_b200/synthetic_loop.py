"""Synthetic inputs of the loop-closure query path (BASELINE configs[4]): a k-ary BRIEF vocabulary tree of the shape of the
shipped `brief_k10L6.bin` (k = 10, L = 6; the file itself is not in the reference tree) in the flattened layout of
`SvinVocabulary` (include/svin_b200.h), and keyframe descriptor sets of a trajectory that revisits places."""
from __future__ import annotations

import numpy as np


def random_vocabulary(k: int, L: int, seed: int = 0):
    """-> dict(first_child, num_children, descriptor [n][32], weight, word_id): node 0 is the root, the children of a node are
    contiguous, leaves carry a word id (in node order) and an idf-like weight ln(N / Ni)."""
    rng = np.random.default_rng(seed)
    first, num, word = [0], [0], [-1]
    level = [0]
    for _ in range(L):
        nxt = []
        for node in level:
            first[node] = len(first)
            num[node] = k
            for _ in range(k):
                nxt.append(len(first))
                first.append(0)
                num.append(0)
                word.append(-1)
        level = nxt
    n = len(first)
    wid = 0
    for node in range(n):
        if num[node] == 0:
            word[node] = wid
            wid += 1
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    weight = np.where(np.array(num) == 0, np.log(rng.uniform(2.0, 2000.0, n)), 0.0)
    return dict(first_child=np.array(first, np.int32), num_children=np.array(num, np.int32), descriptor=desc,
                weight=weight.astype(np.float64), word_id=np.array(word, np.int32))


def make_keyframes(n_keyframes, n_places, per_image=500, seed=0, revisit_after=None, flip_bits=6, fresh=0.2):
    """A place owns a pool of BRIEF-256 descriptors; a keyframe at a place sees a random subset of the pool with a few flipped
    bits plus fresh random descriptors (new texture), so revisits share words with the first visit and other places share
    almost none.  -> ([descriptors [per_image][32] per keyframe], place of every keyframe)."""
    rng = np.random.default_rng(seed)
    pools = rng.integers(0, 256, (n_places, per_image, 32), dtype=np.uint8)
    revisit_after = n_places if revisit_after is None else revisit_after
    frames, place_of = [], []
    for i in range(n_keyframes):
        p = i % revisit_after % n_places
        n_old = int(per_image * (1.0 - fresh))
        pick = rng.permutation(per_image)[:n_old]
        d = pools[p, pick].copy()
        for _ in range(flip_bits):
            byte = rng.integers(0, 32, n_old)
            bit = rng.integers(0, 8, n_old)
            d[np.arange(n_old), byte] ^= (1 << bit).astype(np.uint8)
        new = rng.integers(0, 256, (per_image - n_old, 32), dtype=np.uint8)
        frames.append(np.concatenate([d, new])[rng.permutation(per_image)])
        place_of.append(p)
    return frames, np.array(place_of)
