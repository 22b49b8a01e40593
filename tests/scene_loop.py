"""Synthetic keyframe descriptor sets for the loop-closure tests: a trajectory that revisits places.  A place owns a pool of
BRIEF-256 descriptors; a keyframe at a place sees a random subset of the pool with a few flipped bits plus fresh random
descriptors (new texture), so revisits share words with the first visit and other places share almost none."""
import numpy as np


def make_keyframes(n_keyframes, n_places, per_image=500, seed=0, revisit_after=None, flip_bits=6, fresh=0.2):
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
