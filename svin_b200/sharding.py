"""Host-side partitioning for the sharded single-window BA (BASELINE configs[3]).

Every rank keeps all pose / speed-bias blocks and all dense terms (IMU, priors, marginalisation prior) and a
disjoint subset of the landmarks together with their observations.  Landmark l goes to rank l % world, which
keeps the per-rank mix of track lengths balanced.  The engine all-reduces the reduced system each iteration."""
from __future__ import annotations

import numpy as np

from .window import BaWindow


def landmark_owner(num_landmarks: int, world: int) -> np.ndarray:
    return np.arange(num_landmarks) % world


def shard_window(w: BaWindow, rank: int, world: int) -> BaWindow:
    s = w.copy()
    owner = landmark_owner(w.num_landmarks, world)
    keep_lm = np.nonzero(owner == rank)[0]
    remap = -np.ones(w.num_landmarks, dtype=np.int64)
    remap[keep_lm] = np.arange(len(keep_lm))
    keep_obs = owner[w.obs_landmark] == rank
    s.landmarks = w.landmarks[keep_lm].copy()
    s.landmark_fixed = w.landmark_fixed[keep_lm].copy() if len(w.landmark_fixed) else w.landmark_fixed
    for name in ("obs_pose", "obs_extrinsics", "obs_camera", "obs_measurement", "obs_information"):
        setattr(s, name, getattr(w, name)[keep_obs].copy())
    s.obs_landmark = remap[w.obs_landmark[keep_obs]].astype(np.int32)
    return s.finalize()


def merge_landmarks(full: BaWindow, shards: list[BaWindow]) -> None:
    """Write the landmark solutions of all shards back into `full` (poses / speed-bias are identical on every rank)."""
    world = len(shards)
    owner = landmark_owner(full.num_landmarks, world)
    for r, s in enumerate(shards):
        full.landmarks[owner == r] = s.landmarks
    full.pose_blocks[:] = shards[0].pose_blocks
    full.speedbias[:] = shards[0].speedbias
