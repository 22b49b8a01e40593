"""CPU: the upload path's ordering logic (svin_ba_plan = validate + order_window, no device needed).  Every
landmark lands in exactly one Schur chunk, chunks hold one (pose, camera) pattern, and each chunk kind respects the
limits of the kernel that will take it (svin_b200/csrc/ba_kernels.cu: k_schur_wr / k_schur_lr / k_schur_mma)."""
import ctypes as C

import numpy as np
import pytest

from svin_b200 import capi
from svin_b200.synthetic import make_window


def plan(w):
    lib = capi.load()
    s = w.c_struct()
    L = w.num_landmarks
    order = np.zeros(max(L, 1), dtype=np.int32)
    cap = max(L, 1)
    kind, cnt, runs = (np.zeros(cap, dtype=np.int32) for _ in range(3))
    n = C.c_int32()
    p = lambda a: a.ctypes.data_as(capi.c_int32_p)  # noqa: E731
    capi.check(lib.svin_ba_plan(C.byref(s), p(order), cap, p(kind), p(cnt), p(runs), C.byref(n)), lib)
    return order[:L], kind[:n.value], cnt[:n.value], runs[:n.value]


def patterns(w):
    pat = [[] for _ in range(w.num_landmarks)]
    for p_, l, c in zip(w.obs_pose, w.obs_landmark, w.obs_camera):
        pat[l].append((int(p_), int(c)))
    return [tuple(sorted(x)) for x in pat]


def check(w):
    order, kind, cnt, runs = plan(w)
    L = w.num_landmarks
    assert sorted(order.tolist()) == list(range(L))           # a permutation: every landmark exactly once
    assert cnt.sum() == L and (cnt > 0).all()
    pat = patterns(w)
    k = 0
    for kd, c, r in zip(kind, cnt, runs):
        lms = order[k:k + c]
        k += c
        ps = {pat[l] for l in lms}
        assert len(ps) == 1, "a chunk mixes observation patterns"
        nruns = len({p_ for p_, _ in next(iter(ps))})
        assert nruns == r
        if kd in (4, 5, 6):
            assert r == kd - 2 and 16 <= c <= 32
        elif kd == 3:
            assert r * c <= 32
        elif kd == 7:
            assert 32 < r * c <= 64 and c <= 32
        elif kd == 8:
            assert r * c <= 128 and c <= 32
        else:
            assert kd in (0, 1, 2) and c <= 32
    return kind


def test_bench_window_plan():
    w, _ = make_window(seed=20260925, num_keyframes=10, num_imu_frames=3, num_landmarks=2000, mode="steady")
    kind = check(w)
    # BASELINE configs[1]: two-view tracks dominate -> warp-per-run chunks; the long tracks are run-parallel
    assert (kind == 4).sum() > 20 and np.isin(kind, (3, 7, 8)).sum() > 20
    assert not np.isin(kind, (0, 1, 2)).any()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_ragged_windows_plan(seed):
    w, _ = make_window(seed=700 + seed, num_keyframes=4 + 2 * seed, num_imu_frames=3, num_landmarks=300 * seed,
                       mode="steady")
    rng = np.random.default_rng(seed)
    keep = rng.random(w.num_obs) > 0.25          # mono runs, holes in the tracks, landmarks without observations
    for name in ("obs_pose", "obs_landmark", "obs_extrinsics", "obs_camera", "obs_measurement", "obs_information"):
        setattr(w, name, getattr(w, name)[keep].copy())
    perm = rng.permutation(int(keep.sum()))      # and unsorted
    for name in ("obs_pose", "obs_landmark", "obs_extrinsics", "obs_camera", "obs_measurement", "obs_information"):
        setattr(w, name, getattr(w, name)[perm].copy())
    w.landmark_fixed = (rng.random(w.num_landmarks) < 0.1).astype(np.uint8)
    w._struct = None
    w.finalize()
    check(w)


def test_plan_rejects_malformed_window():
    w, _ = make_window(seed=5, num_keyframes=3, num_imu_frames=3, num_landmarks=40, mode="initial")
    w.obs_pose[0] = 99
    with pytest.raises(capi.SvinError, match="obs_pose"):
        plan(w)


@pytest.mark.parametrize("field,value,msg", [("obs_landmark", -1, "obs_landmark"), ("obs_extrinsics", 1000, "obs_extrinsics"),
                                             ("obs_camera", -3, "obs_camera"), ("obs_pose", -2147483648, "obs_pose")])
def test_plan_rejects_out_of_range_indices_of_either_sign(field, value, msg):
    # validate() checks the four index arrays in one branch-free pass (unsigned compare) and names the culprit in a second
    w, _ = make_window(seed=6, num_keyframes=3, num_imu_frames=3, num_landmarks=40, mode="initial")
    getattr(w, field)[w.num_obs // 2] = value
    with pytest.raises(capi.SvinError, match=msg):
        plan(w)
