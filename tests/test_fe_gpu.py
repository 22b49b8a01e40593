"""GPU parity tests for hot path (A): CUDA front-end through the C ABI vs the CPU oracle.
Bit-exact for Harris scores, keypoint pixels, descriptor bits, best-4 lists and match indices."""
import numpy as np
import pytest

import oracle_lib as ol
from svin_b200 import capi
from svin_b200.synthetic import T_to_pose
from svin_b200.synthetic_images import make_stereo_sequence, random_image, wall_point

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def seq():
    return make_stereo_sequence(seed=20260925, n_frames=2)


@pytest.fixture(scope="module")
def fe():
    from svin_b200.frontend import FeEngine
    e = FeEngine(752, 480, max_images=4)
    yield e
    e.close()


def _inv(T):
    Ti = np.eye(4)
    Ti[:3, :3] = T[:3, :3].T
    Ti[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return Ti


def _check_frame(got, ref):
    (k, d), (kr, dr) = got, ref
    assert len(k) == len(kr)
    # every field exact, the angle included: the fp64 orientation chain is evaluated without FMA contraction and with an
    # arithmetic-only atan2 on both sides, so the 1024-step rotation bin of the descriptor cannot flip
    for name in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(k[name], kr[name]), name
    assert np.array_equal(d, dr)                         # descriptor bits: exact


def test_detect_describe_bit_exact_on_rendered_stereo(fe, seq):
    imgs = [seq["images"][0][0], seq["images"][0][1], seq["images"][1][0], seq["images"][1][1]]
    intr = [seq["intrinsics"][0], seq["intrinsics"][1]] * 2
    edir = [seq["extraction_dir"][0][0], seq["extraction_dir"][0][1], seq["extraction_dir"][1][0],
            seq["extraction_dir"][1][1]]
    out = fe.detect_describe(imgs, intr, edir)
    for i in range(4):
        assert np.array_equal(fe.scores(i), ol.fe_harris(imgs[i]))
        _check_frame(out[i], ol.fe_detect_describe(imgs[i], intr[i], edir[i]))


def test_detect_describe_worst_case_random_image(fe, seq):
    # TestFrame.cpp:65-68 style uniform-random image: tens of thousands of candidates through the sort
    img = random_image(11)
    out = fe.detect_describe([img], [seq["intrinsics"][0]], [seq["extraction_dir"][0][0]])
    _check_frame(out[0], ol.fe_detect_describe(img, seq["intrinsics"][0], seq["extraction_dir"][0][0]))


def test_flat_image_yields_no_keypoints(fe, seq):
    img = np.full((480, 752), 90, dtype=np.uint8)
    out = fe.detect_describe([img], [seq["intrinsics"][0]], [seq["extraction_dir"][0][0]])
    assert len(out[0][0]) == 0


def test_tma_and_plain_tile_paths_agree(seq, monkeypatch):
    from svin_b200.frontend import FeEngine
    img = seq["images"][0][0]
    with FeEngine(752, 480, max_images=1) as a:
        ka, da = a.detect_describe([img], [seq["intrinsics"][0]], [seq["extraction_dir"][0][0]])[0]
    monkeypatch.setenv("SVIN_FE_NO_TMA", "1")
    with FeEngine(752, 480, max_images=1) as b:
        kb, db = b.detect_describe([img], [seq["intrinsics"][0]], [seq["extraction_dir"][0][0]])[0]
    assert np.array_equal(da, db) and np.array_equal(ka, kb)


def _problems(seq):
    im = seq["images"]
    feats = {}
    for f in range(2):
        for c in range(2):
            feats[(f, c)] = ol.fe_detect_describe(im[f][c], seq["intrinsics"][c], seq["extraction_dir"][f][c])
    W, H = 752, 480
    probs = []
    # stereo 2D-2D of frame 0 and frame 1
    for f in range(2):
        (k0, d0), (k1, d1) = feats[(f, 0)], feats[(f, 1)]
        T = _inv(seq["T_WC"][f][0]) @ seq["T_WC"][f][1]
        probs.append(ol.MatchArgs(capi.SVIN_MATCH_2D2D, d0, d1, k0, k1, seq["intrinsics"][0], seq["intrinsics"][1], W,
                                  H, T_CaCb=T_to_pose(T)))
    # temporal 3D-2D per camera, with some host-side skips
    rng = np.random.default_rng(5)
    for c in range(2):
        (kA, dA), (kB, dB) = feats[(0, c)], feats[(1, c)]
        pw = wall_point(seq["T_WC"][0][c], c, np.stack([kA["x"], kA["y"]], axis=1))
        lm = np.concatenate([pw, np.ones((len(pw), 1))], axis=1)
        lm[::9] *= -2.0  # homogeneous points with negative / non-unit w
        probs.append(ol.MatchArgs(capi.SVIN_MATCH_3D2D, dA, dB, kA, kB, seq["intrinsics"][c], seq["intrinsics"][c], W,
                                  H, landmarksA=lm, T_CbW=T_to_pose(_inv(seq["T_WC"][1][c])), pose_uncertainty=1e-2,
                                  skipA=(rng.uniform(size=len(kA)) < 0.1), skipB=(rng.uniform(size=len(kB)) < 0.1)))
    # temporal 2D-2D
    (kA, dA), (kB, dB) = feats[(0, 0)], feats[(1, 0)]
    T = _inv(seq["T_WC"][0][0]) @ seq["T_WC"][1][0]
    probs.append(ol.MatchArgs(capi.SVIN_MATCH_2D2D, dA, dB, kA, kB, seq["intrinsics"][0], seq["intrinsics"][0], W, H,
                              T_CaCb=T_to_pose(T)))
    return probs


def test_match_lists_and_assignment_bit_exact(fe, seq):
    probs = _problems(seq)
    got = fe.match(probs)
    for p, g in zip(probs, got):
        r = ol.fe_match(p)
        assert np.array_equal(g["skipA"], r["skipA"])
        assert np.array_equal(g["best_index"], r["best_index"])
        assert np.array_equal(g["best_distance"], r["best_distance"])
        assert np.array_equal(g["match_of_B"], r["match_of_B"])
        sel = r["match_of_B"] >= 0
        assert np.array_equal(g["match_distance"][sel], r["match_distance"][sel])
        assert sel.sum() > 30


def test_match_edge_cases(fe, seq):
    k, d = ol.fe_detect_describe(seq["images"][0][0], seq["intrinsics"][0], seq["extraction_dir"][0][0])
    W, H = 752, 480
    ident = T_to_pose(np.eye(4))
    # empty A, empty B, everything skipped
    e = np.zeros(0, dtype=k.dtype)
    ed = np.zeros((0, 48), np.uint8)
    cases = [ol.MatchArgs(capi.SVIN_MATCH_2D2D, ed, d, e, k, seq["intrinsics"][0], seq["intrinsics"][0], W, H, T_CaCb=ident),
             ol.MatchArgs(capi.SVIN_MATCH_2D2D, d, ed, k, e, seq["intrinsics"][0], seq["intrinsics"][0], W, H, T_CaCb=ident),
             ol.MatchArgs(capi.SVIN_MATCH_2D2D, d, d, k, k, seq["intrinsics"][0], seq["intrinsics"][0], W, H, T_CaCb=ident,
                          skipA=np.ones(len(k), np.uint8))]
    for p, g in zip(cases, fe.match(cases)):
        r = ol.fe_match(p)
        assert np.array_equal(g["match_of_B"], r["match_of_B"])
        assert np.array_equal(g["best_index"], r["best_index"])
        assert (g["match_of_B"] >= 0).sum() == 0


def test_orientation_and_descriptors_bit_exact_on_1e5_keypoints(seq):
    # angle -> float -> 1024-step rotation bin must not flip.  256 uniform-random images (400 keypoints
    # each, > 10^5 keypoints), varying extraction direction and both cameras' intrinsics: angles and descriptor bits exact.
    from svin_b200.frontend import FeEngine
    rng = np.random.default_rng(77)
    n_img, total = 256, 0
    with FeEngine(752, 480, max_images=64) as e:
        for start in range(0, n_img, 64):
            imgs = [random_image(1000 + start + i) for i in range(64)]
            intr = [seq["intrinsics"][(start + i) % 2] for i in range(64)]
            edir = []
            for i in range(64):
                g = rng.standard_normal(3)
                edir.append(g / np.linalg.norm(g))
            out = e.detect_describe(imgs, intr, edir)
            for i in range(64):
                kr, dr = ol.fe_detect_describe(imgs[i], intr[i], edir[i])
                k, d = out[i]
                assert len(k) == len(kr)
                assert np.array_equal(k["angle"], kr["angle"]) and np.array_equal(d, dr)
                total += len(k)
    assert total >= 100000
