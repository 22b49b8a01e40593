"""GPU parity tests for hot path (B): CUDA engine (through the C ABI) vs the CPU oracle on the
same seeded windows.  fp64 throughout; tolerances are written next to each assertion."""
import numpy as np
import pytest

import oracle_lib
from svin_b200.synthetic import make_window
from svin_b200.window import default_options

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from svin_b200.engine import BaEngine
    e = BaEngine(0)
    yield e
    e.close()


def _rel(a, b):
    return np.abs(a - b).max() / max(1e-300, np.abs(b).max())


@pytest.mark.parametrize("kw", [
    dict(seed=101, num_keyframes=4, num_imu_frames=3, num_landmarks=300, mode="initial"),
    dict(seed=102, num_keyframes=5, num_imu_frames=3, num_landmarks=500, mode="steady"),
    dict(seed=103, num_keyframes=3, num_imu_frames=3, num_landmarks=200, mode="steady", extrinsics="random_walk",
         sonar=True, depth=True),
])
def test_term_evaluation_matches_oracle(engine, kw):
    # EvaluateWithMinimalJacobians seam: residuals and minimal Jacobians of every reprojection and IMU
    # term.  Reprojection: <= 1e-12 relative.  IMU: the 15x15 covariance inverse has condition ~1e8,
    # so fused-multiply-add rounding shows up at ~1e-8 relative in the square-root information.
    w, _ = make_window(**kw)
    ref = oracle_lib.evaluate(w)
    engine.upload([w])
    got = engine.evaluate(0)
    for k in ("reproj_residuals", "reproj_J_pose", "reproj_J_landmark", "reproj_J_extrinsics"):
        assert _rel(got[k], ref[k]) < 1e-12, k
    for k in ("imu_residuals", "imu_J_pose0", "imu_J_speedbias0", "imu_J_pose1", "imu_J_speedbias1"):
        assert _rel(got[k], ref[k]) < 1e-6, (k, _rel(got[k], ref[k]))
    assert abs(got["cost"][0] - ref["cost"][0]) < 1e-7 * abs(ref["cost"][0])


@pytest.mark.parametrize("kw", [
    dict(seed=201, num_keyframes=4, num_imu_frames=3, num_landmarks=300, mode="initial"),
    dict(seed=202, num_keyframes=5, num_imu_frames=3, num_landmarks=600, mode="steady"),
    dict(seed=203, num_keyframes=10, num_imu_frames=3, num_landmarks=2000, mode="steady"),
    dict(seed=204, num_keyframes=3, num_imu_frames=3, num_landmarks=200, mode="steady", extrinsics="random_walk",
         sonar=True, depth=True),
])
def test_solve_matches_oracle_after_same_iteration_count(engine, kw):
    # north_star: pose/landmark solutions within 1e-6 relative after the same iteration count.
    w_gpu, _ = make_window(**kw)
    w_ref = w_gpu.copy()
    opt = default_options(max_num_iterations=10)
    s_ref, q_ref = oracle_lib.solve(w_ref, opt)
    s_gpu, q_gpu = engine.optimize([w_gpu], opt)
    s_gpu, q_gpu = s_gpu[0], q_gpu[0]
    assert s_gpu["iterations"] == s_ref["iterations"]
    assert s_gpu["num_successful_steps"] == s_ref["num_successful_steps"]
    assert s_gpu["termination"] == s_ref["termination"]
    assert abs(s_gpu["initial_cost"] - s_ref["initial_cost"]) < 1e-9 * s_ref["initial_cost"]
    assert abs(s_gpu["final_cost"] - s_ref["final_cost"]) < 1e-6 * s_ref["final_cost"]
    assert _rel(w_gpu.pose_blocks, w_ref.pose_blocks) < 1e-6
    assert _rel(w_gpu.speedbias, w_ref.speedbias) < 1e-6
    assert _rel(w_gpu.landmarks, w_ref.landmarks) < 1e-6
    assert np.abs(q_gpu - q_ref).max() < 1e-6


def test_batch_of_windows_is_solved_independently(engine):
    # ragged batch: different sizes, one window without IMU/prior terms, one empty of observations
    ws = [make_window(seed=300 + i, num_keyframes=3 + i, num_imu_frames=3, num_landmarks=150 + 70 * i,
                      mode="steady" if i % 2 else "initial")[0] for i in range(5)]
    refs = [w.copy() for w in ws]
    opt = default_options(max_num_iterations=6)
    s_gpu, _ = engine.optimize(ws, opt)
    for w, r, s in zip(ws, refs, s_gpu):
        s_ref, _ = oracle_lib.solve(r, opt)
        assert s["iterations"] == s_ref["iterations"]
        assert _rel(w.pose_blocks, r.pose_blocks) < 1e-6
        assert _rel(w.landmarks, r.landmarks) < 1e-6


def test_unsorted_observations_and_fixed_blocks(engine):
    w, _ = make_window(seed=401, num_keyframes=4, num_imu_frames=3, num_landmarks=250, mode="initial")
    rng = np.random.default_rng(0)
    perm = rng.permutation(w.num_obs)
    for name in ("obs_pose", "obs_landmark", "obs_extrinsics", "obs_camera", "obs_measurement", "obs_information"):
        setattr(w, name, getattr(w, name)[perm].copy())
    w.pose_fixed[2] = 1
    w.landmark_fixed = np.zeros(w.num_landmarks, dtype=np.uint8)
    w.landmark_fixed[::7] = 1
    w.finalize()
    r = w.copy()
    opt = default_options(max_num_iterations=8)
    s_ref, _ = oracle_lib.solve(r, opt)
    s_gpu, _ = engine.optimize([w], opt)
    assert s_gpu[0]["iterations"] == s_ref["iterations"]
    assert _rel(w.pose_blocks, r.pose_blocks) < 1e-6
    assert _rel(w.landmarks, r.landmarks) < 1e-6
    assert np.array_equal(w.pose_blocks[2], r.pose_blocks[2])


def test_per_observation_information_and_uniform_windows_in_one_batch(engine):
    # The upload sends one information matrix per window when all observations share it (one keypoint size) and the
    # per-observation slice otherwise; a batch mixing both kinds, the second with full 2x2 matrices (off-diagonal terms).
    wa, _ = make_window(seed=411, num_keyframes=4, num_imu_frames=3, num_landmarks=260, mode="steady")
    wb, _ = make_window(seed=412, num_keyframes=5, num_imu_frames=3, num_landmarks=300, mode="steady")
    rng = np.random.default_rng(9)
    n = wb.num_obs
    a, c = rng.uniform(0.3, 2.0, n), rng.uniform(0.3, 2.0, n)
    b_ = rng.uniform(-0.4, 0.4, n) * np.sqrt(a * c)
    wb.obs_information = np.stack([a, b_, b_, c], axis=1)
    wb.finalize()
    ra, rb = wa.copy(), wb.copy()
    opt = default_options(max_num_iterations=8)
    sa, _ = oracle_lib.solve(ra, opt)
    sb, _ = oracle_lib.solve(rb, opt)
    s, _ = engine.optimize([wa, wb], opt)
    assert [x["iterations"] for x in s] == [sa["iterations"], sb["iterations"]]
    for w, r in ((wa, ra), (wb, rb)):
        assert _rel(w.pose_blocks, r.pose_blocks) < 1e-6 and _rel(w.landmarks, r.landmarks) < 1e-6
    t = engine.timings()
    assert t["h2d_bytes"] < 60 * (wa.num_obs + wb.num_obs) + 200000   # packed indices: well below the 75 B/obs of round 1


def test_reset_and_resolve_is_repeatable(engine):
    w, _ = make_window(seed=501, num_keyframes=5, num_imu_frames=3, num_landmarks=400, mode="steady")
    engine.upload([w])
    a = engine.solve(default_options())
    engine.reset()
    b = engine.solve(default_options())
    assert a[0]["iterations"] == b[0]["iterations"]
    assert abs(a[0]["final_cost"] - b[0]["final_cost"]) < 1e-9 * abs(a[0]["final_cost"])


def test_invalid_window_is_rejected_with_message(engine):
    from svin_b200.capi import SvinError
    w, _ = make_window(seed=601, num_keyframes=3, num_imu_frames=3, num_landmarks=50, mode="initial")
    w.obs_landmark[3] = 10 ** 6
    with pytest.raises(SvinError, match="obs_landmark"):
        engine.upload([w])


def test_pipeline_matches_blocking_optimize():
    """BaPipeline (upload / solve / download_all on two contexts) gives what svin_ba_optimize gives."""
    from svin_b200.engine import BaEngine, BaPipeline
    base = [make_window(seed=900 + k, num_keyframes=4, num_imu_frames=2, num_landmarks=150)[0] for k in range(3)]
    batches = [[w.copy() for w in base[:2]], [w.copy() for w in base[1:]], [w.copy() for w in base]]
    ref = [[w.copy() for w in b] for b in batches]
    with BaEngine(0) as eng:
        ref_out = [eng.optimize(b) for b in ref]
    with BaPipeline(0) as pipe:
        out = pipe.optimize_many(batches)
    for (s0, q0), (s1, q1), b0, b1 in zip(ref_out, out, ref, batches):
        assert [s["iterations"] for s in s0] == [s["iterations"] for s in s1]
        for w0, w1, a, b in zip(b0, b1, q0, q1):
            # not bit-identical: the order of the fp64 atomics differs from run to run
            assert np.allclose(w0.pose_blocks, w1.pose_blocks, rtol=1e-7, atol=1e-9)
            assert np.allclose(w0.landmarks, w1.landmarks, rtol=1e-7, atol=1e-9)
            assert np.allclose(a, b, rtol=1e-5, atol=1e-9)


def _drop_observations(w, keep):
    for name in ("obs_pose", "obs_landmark", "obs_extrinsics", "obs_camera", "obs_measurement", "obs_information"):
        setattr(w, name, getattr(w, name)[keep].copy())
    w.finalize()


def test_mixed_track_shapes_single_camera_runs_and_single_pose_tracks(engine):
    # The Schur chunk kernels are keyed by the (pose, camera) pattern of a landmark: make the patterns ragged -
    # mono observations (one residual block per pose run), landmarks seen from one pose only - and compare with the oracle as above (1e-6 relative, same iteration count).
    w, _ = make_window(seed=777, num_keyframes=8, num_imu_frames=3, num_landmarks=900, mode="steady")
    rng = np.random.default_rng(5)
    keep = np.ones(w.num_obs, dtype=bool)
    keep &= ~((w.obs_camera == 1) & (rng.random(w.num_obs) < 0.3))          # mono runs
    first_pose = np.full(w.num_landmarks, 10 ** 9)
    np.minimum.at(first_pose, w.obs_landmark, w.obs_pose)
    single = rng.random(w.num_landmarks) < 0.1                                 # tracks of one pose
    keep &= ~(single[w.obs_landmark] & (w.obs_pose != first_pose[w.obs_landmark]))
    _drop_observations(w, keep)
    r = w.copy()
    opt = default_options(max_num_iterations=8)
    s_ref, _ = oracle_lib.solve(r, opt)
    s_gpu, _ = engine.optimize([w], opt)
    assert s_gpu[0]["iterations"] == s_ref["iterations"]
    assert s_gpu[0]["termination"] == s_ref["termination"]
    assert abs(s_gpu[0]["final_cost"] - s_ref["final_cost"]) < 1e-6 * s_ref["final_cost"]
    assert _rel(w.pose_blocks, r.pose_blocks) < 1e-6
    assert _rel(w.landmarks, r.landmarks) < 1e-6


_VARIANT_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from svin_b200.engine import BaEngine
from svin_b200.synthetic import make_window
from svin_b200.window import default_options
ws = [make_window(seed=202, num_keyframes=5, num_imu_frames=3, num_landmarks=600, mode="steady")[0],
      make_window(seed=203, num_keyframes=10, num_imu_frames=3, num_landmarks=2000, mode="steady")[0]]
with BaEngine(0) as e:
    s, q = e.optimize(ws, default_options(max_num_iterations=10))
np.savez(sys.argv[2], it=np.array([x["iterations"] for x in s]), cost=np.array([x["final_cost"] for x in s]),
         p0=ws[0].pose_blocks, p1=ws[1].pose_blocks, l0=ws[0].landmarks, l1=ws[1].landmarks)
"""


@pytest.mark.parametrize("env", [
    {"SVIN_SCHUR_LR": "0"},        # lane = landmark chunk kernels only (k_schur_mma)
    {"SVIN_SCHUR_LR": "1"},        # + single-warp run-parallel chunks
    {"SVIN_SOLVE_SMEM": "1", "SVIN_GRAM_FMA": "1"},  # shared-memory Cholesky, FMA Gram matrix
    {"SVIN_BA_GRAPH": "0", "SVIN_BA_NO_FORK": "1"},  # no CUDA graph, single stream
])
def test_kernel_variants_agree(tmp_path, env):
    # Every kernel variant behind an A/B knob computes the same solution (different summation orders: 1e-7).
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "variant.py"
    script.write_text(_VARIANT_SCRIPT)
    outs = []
    for k, e in enumerate(({}, env)):
        out = tmp_path / f"v{k}.npz"
        subprocess.run([sys.executable, str(script), root, str(out)], check=True, env={**os.environ, **e}, timeout=600)
        outs.append(np.load(out))
    a, b = outs
    assert np.array_equal(a["it"], b["it"])
    assert np.allclose(a["cost"], b["cost"], rtol=1e-9)
    for k in ("p0", "p1", "l0", "l1"):
        assert np.allclose(a[k], b[k], rtol=1e-7, atol=1e-9), k


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_TestMarginalization_scene_estimated_extrinsics_no_loss(engine, seed):
    # The scene of okvis_ceres/test/TestMarginalization.cpp:60-231 (two constant poses, third pose + camera extrinsics
    # estimated with a PoseError prior, 100 points, no loss function): the CUDA engine takes the general Schur path
    # (estimated extrinsics), must agree with the oracle and meet the reference test's own tolerances.
    from scene_marginalization import make_scene, pose_errors
    w, truth = make_scene(seed)
    r = w.copy()
    opt = default_options(max_num_iterations=50)
    s_ref, _ = oracle_lib.solve(r, opt, quality=False)
    s, _ = engine.optimize([w], opt)
    assert s[0]["iterations"] == s_ref["iterations"] and s[0]["termination"] == s_ref["termination"]
    assert _rel(w.pose_blocks, r.pose_blocks) < 1e-6 and _rel(w.landmarks, r.landmarks) < 1e-6
    rot, trans = pose_errors(w.pose_blocks[2], truth)
    assert rot < 1.0e-2 and trans < 1.0e-1          # TestMarginalization.cpp:226-231
