"""Closed-loop sequence: add frame -> optimize -> applyMarginalizationStrategy -> next frame, the loop of
ThreadedKFVio::optimizationLoop (ThreadedKFVio.cpp:1086,1115) on the scene of okvis_ceres/test/TestEstimator.cpp:59-192.

CPU: the oracle driven through svin_b200.sequence.SlidingWindow meets the reference test's own tolerances
(TestEstimator.cpp:209-212) in all four extrinsics cases.  GPU: the CUDA engine and the oracle run the same sequence in
lock-step, each feeding ITS OWN solutions and priors forward; every frame's solution agrees to 1e-6 relative."""
import numpy as np
import pytest

import oracle_lib
from scene_estimator import K, add_frame, final_errors, make_sequence, new_window
from svin_b200.sequence import OracleBackend
from svin_b200.window import default_options


def _rel(a, b):
    return np.abs(a - b).max() / max(1e-300, np.abs(b).max())


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_oracle_closed_loop_meets_TestEstimator_tolerances(case):
    seq = make_sequence(case)
    sw, be, ids = new_window(seq), OracleBackend(oracle_lib), {}
    opt = default_options(max_num_iterations=10)
    dims = []
    for k in range(K + 1):
        add_frame(sw, seq, k, ids)
        sw.optimize(be, opt)                                    # TestEstimator.cpp:187
        sw.apply_marginalization_strategy(be, 2, 3)             # :192 (here after every frame, as ThreadedKFVio does)
        dims.append(0 if sw.prior is None else len(sw.prior["e0"]))
        assert len(sw.frames) <= 2 + 3 + 1
    sw.optimize(be, opt)
    sb_err, rot, trans = final_errors(sw, seq)
    assert sb_err < 0.04 and rot < 1e-2 and trans < 1e-1      # TestEstimator.cpp:209-212
    assert dims[2] == 0 and dims[3] > 0                        # the prior appears once a frame leaves the IMU window
    # the prior is a proper linear term: H = J^T J, b0 = -J^T e0 (MarginalizationError.cpp:725-758)
    p = sw.prior
    assert _rel(p["J"].T @ p["J"], p["H"]) < 1e-8
    assert np.abs(p["J"].T @ p["e0"] + p["b0"]).max() < 1e-6 * max(1.0, np.abs(p["b0"]).max())


def test_marginalization_bookkeeping_follows_the_reference_rules():
    # Estimator.cpp:528-538: beyond the newest numImuFrames frames only keyframes survive, at most numKeyframes of them;
    # :616-620 the PoseError prior of a removed first frame is dropped and the new first pose is re-fixed (:800-811)
    seq = make_sequence(0)
    sw, be, ids = new_window(seq), OracleBackend(oracle_lib), {}
    opt = default_options(max_num_iterations=4)
    for k in range(K + 1):
        add_frame(sw, seq, k, ids)
        sw.optimize(be, opt)
        sw.apply_marginalization_strategy(be, 1, 3)
        old = sw.frames[:-3]
        assert all(f.keyframe for f in old) and len(old) <= 1
        assert all(f.sb_id is None for f in old)              # speed/bias of every frame behind the IMU window is gone
    first = sw.frames[0].pose_id
    assert any(b == first and info.reshape(6, 6)[0, 0] == 1.0e14 for b, _, info in sw.pose_priors)
    used = {o[0] for o in sw.obs}
    assert used <= set(sw.landmarks)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [0, 3])
def test_cuda_closed_loop_matches_oracle_frame_by_frame(case):
    from svin_b200.engine import BaEngine
    from svin_b200.sequence import CudaBackend
    seq = make_sequence(case)
    opt = default_options(max_num_iterations=10)
    with BaEngine(0) as eng:
        gpu, ref = new_window(seq), new_window(seq)
        bg, br = CudaBackend(eng), OracleBackend(oracle_lib)
        ig, ir = {}, {}
        for k in range(K + 1):
            add_frame(gpu, seq, k, ig)
            add_frame(ref, seq, k, ir)
            sg, wg = gpu.optimize(bg, opt)
            sr, wr = ref.optimize(br, opt)
            assert sg["iterations"] == sr["iterations"] and sg["termination"] == sr["termination"], k
            assert abs(sg["final_cost"] - sr["final_cost"]) < 1e-6 * sr["final_cost"]
            assert _rel(wg.pose_blocks, wr.pose_blocks) < 1e-6, k      # north_star tolerance, every frame
            assert _rel(wg.speedbias, wr.speedbias) < 1e-6, k
            assert _rel(wg.landmarks, wr.landmarks) < 1e-6, k
            rg = gpu.apply_marginalization_strategy(bg, 2, 3)
            rr = ref.apply_marginalization_strategy(br, 2, 3)
            assert rg == rr
            assert (gpu.prior is None) == (ref.prior is None)
            if gpu.prior is not None:
                assert gpu.prior["blocks"] == ref.prior["blocks"]
                sH = np.abs(ref.prior["H"]).max()
                assert np.abs(gpu.prior["H"] - ref.prior["H"]).max() < 1e-6 * sH, k
                assert np.abs(gpu.prior["b0"] - ref.prior["b0"]).max() < 1e-6 * max(1.0, np.abs(ref.prior["b0"]).max())
        gpu.optimize(bg, opt)
        sb_err, rot, trans = final_errors(gpu, seq)
        assert sb_err < 0.04 and rot < 1e-2 and trans < 1e-1  # TestEstimator.cpp:209-212, on the CUDA chain
