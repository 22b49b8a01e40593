"""CPU tests of the marginalisation oracle against an independent numpy implementation of the same
in-tree arithmetic (MarginalizationError.cpp:126-397, 463-758), built from the oracle's own term dump."""
import numpy as np

import oracle_lib
from svin_b200.marginalization import MargSpec, marginalization_subwindow
from svin_b200.synthetic import make_window


def _numpy_marginalize(sub, ev, mp, ms, prior_on_pose0=None):
    """Dense H, b from the raw reprojection dump (+ Cauchy correction) — only used with windows whose dense terms
    are absent, so that everything can be rebuilt from svin_oracle_ba_evaluate."""
    P = len(sub.pose_blocks)
    off = {}
    n = 0
    for i in range(P):
        if not sub.pose_fixed[i]:
            off[i] = n
            n += 6
    L = sub.num_landmarks
    N = n + 3 * L
    H, b = np.zeros((N, N)), np.zeros(N)
    for o in range(sub.num_obs):
        r = ev["reproj_residuals"][o].copy()
        s = r @ r
        rho1 = 1.0 / (1.0 + s)          # Cauchy(1): rho' ; rho'' < 0 -> plain sqrt(rho') scaling
        sq = np.sqrt(rho1)
        Jp, Jl = sq * ev["reproj_J_pose"][o], sq * ev["reproj_J_landmark"][o]
        r = sq * r
        cols, blocks = [], []
        if sub.obs_pose[o] in off:
            cols.append(np.arange(off[sub.obs_pose[o]], off[sub.obs_pose[o]] + 6)); blocks.append(Jp)
        cols.append(np.arange(n + 3 * sub.obs_landmark[o], n + 3 * sub.obs_landmark[o] + 3)); blocks.append(Jl)
        c = np.concatenate(cols)
        J = np.hstack(blocks)
        H[np.ix_(c, c)] += J.T @ J
        b[c] -= J.T @ r
    if prior_on_pose0 is not None:
        H[:6, :6] += prior_on_pose0
    # landmark part with preconditioner, block by block
    p = np.where(np.diag(H) > 1e-9, np.sqrt(np.abs(np.diag(H))), 1e-3)
    Hs, bs = H / np.outer(p, p), b / p
    U, W, V = Hs[:n, :n], Hs[:n, n:], Hs[n:, n:]
    ba, bb = bs[:n], bs[n:]
    for l in range(L):
        sl = slice(3 * l, 3 * l + 3)
        Vl = V[sl, sl]
        ev_, Uv = np.linalg.eigh(Vl)
        tol = np.finfo(float).eps * 3 * ev_.max()
        Vp = (Uv * np.where(ev_ > tol, 1.0 / np.where(ev_ > tol, ev_, 1), 0)) @ Uv.T
        U = U - W[:, sl] @ Vp @ W[:, sl].T
        ba = ba - W[:, sl] @ Vp @ bb[sl]
    pa = p[:n]
    Hd, bd = U * np.outer(pa, pa), ba * pa
    # dense part
    keep = np.array([off[i] + c for i in sorted(off) if not mp[i] for c in range(6)], dtype=int)
    marg = np.array([off[i] + c for i in sorted(off) if mp[i] for c in range(6)], dtype=int)
    if len(marg):
        p2 = np.where(np.diag(Hd) > 1e-9, np.sqrt(np.abs(np.diag(Hd))), 1e-3)
        Hs2, bs2 = Hd / np.outer(p2, p2), bd / p2
        Vm = 0.5 * (Hs2[np.ix_(marg, marg)] + Hs2[np.ix_(marg, marg)].T)
        e2, U2 = np.linalg.eigh(Vm)
        tol = np.finfo(float).eps * len(marg) * e2.max()
        Vp = (U2 * np.where(e2 > tol, 1.0 / np.where(e2 > tol, e2, 1), 0)) @ U2.T
        Wm = Hs2[np.ix_(keep, marg)]
        Hk = (Hs2[np.ix_(keep, keep)] - Wm @ Vp @ Wm.T) * np.outer(p2[keep], p2[keep])
        bk = (bs2[keep] - Wm @ Vp @ bs2[marg]) * p2[keep]
    else:
        Hk, bk = Hd[np.ix_(keep, keep)], bd[keep]
    return Hk, bk


def test_marginalization_matches_numpy_restatement():
    w, _ = make_window(seed=41, num_keyframes=4, num_imu_frames=3, num_landmarks=250, mode="initial")
    sub, mp, ms = marginalization_subwindow(w, frames_removed=1)
    # strip the IMU / speed-bias terms so that the numpy side can rebuild everything from the reprojection dump;
    # the PoseError on frame 0 stays (it fixes the gauge): at its own measurement r = 0 and J^T J = U^T U
    for name in ("imu_pose0", "imu_speedbias0", "imu_pose1", "imu_speedbias1", "imu_t0_ns", "imu_t1_ns",
                 "speedbias_prior_block", "speedbias_prior_measurement", "speedbias_prior_information"):
        setattr(sub, name, getattr(sub, name)[:0])
    sub.imu_meas_offset = np.zeros(1, np.int32)
    sub.speedbias_fixed = np.ones(len(sub.speedbias), np.uint8)  # no term touches them any more
    sub.finalize()
    assert sub.num_landmarks > 10
    none = np.zeros(len(sub.speedbias), np.uint8)
    ev = oracle_lib.evaluate(sub)
    UtU = np.diag([1e8, 1e8, 1e8, 0.0, 0.0, 1e16])  # Eigen early-exit LLT of diag(1e8,1e8,1e8,0,0,1e8)
    # (a) landmark part only
    out = oracle_lib.marginalize(sub, MargSpec(sub, np.zeros_like(mp), none))
    Hk, bk = _numpy_marginalize(sub, ev, np.zeros_like(mp), ms, prior_on_pose0=UtU)
    assert out["dim"] == Hk.shape[0] == 6 * (len(sub.pose_blocks) - 2)     # the two extrinsics blocks are fixed
    assert np.abs(out["H"] - Hk).max() < 1e-9 * np.abs(Hk).max()
    assert np.abs(out["b0"] - bk).max() < 1e-8 * np.abs(bk).max()
    # (b) landmark part + dense part (pose 0 marginalised)
    out = oracle_lib.marginalize(sub, MargSpec(sub, mp, none))
    Hk, bk = _numpy_marginalize(sub, ev, mp, ms, prior_on_pose0=UtU)
    assert out["dim"] == Hk.shape[0] == 6 * (len(sub.pose_blocks) - 3)
    sel = slice(0, 6)  # pose 1 is the only block carrying information
    assert np.abs(out["H"][sel, sel] - Hk[sel, sel]).max() < 1e-7 * np.abs(Hk[sel, sel]).max()
    assert np.abs(out["b0"][sel] - bk[sel]).max() < 1e-6 * np.abs(bk[sel]).max()


def test_error_computation_reproduces_H_and_b():
    # updateErrorComputation: J^T J == H (on the numerically non-null space) and J^T e0 == -b0
    w, _ = make_window(seed=42, num_keyframes=4, num_imu_frames=3, num_landmarks=300, mode="initial")
    sub, mp, ms = marginalization_subwindow(w, frames_removed=1)
    out = oracle_lib.marginalize(sub, MargSpec(sub, mp, ms))
    H, b0, J, e0 = out["H"], out["b0"], out["J"], out["e0"]
    assert np.abs(H - H.T).max() < 1e-9 * np.abs(H).max()
    assert np.abs(J.T @ J - H).max() < 1e-8 * np.abs(H).max()
    assert np.abs(J.T @ e0 + b0).max() < 1e-6 * max(1.0, np.abs(b0).max())
    assert np.linalg.eigvalsh(0.5 * (H + H.T)).min() > -1e-6 * np.abs(H).max()
    # second marginalisation on top of the first: feed the result back as the existing prior
    kind, index = out["kind"], out["index"]
    sub2, mp2, ms2 = marginalization_subwindow(w, frames_removed=2)
    # frame 0 is gone: drop everything that still references it
    keep_obs = sub2.obs_pose != 0
    for name in ("obs_pose", "obs_landmark", "obs_extrinsics", "obs_camera", "obs_measurement", "obs_information"):
        setattr(sub2, name, getattr(sub2, name)[keep_obs])
    ki = np.nonzero(sub2.imu_pose0 != 0)[0]
    a, b = sub2.imu_meas_offset[ki[0]], sub2.imu_meas_offset[ki[-1] + 1]
    for name in ("imu_pose0", "imu_speedbias0", "imu_pose1", "imu_speedbias1", "imu_t0_ns", "imu_t1_ns"):
        setattr(sub2, name, getattr(sub2, name)[ki])
    sub2.imu_meas_offset = (sub2.imu_meas_offset[ki[0]:ki[-1] + 2] - a).astype(np.int32)
    sub2.imu_meas_t_ns, sub2.imu_meas_gyro, sub2.imu_meas_accel = (sub2.imu_meas_t_ns[a:b], sub2.imu_meas_gyro[a:b],
                                                                     sub2.imu_meas_accel[a:b])
    for name in ("pose_prior_block", "pose_prior_measurement", "pose_prior_information", "speedbias_prior_block",
                 "speedbias_prior_measurement", "speedbias_prior_information"):
        setattr(sub2, name, getattr(sub2, name)[:0])
    sub2.pose_fixed[0] = 1
    sub2.speedbias_fixed[0] = 1
    sub2.finalize()
    mp2[0] = 0
    ms2[0] = 0
    out2 = oracle_lib.marginalize(sub2, MargSpec(sub2, mp2, ms2, kind, index, H, b0))
    assert out2["dim"] == out["dim"] - 15 + 0 or out2["dim"] > 0
    H2, J2 = out2["H"], out2["J"]
    assert np.abs(J2.T @ J2 - H2).max() < 1e-8 * np.abs(H2).max()
