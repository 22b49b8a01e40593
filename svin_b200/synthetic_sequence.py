"""EuRoC-shape synthetic stereo + IMU *sequence* (BASELINE configs[0]/[1]): host-side input generation
for the closed-loop driver (svin_b200.sequence.SlidingWindow).  No engine or oracle code is involved.

20 Hz stereo frames on the trajectory generator of synthetic.py (TestImuError.cpp:91-186 style, 200 Hz IMU with the
noise densities of config_fpga_p2_euroc.yaml), a static landmark cloud at 2..15 m, at most `max_kp` observations per image
(config detection_options.maxNoKeypoints = 400), N(0, 1) px keypoint noise, a keyframe every `kf_every` frames."""
from __future__ import annotations

import numpy as np

from .sequence import SlidingWindow, propagate
from .synthetic import (EUROC_IMAGE, EUROC_IMU, EUROC_INTRINSICS, EUROC_T_SC, T_to_pose, backproject, project,
                        quat_to_rot, simulate_trajectory)


def make_euroc_sequence(seed=20260925, n_frames=30, frame_dt=0.05, kf_every=3, n_points=4000, max_kp=400,
                        pixel_noise=1.0, track_p=None, stereo_only=False):
    """track_p: None = a point is tracked for as long as it is in view (long tracks); a float = every point lives for
    clip(Geometric(track_p), 2, 13) consecutive frames from a random birth frame (k ~ clip(Geom(0.35), 2, P)),
    which gives the short-track mix of the BA bench windows."""
    rng = np.random.default_rng(seed)
    t0_ms = 500 + 2
    t_ms = [t0_ms + int(round(k * frame_dt * 1000)) for k in range(n_frames)]
    traj = simulate_trajectory(rng, t_ms[-1] / 1000.0 + 0.2, EUROC_IMU)
    bg, ba = rng.normal(0, 0.002, 3), rng.normal(0, 0.02, 3)
    gyro, accel = traj["gyr"] + bg, traj["acc"] + ba
    W, H = EUROC_IMAGE
    intr = EUROC_INTRINSICS
    T_SC = [T.copy() for T in EUROC_T_SC]

    def T_of(ms):
        T = np.eye(4)
        T[:3, :3] = quat_to_rot(traj["q"][ms])
        T[:3, 3] = traj["r"][ms]
        return T
    # landmark cloud in the frustum of camera 0 at the middle of the sequence, widened
    Tm = T_of(t_ms[n_frames // 2]) @ T_SC[0]
    px = np.stack([rng.uniform(-150, W + 150, n_points), rng.uniform(-120, H + 120, n_points)], axis=-1)
    dirs = backproject(intr[0], np.clip(px, [1, 1], [W - 2, H - 2]))
    dirs[:, 0] += (px[:, 0] - np.clip(px[:, 0], 1, W - 2)) / intr[0][0]
    dirs[:, 1] += (px[:, 1] - np.clip(px[:, 1], 1, H - 2)) / intr[0][1]
    p_c = dirs * rng.uniform(2.0, 15.0, (n_points, 1))
    points = p_c @ Tm[:3, :3].T + Tm[:3, 3]
    if track_p is not None:
        birth = rng.integers(-12, n_frames, n_points)
        death = birth + np.clip(rng.geometric(track_p, n_points), 2, 13)
    frames = []
    for k, ms in enumerate(t_ms):
        T_WS = T_of(ms)
        obs = []
        oks, ips = [], []
        for c in range(2):
            T_CW = np.linalg.inv(T_WS @ T_SC[c])
            pc = points @ T_CW[:3, :3].T + T_CW[:3, 3]
            ok = pc[:, 2] > 0.5
            ip = project(intr[c], np.where(ok[:, None], pc, np.array([0, 0, 1.0])))
            ok &= (ip[:, 0] > 2) & (ip[:, 0] < W - 3) & (ip[:, 1] > 2) & (ip[:, 1] < H - 3)
            if track_p is not None:
                ok &= (birth <= k) & (k < death)
            oks.append(ok)
            ips.append(ip)
        if stereo_only:                                           # a point is tracked only while both cameras see it
            oks = [oks[0] & oks[1]] * 2
        for c in range(2):
            vis = np.nonzero(oks[c])[0][:max_kp]                  # lowest ids first: persistent tracks
            for j in vis:
                # keypoint coordinates are floats in the reference (cv::KeyPoint::pt, widened by Frame::getKeypoint)
                z = (ips[c][j] + rng.normal(0, pixel_noise, 2)).astype(np.float32).astype(np.float64)
                obs.append((int(j), c, z))
        frames.append(dict(t_ns=ms * 1000000, keyframe=(k % kf_every == 0), pose=np.concatenate([traj["r"][ms], traj["q"][ms]]),
                           vel=traj["v"][ms].copy(), obs=obs))
    return dict(points=points, frames=frames, imu=(traj["t_imu"], gyro, accel), bias=(bg, ba), intrinsics=intr,
                T_SC=[T_to_pose(T) for T in T_SC], imu_params=dict(EUROC_IMU), rng_seed=seed)


def new_window(seq) -> SlidingWindow:
    return SlidingWindow(seq["intrinsics"], seq["T_SC"], seq["imu_params"], estimate_extrinsics=False)


def add_frame(sw: SlidingWindow, seq, k: int, lm_ids: dict, rng: np.random.Generator, landmark_noise=0.05):
    """addStates with IMU-propagated initial values + one observation per visible point; a point seen for the first time
    is created at truth (+) N(0, 5 cm) - the stand-in for the front-end's triangulation."""
    f = seq["frames"][k]
    t_imu, gyro, accel = seq["imu"]
    if k == 0:
        sb0 = np.concatenate([f["vel"] + rng.normal(0, 0.02, 3), np.zeros(6)])
        fid = sw.add_states(f["t_ns"], f["keyframe"], f["pose"], sb0)
    else:
        prev = sw.frames[-1]
        pose, sb = propagate(sw.pose[prev.pose_id], sw.sb[prev.sb_id], t_imu, gyro, accel, prev.t_ns, f["t_ns"],
                             seq["imu_params"]["g"])
        sel = (t_imu >= prev.t_ns - 20_000_000) & (t_imu <= f["t_ns"] + 20_000_000)
        fid = sw.add_states(f["t_ns"], f["keyframe"], pose, sb, (t_imu[sel], gyro[sel], accel[sel]))
    for j, c, z in f["obs"]:
        if j not in lm_ids or lm_ids[j] not in sw.landmarks:
            lm_ids[j] = sw.add_landmark(np.append(seq["points"][j] + rng.normal(0, landmark_noise, 3), 1.0))
        sw.add_observation(lm_ids[j], fid, c, z)
    return fid


def make_chain_windows(backend, seed=20260925, n_windows=8, num_keyframes=10, num_imu_frames=2, max_kp=400,
                       pose_noise=(0.0, 0.0), landmark_noise=0.0, options=None):
    """BA windows of the BASELINE configs[1] shape whose marginalisation prior comes from ACTUALLY RUNNING the window
    forward: a closed-loop chain (every frame a keyframe, so the steady-state window is num_keyframes +
    num_imu_frames consecutive frames) is driven through `backend` (CudaBackend on the GPU arm, OracleBackend on the CPU
    arm) for num_keyframes + num_imu_frames + 5 frames; then `n_windows` consecutive windows are captured right after
    addStates / addObservation, i.e. exactly what Estimator::optimize is handed: the new frame's states come from the IMU
    propagation, landmarks seen for the first time from the (5 cm noisy) triangulation stand-in, everything else from the
    previous solve.  num_imu_frames = 2 frames are kept behind the new one, so a window holds num_keyframes + 3 poses, 3
    speed/bias blocks and 2 IMU terms - the 13-pose / n = 105 shape of BASELINE configs[1].  The states are NOT perturbed
    by default (SURVEY 8(d)'s "truth (+) noise" cannot be combined with a real prior: the chain's prior pins the old
    keyframes with information up to 1e14, a 2 cm offset there is a 1e22 cost); pose_noise / landmark_noise add it anyway.
    The prior's linearisation points, J and e0 are the chain's own (svin_ba_marginalize / the oracle)."""
    from .synthetic import pose_oplus
    from .window import default_options
    P = num_keyframes + num_imu_frames
    warm = P + 5
    n_frames = warm + n_windows
    # ~max_kp live points per image: lifetime ~3.3 frames, ~60 % of the cloud projects into an image
    n_points = int(max_kp * (n_frames + 12) / 3.3 / 0.30)
    seq = make_euroc_sequence(seed=seed, n_frames=n_frames, kf_every=1, n_points=n_points, max_kp=max_kp, track_p=0.35, stereo_only=True)
    rng = np.random.default_rng(seed + 7)
    sw, ids = new_window(seq), {}
    opt = options or default_options()
    out = []
    for k in range(n_frames):
        add_frame(sw, seq, k, ids, rng)
        if k >= warm:
            w, _ = sw.flatten()
            nf = len(sw.frames)
            sig_t, sig_r = pose_noise[0], np.deg2rad(pose_noise[1])
            if sig_t > 0 or sig_r > 0:
                for f in range(nf):
                    d = np.concatenate([rng.normal(0, sig_t, 3), rng.normal(0, sig_r, 3)])
                    w.pose_blocks[f] = pose_oplus(w.pose_blocks[f], d)
            if landmark_noise > 0:
                w.landmarks[:, :3] += rng.normal(0, landmark_noise, w.landmarks[:, :3].shape)
            out.append(w.finalize())
        sw.optimize(backend, opt)
        sw.apply_marginalization_strategy(backend, num_keyframes, num_imu_frames)
    return out


def run_chain(backend, seed=20260925, n_frames=20, num_keyframes=10, num_imu_frames=2, max_kp=400, options=None):
    """Drive a closed-loop chain of the make_chain_windows shape for n_frames and return the SlidingWindow (its
    `last_marg` holds the window + spec of the last applyMarginalizationStrategy call - the steady-state B9 workload)."""
    from .window import default_options
    n_points = int(max_kp * (n_frames + 12) / 3.3 / 0.30)
    seq = make_euroc_sequence(seed=seed, n_frames=n_frames, kf_every=1, n_points=n_points, max_kp=max_kp, track_p=0.35,
                              stereo_only=True)
    rng = np.random.default_rng(seed + 7)
    sw, ids = new_window(seq), {}
    opt = options or default_options()
    for k in range(n_frames):
        add_frame(sw, seq, k, ids, rng)
        sw.optimize(backend, opt)
        sw.apply_marginalization_strategy(backend, num_keyframes, num_imu_frames)
    return sw
