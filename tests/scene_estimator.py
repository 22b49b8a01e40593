"""The synthetic sequence of okvis_ceres/test/TestEstimator.cpp:59-192, rebuilt over svin_b200.sequence.SlidingWindow.

Constant-velocity stereo rig (v = (0, 1, 0) m/s, no rotation) in front of a landmark wall (x = 3, y in [-10, 11],
z in [-10, 10], step 0.5), 100 Hz IMU, K + 1 = 7 frames 10/6 s apart, every third one a keyframe, uniform [-1, 1] px
keypoint noise, keypoint size 8, the four extrinsics-uncertainty cases `c` of the reference test.  The reference uses the
Equidistant test camera (PinholeCamera.hpp:276-280: 752x480, f = 350/360, c = 378/238); the hot path here is
PinholeCamera<RadialTangentialDistortion>, so the same pinhole parameters are used with a radial-tangential model."""
import numpy as np

from svin_b200.sequence import SlidingWindow, propagate
from svin_b200.synthetic import project

DURATION, IMU_RATE, K = 10.0, 100.0, 6
INTR = np.array([[350.0, 360.0, 378.0, 238.0, -0.05, 0.01, 2.0e-4, -1.0e-4]] * 2)
IMU = dict(sigma_g_c=6.0e-4, sigma_a_c=2.0e-3, sigma_gw_c=3.0e-6, sigma_aw_c=2.0e-5, g=9.81, g_max=1000.0, a_max=1000.0)
T_SC = [np.array([0, 0, 0, 0, 0, 0, 1.0]), np.array([0, 0.1, 0, 0, 0, 0, 1.0])]
VEL = np.array([0.0, 1.0, 0.0])


def make_sequence(case: int, seed: int = 0):
    """-> dict(window factory inputs, frames[k] = dict(t_ns, keyframe, truth pose, observations))."""
    rng = np.random.default_rng(seed)
    dt = 1.0 / IMU_RATE
    n = int(DURATION * IMU_RATE) + 1
    t_imu = (np.arange(n) * dt * 1e9).round().astype(np.int64)
    gyro = rng.uniform(-1, 1, (n, 3)) * IMU["sigma_g_c"] * np.sqrt(dt)
    accel = np.array([0, 0, IMU["g"]]) + rng.uniform(-1, 1, (n, 3)) * IMU["sigma_a_c"] * np.sqrt(dt)
    pts = np.array([[3.0, y, z] for y in np.arange(-10.0, DURATION * 0.1 + 10.0 + 1e-9, 0.5)
                    for z in np.arange(-10.0, 10.0 + 1e-9, 0.5)])
    frames = []
    for k in range(K + 1):
        t = k * DURATION / K
        r = VEL * t
        obs = []
        for c in range(2):
            pc = pts - r - T_SC[c][:3]
            ok = pc[:, 2] > 1e-3
            ip = project(INTR[c], np.where(ok[:, None], pc, np.array([0, 0, 1.0])))
            ok &= (ip[:, 0] >= 0) & (ip[:, 0] < 752) & (ip[:, 1] >= 0) & (ip[:, 1] < 480)
            for j in np.nonzero(ok)[0]:
                obs.append((int(j), c, ip[j] + rng.uniform(-1, 1, 2)))
        frames.append(dict(t_ns=int(round(t * 1e9)), keyframe=(k % 3 == 0), pose=np.concatenate([r, [0, 0, 0, 1.0]]),
                           obs=obs))
    sig_abs = (1.0e-3 * (case % 2), 1.0e-4 * (case % 2))
    sig_rel = (1.0e-8 * (case // 2), 1.0e-7 * (case // 2))
    return dict(points=pts, frames=frames, imu=(t_imu, gyro, accel), estimate_extrinsics=(case % 2 == 1),
                sigma_abs=sig_abs, sigma_rel=sig_rel)


def new_window(seq):
    return SlidingWindow(INTR, T_SC, IMU, estimate_extrinsics=seq["estimate_extrinsics"], sigma_abs=seq["sigma_abs"],
                         sigma_rel=seq["sigma_rel"])


def add_frame(sw: SlidingWindow, seq, k: int, lm_ids: dict):
    """TestEstimator.cpp:141-186: addStates (initial values from the IMU) + addObservation for every visible point."""
    f = seq["frames"][k]
    t_imu, gyro, accel = seq["imu"]
    if k == 0:
        fid = sw.add_states(f["t_ns"], f["keyframe"], np.array([0, 0, 0, 0, 0, 0, 1.0]), np.zeros(9))
    else:
        prev = sw.frames[-1]
        pose, sb = propagate(sw.pose[prev.pose_id], sw.sb[prev.sb_id], t_imu, gyro, accel, prev.t_ns, f["t_ns"],
                             IMU["g"])
        sel = (t_imu >= prev.t_ns - 20_000_000) & (t_imu <= f["t_ns"] + 20_000_000)
        fid = sw.add_states(f["t_ns"], f["keyframe"], pose, sb, (t_imu[sel], gyro[sel], accel[sel]))
    for j, c, z in f["obs"]:
        if j not in lm_ids:
            lm_ids[j] = sw.add_landmark(np.append(seq["points"][j], 1.0))
        sw.add_observation(lm_ids[j], fid, c, z)
    return fid


def final_errors(sw: SlidingWindow, seq):
    """TestEstimator.cpp:195-212 on the newest frame."""
    f = sw.frames[-1]
    truth = seq["frames"][K]["pose"]
    est = sw.pose[f.pose_id]
    sb = sw.sb[f.sb_id]
    sb_true = np.concatenate([VEL, np.zeros(6)])
    q, qe = truth[3:7], est[3:7]
    rot = 2 * np.linalg.norm(qe[:3] * q[3] - q[:3] * qe[3] - np.cross(q[:3], qe[:3]))
    return np.linalg.norm(sb - sb_true), rot, np.linalg.norm(truth[:3] - est[:3])
