"""Seeded synthetic sliding windows of the shapes BASELINE.json names.

Recipe: stereo rig with the EuRoC calibration of
config/config_fpga_p2_euroc.yaml, a smooth sinusoidal trajectory generated like
okvis_ceres/test/TestImuError.cpp:91-186 (1 kHz truth, 200 Hz IMU with the config's
noise densities), landmarks in the camera frusta at depth U[2,15] m observed in both
cameras of consecutive frames, N(0,1) px keypoint noise, keypoint size 8 (information
64/size^2 = I, Estimator.hpp impl:64-67), Cauchy(1) loss, perturbed initial values.

This is host-side input generation only; no engine or oracle code is involved.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .window import BaWindow

# config/config_fpga_p2_euroc.yaml:3-23
EUROC_T_SC = [
    np.array([[0.0148655429818, -0.999880929698, 0.00414029679422, -0.0216401454975],
              [0.999557249008, 0.0149672133247, 0.025715529948, -0.064676986768],
              [-0.0257744366974, 0.00375618835797, 0.999660727178, 0.00981073058949],
              [0, 0, 0, 1.0]]),
    np.array([[0.0125552670891, -0.999755099723, 0.0182237714554, -0.0198435579556],
              [0.999598781151, 0.0130119051815, 0.0251588363115, 0.0453689425024],
              [-0.0253898008918, 0.0179005838253, 0.999517347078, 0.00786212447038],
              [0, 0, 0, 1.0]]),
]
EUROC_INTRINSICS = np.array([
    [458.654880721, 457.296696463, 367.215803962, 248.37534061,
     -0.28340811217, 0.0739590738929, 0.000193595028569, 1.76187114545e-05],
    [457.587426604, 456.13442556, 379.99944652, 255.238185386,
     -0.283683654496, 0.0745128430929, -0.000104738949098, -3.55590700274e-05],
])
EUROC_IMAGE = (752, 480)
EUROC_IMU = dict(sigma_g_c=12.0e-4, sigma_a_c=8.0e-3, sigma_gw_c=4.0e-6, sigma_aw_c=4.0e-5, g=9.81007,
                 g_max=7.8, a_max=176.0)
# config/config_stereorig_v2.yaml sonar extrinsics (T_SSo), row-major 4x4 -> we keep [r, q]
STEREORIG_T_SSO = np.array([0.365, 0.095, 0.165, 0.0, 0.0, 0.0, 1.0])


# ------------------------------------------------------------------ kinematics (numpy)
def quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def quat_to_rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def rot_to_quat(R):
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    return q / np.linalg.norm(q)


def delta_q(dalpha):
    half = 0.5 * np.linalg.norm(dalpha)
    s = np.sinc(half / np.pi) * 0.5
    return np.array([s * dalpha[0], s * dalpha[1], s * dalpha[2], np.cos(half)])


def pose_oplus(pose, delta):
    """okvis Transformation::oplus on a [x y z qx qy qz qw] pose."""
    out = np.array(pose, dtype=np.float64)
    out[:3] += delta[:3]
    q = quat_mul(delta_q(delta[3:6]), out[3:7])
    out[3:7] = q / np.linalg.norm(q)
    return out


def T_to_pose(T):
    return np.concatenate([T[:3, 3], rot_to_quat(T[:3, :3])])


# ------------------------------------------------------------------ camera (numpy, vectorised)
def distort(intr, u):
    k1, k2, p1, p2 = intr[4:8]
    u0, u1 = u[..., 0], u[..., 1]
    mx, my, mxy = u0 * u0, u1 * u1, u0 * u1
    rho = mx + my
    rad = k1 * rho + k2 * rho * rho
    return np.stack([u0 + u0 * rad + 2 * p1 * mxy + p2 * (rho + 2 * mx),
                     u1 + u1 * rad + 2 * p2 * mxy + p1 * (rho + 2 * my)], axis=-1)


def project(intr, p):
    und = p[..., :2] / p[..., 2:3]
    d = distort(intr, und)
    return np.stack([intr[0] * d[..., 0] + intr[2], intr[1] * d[..., 1] + intr[3]], axis=-1)


def backproject(intr, ip, iters=8):
    y = np.stack([(ip[..., 0] - intr[2]) / intr[0], (ip[..., 1] - intr[3]) / intr[1]], axis=-1)
    x = y.copy()
    for _ in range(iters):
        eps = 1e-6
        f0 = distort(intr, x)
        fx = (distort(intr, x + np.array([eps, 0])) - f0) / eps
        fy = (distort(intr, x + np.array([0, eps])) - f0) / eps
        e = y - f0
        det = fx[..., 0] * fy[..., 1] - fy[..., 0] * fx[..., 1]
        dx0 = (fy[..., 1] * e[..., 0] - fy[..., 0] * e[..., 1]) / det
        dx1 = (-fx[..., 1] * e[..., 0] + fx[..., 0] * e[..., 1]) / det
        x = x + np.stack([dx0, dx1], axis=-1)
    return np.concatenate([x, np.ones(x.shape[:-1] + (1,))], axis=-1)


# ------------------------------------------------------------------ trajectory + IMU
def simulate_trajectory(rng, duration, imu_params, fine_rate=1000, imu_div=5):
    """TestImuError.cpp:91-186 style generator.  Returns fine-rate truth and 200 Hz measurements."""
    w_om, p_om = rng.uniform(0.5, 3.0, 3), rng.uniform(0.0, np.pi, 3)
    m_om = rng.uniform(0.05, 0.3, 3)
    w_a, p_a = rng.uniform(0.5, 3.0, 3), rng.uniform(0.1, np.pi, 3)
    m_a = rng.uniform(0.1, 1.0, 3)
    n = int(duration * fine_rate)
    dt = 1.0 / fine_rate
    q = np.array([0, 0, 0, 1.0])
    r = np.zeros(3)
    v = np.zeros(3)
    R_all, q_all, r_all, v_all = [], np.zeros((n, 4)), np.zeros((n, 3)), np.zeros((n, 3))
    gyr, acc, t_imu = [], [], []
    imu_dt = dt * imu_div
    for i in range(n):
        time = i * dt
        omega_S = m_om * np.sin(w_om * time + p_om)
        a_W = m_a * np.sin(w_a * time + p_a)
        q = quat_mul(q, delta_q(omega_S * dt))
        q = q / np.linalg.norm(q)
        v = v + dt * a_W
        r = r + dt * v
        q_all[i], r_all[i], v_all[i] = q, r, v
        if i % imu_div == 0:
            C_WS = quat_to_rot(q)
            gyr.append(omega_S + imu_params["sigma_g_c"] / np.sqrt(imu_dt) * rng.standard_normal(3))
            acc.append(C_WS.T @ (a_W + np.array([0, 0, imu_params["g"]])) +
                       imu_params["sigma_a_c"] / np.sqrt(imu_dt) * rng.standard_normal(3))
            t_imu.append(int(round(time * 1e9)))
    return dict(q=q_all, r=r_all, v=v_all, dt=dt, t_imu=np.array(t_imu, dtype=np.int64), gyr=np.array(gyr),
                acc=np.array(acc))


def make_marg_prior(rng, dims, scale_lo=1e2, scale_hi=1e5, b_scale=0.5):
    """A dense PSD prior H,b0 and its (J, e0) via MarginalizationError::updateErrorComputation
    (MarginalizationError.cpp:725-758), computed with numpy's eigh."""
    n = int(sum(dims))
    A = rng.standard_normal((n, n))
    Q, _ = np.linalg.qr(A)
    lam = np.exp(rng.uniform(np.log(scale_lo), np.log(scale_hi), n))
    H = (Q * lam) @ Q.T
    H = 0.5 * (H + H.T)
    b0 = b_scale * rng.standard_normal(n) * np.sqrt(np.diag(H)) * 1e-2
    d = np.diag(H)
    p = np.where(d > 1.0e-9, np.sqrt(np.abs(d)), 1.0e-3)
    p_inv = 1.0 / p
    Hs = 0.5 * (p_inv[:, None] * (H + H.T) * p_inv[None, :])
    ev, U = np.linalg.eigh(Hs)
    tol = np.finfo(np.float64).eps * n * ev.max()
    S = np.where(ev > tol, ev, 0.0)
    S_pinv = np.where(ev > tol, 1.0 / np.where(ev > tol, ev, 1.0), 0.0)
    J = ((p[:, None] * U) * np.sqrt(S)[None, :]).T
    J_pinv_T = (np.sqrt(S_pinv)[:, None] * U.T) * p_inv[None, :]
    e0 = -J_pinv_T @ b0
    return np.ascontiguousarray(J), np.ascontiguousarray(e0)


def make_window(seed=20260925, num_keyframes=10, num_imu_frames=3, num_landmarks=2000, mode="steady",
                extrinsics="fixed", sonar=False, depth=False, pose_noise=(0.02, 0.5), landmark_noise=0.05,
                pixel_noise=1.0, track_p=0.35, image=EUROC_IMAGE, keyframe_dt=0.25, frame_dt=0.05):
    """Build one window.  mode='steady': K keyframe poses + I recent frames with speed/bias, IMU terms among
    the recent frames and a dense marginalisation prior (the steady-state OKVIS graph).  mode='initial': every
    frame has speed/bias and IMU links, with the PoseError/SpeedAndBiasError priors of the very first frame
    (Estimator.cpp:319-361), as in okvis_ceres/test/TestEstimator.cpp.
    Returns (BaWindow, truth) where truth holds ground-truth poses/landmarks."""
    rng = np.random.default_rng(seed)
    K, I = num_keyframes, num_imu_frames
    P = K + I
    imu_params = dict(EUROC_IMU)
    # frame times in ms on the 1 kHz truth grid, offset 2 ms from the 200 Hz IMU grid
    t0_ms = 500 + 2
    t_ms = [t0_ms + int(round(i * keyframe_dt * 1000)) for i in range(K)]
    for j in range(I):
        t_ms.append(t_ms[K - 1] + int(round((j + 1) * frame_dt * 1000)))
    duration = t_ms[-1] / 1000.0 + 0.2
    traj = simulate_trajectory(rng, duration, imu_params)
    t_ns = np.array([ms * 1000000 for ms in t_ms], dtype=np.int64)
    poses_true = np.array([np.concatenate([traj["r"][ms], traj["q"][ms]]) for ms in t_ms])
    vel_true = np.array([traj["v"][ms] for ms in t_ms])

    W, Hh = image
    ncam = 2
    T_SC = [T.copy() for T in EUROC_T_SC]
    intr = EUROC_INTRINSICS.copy()

    # world-from-camera for each frame/cam
    def T_of(pose):
        T = np.eye(4)
        T[:3, :3] = quat_to_rot(pose[3:7])
        T[:3, 3] = pose[:3]
        return T

    T_WS = [T_of(p) for p in poses_true]
    T_CW = [[np.linalg.inv(T_WS[f] @ T_SC[c]) for c in range(ncam)] for f in range(P)]

    # ---- landmarks + tracks
    L = num_landmarks
    lm_true = np.zeros((0, 3))
    tracks = []  # (f0, k)
    while len(lm_true) < L:
        ncand = max(2 * (L - len(lm_true)), 64)
        f0 = rng.integers(0, P - 1, ncand)
        px = np.stack([rng.uniform(20, W - 20, ncand), rng.uniform(20, Hh - 20, ncand)], axis=-1)
        depth_c = rng.uniform(2.0, 15.0, ncand)
        dirs = backproject(intr[0], px)
        p_C = dirs * depth_c[:, None]
        p_W = np.zeros((ncand, 3))
        for f in range(P):
            sel = f0 == f
            if sel.any():
                Twc = T_WS[f] @ T_SC[0]
                p_W[sel] = p_C[sel] @ Twc[:3, :3].T + Twc[:3, 3]
        vis = np.zeros((ncand, P), dtype=bool)
        for f in range(P):
            ok = np.ones(ncand, dtype=bool)
            for c in range(ncam):
                pc = p_W @ T_CW[f][c][:3, :3].T + T_CW[f][c][:3, 3]
                z_ok = pc[:, 2] > 0.5
                ip = project(intr[c], np.where(z_ok[:, None], pc, np.array([0, 0, 1.0])))
                ok &= z_ok & (ip[:, 0] > 2) & (ip[:, 0] < W - 3) & (ip[:, 1] > 2) & (ip[:, 1] < Hh - 3)
            vis[:, f] = ok
        klen = np.clip(rng.geometric(track_p, ncand), 2, P)
        for i in range(ncand):
            k = 0
            while f0[i] + k < P and k < klen[i] and vis[i, f0[i] + k]:
                k += 1
            if k >= 2:
                tracks.append((int(f0[i]), k))
                lm_true = np.vstack([lm_true, p_W[i]])
                if len(lm_true) >= L:
                    break

    # ---- parameter blocks
    w = BaWindow()
    w.imu_params = imu_params
    ext_per_frame = extrinsics == "random_walk"
    n_ext = ncam * P if ext_per_frame else ncam
    ext_true = np.array([T_to_pose(T_SC[c]) for c in range(ncam)])
    pose_blocks_true = np.vstack([poses_true] + [ext_true] * (P if ext_per_frame else 1))

    def ext_block(f, c):
        return P + (f * ncam + c if ext_per_frame else c)

    sig_t, sig_r = pose_noise[0], np.deg2rad(pose_noise[1])
    pose_blocks = pose_blocks_true.copy()
    for f in range(P):
        d = np.concatenate([rng.normal(0, sig_t, 3), rng.normal(0, sig_r, 3)])
        pose_blocks[f] = pose_oplus(poses_true[f], d)
    pose_fixed = np.zeros(P + n_ext, dtype=np.uint8)
    if ext_per_frame:
        for b in range(P, P + n_ext):
            d = np.concatenate([rng.normal(0, 1e-3, 3), rng.normal(0, 1e-3, 3)])
            pose_blocks[b] = pose_oplus(pose_blocks_true[b], d)
    else:
        pose_fixed[P:] = 1  # sigma_absolute_* = 0 -> setParameterBlockConstant (Estimator.cpp:345-348)
    w.pose_blocks, w.pose_fixed = pose_blocks, pose_fixed

    sb_frames = list(range(P)) if mode == "initial" else list(range(K, P))
    sb_true = np.zeros((len(sb_frames), 9))
    bg_true, ba_true = rng.normal(0, 0.002, 3), rng.normal(0, 0.02, 3)
    for i, f in enumerate(sb_frames):
        sb_true[i, :3] = vel_true[f]
        sb_true[i, 3:6] = bg_true
        sb_true[i, 6:9] = ba_true
    sb = sb_true.copy()
    sb[:, :3] += rng.normal(0, 0.02, sb[:, :3].shape)
    sb[:, 3:6] += rng.normal(0, 5e-4, sb[:, 3:6].shape)
    sb[:, 6:9] += rng.normal(0, 5e-3, sb[:, 6:9].shape)
    w.speedbias = sb
    w.speedbias_fixed = np.zeros(len(sb), dtype=np.uint8)
    # the measurements must carry the true biases
    gyr = traj["gyr"] + bg_true
    acc = traj["acc"] + ba_true

    lms = np.concatenate([lm_true + rng.normal(0, landmark_noise, lm_true.shape), np.ones((L, 1))], axis=1)
    # the reference keeps homogeneous points; exercise a non-unit w on a few of them
    scale = np.where(rng.uniform(size=L) < 0.1, rng.uniform(0.5, 2.0, L), 1.0)
    lms = lms * scale[:, None]
    w.landmarks = lms
    w.intrinsics = intr

    # ---- observations, landmark-major
    obs_pose, obs_lm, obs_ext, obs_cam, obs_z = [], [], [], [], []
    for l, (f0_, k) in enumerate(tracks):
        for f in range(f0_, f0_ + k):
            for c in range(ncam):
                pc = T_CW[f][c][:3, :3] @ lm_true[l] + T_CW[f][c][:3, 3]
                ip = project(intr[c], pc)
                obs_pose.append(f)
                obs_lm.append(l)
                obs_ext.append(ext_block(f, c))
                obs_cam.append(c)
                obs_z.append(ip)
    obs_z = np.array(obs_z) + rng.normal(0, pixel_noise, (len(obs_z), 2))
    # a few gross outliers for the Cauchy loss to chew on
    n_out = max(1, len(obs_z) // 100)
    idx = rng.choice(len(obs_z), n_out, replace=False)
    obs_z[idx] += rng.normal(0, 25.0, (n_out, 2))
    w.obs_pose, w.obs_landmark, w.obs_extrinsics, w.obs_camera = obs_pose, obs_lm, obs_ext, obs_cam
    w.obs_measurement = obs_z
    size = 8.0
    w.obs_information = np.tile(np.array([64.0 / (size * size), 0, 0, 64.0 / (size * size)]), (len(obs_z), 1))

    # ---- IMU terms
    links = [(sb_frames[i], sb_frames[i + 1]) for i in range(len(sb_frames) - 1)]
    sb_index = {f: i for i, f in enumerate(sb_frames)}
    off = [0]
    mt, mg, ma = [], [], []
    for (fa, fb) in links:
        ta, tb = t_ns[fa], t_ns[fb]
        sel = (traj["t_imu"] >= ta - 20000000) & (traj["t_imu"] <= tb + 20000000)
        mt.append(traj["t_imu"][sel])
        mg.append(gyr[sel])
        ma.append(acc[sel])
        off.append(off[-1] + int(sel.sum()))
    w.imu_pose0 = [a for a, _ in links]
    w.imu_pose1 = [b for _, b in links]
    w.imu_speedbias0 = [sb_index[a] for a, _ in links]
    w.imu_speedbias1 = [sb_index[b] for _, b in links]
    w.imu_t0_ns = [t_ns[a] for a, _ in links]
    w.imu_t1_ns = [t_ns[b] for _, b in links]
    w.imu_meas_offset = off
    if links:
        w.imu_meas_t_ns = np.concatenate(mt)
        w.imu_meas_gyro = np.vstack(mg)
        w.imu_meas_accel = np.vstack(ma)

    # ---- priors
    if mode == "initial":
        info = np.zeros((6, 6))
        info[0, 0] = info[1, 1] = info[2, 2] = info[5, 5] = 1.0e8  # Estimator.cpp:321-326
        w.pose_prior_block = [0]
        w.pose_prior_measurement = [pose_blocks[0]]
        w.pose_prior_information = [info.ravel()]
        sinfo = np.diag([1.0] * 3 + [1.0 / 0.03 ** 2] * 3 + [1.0 / 0.1 ** 2] * 3)  # Estimator.cpp:351-355
        w.speedbias_prior_block = [0]
        w.speedbias_prior_measurement = [sb[0]]
        w.speedbias_prior_information = [sinfo.ravel()]
    else:
        kinds = [capi.SVIN_BLOCK_POSE] * (K + 1) + [capi.SVIN_BLOCK_SPEEDBIAS]
        idxs = list(range(K + 1)) + [0]
        dims = [6] * (K + 1) + [9]
        lin = []
        for f in range(K + 1):
            d = np.concatenate([rng.normal(0, 0.01, 3), rng.normal(0, np.deg2rad(0.2), 3)])
            lin.append(pose_oplus(poses_true[f], d))
        lin.append(sb_true[0] + np.concatenate([rng.normal(0, 0.01, 3), rng.normal(0, 2e-4, 3), rng.normal(0, 2e-3, 3)]))
        J, e0 = make_marg_prior(rng, dims)
        w.marg_block_kind, w.marg_block_index = kinds, idxs
        w.marg_linearization_points = np.concatenate(lin)
        w.marg_J, w.marg_e0 = J.ravel(), e0
        w.marg_dim = int(sum(dims))
    if ext_per_frame:
        b0, b1, infos = [], [], []
        for f in range(P - 1):
            dtf = (t_ns[f + 1] - t_ns[f]) * 1e-9
            for c in range(ncam):
                b0.append(ext_block(f, c))
                b1.append(ext_block(f + 1, c))
                tv, rv = (1e-3 ** 2) * dtf, (1e-3 ** 2) * dtf  # sigma_c_relative_* = 1e-3
                infos.append(np.diag([1 / tv] * 3 + [1 / rv] * 3).ravel())
        w.relative_pose_block0, w.relative_pose_block1, w.relative_pose_information = b0, b1, infos
        # absolute extrinsics priors on the first frame's blocks (Estimator.cpp:329-344)
        pb = list(w.pose_prior_block) if len(np.atleast_1d(w.pose_prior_block)) else []
        pm = [np.asarray(m) for m in np.asarray(w.pose_prior_measurement).reshape(-1, 7)]
        pi = [np.asarray(m) for m in np.asarray(w.pose_prior_information).reshape(-1, 36)]
        for c in range(ncam):
            pb.append(ext_block(0, c))
            pm.append(pose_blocks_true[ext_block(0, c)])
            pi.append(np.diag([1 / 1e-2 ** 2] * 3 + [1 / 1e-2 ** 2] * 3).ravel())
        w.pose_prior_block, w.pose_prior_measurement, w.pose_prior_information = pb, pm, pi
    if sonar:
        sp, sr, sh, si, sm = [], [], [], [], []
        for f in range(P):
            sel = rng.choice(L, 3, replace=False)
            mean = lm_true[sel].mean(axis=0)
            rng_true = np.linalg.norm(poses_true[f][:3] - mean)
            sp.append(f)
            sr.append(rng_true + rng.normal(0, 0.02))
            sh.append(rng.uniform(-np.pi, np.pi))
            si.append(1.0)
            sm.append(mean)
        w.sonar_pose, w.sonar_range, w.sonar_heading, w.sonar_information, w.sonar_landmark_mean = sp, sr, sh, si, sm
        w.sonar_T_SSo = STEREORIG_T_SSO
    if depth:
        first = 0.3
        w.depth_pose = list(range(P))
        w.depth_measurement = [first - poses_true[f][2] + rng.normal(0, 0.01) for f in range(P)]
        w.depth_first = [first] * P
        w.depth_information = [5.0] * P  # Estimator.cpp:256
    w.finalize()
    truth = dict(pose_blocks=pose_blocks_true, landmarks=lm_true, speedbias=sb_true, t_ns=t_ns, tracks=tracks)
    return w, truth
