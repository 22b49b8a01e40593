"""The scene of okvis_ceres/test/TestMarginalization.cpp:60-231 as a BaWindow: three sensor poses (the first two
constant), an estimated camera extrinsics block with a tight PoseError prior (variance 1e-4 / 1e-4), N = 100 points
seen from all three poses with 1 px uniform noise, identity information, NO loss function (the test passes NULL);
the third pose starts from T_WS2 * T_disturb.  The reference test uses the equidistant test camera; the restatement
has the radial-tangential model of the shipped configs, so the 752x480 test intrinsics are used with it.
The test asserts, after Map::solve(): 2 |vec(q_true * q_est^-1)| < 1e-2 and |r_true - r_est| < 1e-1."""
import numpy as np

from svin_b200 import capi
from svin_b200.synthetic import EUROC_INTRINSICS, T_to_pose, project, quat_to_rot
from svin_b200.window import BaWindow


def _rand_T(rng, trans, rot):
    """Transformation::setRandom(translationMaxMeters, rotationMaxRadians) (Transformation.hpp impl:105-121)."""
    r = rng.uniform(-1, 1, 3) * trans
    axis = rng.uniform(-1, 1, 3)
    axis /= np.linalg.norm(axis)
    ang = rng.uniform(-1, 1) * rot
    q = np.concatenate([np.sin(ang / 2) * axis, [np.cos(ang / 2)]])
    T = np.eye(4)
    T[:3, :3] = quat_to_rot(q)
    T[:3, 3] = r
    return T


def make_scene(seed=0, n_points=100):
    rng = np.random.default_rng(seed)
    intr = EUROC_INTRINSICS[0].copy()
    T_WS0 = _rand_T(rng, 10.0, np.pi)
    T_WS1 = T_WS0 @ _rand_T(rng, 1.0, 0.01)
    T_WS2 = T_WS1 @ _rand_T(rng, 1.0, 0.01)
    T_WS2_init = T_WS2 @ _rand_T(rng, 1.0, 0.01)
    T_SC = _rand_T(rng, 0.2, np.pi)
    w = BaWindow()
    w.pose_blocks = np.stack([T_to_pose(T) for T in (T_WS0, T_WS1, T_WS2_init, T_SC)])
    w.pose_fixed = np.array([1, 1, 0, 0], dtype=np.uint8)
    w.intrinsics = intr[None, :]
    lms, obs = [], []
    while len(lms) < n_points:
        # createRandomVisibleHomogeneousPoint(10.0): a random pixel back-projected to a random depth <= 10 m
        ip = np.array([rng.uniform(0, 751), rng.uniform(0, 479)])
        und = np.array([(ip[0] - intr[2]) / intr[0], (ip[1] - intr[3]) / intr[1], 1.0])
        pC0 = und / np.linalg.norm(und) * rng.uniform(0.5, 10.0)
        pW = (T_WS0 @ T_SC @ np.append(pC0, 1.0))
        k = len(lms)
        seen = 0
        for j, T_WS in enumerate((T_WS0, T_WS1, T_WS2)):
            pC = np.linalg.inv(T_WS @ T_SC) @ pW
            if pC[2] < 0.3:
                continue
            z = project(intr, pC[:3])
            if not (0 <= z[0] < 752 and 0 <= z[1] < 480):
                continue
            obs.append((j, k, z + rng.uniform(-1, 1, 2)))
            seen += 1
        if seen == 0:
            obs = [o for o in obs if o[1] != k]
            continue
        lms.append(pW)
    w.landmarks = np.stack(lms)
    w.obs_pose = np.array([o[0] for o in obs], dtype=np.int32)
    w.obs_landmark = np.array([o[1] for o in obs], dtype=np.int32)
    w.obs_extrinsics = np.full(len(obs), 3, dtype=np.int32)
    w.obs_camera = np.zeros(len(obs), dtype=np.int32)
    w.obs_measurement = np.stack([o[2] for o in obs])
    w.obs_information = np.tile(np.eye(2).reshape(1, 4), (len(obs), 1))
    w.pose_prior_block = np.array([3], dtype=np.int32)
    w.pose_prior_measurement = T_to_pose(T_SC)[None, :]
    w.pose_prior_information = (np.eye(6) * 1.0e4).reshape(1, 36)     # PoseError(T_SC, 1e-4, 1e-4)
    w.loss_type = capi.SVIN_LOSS_NONE
    w.finalize()
    return w, T_to_pose(T_WS2)


def pose_errors(est, true):
    """(rotation error 2 |vec(q_true * q_est^-1)|, translation error) as asserted at TestMarginalization.cpp:226-231."""
    from svin_b200.synthetic import quat_mul
    qe = est[3:7] * np.array([-1, -1, -1, 1])
    dq = quat_mul(true[3:7], qe / np.dot(qe, qe))
    return 2 * np.linalg.norm(dq[:3]), np.linalg.norm(true[:3] - est[:3])
