"""Synthetic correspondence sets for the RANSAC row (stereo rig with the EuRoC extrinsics, points at 2..15 m)."""
import numpy as np

from svin_b200.synthetic import EUROC_INTRINSICS, EUROC_T_SC, quat_to_rot, delta_q


def _rand_rot(rng, max_angle):
    return quat_to_rot(delta_q(rng.normal(0, max_angle / 2, 3)))


def absolute_scene(seed, n_per_cam=150, outlier_ratio=0.25, pixel_noise=0.5, n_samples=64):
    rng = np.random.default_rng(seed)
    R_ws, t_ws = _rand_rot(rng, 0.6), rng.normal(0, 2.0, 3)
    cam_R = np.stack([T[:3, :3] for T in EUROC_T_SC])
    cam_t = np.stack([T[:3, 3] for T in EUROC_T_SC])
    pts, brs, cams, sig = [], [], [], []
    for c in range(2):
        fu = EUROC_INTRINSICS[c][0]
        d = np.stack([rng.uniform(-0.6, 0.6, n_per_cam), rng.uniform(-0.4, 0.4, n_per_cam), np.ones(n_per_cam)], axis=1)
        p_c = d * rng.uniform(2.0, 15.0, (n_per_cam, 1))
        p_b = p_c @ cam_R[c].T + cam_t[c]
        p_w = p_b @ R_ws.T + t_ws
        b = d + np.concatenate([rng.normal(0, pixel_noise / fu, (n_per_cam, 2)), np.zeros((n_per_cam, 1))], axis=1)
        pts.append(p_w)
        brs.append(b / np.linalg.norm(b, axis=1, keepdims=True))
        cams.append(np.full(n_per_cam, c, np.int32))
        sd = 0.8 * 12.0 / 12.0                                  # keypoint size 12 (the detector's)
        sig.append(np.full(n_per_cam, np.sqrt(2.0) * sd * sd / (fu * fu)))
    points, bearings = np.vstack(pts), np.vstack(brs)
    cam_index, sigma = np.concatenate(cams), np.concatenate(sig)
    n = len(points)
    outlier = rng.uniform(size=n) < outlier_ratio
    points[outlier] += rng.normal(0, 1.5, (int(outlier.sum()), 3))
    # samples: 3 + 1 indices, the first three inside one camera (the adapter draws them so)
    samples = np.zeros((n_samples, 4), np.int32)
    for j in range(n_samples):
        c = rng.integers(0, 2)
        samples[j, :3] = rng.choice(np.nonzero(cam_index == c)[0], 3, replace=False)
        rest = np.setdiff1d(np.arange(n), samples[j, :3])
        samples[j, 3] = rng.choice(rest)
    return dict(points=points, bearings=bearings, cam_index=cam_index, cam_R=cam_R, cam_t=cam_t, sigma=sigma,
                samples=samples, truth=(R_ws, t_ws), outlier=outlier)


def relative_scene(seed, n=200, outlier_ratio=0.25, pixel_noise=0.5, n_samples=64, rotation_only=False):
    rng = np.random.default_rng(seed)
    fu = EUROC_INTRINSICS[0][0]
    R12 = _rand_rot(rng, 0.3)
    t12 = np.zeros(3) if rotation_only else rng.normal(0, 0.4, 3)
    d = np.stack([rng.uniform(-0.5, 0.5, n), rng.uniform(-0.35, 0.35, n), np.ones(n)], axis=1)
    p2 = d * rng.uniform(2.0, 15.0, (n, 1))                      # in frame 2
    p1 = p2 @ R12.T + t12
    ok = p1[:, 2] > 0.5
    p1, p2 = p1[ok], p2[ok]
    n = len(p1)

    def bearing(p):
        u = p[:, :2] / p[:, 2:3] + rng.normal(0, pixel_noise / fu, (len(p), 2))
        b = np.concatenate([u, np.ones((len(p), 1))], axis=1)
        return b / np.linalg.norm(b, axis=1, keepdims=True)
    f1, f2 = bearing(p1), bearing(p2)
    outlier = rng.uniform(size=n) < outlier_ratio
    f2[outlier] = bearing(np.stack([rng.uniform(-0.5, 0.5, outlier.sum()), rng.uniform(-0.35, 0.35, outlier.sum()),
                                    np.ones(outlier.sum())], axis=1))
    sd = 0.8 * 12.0 / 12.0
    sigma = np.full(n, np.sqrt(2.0) * sd * sd / (fu * fu))
    samples_rot = np.stack([rng.choice(n, 2, replace=False) for _ in range(n_samples)]).astype(np.int32)
    samples_rel = np.stack([rng.choice(n, 8, replace=False) for _ in range(n_samples)]).astype(np.int32)
    tn = np.linalg.norm(t12)
    return dict(f1=f1, f2=f2, sigma1=sigma, sigma2=sigma.copy(), samples_rot=samples_rot, samples_rel=samples_rel,
                truth=(R12, t12 / tn if tn > 0 else t12), outlier=outlier)
