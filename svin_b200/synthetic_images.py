"""Seeded synthetic stereo imagery with ground-truth geometry (SURVEY.md §8d front-end inputs).

A textured wall (multi-scale random blobs and corners) at x_W = 4 m is rendered through the EuRoC
PinholeCamera<RadialTangentialDistortion> models of config/config_fpga_p2_euroc.yaml for both cameras of
a slowly moving rig, plus sigma = 2 grey-level sensor noise.  Because the scene is a plane, the 3-D
point behind any pixel is known exactly, which gives landmarks for the 3D-2D matcher tests.
Host-side input generation only.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

from .synthetic import EUROC_IMAGE, EUROC_INTRINSICS, EUROC_T_SC, T_to_pose, backproject

WALL_X = 4.0
TEX_SCALE = 250.0  # texture pixels per metre
TEX_SIZE = 3000

# sensor z (optical axis, roughly) -> world +x ; sensor x -> world -y ; sensor y -> world -z  (gravity = -z_W)
R_WS0 = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])


def make_texture(seed: int = 0, size: int = TEX_SIZE) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.zeros((size, size), dtype=np.float32)
    for sigma, amp in ((24.0, 60.0), (8.0, 50.0), (3.0, 45.0)):
        n = rng.standard_normal((size, size)).astype(np.float32)
        n = ndimage.gaussian_filter(n, sigma)
        t += amp * n / (n.std() + 1e-9)
    # rectangles give strong corners
    for _ in range(1500):
        x, y = rng.integers(0, size - 40, 2)
        w, h = rng.integers(6, 40, 2)
        t[y:y + h, x:x + w] += rng.uniform(-90, 90)
    t = ndimage.gaussian_filter(t, 1.0)
    t = 128.0 + t
    return np.clip(t, 0, 255).astype(np.float32)


_RAYS = {}


def pixel_rays(cam: int, W: int, H: int) -> np.ndarray:
    key = (cam, W, H)
    if key not in _RAYS:
        xs, ys = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
        ip = np.stack([xs.ravel(), ys.ravel()], axis=-1)
        _RAYS[key] = backproject(EUROC_INTRINSICS[cam], ip)
    return _RAYS[key]


def render(texture: np.ndarray, T_WC: np.ndarray, cam: int, rng, noise_sigma: float = 2.0, image=EUROC_IMAGE):
    W, H = image
    rays = pixel_rays(cam, W, H)
    d = rays @ T_WC[:3, :3].T
    o = T_WC[:3, 3]
    t = (WALL_X - o[0]) / d[:, 0]
    p = o[None, :] + t[:, None] * d
    u = p[:, 1] * TEX_SCALE + texture.shape[1] / 2
    v = p[:, 2] * TEX_SCALE + texture.shape[0] / 2
    img = ndimage.map_coordinates(texture, [v, u], order=1, mode="reflect").reshape(H, W)
    img = img + noise_sigma * rng.standard_normal(img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def wall_point(T_WC: np.ndarray, cam: int, xy: np.ndarray) -> np.ndarray:
    """3-D world points on the wall behind pixels xy [n,2] of camera `cam` at pose T_WC."""
    rays = backproject(EUROC_INTRINSICS[cam], np.asarray(xy, dtype=np.float64))
    d = rays @ T_WC[:3, :3].T
    o = T_WC[:3, 3]
    t = (WALL_X - o[0]) / d[:, 0]
    return o[None, :] + t[:, None] * d


def rig_pose(rng, k: int) -> np.ndarray:
    """T_WS of frame k: small motion in front of the wall."""
    from .synthetic import delta_q, quat_to_rot
    T = np.eye(4)
    ang = rng.normal(0, 0.03, 3) + np.array([0.0, 0.01 * k, 0.0])
    R = quat_to_rot(delta_q(ang))
    T[:3, :3] = R @ R_WS0
    T[:3, 3] = np.array([0.02 * k, -0.05 * k, 0.01 * k]) + rng.normal(0, 0.01, 3)
    return T


def make_stereo_sequence(seed: int = 20260925, n_frames: int = 2, image=EUROC_IMAGE, texture=None):
    """Returns dict(images [n_frames][2] uint8, T_WS [n_frames], T_WC [n_frames][2], intrinsics [2][8],
    extraction_dir [n_frames][2][3])."""
    rng = np.random.default_rng(seed)
    tex = texture if texture is not None else make_texture(seed)
    images, T_WS, T_WC, ed = [], [], [], []
    for k in range(n_frames):
        Tws = rig_pose(rng, k)
        T_WS.append(Tws)
        row, rowT, rowE = [], [], []
        for c in range(2):
            Twc = Tws @ EUROC_T_SC[c]
            row.append(render(tex, Twc, c, rng, image=image))
            rowT.append(Twc)
            rowE.append(Twc[:3, :3].T @ np.array([0.0, 0.0, -1.0]))  # Frontend.cpp:107-108
        images.append(row)
        T_WC.append(rowT)
        ed.append(rowE)
    return dict(images=images, T_WS=T_WS, T_WC=T_WC, intrinsics=EUROC_INTRINSICS.copy(), extraction_dir=ed,
                texture=tex)


def random_image(seed: int, image=EUROC_IMAGE) -> np.ndarray:
    """Uniform random 8-bit image, the worst-case detector load of okvis_cv/test/TestFrame.cpp:65-68."""
    W, H = image
    return np.random.default_rng(seed).integers(0, 256, (H, W), dtype=np.uint8)
