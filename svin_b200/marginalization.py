"""Host-side containers for the marginalisation core (SvinMargSpec / SvinMargResult in include/svin_b200.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .window import BaWindow


class MargSpec:
    def __init__(self, window: BaWindow, marginalize_pose, marginalize_speedbias, prior_kind=(), prior_index=(),
                 prior_H=None, prior_b0=None):
        c = np.ascontiguousarray
        self.mp = c(marginalize_pose, dtype=np.uint8)
        self.ms = c(marginalize_speedbias, dtype=np.uint8)
        assert len(self.mp) == len(window.pose_blocks) and len(self.ms) == len(window.speedbias)
        self.pk, self.pi = c(prior_kind, dtype=np.int32), c(prior_index, dtype=np.int32)
        self.pH = c(prior_H, dtype=np.float64) if prior_H is not None else np.zeros((0, 0))
        self.pb = c(prior_b0, dtype=np.float64) if prior_b0 is not None else np.zeros(0)
        s = capi.SvinMargSpec()
        s.prior_num_blocks = len(self.pk)
        s.prior_dim = len(self.pb)
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t)) if a.size else C.POINTER(t)()
        s.prior_block_kind, s.prior_block_index = p(self.pk, C.c_int32), p(self.pi, C.c_int32)
        s.prior_H, s.prior_b0 = p(self.pH, C.c_double), p(self.pb, C.c_double)
        s.marginalize_pose, s.marginalize_speedbias = p(self.mp, C.c_uint8), p(self.ms, C.c_uint8)
        self.c = s


class MargResult:
    def __init__(self, window: BaWindow):
        n = window.dense_dim()
        nb = len(window.pose_blocks) + len(window.speedbias)
        self.kind, self.index = np.zeros(nb, np.int32), np.zeros(nb, np.int32)
        self._H, self._b0 = np.zeros(n * n), np.zeros(n)
        self._J, self._e0 = np.zeros(n * n), np.zeros(n)
        r = capi.SvinMargResult()
        r.block_kind = self.kind.ctypes.data_as(capi.c_int32_p)
        r.block_index = self.index.ctypes.data_as(capi.c_int32_p)
        r.H, r.b0 = self._H.ctypes.data_as(capi.c_double_p), self._b0.ctypes.data_as(capi.c_double_p)
        r.J, r.e0 = self._J.ctypes.data_as(capi.c_double_p), self._e0.ctypes.data_as(capi.c_double_p)
        self.c = r

    def unpack(self) -> dict:
        d, nb = self.c.dim, self.c.num_blocks
        return dict(dim=d, kind=self.kind[:nb].copy(), index=self.index[:nb].copy(),
                    H=self._H[:d * d].reshape(d, d).copy(), b0=self._b0[:d].copy(),
                    J=self._J[:d * d].reshape(d, d).copy(), e0=self._e0[:d].copy())


def marginalization_subwindow(w: BaWindow, frames_removed: int = 1):
    """From an 'initial'-mode synthetic window build the window the marginalisation of its oldest
    `frames_removed` frames linearises: the prior terms and IMU links of those frames and the reprojection
    terms of every landmark whose observations all lie in the first frames_removed + 1 frames
    (Estimator.cpp:700-741: landmarks without newer observations are marginalised with all their residuals)."""
    s = w.copy()
    last = frames_removed  # highest frame index still touched
    obs_frame = w.obs_pose
    lm_max = np.full(w.num_landmarks, -1)
    np.maximum.at(lm_max, w.obs_landmark, obs_frame)
    lm_cnt = np.bincount(w.obs_landmark, minlength=w.num_landmarks)
    keep_lm = np.nonzero((lm_max <= last) & (lm_cnt >= 2))[0]
    remap = -np.ones(w.num_landmarks, dtype=np.int64)
    remap[keep_lm] = np.arange(len(keep_lm))
    keep_obs = remap[w.obs_landmark] >= 0
    s.landmarks = w.landmarks[keep_lm].copy()
    s.landmark_fixed = np.zeros(len(keep_lm), dtype=np.uint8)
    for name in ("obs_pose", "obs_extrinsics", "obs_camera", "obs_measurement", "obs_information"):
        setattr(s, name, getattr(w, name)[keep_obs].copy())
    s.obs_landmark = remap[w.obs_landmark[keep_obs]].astype(np.int32)
    # IMU links that start in a removed frame
    ki = [i for i in range(len(w.imu_pose0)) if w.imu_pose0[i] < frames_removed]
    off, mt, mg, ma = [0], [], [], []
    for i in ki:
        a, b = w.imu_meas_offset[i], w.imu_meas_offset[i + 1]
        mt.append(w.imu_meas_t_ns[a:b]); mg.append(w.imu_meas_gyro[a:b]); ma.append(w.imu_meas_accel[a:b])
        off.append(off[-1] + (b - a))
    for name in ("imu_pose0", "imu_speedbias0", "imu_pose1", "imu_speedbias1", "imu_t0_ns", "imu_t1_ns"):
        setattr(s, name, getattr(w, name)[ki].copy())
    s.imu_meas_offset = np.array(off, dtype=np.int32)
    s.imu_meas_t_ns = np.concatenate(mt) if mt else np.zeros(0, np.int64)
    s.imu_meas_gyro = np.vstack(mg) if mg else np.zeros((0, 3))
    s.imu_meas_accel = np.vstack(ma) if ma else np.zeros((0, 3))
    s.finalize()
    mp = np.zeros(len(w.pose_blocks), dtype=np.uint8)
    mp[:frames_removed] = 1
    ms = np.zeros(len(w.speedbias), dtype=np.uint8)
    ms[:frames_removed] = 1
    return s, mp, ms
