"""Host-side container for one sliding window (the flattened okvis::ceres::Map).

Mirrors ``SvinBaWindow`` in include/svin_b200.h field for field; numpy arrays own
the memory and :meth:`BaWindow.c_struct` exposes them to the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

_F64 = np.float64
_I32 = np.int32
_I64 = np.int64
_U8 = np.uint8


def _arr(a, dtype, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=dtype))
    if shape is not None:
        a = a.reshape(shape)
    return a


class BaWindow:
    """All parameter blocks and error terms of one window, as contiguous numpy arrays."""

    # (name, dtype, trailing shape) ; leading dim is the count
    _ARRAYS = [
        ("pose_blocks", _F64, (7,)), ("speedbias", _F64, (9,)), ("landmarks", _F64, (4,)),
        ("pose_fixed", _U8, ()), ("speedbias_fixed", _U8, ()), ("landmark_fixed", _U8, ()),
        ("intrinsics", _F64, (8,)),
        ("obs_pose", _I32, ()), ("obs_landmark", _I32, ()), ("obs_extrinsics", _I32, ()), ("obs_camera", _I32, ()),
        ("obs_measurement", _F64, (2,)), ("obs_information", _F64, (4,)),
        ("imu_pose0", _I32, ()), ("imu_speedbias0", _I32, ()), ("imu_pose1", _I32, ()), ("imu_speedbias1", _I32, ()),
        ("imu_t0_ns", _I64, ()), ("imu_t1_ns", _I64, ()), ("imu_meas_offset", _I32, ()),
        ("imu_meas_t_ns", _I64, ()), ("imu_meas_gyro", _F64, (3,)), ("imu_meas_accel", _F64, (3,)),
        ("pose_prior_block", _I32, ()), ("pose_prior_measurement", _F64, (7,)), ("pose_prior_information", _F64, (36,)),
        ("speedbias_prior_block", _I32, ()), ("speedbias_prior_measurement", _F64, (9,)),
        ("speedbias_prior_information", _F64, (81,)),
        ("relative_pose_block0", _I32, ()), ("relative_pose_block1", _I32, ()),
        ("relative_pose_information", _F64, (36,)),
        ("sonar_pose", _I32, ()), ("sonar_range", _F64, ()), ("sonar_heading", _F64, ()),
        ("sonar_information", _F64, ()), ("sonar_landmark_mean", _F64, (3,)), ("sonar_T_SSo", _F64, ()),
        ("depth_pose", _I32, ()), ("depth_measurement", _F64, ()), ("depth_first", _F64, ()),
        ("depth_information", _F64, ()),
        ("marg_block_kind", _I32, ()), ("marg_block_index", _I32, ()), ("marg_linearization_points", _F64, ()),
        ("marg_J", _F64, ()), ("marg_e0", _F64, ()),
    ]

    _ARRAY_NAMES = frozenset(n for n, _, _ in _ARRAYS)

    def __setattr__(self, name, value):
        # re-assigning an array (w.landmarks = ..., w.obs_* = ...) invalidates the cached C view: its raw pointers would
        # otherwise keep pointing at the old (possibly freed) buffers
        if name in self._ARRAY_NAMES or name in ("loss_type", "loss_scale", "marg_dim", "imu_params"):
            object.__setattr__(self, "_struct", None)
        object.__setattr__(self, name, value)

    def __init__(self):
        for name, dt, tail in self._ARRAYS:
            setattr(self, name, np.zeros((0,) + tail, dtype=dt))
        self.sonar_T_SSo = np.array([0, 0, 0, 0, 0, 0, 1], dtype=_F64)
        self.imu_meas_offset = np.zeros(1, dtype=_I32)
        self.loss_type = capi.SVIN_LOSS_CAUCHY
        self.loss_scale = 1.0
        self.imu_params = dict(sigma_g_c=12.0e-4, sigma_a_c=8.0e-3, sigma_gw_c=4.0e-6, sigma_aw_c=4.0e-5,
                               g=9.81007, g_max=7.8, a_max=176.0)  # config_fpga_p2_euroc.yaml:35-45
        self.marg_dim = 0
        self._struct = None

    # ---- normalisation --------------------------------------------------------------
    def finalize(self) -> "BaWindow":
        for name, dt, tail in self._ARRAYS:
            a = getattr(self, name)
            a = np.ascontiguousarray(np.asarray(a, dtype=dt))
            if tail and a.size:
                a = a.reshape((-1,) + tail)
            elif tail:
                a = a.reshape((0,) + tail)
            setattr(self, name, a)
        if self.pose_fixed.size == 0:
            self.pose_fixed = np.zeros(len(self.pose_blocks), dtype=_U8)
        if self.speedbias_fixed.size == 0:
            self.speedbias_fixed = np.zeros(len(self.speedbias), dtype=_U8)
        if self.landmark_fixed.size == 0:
            self.landmark_fixed = np.zeros(len(self.landmarks), dtype=_U8)
        self._struct = None
        return self

    def copy(self) -> "BaWindow":
        w = BaWindow()
        for name, _, _ in self._ARRAYS:
            setattr(w, name, getattr(self, name).copy())
        w.loss_type, w.loss_scale = self.loss_type, self.loss_scale
        w.imu_params = dict(self.imu_params)
        w.marg_dim = self.marg_dim
        return w.finalize()

    # ---- sizes ----------------------------------------------------------------------
    @property
    def num_obs(self):
        return len(self.obs_pose)

    @property
    def num_landmarks(self):
        return len(self.landmarks)

    def num_pose_blocks_free(self) -> int:
        return int((np.asarray(self.pose_fixed) == 0).sum())

    def dense_dim(self) -> int:
        return int(6 * np.count_nonzero(self.pose_fixed == 0) + 9 * np.count_nonzero(self.speedbias_fixed == 0))

    def h2d_bytes(self) -> int:
        return int(sum(getattr(self, n).nbytes for n, _, _ in self._ARRAYS))

    def d2h_bytes(self) -> int:
        return int(self.pose_blocks.nbytes + self.speedbias.nbytes + self.landmarks.nbytes + 8 * len(self.landmarks))

    # ---- C view -----------------------------------------------------------------------
    def c_struct(self) -> capi.SvinBaWindow:
        if self._struct is not None:
            return self._struct
        self.finalize()
        s = capi.SvinBaWindow()

        def ptr(a, ctype):
            return a.ctypes.data_as(C.POINTER(ctype)) if a.size else C.POINTER(ctype)()

        d, i32, i64, u8 = C.c_double, C.c_int32, C.c_int64, C.c_uint8
        s.num_pose_blocks = len(self.pose_blocks)
        s.num_speedbias = len(self.speedbias)
        s.num_landmarks = len(self.landmarks)
        s.num_cameras = len(self.intrinsics)
        s.pose_blocks, s.speedbias, s.landmarks = ptr(self.pose_blocks, d), ptr(self.speedbias, d), ptr(self.landmarks, d)
        s.pose_fixed, s.speedbias_fixed = ptr(self.pose_fixed, u8), ptr(self.speedbias_fixed, u8)
        s.landmark_fixed = ptr(self.landmark_fixed, u8)
        s.intrinsics = ptr(self.intrinsics, d)
        s.num_obs = len(self.obs_pose)
        s.loss_type, s.loss_scale = int(self.loss_type), float(self.loss_scale)
        s.obs_pose, s.obs_landmark = ptr(self.obs_pose, i32), ptr(self.obs_landmark, i32)
        s.obs_extrinsics, s.obs_camera = ptr(self.obs_extrinsics, i32), ptr(self.obs_camera, i32)
        s.obs_measurement, s.obs_information = ptr(self.obs_measurement, d), ptr(self.obs_information, d)
        s.num_imu = len(self.imu_pose0)
        for k, v in self.imu_params.items():
            setattr(s.imu_params, k, float(v))
        s.imu_pose0, s.imu_speedbias0 = ptr(self.imu_pose0, i32), ptr(self.imu_speedbias0, i32)
        s.imu_pose1, s.imu_speedbias1 = ptr(self.imu_pose1, i32), ptr(self.imu_speedbias1, i32)
        s.imu_t0_ns, s.imu_t1_ns = ptr(self.imu_t0_ns, i64), ptr(self.imu_t1_ns, i64)
        s.imu_meas_offset, s.imu_meas_t_ns = ptr(self.imu_meas_offset, i32), ptr(self.imu_meas_t_ns, i64)
        s.imu_meas_gyro, s.imu_meas_accel = ptr(self.imu_meas_gyro, d), ptr(self.imu_meas_accel, d)
        s.num_pose_priors = len(self.pose_prior_block)
        s.pose_prior_block = ptr(self.pose_prior_block, i32)
        s.pose_prior_measurement = ptr(self.pose_prior_measurement, d)
        s.pose_prior_information = ptr(self.pose_prior_information, d)
        s.num_speedbias_priors = len(self.speedbias_prior_block)
        s.speedbias_prior_block = ptr(self.speedbias_prior_block, i32)
        s.speedbias_prior_measurement = ptr(self.speedbias_prior_measurement, d)
        s.speedbias_prior_information = ptr(self.speedbias_prior_information, d)
        s.num_relative_pose = len(self.relative_pose_block0)
        s.relative_pose_block0 = ptr(self.relative_pose_block0, i32)
        s.relative_pose_block1 = ptr(self.relative_pose_block1, i32)
        s.relative_pose_information = ptr(self.relative_pose_information, d)
        s.num_sonar = len(self.sonar_pose)
        s.sonar_pose, s.sonar_range = ptr(self.sonar_pose, i32), ptr(self.sonar_range, d)
        s.sonar_heading, s.sonar_information = ptr(self.sonar_heading, d), ptr(self.sonar_information, d)
        s.sonar_landmark_mean, s.sonar_T_SSo = ptr(self.sonar_landmark_mean, d), ptr(self.sonar_T_SSo, d)
        s.num_depth = len(self.depth_pose)
        s.depth_pose, s.depth_measurement = ptr(self.depth_pose, i32), ptr(self.depth_measurement, d)
        s.depth_first, s.depth_information = ptr(self.depth_first, d), ptr(self.depth_information, d)
        s.marg_num_blocks = len(self.marg_block_kind)
        s.marg_dim = int(self.marg_dim)
        s.marg_block_kind, s.marg_block_index = ptr(self.marg_block_kind, i32), ptr(self.marg_block_index, i32)
        s.marg_linearization_points = ptr(self.marg_linearization_points, d)
        s.marg_J, s.marg_e0 = ptr(self.marg_J, d), ptr(self.marg_e0, d)
        self._struct = s
        return s


def default_options(**overrides) -> capi.SvinBaOptions:
    """Ceres 2.2 defaults as left by Estimator::optimize (Estimator.cpp:876-899)."""
    o = capi.SvinBaOptions()
    o.max_num_iterations = 10
    o.min_num_iterations = 3
    o.time_limit_seconds = -1.0
    o.initial_trust_region_radius = 1e4
    o.max_trust_region_radius = 1e16
    o.min_trust_region_radius = 1e-32
    o.min_relative_decrease = 1e-3
    o.min_lm_diagonal = 1e-6
    o.max_lm_diagonal = 1e32
    o.function_tolerance = 1e-6
    o.gradient_tolerance = 1e-10
    o.parameter_tolerance = 1e-8
    o.max_num_consecutive_invalid_steps = 5
    o.jacobi_scaling = 1
    o.compute_landmark_quality = 1
    for k, v in overrides.items():
        setattr(o, k, v)
    return o
