"""ctypes view of ``include/svin_b200.h`` — the only way Python reaches the engine.

The product library is ``svin_b200/lib/libsvin_b200.so`` (CUDA, sm_100a).  There is
no CPU fallback: if the library is missing :func:`load` raises, and every engine
call fails loudly when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsvin_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint8_p = C.POINTER(C.c_uint8)

SVIN_LOSS_NONE, SVIN_LOSS_CAUCHY, SVIN_LOSS_HUBER = 0, 1, 2
SVIN_BLOCK_POSE, SVIN_BLOCK_SPEEDBIAS, SVIN_BLOCK_LANDMARK = 0, 1, 2
SVIN_TERM_NO_CONVERGENCE, SVIN_TERM_CONVERGENCE, SVIN_TERM_FAILURE, SVIN_TERM_USER_SUCCESS = 0, 1, 2, 3


class SvinImuParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("sigma_g_c", "sigma_a_c", "sigma_gw_c", "sigma_aw_c", "g", "g_max", "a_max")]


class SvinBaWindow(C.Structure):
    _fields_ = [
        ("num_pose_blocks", C.c_int32), ("num_speedbias", C.c_int32), ("num_landmarks", C.c_int32),
        ("num_cameras", C.c_int32),
        ("pose_blocks", c_double_p), ("speedbias", c_double_p), ("landmarks", c_double_p),
        ("pose_fixed", c_uint8_p), ("speedbias_fixed", c_uint8_p), ("landmark_fixed", c_uint8_p),
        ("intrinsics", c_double_p),
        ("num_obs", C.c_int32), ("loss_type", C.c_int32), ("loss_scale", C.c_double),
        ("obs_pose", c_int32_p), ("obs_landmark", c_int32_p), ("obs_extrinsics", c_int32_p),
        ("obs_camera", c_int32_p), ("obs_measurement", c_double_p), ("obs_information", c_double_p),
        ("num_imu", C.c_int32), ("imu_params", SvinImuParams),
        ("imu_pose0", c_int32_p), ("imu_speedbias0", c_int32_p), ("imu_pose1", c_int32_p),
        ("imu_speedbias1", c_int32_p), ("imu_t0_ns", c_int64_p), ("imu_t1_ns", c_int64_p),
        ("imu_meas_offset", c_int32_p), ("imu_meas_t_ns", c_int64_p), ("imu_meas_gyro", c_double_p),
        ("imu_meas_accel", c_double_p),
        ("num_pose_priors", C.c_int32), ("pose_prior_block", c_int32_p), ("pose_prior_measurement", c_double_p),
        ("pose_prior_information", c_double_p),
        ("num_speedbias_priors", C.c_int32), ("speedbias_prior_block", c_int32_p),
        ("speedbias_prior_measurement", c_double_p), ("speedbias_prior_information", c_double_p),
        ("num_relative_pose", C.c_int32), ("relative_pose_block0", c_int32_p), ("relative_pose_block1", c_int32_p),
        ("relative_pose_information", c_double_p),
        ("num_sonar", C.c_int32), ("sonar_pose", c_int32_p), ("sonar_range", c_double_p),
        ("sonar_heading", c_double_p), ("sonar_information", c_double_p), ("sonar_landmark_mean", c_double_p),
        ("sonar_T_SSo", c_double_p),
        ("num_depth", C.c_int32), ("depth_pose", c_int32_p), ("depth_measurement", c_double_p),
        ("depth_first", c_double_p), ("depth_information", c_double_p),
        ("marg_num_blocks", C.c_int32), ("marg_dim", C.c_int32), ("marg_block_kind", c_int32_p),
        ("marg_block_index", c_int32_p), ("marg_linearization_points", c_double_p), ("marg_J", c_double_p),
        ("marg_e0", c_double_p),
    ]


class SvinBaOptions(C.Structure):
    _fields_ = [
        ("max_num_iterations", C.c_int32), ("min_num_iterations", C.c_int32), ("time_limit_seconds", C.c_double),
        ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double), ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
        ("max_num_consecutive_invalid_steps", C.c_int32), ("jacobi_scaling", C.c_int32),
        ("compute_landmark_quality", C.c_int32),
    ]


class SvinBaSummary(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32), ("num_successful_steps", C.c_int32), ("termination", C.c_int32),
        ("imu_repropagations", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double),
        ("final_trust_region_radius", C.c_double),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SvinBaEvaluation(C.Structure):
    _fields_ = [(n, c_double_p) for n in (
        "reproj_residuals", "reproj_J_pose", "reproj_J_landmark", "reproj_J_extrinsics", "imu_residuals",
        "imu_J_pose0", "imu_J_speedbias0", "imu_J_pose1", "imu_J_speedbias1", "cost")]


class SvinBaTimings(C.Structure):
    _fields_ = [("solve_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("host_order_ms", C.c_double), ("host_fill_ms", C.c_double), ("host_upload_ms", C.c_double),
                ("host_scatter_ms", C.c_double)]


class SvinKeypoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32), ("class_id", C.c_int32)]


class SvinFeOptions(C.Structure):
    _fields_ = [("image_width", C.c_int32), ("image_height", C.c_int32), ("detection_threshold", C.c_double),
                ("detection_octaves", C.c_int32), ("absolute_threshold", C.c_double), ("max_keypoints", C.c_int32),
                ("rotation_invariance", C.c_int32), ("scale_invariance", C.c_int32), ("max_images", C.c_int32)]


SVIN_HIST_NONE, SVIN_HIST_EQUALIZE, SVIN_HIST_CLAHE = 0, 1, 2


class SvinPreOptions(C.Structure):
    _fields_ = [("src_width", C.c_int32), ("src_height", C.c_int32), ("resize_factor", C.c_double),
                ("median_filter", C.c_int32), ("histogram_method", C.c_int32), ("clahe_clip_limit", C.c_double),
                ("clahe_tiles", C.c_int32), ("max_images", C.c_int32)]


class SvinPreTimings(C.Structure):
    _fields_ = [("run_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("kernel_launches", C.c_int64), ("kernel_ms", C.c_double * 5)]


SVIN_MATCH_3D2D, SVIN_MATCH_2D2D = 0, 1
c_float_p = C.POINTER(C.c_float)


class SvinMatchProblem(C.Structure):
    _fields_ = [("type", C.c_int32), ("nA", C.c_int32), ("nB", C.c_int32), ("descA", c_uint8_p), ("descB", c_uint8_p),
                ("skipA", c_uint8_p), ("skipB", c_uint8_p), ("kpA", C.POINTER(SvinKeypoint)),
                ("kpB", C.POINTER(SvinKeypoint)), ("distance_threshold", C.c_float), ("landmarksA", c_double_p),
                ("T_CbW", c_double_p), ("pose_uncertainty", C.c_double), ("intrA", c_double_p), ("intrB", c_double_p),
                ("T_CaCb", c_double_p), ("image_width", C.c_int32), ("image_height", C.c_int32)]


class SvinMatchResult(C.Structure):
    _fields_ = [("best_index", c_int32_p), ("best_distance", c_float_p), ("match_of_B", c_int32_p),
                ("match_distance", c_float_p), ("skipA_effective", c_uint8_p)]


class SvinFeTimings(C.Structure):
    _fields_ = [("run_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("kernel_ms", C.c_double * 8)]


FE_KERNEL_NAMES = ["harris", "nms_compact", "sort", "uniformity", "orient_describe", "match", "assign", "unused"]

class SvinMargSpec(C.Structure):
    _fields_ = [("prior_num_blocks", C.c_int32), ("prior_block_kind", c_int32_p), ("prior_block_index", c_int32_p),
                ("prior_dim", C.c_int32), ("prior_H", c_double_p), ("prior_b0", c_double_p),
                ("marginalize_pose", c_uint8_p), ("marginalize_speedbias", c_uint8_p)]


class SvinMargResult(C.Structure):
    _fields_ = [("dim", C.c_int32), ("num_blocks", C.c_int32), ("block_kind", c_int32_p), ("block_index", c_int32_p),
                ("H", c_double_p), ("b0", c_double_p), ("J", c_double_p), ("e0", c_double_p)]


BA_KERNEL_NAMES = ["linearize", "dense_eval", "schur", "dense_solve", "backsub", "step_dense", "step_lm", "decide",
                   "clear"]


class SvinBaKernelTimes(C.Structure):
    _fields_ = [("ms", C.c_double * 9), ("launches", C.c_int64 * 9)]


class SvinRansacAbsProblem(C.Structure):
    _fields_ = [("num_correspondences", C.c_int32), ("points", c_double_p), ("bearings", c_double_p),
                ("camera_index", c_int32_p), ("sigma_angle", c_double_p), ("num_cameras", C.c_int32),
                ("camera_rotation", c_double_p), ("camera_offset", c_double_p), ("num_samples", C.c_int32),
                ("samples", c_int32_p), ("threshold", C.c_double), ("max_iterations", C.c_int32)]


class SvinRansacRelProblem(C.Structure):
    _fields_ = [("num_correspondences", C.c_int32), ("bearings1", c_double_p), ("bearings2", c_double_p),
                ("sigma_angle1", c_double_p), ("sigma_angle2", c_double_p), ("num_samples", C.c_int32),
                ("samples_rotation", c_int32_p), ("samples_relative", c_int32_p), ("threshold", C.c_double),
                ("max_iterations", C.c_int32)]


class SvinRansacResult(C.Structure):
    _fields_ = [("best_sample", C.c_int32), ("num_inliers", C.c_int32), ("iterations", C.c_int32),
                ("model", C.c_double * 12), ("inliers", c_uint8_p), ("hypothesis_inliers", c_int32_p),
                ("hypothesis_valid", c_uint8_p)]


class SvinVocabulary(C.Structure):
    _fields_ = [("num_nodes", C.c_int32), ("first_child", c_int32_p), ("num_children", c_int32_p),
                ("descriptor", c_uint8_p), ("weight", c_double_p), ("word_id", c_int32_p)]


class SvinError(RuntimeError):
    pass


_lib = None

# every symbol include/svin_b200.h declares (tests check the .so exports each one)
EXPORTED_SYMBOLS = [
    "svin_last_error", "svin_version", "svin_ba_default_options", "svin_ba_create", "svin_ba_destroy",
    "svin_ba_upload", "svin_ba_evaluate", "svin_ba_solve", "svin_ba_download", "svin_ba_download_all", "svin_ba_reset", "svin_ba_optimize",
    "svin_ba_timings", "svin_ba_set_profiling", "svin_ba_kernel_times", "svin_nccl_unique_id", "svin_ba_comm_init",
    "svin_ba_comm_init_local",
    "svin_ba_marginalize", "svin_ba_plan", "svin_ba_plan_observations", "svin_ba_uploaded_plan",
    "svin_fe_default_options", "svin_fe_create", "svin_fe_destroy", "svin_fe_detect_describe", "svin_fe_upload",
    "svin_fe_run", "svin_fe_download", "svin_fe_scores", "svin_match", "svin_fe_timings", "svin_fe_upload_device",
    "svin_pre_create", "svin_pre_destroy", "svin_pre_output_size", "svin_pre_process", "svin_pre_upload",
    "svin_pre_run", "svin_pre_download", "svin_pre_device_output", "svin_pre_timings",
    "svin_ransac_create", "svin_ransac_destroy", "svin_ransac_absolute", "svin_ransac_relative", "svin_ransac_timings",
    "svin_loop_create", "svin_loop_destroy", "svin_loop_transform", "svin_loop_add", "svin_loop_query",
    "svin_loop_brief_search", "svin_loop_stats",
]


def load(path: str | None = None) -> C.CDLL:
    """Load the CUDA engine.  Raises if the shared library has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise SvinError(
            f"{p} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback.")
    lib = C.CDLL(p)
    lib.svin_last_error.restype = C.c_char_p
    lib.svin_version.restype = C.c_char_p
    lib.svin_ba_default_options.argtypes = [C.POINTER(SvinBaOptions)]
    lib.svin_ba_default_options.restype = None
    lib.svin_ba_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.svin_ba_destroy.argtypes = [C.c_void_p]
    lib.svin_ba_destroy.restype = None
    lib.svin_ba_upload.argtypes = [C.c_void_p, C.POINTER(SvinBaWindow), C.c_int32]
    lib.svin_ba_evaluate.argtypes = [C.c_void_p, C.c_int32, C.POINTER(SvinBaEvaluation)]
    lib.svin_ba_solve.argtypes = [C.c_void_p, C.POINTER(SvinBaOptions), C.POINTER(SvinBaSummary)]
    lib.svin_ba_download.argtypes = [C.c_void_p, C.c_int32, C.POINTER(SvinBaWindow), c_double_p]
    lib.svin_ba_download_all.argtypes = [C.c_void_p, C.POINTER(SvinBaWindow), C.c_int32, C.POINTER(c_double_p)]
    lib.svin_ba_reset.argtypes = [C.c_void_p]
    lib.svin_ba_optimize.argtypes = [C.c_void_p, C.POINTER(SvinBaWindow), C.c_int32, C.POINTER(SvinBaOptions),
                                     C.POINTER(SvinBaSummary), C.POINTER(c_double_p)]
    lib.svin_ba_timings.argtypes = [C.c_void_p, C.POINTER(SvinBaTimings)]
    lib.svin_ba_marginalize.argtypes = [C.c_void_p, C.c_int32, C.POINTER(SvinMargSpec), C.POINTER(SvinMargResult)]
    lib.svin_nccl_unique_id.argtypes = [c_uint8_p]
    lib.svin_ba_comm_init.argtypes = [C.c_void_p, c_uint8_p, C.c_int32, C.c_int32]
    lib.svin_ba_comm_init_local.argtypes = [C.POINTER(C.c_void_p), C.c_int32]
    lib.svin_ba_set_profiling.argtypes = [C.c_void_p, C.c_int]
    lib.svin_ba_kernel_times.argtypes = [C.c_void_p, C.POINTER(SvinBaKernelTimes)]
    lib.svin_ba_plan.argtypes = [C.POINTER(SvinBaWindow), c_int32_p, C.c_int32, c_int32_p, c_int32_p, c_int32_p, c_int32_p]
    lib.svin_ba_plan_observations.argtypes = [C.POINTER(SvinBaWindow), c_int32_p]
    lib.svin_ba_uploaded_plan.argtypes = [C.c_void_p, C.c_int32, c_int32_p, C.c_int32, c_int32_p, c_int32_p, c_int32_p,
                                          c_int32_p, c_int32_p, c_int32_p]
    lib.svin_fe_default_options.argtypes = [C.POINTER(SvinFeOptions)]
    lib.svin_fe_default_options.restype = None
    lib.svin_fe_create.argtypes = [C.c_int, C.POINTER(SvinFeOptions), C.POINTER(C.c_void_p)]
    lib.svin_fe_destroy.argtypes = [C.c_void_p]
    lib.svin_fe_destroy.restype = None
    pp_u8 = C.POINTER(c_uint8_p)
    lib.svin_fe_detect_describe.argtypes = [C.c_void_p, C.c_int32, pp_u8, C.c_int32, c_double_p, c_double_p,
                                            C.POINTER(SvinKeypoint), c_uint8_p, c_int32_p]
    lib.svin_fe_upload.argtypes = [C.c_void_p, C.c_int32, pp_u8, C.c_int32, c_double_p, c_double_p]
    lib.svin_fe_run.argtypes = [C.c_void_p]
    lib.svin_fe_download.argtypes = [C.c_void_p, C.POINTER(SvinKeypoint), c_uint8_p, c_int32_p]
    lib.svin_fe_scores.argtypes = [C.c_void_p, C.c_int32, c_int32_p]
    lib.svin_match.argtypes = [C.c_void_p, C.c_int32, C.POINTER(SvinMatchProblem), C.POINTER(SvinMatchResult)]
    lib.svin_fe_timings.argtypes = [C.c_void_p, C.POINTER(SvinFeTimings)]
    lib.svin_fe_upload_device.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, c_double_p, c_double_p]
    lib.svin_pre_create.argtypes = [C.c_int, C.POINTER(SvinPreOptions), C.POINTER(C.c_void_p)]
    lib.svin_pre_destroy.argtypes = [C.c_void_p]
    lib.svin_pre_destroy.restype = None
    lib.svin_pre_output_size.argtypes = [C.c_void_p, c_int32_p, c_int32_p]
    lib.svin_pre_process.argtypes = [C.c_void_p, C.c_int32, pp_u8, C.c_int32, pp_u8]
    lib.svin_pre_upload.argtypes = [C.c_void_p, C.c_int32, pp_u8, C.c_int32]
    lib.svin_pre_run.argtypes = [C.c_void_p]
    lib.svin_pre_download.argtypes = [C.c_void_p, pp_u8]
    lib.svin_pre_device_output.argtypes = [C.c_void_p]
    lib.svin_pre_device_output.restype = C.c_void_p
    lib.svin_pre_timings.argtypes = [C.c_void_p, C.POINTER(SvinPreTimings)]
    lib.svin_ransac_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.svin_ransac_destroy.argtypes = [C.c_void_p]
    lib.svin_ransac_destroy.restype = None
    lib.svin_ransac_absolute.argtypes = [C.c_void_p, C.c_int32, C.POINTER(SvinRansacAbsProblem),
                                         C.POINTER(SvinRansacResult)]
    lib.svin_ransac_relative.argtypes = [C.c_void_p, C.c_int32, C.POINTER(SvinRansacRelProblem),
                                         C.POINTER(SvinRansacResult), C.POINTER(SvinRansacResult)]
    lib.svin_ransac_timings.argtypes = [C.c_void_p, c_double_p, C.POINTER(C.c_int64)]
    lib.svin_loop_create.argtypes = [C.c_int, C.POINTER(SvinVocabulary), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    lib.svin_loop_destroy.argtypes = [C.c_void_p]
    lib.svin_loop_destroy.restype = None
    lib.svin_loop_transform.argtypes = [C.c_void_p, C.c_int32, c_uint8_p, c_int32_p, c_int32_p, c_double_p, c_int32_p]
    lib.svin_loop_add.argtypes = [C.c_void_p, C.c_int32, c_uint8_p, c_int32_p]
    lib.svin_loop_query.argtypes = [C.c_void_p, c_uint8_p, C.c_int32, C.c_int32, C.c_int32, c_int32_p, c_double_p, c_int32_p]
    lib.svin_loop_brief_search.argtypes = [C.c_void_p, c_uint8_p, C.c_int32, c_uint8_p, C.c_int32, c_int32_p, c_int32_p,
                                           c_uint8_p]
    lib.svin_loop_stats.argtypes = [C.c_void_p, c_int32_p, c_int32_p, c_double_p]
    if path is None:
        _lib = lib
    return lib


def check(status: int, lib: C.CDLL | None = None) -> None:
    if status != 0:
        lib = lib or load()
        msg = lib.svin_last_error()
        raise SvinError(f"svin_b200 error {status}: {msg.decode() if msg else '?'}")
