"""TEST INFRASTRUCTURE: loads oracle/libsvin_oracle.so (building it with gcc if needed)."""
import ctypes as C
import os
import subprocess

import numpy as np

from svin_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libsvin_oracle.so")

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "libsvin_oracle.so"])


def load():
    global _lib
    if _lib is not None:
        return _lib
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".hpp"))]
    if (not os.path.exists(ORACLE_SO)) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs):
        try:
            build()
        except Exception:
            if not os.path.exists(ORACLE_SO):
                raise
    lib = C.CDLL(ORACLE_SO)
    dp = capi.c_double_p
    lib.svin_oracle_default_options.argtypes = [C.POINTER(capi.SvinBaOptions)]
    lib.svin_oracle_ba_evaluate.argtypes = [C.POINTER(capi.SvinBaWindow), C.POINTER(capi.SvinBaEvaluation)]
    lib.svin_oracle_ba_solve.argtypes = [C.POINTER(capi.SvinBaWindow), C.POINTER(capi.SvinBaOptions),
                                         C.POINTER(capi.SvinBaSummary), dp]
    lib.svin_oracle_reprojection.argtypes = [dp] * 10 + [C.POINTER(C.c_int)]
    lib.svin_oracle_project.argtypes = [dp, dp, dp, dp, C.c_int, C.c_int]
    lib.svin_oracle_backproject.argtypes = [dp, dp, dp]
    lib.svin_oracle_imu.argtypes = [C.c_int, capi.c_int64_p, dp, dp, C.POINTER(capi.SvinImuParams), C.c_int64,
                                    C.c_int64, dp, dp, dp, dp, dp, dp, dp, dp, dp]
    lib.svin_oracle_pose_plus.argtypes = [dp, dp, dp]
    lib.svin_oracle_pose_minus.argtypes = [dp, dp, dp]
    lib.svin_oracle_sqrt_information.argtypes = [dp, dp, C.c_int]
    lib.svin_oracle_pose_error.argtypes = [dp] * 5
    lib.svin_oracle_relative_pose_error.argtypes = [dp] * 6
    lib.svin_oracle_sonar_error.argtypes = [C.c_double, C.c_double, C.c_double, dp, dp, dp, dp, dp]
    lib.svin_oracle_sym3_eigenvalues.argtypes = [dp, dp]
    _lib = lib
    return lib


def P(a):
    """double* of a contiguous float64 array (or NULL for None)."""
    if a is None:
        return capi.c_double_p()
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(capi.c_double_p)


def evaluate(window):
    """Oracle term dump for a BaWindow -> dict of numpy arrays."""
    lib = load()
    n, m = window.num_obs, len(window.imu_pose0)
    out = dict(reproj_residuals=np.zeros((n, 2)), reproj_J_pose=np.zeros((n, 2, 6)),
               reproj_J_landmark=np.zeros((n, 2, 3)), reproj_J_extrinsics=np.zeros((n, 2, 6)),
               imu_residuals=np.zeros((m, 15)), imu_J_pose0=np.zeros((m, 15, 6)), imu_J_speedbias0=np.zeros((m, 15, 9)),
               imu_J_pose1=np.zeros((m, 15, 6)), imu_J_speedbias1=np.zeros((m, 15, 9)), cost=np.zeros(1))
    ev = capi.SvinBaEvaluation()
    for k, v in out.items():
        setattr(ev, k, P(v))
    s = window.c_struct()
    assert lib.svin_oracle_ba_evaluate(C.byref(s), C.byref(ev)) == 0
    return out


def solve(window, options=None, quality=True):
    """Oracle trust-region solve IN PLACE on `window`; returns (summary dict, landmark_quality)."""
    from svin_b200.window import default_options
    lib = load()
    opt = options or default_options()
    summ = capi.SvinBaSummary()
    q = np.zeros(window.num_landmarks) if quality else None
    s = window.c_struct()
    rc = lib.svin_oracle_ba_solve(C.byref(s), C.byref(opt), C.byref(summ), P(q))
    assert rc == 0, rc
    return summ.as_dict(), q
