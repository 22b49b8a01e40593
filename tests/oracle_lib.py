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
    override = os.environ.get("SVIN_ORACLE_SO")   # bench.py's CPU arm: the -O3 -march=native timing build
    if override and os.path.exists(override):
        return _bind(C.CDLL(override))
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".hpp"))]
    if (not os.path.exists(ORACLE_SO)) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs):
        try:
            build()
        except Exception:
            if not os.path.exists(ORACLE_SO):
                raise
    return _bind(C.CDLL(ORACLE_SO))


def _bind(lib):
    global _lib
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


# ------------------------------------------------------------------------------ front-end oracle
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref_matcher.so")
_ref = None


def load_fe():
    lib = load()
    if getattr(lib, "_fe_ready", False):
        return lib
    u8p, i32p, fp, dp = capi.c_uint8_p, capi.c_int32_p, capi.c_float_p, capi.c_double_p
    lib.svin_oracle_fe_detect_describe.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, dp,
                                                   dp, C.POINTER(capi.SvinKeypoint), u8p]
    lib.svin_oracle_fe_harris.argtypes = [u8p, C.c_int, C.c_int, C.c_int, i32p]
    lib.svin_oracle_hamming48.argtypes = [u8p, u8p]
    lib.svin_oracle_hamming48.restype = C.c_uint32
    lib.svin_oracle_atan2.argtypes = [C.c_double, C.c_double]
    lib.svin_oracle_atan2.restype = C.c_double
    lib.svin_oracle_match.argtypes = [C.POINTER(capi.SvinMatchProblem), i32p, fp, i32p, fp, u8p]
    lib.svin_oracle_match_matrix.argtypes = [C.c_int, C.c_int, fp, u8p, u8p, C.c_float, i32p, fp, i32p, fp]
    lib._fe_ready = True
    return lib


def load_ref_matcher():
    """The reference's own DenseMatcher (oracle/_ref, built from /root/reference when present); None if absent."""
    global _ref
    if _ref is not None:
        return _ref
    if not os.path.exists(REF_SO):
        try:
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])
        except Exception:
            pass
    if not os.path.exists(REF_SO):
        return None
    lib = C.CDLL(REF_SO)
    lib.svin_ref_dense_match.argtypes = [C.c_int, C.c_int, capi.c_float_p, capi.c_uint8_p, capi.c_uint8_p, C.c_float,
                                         C.c_float, C.c_int, C.c_int, capi.c_int32_p, capi.c_float_p]
    _ref = lib
    return lib


def _u8(a):
    return a.ctypes.data_as(capi.c_uint8_p) if a is not None else capi.c_uint8_p()


def fe_detect_describe(img, intr, edir, radius=40.0, abs_thr=800.0, max_kp=400):
    lib = load_fe()
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    kps = (capi.SvinKeypoint * max_kp)()
    desc = np.zeros((max_kp, 48), dtype=np.uint8)
    intr = np.ascontiguousarray(intr, dtype=np.float64)
    edir = np.ascontiguousarray(edir, dtype=np.float64)
    n = lib.svin_oracle_fe_detect_describe(_u8(img), W, W, H, radius, abs_thr, max_kp, P(intr), P(edir), kps, _u8(desc))
    arr = np.frombuffer(kps, dtype=KP_DTYPE)[:n].copy()
    return arr, desc[:n].copy()


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])


def fe_harris(img):
    lib = load_fe()
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    out = np.zeros((H, W), dtype=np.int32)
    lib.svin_oracle_fe_harris(_u8(img), W, W, H, out.ctypes.data_as(capi.c_int32_p))
    return out


def match_matrix(D, skipA=None, skipB=None, thr=60.0):
    lib = load_fe()
    D = np.ascontiguousarray(D, dtype=np.float32)
    nA, nB = D.shape
    bi, bd = np.zeros((nA, 4), np.int32), np.zeros((nA, 4), np.float32)
    mb, md = np.zeros(nB, np.int32), np.zeros(nB, np.float32)
    lib.svin_oracle_match_matrix(nA, nB, D.ctypes.data_as(capi.c_float_p), _u8(skipA), _u8(skipB), thr,
                                 bi.ctypes.data_as(capi.c_int32_p), bd.ctypes.data_as(capi.c_float_p),
                                 mb.ctypes.data_as(capi.c_int32_p), md.ctypes.data_as(capi.c_float_p))
    return bi, bd, mb, md


def ref_match_matrix(D, skipA=None, skipB=None, thr=60.0, ratio=3.0, use_ratio=False, threads=1):
    lib = load_ref_matcher()
    assert lib is not None
    D = np.ascontiguousarray(D, dtype=np.float32)
    nA, nB = D.shape
    mb, md = np.zeros(nB, np.int32), np.zeros(nB, np.float32)
    lib.svin_ref_dense_match(nA, nB, D.ctypes.data_as(capi.c_float_p), _u8(skipA), _u8(skipB), thr, ratio,
                             int(use_ratio), threads, mb.ctypes.data_as(capi.c_int32_p),
                             md.ctypes.data_as(capi.c_float_p))
    return mb, md


from svin_b200.frontend import MatchProblem as MatchArgs  # noqa: E402  (same container for oracle and engine)


def fe_match(args: MatchArgs):
    lib = load_fe()
    nA, nB = args.c.nA, args.c.nB
    bi, bd = np.zeros((nA, 4), np.int32), np.zeros((nA, 4), np.float32)
    mb, md = np.zeros(nB, np.int32), np.zeros(nB, np.float32)
    sk = np.zeros(nA, np.uint8)
    lib.svin_oracle_match(C.byref(args.c), bi.ctypes.data_as(capi.c_int32_p), bd.ctypes.data_as(capi.c_float_p),
                          mb.ctypes.data_as(capi.c_int32_p), md.ctypes.data_as(capi.c_float_p), _u8(sk))
    return dict(best_index=bi, best_distance=bd, match_of_B=mb, match_distance=md, skipA=sk)


# ------------------------------------------------------------------------------ marginalisation oracle
def marginalize(window, spec):
    from svin_b200.marginalization import MargResult
    lib = load()
    lib.svin_oracle_marginalize.argtypes = [C.POINTER(capi.SvinBaWindow), C.POINTER(capi.SvinMargSpec),
                                            C.POINTER(capi.SvinMargResult)]
    res = MargResult(window)
    s = window.c_struct()
    assert lib.svin_oracle_marginalize(C.byref(s), C.byref(spec.c), C.byref(res.c)) == 0
    return res.unpack()
