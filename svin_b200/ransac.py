"""Host-side mirror of the RANSAC calls of okvis::Frontend over the C ABI (svin_ransac_*, include/svin_b200.h).

  run_ransac_3d2d  <->  Frontend::runRansac3d2d               (okvis_frontend/src/Frontend.cpp:617-676)
  run_ransac_2d2d  <->  Frontend::runRansac2d2d (:832-980) and runRansac2d2dToRefineScale (:680-830)

Same names, argument meaning and decisions (>= 10 / > 10 inliers, rotation-only vs relative-pose ratio rule); the
consensus itself runs on the device.  There is no CPU path: RansacEngine raises SvinError without the CUDA library."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

THRESHOLD, MAX_ITERATIONS = 9.0, 50          # Frontend.cpp:643-644


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a.size else C.POINTER(t)()


class _Result:
    def __init__(self, n, ns):
        self.inliers = np.zeros(n, np.uint8)
        self.counts = np.zeros(ns, np.int32)
        self.valid = np.zeros(ns, np.uint8)

    def bind(self, r: capi.SvinRansacResult):
        r.inliers = _p(self.inliers, C.c_uint8)
        r.hypothesis_inliers = _p(self.counts, C.c_int32)
        r.hypothesis_valid = _p(self.valid, C.c_uint8)

    def unpack(self, r: capi.SvinRansacResult) -> dict:
        M = np.array(list(r.model)).reshape(3, 4)
        return dict(best=int(r.best_sample), num_inliers=int(r.num_inliers), iterations=int(r.iterations),
                    R=M[:, :3].copy(), t=M[:, 3].copy(), inliers=self.inliers.astype(bool), counts=self.counts.copy(),
                    valid=self.valid.astype(bool))


class RansacEngine:
    def __init__(self, device: int = 0):
        self._lib = capi.load()
        self._ctx = C.c_void_p()
        capi.check(self._lib.svin_ransac_create(device, C.byref(self._ctx)), self._lib)

    def close(self):
        if self._ctx:
            self._lib.svin_ransac_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def absolute(self, problems: list[dict], threshold=THRESHOLD, max_iterations=MAX_ITERATIONS) -> list[dict]:
        """problems: dict(points, bearings, cam_index, cam_R [c][3][3], cam_t [c][3], sigma, samples [ns][4])."""
        n = len(problems)
        arr = (capi.SvinRansacAbsProblem * n)()
        res = (capi.SvinRansacResult * n)()
        keep, outs = [], []
        c = np.ascontiguousarray
        for k, p in enumerate(problems):
            a = dict(points=c(p["points"], np.float64), bearings=c(p["bearings"], np.float64),
                     cam=c(p["cam_index"], np.int32), sigma=c(p["sigma"], np.float64),
                     cR=c(np.asarray(p["cam_R"]).reshape(-1, 9), np.float64), ct=c(p["cam_t"], np.float64),
                     smp=c(p["samples"], np.int32))
            keep.append(a)
            q = arr[k]
            q.num_correspondences, q.num_cameras, q.num_samples = len(a["points"]), len(a["ct"]), len(a["smp"])
            q.points, q.bearings = _p(a["points"], C.c_double), _p(a["bearings"], C.c_double)
            q.camera_index, q.sigma_angle = _p(a["cam"], C.c_int32), _p(a["sigma"], C.c_double)
            q.camera_rotation, q.camera_offset = _p(a["cR"], C.c_double), _p(a["ct"], C.c_double)
            q.samples, q.threshold, q.max_iterations = _p(a["smp"], C.c_int32), threshold, max_iterations
            o = _Result(q.num_correspondences, q.num_samples)
            o.bind(res[k])
            outs.append(o)
        capi.check(self._lib.svin_ransac_absolute(self._ctx, n, arr, res), self._lib)
        return [o.unpack(res[k]) for k, o in enumerate(outs)]

    def relative(self, problems: list[dict], threshold=THRESHOLD, max_iterations=MAX_ITERATIONS):
        """problems: dict(f1, f2, sigma1, sigma2, samples_rot [ns][2], samples_rel [ns][8]) -> [(rotation-only, relative)]."""
        n = len(problems)
        arr = (capi.SvinRansacRelProblem * n)()
        r_rot = (capi.SvinRansacResult * n)()
        r_rel = (capi.SvinRansacResult * n)()
        keep, outs = [], []
        c = np.ascontiguousarray
        for k, p in enumerate(problems):
            a = dict(f1=c(p["f1"], np.float64), f2=c(p["f2"], np.float64), s1=c(p["sigma1"], np.float64),
                     s2=c(p["sigma2"], np.float64), sr=c(p["samples_rot"], np.int32), sp=c(p["samples_rel"], np.int32))
            assert len(a["sr"]) == len(a["sp"])
            keep.append(a)
            q = arr[k]
            q.num_correspondences, q.num_samples = len(a["f1"]), len(a["sr"])
            q.bearings1, q.bearings2 = _p(a["f1"], C.c_double), _p(a["f2"], C.c_double)
            q.sigma_angle1, q.sigma_angle2 = _p(a["s1"], C.c_double), _p(a["s2"], C.c_double)
            q.samples_rotation, q.samples_relative = _p(a["sr"], C.c_int32), _p(a["sp"], C.c_int32)
            q.threshold, q.max_iterations = threshold, max_iterations
            o = (_Result(q.num_correspondences, q.num_samples), _Result(q.num_correspondences, q.num_samples))
            o[0].bind(r_rot[k])
            o[1].bind(r_rel[k])
            outs.append(o)
        capi.check(self._lib.svin_ransac_relative(self._ctx, n, arr, r_rot, r_rel), self._lib)
        return [(o[0].unpack(r_rot[k]), o[1].unpack(r_rel[k])) for k, o in enumerate(outs)]

    def timings(self) -> dict:
        ms, ln = C.c_double(), C.c_int64()
        capi.check(self._lib.svin_ransac_timings(self._ctx, C.byref(ms), C.byref(ln)), self._lib)
        return dict(device_ms=ms.value, kernel_launches=ln.value)


def run_ransac_3d2d(engine: RansacEngine, problem: dict):
    """Frontend::runRansac3d2d: -> (numInliers, keep mask or None).  Fewer than 5 correspondences: nothing is run
    (:634); outliers are only kicked out when the consensus has >= 10 inliers (:650-672)."""
    n = len(problem["points"])
    if n < 5:
        return n, None
    r = engine.absolute([problem])[0]
    return r["num_inliers"], (r["inliers"] if r["num_inliers"] >= 10 else None), r


def run_ransac_2d2d(engine: RansacEngine, problem: dict):
    """Frontend::runRansac2d2d for one camera pair: -> dict(inliers mask or None, rotation_only, success,
    num_inliers, T = winning 3x4 model).  < 10 correspondences: skipped (:855-856)."""
    n = len(problem["f1"])
    if n < 10:
        return dict(inliers=None, rotation_only=False, success=False, num_inliers=0, T=None)
    rot, rel = engine.relative([problem])[0]
    rr = np.float32(rot["num_inliers"]) / np.float32(n)
    pr = np.float32(rel["num_inliers"]) / np.float32(n)
    if rr > pr or rr > np.float32(0.8):                              # :878
        win, rotation_only, success = rot, True, rot["num_inliers"] > 10
    else:
        win, rotation_only, success = rel, False, rel["num_inliers"] > 10
    return dict(inliers=win["inliers"] if success else None, rotation_only=rotation_only, success=success,
                num_inliers=win["num_inliers"], T=np.concatenate([win["R"], win["t"][:, None]], axis=1),
                rotation=rot, relative=rel)
