"""Python handle on the CUDA front-end (ctypes over include/svin_b200.h): Frame::detect/describe and
DenseMatcher::match replacements.  No CPU path: raises SvinError without the CUDA library / device."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])


def _u8(a):
    return a.ctypes.data_as(capi.c_uint8_p) if a is not None else capi.c_uint8_p()


def _dp(a):
    return a.ctypes.data_as(capi.c_double_p) if a is not None else capi.c_double_p()


class MatchProblem:
    """One DenseMatcher::match<VioKeyframeWindowMatchingAlgorithm> call; keeps its numpy arrays alive."""

    def __init__(self, type_, descA, descB, kpA, kpB, intrA, intrB, W, H, skipA=None, skipB=None, landmarksA=None,
                 T_CbW=None, pose_uncertainty=4e-8, T_CaCb=None, thr=60.0):
        c = np.ascontiguousarray
        self.descA, self.descB = c(descA, dtype=np.uint8), c(descB, dtype=np.uint8)
        self.kpA, self.kpB = c(kpA), c(kpB)
        self.intrA, self.intrB = c(intrA, dtype=np.float64), c(intrB, dtype=np.float64)
        self.skipA = c(skipA, dtype=np.uint8) if skipA is not None else None
        self.skipB = c(skipB, dtype=np.uint8) if skipB is not None else None
        self.landmarksA = c(landmarksA, dtype=np.float64) if landmarksA is not None else None
        self.T_CbW = c(T_CbW, dtype=np.float64) if T_CbW is not None else None
        self.T_CaCb = c(T_CaCb, dtype=np.float64) if T_CaCb is not None else None
        p = capi.SvinMatchProblem()
        p.type, p.nA, p.nB = type_, len(self.kpA), len(self.kpB)
        p.descA, p.descB = _u8(self.descA), _u8(self.descB)
        p.skipA, p.skipB = _u8(self.skipA), _u8(self.skipB)
        p.kpA = self.kpA.ctypes.data_as(C.POINTER(capi.SvinKeypoint))
        p.kpB = self.kpB.ctypes.data_as(C.POINTER(capi.SvinKeypoint))
        p.distance_threshold = thr
        p.landmarksA, p.T_CbW = _dp(self.landmarksA), _dp(self.T_CbW)
        p.pose_uncertainty = pose_uncertainty
        p.intrA, p.intrB, p.T_CaCb = _dp(self.intrA), _dp(self.intrB), _dp(self.T_CaCb)
        p.image_width, p.image_height = W, H
        self.c = p


class FeEngine:
    def __init__(self, width=752, height=480, max_images=2, device=0, **opts):
        self._lib = capi.load()
        o = capi.SvinFeOptions()
        self._lib.svin_fe_default_options(C.byref(o))
        o.image_width, o.image_height, o.max_images = width, height, max_images
        for k, v in opts.items():
            setattr(o, k, v)
        self.opt = o
        self._ctx = C.c_void_p()
        capi.check(self._lib.svin_fe_create(device, C.byref(o), C.byref(self._ctx)), self._lib)

    def close(self):
        if self._ctx:
            self._lib.svin_fe_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- detect + describe ---------------------------------------------------------------------------
    def _pack(self, images, intrinsics, extraction_dirs):
        self._imgs = [np.ascontiguousarray(i, dtype=np.uint8) for i in images]
        n = len(self._imgs)
        ptrs = (capi.c_uint8_p * n)(*[_u8(i) for i in self._imgs])
        self._intr = np.ascontiguousarray(intrinsics, dtype=np.float64).reshape(n, 8)
        self._edir = np.ascontiguousarray(extraction_dirs, dtype=np.float64).reshape(n, 3)
        return n, ptrs

    def _alloc_out(self, n):
        K = self.opt.max_keypoints
        return np.zeros((n, K), dtype=KP_DTYPE), np.zeros((n, K, 48), dtype=np.uint8), np.zeros(n, dtype=np.int32)

    def _split(self, kps, desc, counts):
        return [(kps[i, :counts[i]].copy(), desc[i, :counts[i]].copy()) for i in range(len(counts))]

    def detect_describe(self, images, intrinsics, extraction_dirs):
        """Host images in, host keypoints/descriptors out (the drop-in call)."""
        n, ptrs = self._pack(images, intrinsics, extraction_dirs)
        kps, desc, counts = self._alloc_out(n)
        capi.check(self._lib.svin_fe_detect_describe(
            self._ctx, n, ptrs, self._imgs[0].shape[1], _dp(self._intr), _dp(self._edir),
            kps.ctypes.data_as(C.POINTER(capi.SvinKeypoint)), _u8(desc), counts.ctypes.data_as(capi.c_int32_p)),
            self._lib)
        return self._split(kps, desc, counts)

    def upload(self, images, intrinsics, extraction_dirs):
        n, ptrs = self._pack(images, intrinsics, extraction_dirs)
        capi.check(self._lib.svin_fe_upload(self._ctx, n, ptrs, self._imgs[0].shape[1], _dp(self._intr),
                                            _dp(self._edir)), self._lib)
        self._n = n

    def run(self):
        capi.check(self._lib.svin_fe_run(self._ctx), self._lib)

    def download(self):
        kps, desc, counts = self._alloc_out(self._n)
        capi.check(self._lib.svin_fe_download(self._ctx, kps.ctypes.data_as(C.POINTER(capi.SvinKeypoint)), _u8(desc),
                                              counts.ctypes.data_as(capi.c_int32_p)), self._lib)
        return self._split(kps, desc, counts)

    def scores(self, index=0):
        out = np.zeros((self.opt.image_height, self.opt.image_width), dtype=np.int32)
        capi.check(self._lib.svin_fe_scores(self._ctx, index, out.ctypes.data_as(capi.c_int32_p)), self._lib)
        return out

    # ---- matching ---------------------------------------------------------------------------------------
    def match(self, problems: list[MatchProblem]):
        n = len(problems)
        parr = (capi.SvinMatchProblem * n)(*[p.c for p in problems])
        outs, rarr = [], (capi.SvinMatchResult * n)()
        for i, p in enumerate(problems):
            nA, nB = p.c.nA, p.c.nB
            o = dict(best_index=np.zeros((nA, 4), np.int32), best_distance=np.zeros((nA, 4), np.float32),
                     match_of_B=np.zeros(nB, np.int32), match_distance=np.zeros(nB, np.float32),
                     skipA=np.zeros(nA, np.uint8))
            rarr[i].best_index = o["best_index"].ctypes.data_as(capi.c_int32_p)
            rarr[i].best_distance = o["best_distance"].ctypes.data_as(capi.c_float_p)
            rarr[i].match_of_B = o["match_of_B"].ctypes.data_as(capi.c_int32_p)
            rarr[i].match_distance = o["match_distance"].ctypes.data_as(capi.c_float_p)
            rarr[i].skipA_effective = _u8(o["skipA"])
            outs.append(o)
        capi.check(self._lib.svin_match(self._ctx, n, parr, rarr), self._lib)
        return outs

    def timings(self) -> dict:
        t = capi.SvinFeTimings()
        capi.check(self._lib.svin_fe_timings(self._ctx, C.byref(t)), self._lib)
        d = {n: getattr(t, n) for n, _ in t._fields_ if n != "kernel_ms"}
        d["kernel_ms"] = {n: t.kernel_ms[i] for i, n in enumerate(capi.FE_KERNEL_NAMES)}
        return d
