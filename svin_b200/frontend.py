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


class MatchResults:
    """Pooled outputs of one svin_match call; res[i] -> dict of views for problem i."""

    _DT = np.dtype([("best_index", "<u8"), ("best_distance", "<u8"), ("match_of_B", "<u8"), ("match_distance", "<u8"),
                    ("skipA_effective", "<u8")])

    def __init__(self, nA, nB):
        assert self._DT.itemsize == C.sizeof(capi.SvinMatchResult)
        self.offA = np.concatenate([[0], np.cumsum(nA)])
        self.offB = np.concatenate([[0], np.cumsum(nB)])
        tA, tB = int(self.offA[-1]), int(self.offB[-1])
        self.best_index = np.zeros((tA, 4), np.int32)
        self.best_distance = np.zeros((tA, 4), np.float32)
        self.match_of_B = np.zeros(tB, np.int32)
        self.match_distance = np.zeros(tB, np.float32)
        self.skipA = np.zeros(tA, np.uint8)
        t = np.zeros(len(nA), dtype=self._DT)
        t["best_index"] = self.best_index.ctypes.data + 16 * self.offA[:-1]
        t["best_distance"] = self.best_distance.ctypes.data + 16 * self.offA[:-1]
        t["match_of_B"] = self.match_of_B.ctypes.data + 4 * self.offB[:-1]
        t["match_distance"] = self.match_distance.ctypes.data + 4 * self.offB[:-1]
        t["skipA_effective"] = self.skipA.ctypes.data + self.offA[:-1]
        self.table = t

    def __len__(self):
        return len(self.table)

    def __getitem__(self, i):
        a0, a1, b0, b1 = self.offA[i], self.offA[i + 1], self.offB[i], self.offB[i + 1]
        return dict(best_index=self.best_index[a0:a1], best_distance=self.best_distance[a0:a1],
                    match_of_B=self.match_of_B[b0:b1], match_distance=self.match_distance[b0:b1],
                    skipA=self.skipA[a0:a1])

    def __iter__(self):
        return (self[i] for i in range(len(self)))


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

    def upload_device(self, device_images: int, n: int, intrinsics, extraction_dirs):
        """Images already on the device (dense rows), e.g. Preprocessor.device_output(): no host round trip."""
        self._intr = np.ascontiguousarray(intrinsics, dtype=np.float64).reshape(n, 8)
        self._edir = np.ascontiguousarray(extraction_dirs, dtype=np.float64).reshape(n, 3)
        capi.check(self._lib.svin_fe_upload_device(self._ctx, n, C.c_void_p(device_images), _dp(self._intr),
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
        """Returns a MatchResults (indexable like a list of dicts).  Outputs of all problems live in five pooled
        arrays and the SvinMatchResult pointer table is filled with vectorised address arithmetic: with ~1 000
        problems per call a per-problem Python loop costs more than the matching itself."""
        n = len(problems)
        key = tuple(id(p) for p in problems)
        if getattr(self, "_match_key", None) != key:     # same problem list as last call: reuse the struct array
            self._match_parr = (capi.SvinMatchProblem * n)(*[p.c for p in problems])
            self._match_keep = list(problems)
            self._match_nA = np.array([p.c.nA for p in problems], dtype=np.int64)
            self._match_nB = np.array([p.c.nB for p in problems], dtype=np.int64)
            self._match_key = key
        res = MatchResults(self._match_nA, self._match_nB)
        capi.check(self._lib.svin_match(self._ctx, n, self._match_parr,
                                        C.cast(res.table.ctypes.data, C.POINTER(capi.SvinMatchResult))), self._lib)
        return res

    def timings(self) -> dict:
        t = capi.SvinFeTimings()
        capi.check(self._lib.svin_fe_timings(self._ctx, C.byref(t)), self._lib)
        d = {n: getattr(t, n) for n, _ in t._fields_ if n != "kernel_ms"}
        d["kernel_ms"] = {n: t.kernel_ms[i] for i, n in enumerate(capi.FE_KERNEL_NAMES)}
        return d
