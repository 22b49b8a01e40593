"""Host-side mirror of the image pre-processing in front of the detector (svin_pre_*).

Reference: Subscriber::imageCallback (okvis_ros/src/Subscriber.cpp:123-147): cv::resize by miscParams.resizeFactor,
optional cv::medianBlur(3), CLAHE / equalizeHist per histogramParams, then VioInterface::addImage.  The option names
follow the reference's yaml keys (resizeFactor, useMedianFilter, histogramMethod, claheClipLimit, claheTilesGridSize).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

HISTOGRAM_METHODS = {"NONE": capi.SVIN_HIST_NONE, "HISTOGRAM": capi.SVIN_HIST_EQUALIZE, "CLAHE": capi.SVIN_HIST_CLAHE}


def _u8(a):
    return a.ctypes.data_as(capi.c_uint8_p)


class Preprocessor:
    def __init__(self, src_width, src_height, resizeFactor=1.0, useMedianFilter=False, histogramMethod="NONE",
                 claheClipLimit=1.0, claheTilesGridSize=4, max_images=2, device=0):
        self._lib = capi.load()
        o = capi.SvinPreOptions()
        o.src_width, o.src_height, o.resize_factor = src_width, src_height, float(resizeFactor)
        o.median_filter = int(bool(useMedianFilter))
        o.histogram_method = HISTOGRAM_METHODS[histogramMethod] if isinstance(histogramMethod, str) else histogramMethod
        o.clahe_clip_limit, o.clahe_tiles, o.max_images = float(claheClipLimit), int(claheTilesGridSize), max_images
        self.opt = o
        self._ctx = C.c_void_p()
        capi.check(self._lib.svin_pre_create(device, C.byref(o), C.byref(self._ctx)), self._lib)
        w, h = C.c_int32(), C.c_int32()
        capi.check(self._lib.svin_pre_output_size(self._ctx, C.byref(w), C.byref(h)), self._lib)
        self.width, self.height = w.value, h.value
        self._n = 0

    def close(self):
        if self._ctx:
            self._lib.svin_pre_destroy(self._ctx)
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

    def _ptrs(self, images):
        self._imgs = [np.ascontiguousarray(i, dtype=np.uint8) for i in images]
        return (capi.c_uint8_p * len(self._imgs))(*[_u8(i) for i in self._imgs])

    def _outs(self, n):
        out = np.zeros((n, self.height, self.width), dtype=np.uint8)
        return out, (capi.c_uint8_p * n)(*[_u8(out[i]) for i in range(n)])

    def process(self, images):
        """Host images in, host images out (one call: H2D, kernels, D2H)."""
        ptrs = self._ptrs(images)
        n = len(self._imgs)
        out, optrs = self._outs(n)
        capi.check(self._lib.svin_pre_process(self._ctx, n, ptrs, self._imgs[0].shape[1], optrs), self._lib)
        self._n = n
        return out

    def upload(self, images):
        ptrs = self._ptrs(images)
        self._n = len(self._imgs)
        capi.check(self._lib.svin_pre_upload(self._ctx, self._n, ptrs, self._imgs[0].shape[1]), self._lib)

    def run(self):
        capi.check(self._lib.svin_pre_run(self._ctx), self._lib)

    def download(self):
        out, optrs = self._outs(self._n)
        capi.check(self._lib.svin_pre_download(self._ctx, optrs), self._lib)
        return out

    def device_output(self) -> int:
        return self._lib.svin_pre_device_output(self._ctx)

    def timings(self) -> dict:
        t = capi.SvinPreTimings()
        capi.check(self._lib.svin_pre_timings(self._ctx, C.byref(t)), self._lib)
        names = ("resize", "median", "histogram", "lut", "apply")
        return {"run_ms": t.run_ms, "h2d_ms": t.h2d_ms, "d2h_ms": t.d2h_ms, "h2d_bytes": t.h2d_bytes,
                "d2h_bytes": t.d2h_bytes, "kernel_launches": t.kernel_launches,
                "kernel_ms": dict(zip(names, t.kernel_ms))}
