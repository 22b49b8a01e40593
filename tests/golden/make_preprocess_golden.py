#!/usr/bin/env python
"""Golden vectors for the image pre-processing chain, generated with the real OpenCV (cv2 wheel of this image).

The reference (okvis_ros/src/Subscriber.cpp:123-147) calls cv::resize / cv::medianBlur / cv::CLAHE / cv::equalizeHist;
OpenCV itself is a third-party dependency that does not travel to the GPU box, so its outputs on small seeded
images are committed here (IPP disabled: the portable C++ code paths).

    python tests/golden/make_preprocess_golden.py   ->  tests/golden/preprocess_golden.npz
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# (name, height, width, resizeFactor, useMedianFilter, histogramMethod, claheClipLimit, claheTilesGridSize)
CASES = [
    ("stereorig_v1", 120, 160, 0.5, 0, "CLAHE", 1.0, 2),          # config_stereorig_v1.yaml:103-106
    ("flir_gv7", 120, 160, 0.5, 0, "CLAHE", 2.0, 4),              # config_flir_stereo_gv7.yaml:117-120
    ("stereorig_v2", 120, 160, 0.5, 0, "NONE", 1.0, 4),           # config_stereorig_v2.yaml:123-126
    ("resize_0p8_median_hist", 150, 200, 0.8, 1, "HISTOGRAM", 1.0, 4),   # the 0.8 example of the yaml comment
    ("odd_size_clahe_pad", 97, 131, 0.5, 1, "CLAHE", 3.0, 4),     # odd source, tiles do not divide the image
    ("upscale_hist", 60, 80, 1.5, 0, "HISTOGRAM", 1.0, 4),
    ("noresize_clahe8", 96, 128, 1.0, 0, "CLAHE", 40.0, 8),
    ("constant_hist", 48, 64, 1.0, 0, "HISTOGRAM", 1.0, 4),
    ("quarter", 128, 160, 0.25, 0, "NONE", 1.0, 4),
]


def make_image(rng, h, w, name):
    if name.startswith("constant"):
        return np.full((h, w), 77, dtype=np.uint8)
    y, x = np.mgrid[0:h, 0:w]
    img = 110 + 70 * np.sin(x / 9.0) * np.cos(y / 7.0) + rng.normal(0, 18, (h, w))
    img[h // 3:h // 2, w // 4:w // 2] += 60          # a bright patch: clipped histogram bins
    return img.clip(0, 255).astype(np.uint8)


def reference_chain(raw, factor, median, method, clip, tiles):
    """Subscriber::imageCallback, cv2 calls in the reference's order."""
    img = cv2.resize(raw, None, fx=factor, fy=factor) if factor != 1.0 else raw.copy()
    if median:
        img = cv2.medianBlur(img, 3)
    if method == "CLAHE":
        img = cv2.createCLAHE(clipLimit=clip, tileGridSize=(tiles, tiles)).apply(img)
    elif method == "HISTOGRAM":
        img = cv2.equalizeHist(img)
    return img


def main():
    cv2.ipp.setUseIPP(False)
    cv2.setNumThreads(1)
    rng = np.random.default_rng(20260925)
    out = {"cv2_version": np.array(cv2.__version__)}
    for name, h, w, f, med, method, clip, tiles in CASES:
        raw = make_image(rng, h, w, name)
        out[name + "/raw"] = raw
        out[name + "/out"] = reference_chain(raw, f, med, method, clip, tiles)
        out[name + "/resized"] = cv2.resize(raw, None, fx=f, fy=f) if f != 1.0 else raw.copy()
    np.savez_compressed(os.path.join(HERE, "preprocess_golden.npz"), **out)
    print("wrote", len(CASES), "cases, cv2", cv2.__version__)


if __name__ == "__main__":
    main()
