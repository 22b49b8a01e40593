"""GPU parity for the pre-processing chain: CUDA (through the C ABI) vs the OpenCV golden vectors and vs the numpy
oracle on full-size images.  Byte work: bit-exact."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import preprocess_oracle as po  # noqa: E402
from test_oracle_pre import GOLD, METHOD, _cases  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", _cases(), ids=[c[0] for c in _cases()])
def test_cuda_matches_opencv_golden(case):
    from svin_b200.preprocess import Preprocessor
    name, h, w, f, med, method, clip, tiles = case
    raw = GOLD[name + "/raw"]
    with Preprocessor(w, h, resizeFactor=f, useMedianFilter=bool(med), histogramMethod=method, claheClipLimit=clip,
                      claheTilesGridSize=tiles, max_images=3) as pre:
        out = pre.process([raw, raw[::-1].copy(), raw])
        assert (pre.height, pre.width) == GOLD[name + "/out"].shape
        assert np.array_equal(out[0], GOLD[name + "/out"])
        assert np.array_equal(out[2], GOLD[name + "/out"])
        assert np.array_equal(out[1], po.preprocess(raw[::-1].copy(), f, bool(med), METHOD[method], clip, tiles))


@pytest.mark.parametrize("kw", [
    dict(resizeFactor=0.5, histogramMethod="CLAHE", claheClipLimit=1.0, claheTilesGridSize=2),   # stereorig_v1
    dict(resizeFactor=0.5, histogramMethod="CLAHE", claheClipLimit=2.0, claheTilesGridSize=4),   # flir gv7
    dict(resizeFactor=0.5, useMedianFilter=True, histogramMethod="HISTOGRAM"),
    dict(resizeFactor=0.8, histogramMethod="NONE"),
])
def test_full_size_batch_matches_oracle(kw):
    from svin_b200.preprocess import HISTOGRAM_METHODS, Preprocessor
    rng = np.random.default_rng(11)
    y, x = np.mgrid[0:1200, 0:1600]
    imgs = [(120 + 90 * np.sin(x / (31.0 + k)) * np.cos(y / 17.0) + rng.normal(0, 20, x.shape)).clip(0, 255)
            .astype(np.uint8) for k in range(3)]
    imgs.append(rng.integers(0, 256, (1200, 1600), dtype=np.uint8))
    with Preprocessor(1600, 1200, max_images=4, **kw) as pre:
        out = pre.process(imgs)
        # split API: upload / run / download gives the same bytes
        pre.upload(imgs)
        pre.run()
        assert np.array_equal(pre.download(), out)
    for k, img in enumerate(imgs):
        ref = po.preprocess(img, kw.get("resizeFactor", 1.0), kw.get("useMedianFilter", False),
                            HISTOGRAM_METHODS[kw.get("histogramMethod", "NONE")], kw.get("claheClipLimit", 1.0),
                            kw.get("claheTilesGridSize", 4))
        assert np.array_equal(out[k], ref), k


def test_device_handover_to_detector_equals_host_round_trip():
    # pre-processing output stays on the device and feeds svin_fe (svin_fe_upload_device)
    from svin_b200.frontend import FeEngine
    from svin_b200.preprocess import Preprocessor
    from svin_b200.synthetic import EUROC_INTRINSICS
    rng = np.random.default_rng(2)
    raws = [rng.integers(0, 256, (960, 1504), dtype=np.uint8) for _ in range(2)]
    g = np.array([[0.1, 0.99, 0.05], [0.0, 1.0, 0.0]])
    intr = [EUROC_INTRINSICS[0], EUROC_INTRINSICS[1]]
    with Preprocessor(1504, 960, resizeFactor=0.5, histogramMethod="CLAHE", claheClipLimit=2.0, claheTilesGridSize=4,
                      max_images=2) as pre, FeEngine(752, 480, max_images=2) as fe:
        host = pre.process(raws)
        ref = fe.detect_describe([host[0], host[1]], intr, g)
        pre.upload(raws)
        pre.run()
        fe.upload_device(pre.device_output(), 2, intr, g)
        fe.run()
        got = fe.download()
    for (k0, d0), (k1, d1) in zip(ref, got):
        assert len(k0) == len(k1) and np.array_equal(d0, d1) and np.array_equal(k0["x"], k1["x"])


def test_invalid_arguments_are_rejected():
    from svin_b200.capi import SvinError
    from svin_b200.preprocess import Preprocessor
    with pytest.raises(SvinError):
        Preprocessor(4, 4)
    with Preprocessor(64, 48, max_images=1) as pre:
        with pytest.raises(SvinError, match="max_images"):
            pre.process([np.zeros((48, 64), np.uint8)] * 2)
