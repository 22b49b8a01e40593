"""CPU: the numpy restatement of the pre-processing chain (oracle/preprocess_oracle.py) against the golden vectors
generated with the real OpenCV (tests/golden/make_preprocess_golden.py).  Bit-exact."""
import importlib.util
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import preprocess_oracle as po  # noqa: E402

_spec = importlib.util.spec_from_file_location("make_preprocess_golden",
                                               os.path.join(HERE, "golden", "make_preprocess_golden.py"))
CASES = None


def _cases():
    # the case table lives in the generator; read it without importing cv2
    global CASES
    if CASES is None:
        src = open(os.path.join(HERE, "golden", "make_preprocess_golden.py")).read()
        ns = {}
        exec(src[src.index("CASES = ["):src.index("def make_image")], ns)
        CASES = ns["CASES"]
    return CASES


GOLD = np.load(os.path.join(HERE, "golden", "preprocess_golden.npz"))
METHOD = {"NONE": po.HIST_NONE, "HISTOGRAM": po.HIST_EQUALIZE, "CLAHE": po.HIST_CLAHE}


@pytest.mark.parametrize("case", _cases(), ids=[c[0] for c in _cases()])
def test_oracle_reproduces_opencv_bit_exactly(case):
    name, h, w, f, med, method, clip, tiles = case
    raw = GOLD[name + "/raw"]
    assert raw.shape == (h, w)
    if f != 1.0:
        assert np.array_equal(po.resize_linear(raw, f), GOLD[name + "/resized"])
    out = po.preprocess(raw, f, bool(med), METHOD[method], clip, tiles)
    assert out.shape == GOLD[name + "/out"].shape
    assert np.array_equal(out, GOLD[name + "/out"])


def test_oracle_against_live_opencv_when_available():
    # in this container cv2 is importable: full-size images of the shipped configs (1600x1200 -> 800x600)
    cv2 = pytest.importorskip("cv2")
    cv2.ipp.setUseIPP(False)
    rng = np.random.default_rng(3)
    raw = rng.integers(0, 256, (1200, 1600), dtype=np.uint8)
    ref = cv2.createCLAHE(clipLimit=1.0, tileGridSize=(2, 2)).apply(cv2.resize(raw, None, fx=0.5, fy=0.5))
    assert np.array_equal(po.preprocess(raw, 0.5, False, po.HIST_CLAHE, 1.0, 2), ref)


def test_median_and_equalize_edge_cases():
    img = np.zeros((9, 11), dtype=np.uint8)
    img[4, 5] = 255
    assert po.median3(img).max() == 0                      # an isolated pixel disappears
    assert np.array_equal(po.equalize_hist(np.full((5, 7), 9, np.uint8)), np.full((5, 7), 9, np.uint8))
    two = np.zeros((4, 4), np.uint8)
    two[:, 2:] = 200
    eq = po.equalize_hist(two)
    assert set(np.unique(eq)) == {0, 255}
