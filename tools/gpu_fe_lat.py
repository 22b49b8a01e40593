"""Per-image front-end latency probe: one 752x480 image per call (kernel times from the engine's events)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from svin_b200.frontend import FeEngine
from svin_b200.synthetic import EUROC_INTRINSICS
from svin_b200.synthetic_images import make_stereo_sequence, random_image
seq = make_stereo_sequence(seed=20260925, n_frames=1)
g = np.array([0.1, 0.99, 0.05])
for name, img in (("rendered", seq["images"][0][0]), ("noise", random_image(11))):
    with FeEngine(752, 480, max_images=1) as fe:
        for _ in range(5):
            fe.detect_describe([img], [EUROC_INTRINSICS[0]], [g])
        ts = []
        for _ in range(50):
            t0 = time.perf_counter()
            out = fe.detect_describe([img], [EUROC_INTRINSICS[0]], [g])
            ts.append(1e3 * (time.perf_counter() - t0))
        t = fe.timings()
        print(name, "kp", len(out[0][0]), "e2e p50 %.3f ms" % np.median(ts), "device run %.3f ms" % t["run_ms"],
              {k: round(v, 4) for k, v in t["kernel_ms"].items()})
