import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from svin_b200.engine import BaEngine
from svin_b200.window import default_options
import bench
batch = bench.make_batch(256, 8, seed0=20260925)
opt = default_options()
with BaEngine(0) as eng:
    eng.upload(batch)
    for _ in range(3):
        eng.reset(); eng.solve(opt)
    ts = []
    for _ in range(5):
        eng.reset(); eng.solve(opt); ts.append(eng.timings()["solve_ms"])
    print("NO_FORK" if os.environ.get("SVIN_BA_NO_FORK") else "FORK", np.round(ts, 3))
