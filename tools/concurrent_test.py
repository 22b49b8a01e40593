"""Experiment: does solving two half batches concurrently (two contexts, two streams) beat one full batch?"""
import sys, time, threading
sys.path.insert(0, '/root/repo')
import numpy as np
import torch
from svin_b200.engine import BaEngine
from svin_b200.window import default_options
import bench
batch = bench.make_batch(256, 8, seed0=20260925)
opt = default_options()
def run(parts, reps=6):
    engs = [BaEngine(0) for _ in parts]
    for e, p in zip(engs, parts):
        e.upload(p)
        for _ in range(2):
            e.reset(); e.solve(opt)
    ts = []
    for _ in range(reps):
        for e in engs: e.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=lambda e=e: e.solve(opt)) for e in engs]
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    for e in engs: e.close()
    return np.round(ts, 2)
print("1 x 256:", run([batch]))
print("2 x 128:", run([batch[:128], batch[128:]]))
print("4 x 64 :", run([batch[i * 64:(i + 1) * 64] for i in range(4)]))
