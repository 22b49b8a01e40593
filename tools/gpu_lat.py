"""Single-window latency + marginalisation timing probe (GPU box)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from svin_b200.engine import BaEngine
eng = BaEngine(0)
out = {"latency": bench.latency_section(eng, 0), "marg": bench.marginalization_section(eng)}
eng.close()
for k, v in out["latency"].items():
    if isinstance(v, dict):
        print(k, "e2e p50 %.3f ms  device %.3f  upload %.3f  cpu %.1f" % (v["e2e_ms"]["p50"], v["device_solve_ms_p50"], v["host_upload_ms_p50"], v["cpu_1thread_ms_p50"]))
print("marg", json.dumps(out["marg"], indent=1))
