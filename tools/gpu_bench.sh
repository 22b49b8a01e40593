#!/bin/bash
# sequence test + the default bench line (what the driver runs) + the reference arm
export TAG=${1:-r2d}
python -m pytest tests/test_sequence.py -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
tail -15 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err | head -2
python - <<'PY'
import json,os
d=json.load(open("gpurun_out/%s_bench.json"%os.environ["TAG"]))
for k in ("value","ms_per_step","device_ms_per_step","gpu_launches","clocks"): print(k, d.get(k))
print("e2e", {k:v for k,v in d["e2e"].items() if k in ("value","ms_per_step","h2d_bytes_per_step")}, d["e2e"]["serial"])
print("roofline", d["roofline"])
print("kernels", {k:(round(v["ms_total"],3),v["launches"]) for k,v in d["kernels"].items()})
print("latency", json.dumps(d.get("latency"), indent=1))
print("marg", d.get("marginalization"))
print("sharded", d.get("sharded"))
print("cpu", d.get("cpu_baseline"))
fe=d.get("frontend",{})
print("fe", fe.get("value"), fe.get("e2e",{}).get("value"), fe.get("kernels_ms_per_step"))
PY
