#!/bin/bash
# what the driver runs at round end, in its order: GPU tests, smoke, reference arm, bench
export TAG=${1:-r2final}
python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -4 gpurun_out/${TAG}_smoke.log
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<'PY'
import json,os
t=os.environ["TAG"]
r=json.load(open(f"gpurun_out/{t}_bench_reference.json")); print("reference", r["value"], r["cpu_baseline"]["cores"], r.get("frontend",{}).get("value"))
d=json.load(open(f"gpurun_out/{t}_bench.json"))
for k in ("value","ms_per_step","device_ms_per_step","gpu_launches","clocks"): print(k, d.get(k))
print("e2e", {k:v for k,v in d["e2e"].items() if k in ("value","ms_per_step","h2d_bytes_per_step")}, d["e2e"]["serial"])
print("roofline", {k:v for k,v in d["roofline"].items() if k!="note"})
print("kernels", {k:(round(v["ms_total"],3),v["launches"],round(v.get("gbs",0))) for k,v in d["kernels"].items()})
for k,v in d.get("latency",{}).items():
    if isinstance(v,dict): print("latency",k,v["e2e_ms"],v["device_solve_ms_p50"],v["cpu_1thread_ms_p50"],v["speedup_vs_cpu_1thread"])
print("marg", d.get("marginalization"))
print("sharded", d.get("sharded"))
print("cpu", d.get("cpu_baseline"))
fe=d.get("frontend",{})
print("fe", fe.get("value"), fe.get("e2e",{}).get("value"), fe.get("kernels_ms_per_step"), fe.get("cpu_baseline"))
print("loop", d.get("loop_closure"))
print("pre", {k:v for k,v in d.get("preprocess",{}).items() if k in ("value","e2e")})
PY
