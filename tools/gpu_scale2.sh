#!/bin/bash
# 2-GPU run exactly as the driver launches it (torchrun, one rank per GPU)
export TAG=${1:-r2s}
export N=${2:-2}
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
echo "rc=$?"
tail -5 gpurun_out/${TAG}_bench_${N}gpu.err
python - <<'PY'
import json,os
d=json.load(open("gpurun_out/%s_bench_%sgpu.json"%(os.environ["TAG"], os.environ.get("N","2"))))
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "dev ms", d["device_ms_per_step"])
print("loop", d.get("loop_closure"))
print("sharded", json.dumps(d.get("sharded"), indent=1))
PY

