#!/bin/bash
# A/B on the GPU box: BA parity tests, then resident-solve bench fused vs materialised, then single-window latency.
set -x
export TAG=${1:-r2a}
python -m pytest tests/test_ba_gpu.py tests/test_marg_gpu.py -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --skip-e2e --frames 0 --cpu-seconds 1 > gpurun_out/${TAG}_bench_fused.json 2> gpurun_out/${TAG}_bench_fused.err
SVIN_BA_FUSED=0 python bench.py --skip-e2e --frames 0 --cpu-seconds 1 > gpurun_out/${TAG}_bench_mat.json 2> gpurun_out/${TAG}_bench_mat.err
python bench.py --skip-e2e --frames 0 --cpu-seconds 1 --windows 1 --steps 50 > gpurun_out/${TAG}_bench_b1.json 2> gpurun_out/${TAG}_bench_b1.err
python - <<'PY'
import json,sys,os
for n in ("fused","mat","b1"):
    try:
        d=json.load(open(f"gpurun_out/%s_bench_%s.json" % (os.environ["TAG"], n)))
        print(n, round(d["value"],1), d["device_ms_per_step"], {k:(round(v["ms_total"],3),v["launches"]) for k,v in d["kernels"].items()})
    except Exception as e:
        print(n, "ERR", e)
PY
