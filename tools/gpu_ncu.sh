#!/bin/bash
# ncu evidence for the round: launch list of the default solve + one --set full capture of the per-observation kernels.
# Numbers printed by bench.py under ncu are never bench values.
export TAG=${1:-r2p}
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --skip-e2e --frames 0 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/${TAG}_launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt
head -30 gpurun_out/${TAG}_launches_summary.txt
ncu --set full --clock-control none --import-source on -k regex:'k_schur|k_linearize|k_backsub' -s 64 -c 10 -f -o gpurun_out/${TAG}_prof \
  python bench.py --skip-e2e --frames 0 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out/${TAG}_prof.ncu-rep
