#!/bin/bash
# Round profile (run on the GPU box through gpurun): bench line, ncu launch list of the bench command, one
# `ncu --set full` capture of a solver slot.  usage: tools/profile_round.sh <tag>
set -x
tag=${1:-r1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --frames 16 --cpu-seconds 0.5 --skip-e2e > gpurun_out/${tag}_launches_bench.log 2>&1; echo "launch list rc=$?"
SVIN_BA_GRAPH=0 timeout 900 ncu --set full --import-source on --clock-control none \
  -k regex:"k_schur|k_linearize|k_backsub|k_dense|k_pre" -s 25 -c 12 -o gpurun_out/${tag}_prof -f \
  python tools/schur_probe.py --windows 296 --solves 1 --no-prof > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
