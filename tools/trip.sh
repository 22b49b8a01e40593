set -x
mkdir -p gpurun_out
timeout 300 python tools/schur_probe.py >> gpurun_out/t31_probe.log 2>&1
timeout 300 python tools/schur_probe.py --no-prof --solves 3 >> gpurun_out/t31_probe.log 2>&1
cat gpurun_out/t31_probe.log
