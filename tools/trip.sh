set -x
mkdir -p gpurun_out
for v in 12 16; do echo "lr warps $v" >> gpurun_out/t32_probe.log; SVIN_LR_WARPS=$v timeout 300 python tools/schur_probe.py >> gpurun_out/t32_probe.log 2>&1; SVIN_LR_WARPS=$v timeout 300 python tools/schur_probe.py --no-prof --solves 3 >> gpurun_out/t32_probe.log 2>&1; done
cat gpurun_out/t32_probe.log
