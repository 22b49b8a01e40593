set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_marg_gpu.py -m gpu -x -q > gpurun_out/t30_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t30_pytest.log
timeout 300 python tools/schur_probe.py --no-prof --solves 3 >> gpurun_out/t30_probe.log 2>&1
timeout 300 python tools/schur_probe.py --no-prof --solves 3 --windows 296 >> gpurun_out/t30_probe.log 2>&1
cat gpurun_out/t30_probe.log
