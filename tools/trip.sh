set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pre_gpu.py -m gpu -x -q > gpurun_out/t17_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/t17_pytest.log
