set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t36_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t36_pytest.log
timeout 900 python bench.py --frames 0 > gpurun_out/t36_bench.json 2> gpurun_out/t36_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/t36_bench.err
