set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pre_gpu.py -m gpu -x -q > gpurun_out/t20_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/t20_pytest.log
python - <<'PY'
import sys, json, argparse
sys.path.insert(0, '.')
import bench
args = argparse.Namespace(frames=64, steps=5)
print(json.dumps(bench.run_preprocess(args, 0)))
PY
