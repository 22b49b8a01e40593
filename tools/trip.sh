set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r1y_bench.json 2> gpurun_out/r1y_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r1y_bench.err
