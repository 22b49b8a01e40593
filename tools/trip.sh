set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/t23_bench_2gpu.json 2> gpurun_out/t23_bench_2gpu.err; echo "bench2 rc=$?"; tail -5 gpurun_out/t23_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/t23_ref_2gpu.json 2>> gpurun_out/t23_bench_2gpu.err; echo "ref2 rc=$?"
timeout 600 python -m pytest tests/test_sharding.py -m gpu -x -q 2>&1 | tail -3
