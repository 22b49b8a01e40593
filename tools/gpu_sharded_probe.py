"""Runs only bench.sharded_section under torchrun (probe for the N > 1 path); dumps stacks if it stalls."""
import faulthandler, json, os, sys
faulthandler.dump_traceback_later(70, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import bench
args = argparse.Namespace(gpus=int(os.environ.get("WORLD_SIZE", "1")), steps=5, warmup=3)
world, rank, local, dist = bench.dist_setup(args.gpus)
print(f"rank {rank} setup done", flush=True)
out = bench.sharded_section(args, local, rank, world, dist)
if rank == 0:
    print(json.dumps(out), flush=True)
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
