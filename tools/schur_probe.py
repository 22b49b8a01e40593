"""Diagnostics: per-kernel times of one profiled solve of the bench batch (optionally under SVIN_SCHUR_CLASSMASK).

  python tools/schur_probe.py [--windows 256] [--solves 2]
Prints one JSON line with the per-family ms / launches.  Not a benchmark (profiling events serialise the streams).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=256)
    ap.add_argument("--distinct", type=int, default=8)
    ap.add_argument("--solves", type=int, default=2)
    ap.add_argument("--no-prof", action="store_true")
    args = ap.parse_args()
    from bench import make_batch
    from svin_b200.engine import BaEngine
    from svin_b200.window import default_options
    batch = make_batch(args.windows, args.distinct)
    for w in batch:
        w.c_struct()
    opt = default_options()
    with BaEngine(0) as eng:
        eng.upload(batch)
        eng.solve(opt)
        eng.set_profiling(not args.no_prof)
        for _ in range(args.solves):
            eng.reset()
            summ = eng.solve(opt)
        out = {"mask": os.environ.get("SVIN_SCHUR_CLASSMASK", "7"), "solve_ms": eng.timings()["solve_ms"],
               "iterations": summ[0]["iterations"]}
        if not args.no_prof:
            kt = eng.kernel_times()
            out["kernels"] = {k: [round(v["ms"], 3), v["launches"]] for k, v in kt.items()}
        print(json.dumps(out))


if __name__ == "__main__":
    main()
