#!/usr/bin/env python
"""Rank CUDA source lines of one kernel by executed warp instructions / stall samples.
usage: ncu_source_hot.py report.ncu-rep kernel_name [top]"""
import csv
import subprocess
import sys


def main(rep, kernel, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "demangled", "--kernel-name", kernel, "--launch-count", "1",
                          "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, agg = None, None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
        elif hdr and r[0].isdigit() and r[2] == "-":      # a source line row (its SASS rows follow)
            key = (cur_file, int(r[0]))
            a = agg.setdefault(key, [r[1], 0, 0])
            a[1] += int(r[iI] or 0)
            a[2] += int(r[iS] or 0)
    tot_i = sum(a[1] for a in agg.values()) or 1
    tot_s = sum(a[2] for a in agg.values()) or 1
    print(f"# {kernel}: {tot_i} warp instructions, {tot_s} stall samples")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{f}:{ln:5d} inst {100 * a[1] / tot_i:5.1f}%  samples {100 * a[2] / tot_s:5.1f}%  {a[0].strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
