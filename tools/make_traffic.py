#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` capture: DRAM bytes (read + write) per launch of each kernel
family of one solver slot.  The schur family is the sum over its chunk kernels (one launch each per slot).

usage: make_traffic.py report.ncu-rep windows out.json [summary.csv]"""
import csv
import json
import subprocess
import sys

FAMILY = {"k_schur_lr": "schur", "k_schur_wr": "schur", "k_schur_mma": "schur", "k_schur": "schur",
          "k_linearize": "linearize", "k_backsub": "backsub", "k_dense_solve_reg": "dense_solve",
          "k_dense_solve_smem": "dense_solve", "k_dense_eval": "dense_eval", "k_dense_gram_mma": "dense_eval"}
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
USEC = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def main(rep, windows, out, summary=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-kernel-base", "demangled"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    iname, ird, iwr, it = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                                  "gpu__time_duration.sum"))
    per_kernel = {}
    for r in rows[2:]:
        name = r[iname].split("(")[0].replace("void ", "").replace("svin::", "")
        if "<" in r[iname].split("(")[0]:
            name = r[iname][:r[iname].index(">") + 1].replace("void ", "").replace("svin::", "").replace("(int)", "")
        d = per_kernel.setdefault(name, {"n": 0, "bytes": 0.0, "us": 0.0})
        d["n"] += 1
        d["bytes"] += float(r[ird]) * BYTES[units[ird]] + float(r[iwr]) * BYTES[units[iwr]]
        d["us"] += float(r[it]) * USEC[units[it]]
    fam = {}
    for name, d in per_kernel.items():
        f = FAMILY.get(name.split("<")[0])
        if f is None:
            continue
        e = fam.setdefault(f, {"dram_bytes_per_launch": 0.0, "us_per_launch_under_ncu": 0.0, "kernels": {}})
        e["dram_bytes_per_launch"] += d["bytes"] / d["n"]
        e["us_per_launch_under_ncu"] += d["us"] / d["n"]
        e["kernels"][name] = {"launches_captured": d["n"], "dram_bytes": d["bytes"] / d["n"], "us": d["us"] / d["n"]}
    doc = {"source": summary or rep, "windows": int(windows),
           "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none; "
                   "durations under ncu are cold-cache and serialised (shares only)",
           "kernels": fam}
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps({k: round(v["dram_bytes_per_launch"] / 1e6, 1) for k, v in fam.items()}))


if __name__ == "__main__":
    main(*sys.argv[1:])
