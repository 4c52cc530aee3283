#!/usr/bin/env python
"""Opcode histogram of an `ncu --page source --csv` export: executed warp instructions and stall samples per opcode.

    ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > src.csv; python tools/sass_hist.py src.csv
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[h]
    i_s, i_e, i_w = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    ops, stall, total = collections.Counter(), collections.Counter(), 0
    for r in rows[h + 1:]:
        if len(r) <= i_e or not r[i_e].isdigit():
            continue
        parts = r[i_s].split()
        if not parts:
            continue
        op = (parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]).split(".")[0]
        n = int(r[i_e])
        ops[op] += n
        stall[op] += int(r[i_w]) if r[i_w].isdigit() else 0
        total += n
    print("warp instructions executed:", total, " static SASS lines:", len(rows) - h - 1)
    for k, v in ops.most_common(24):
        print(f"{k:10s} {v:12d} {100.0 * v / total:5.1f}%   stall samples {stall[k]}")


if __name__ == "__main__":
    main(sys.argv[1])
