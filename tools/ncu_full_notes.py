#!/usr/bin/env python
"""Key figures of `ncu --set full` reports as a markdown table (one row per captured launch).

    python tools/ncu_full_notes.py a.ncu-rep b.ncu-rep ... > notes.md
"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr")]


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name.split("(")[0][:48]


def main(paths):
    print("| report | kernel | " + " | ".join(h for _, h in WANT) + " | top stalls (cycles per issued instruction) |")
    print("|---|---|" + "---|" * (len(WANT) + 1))
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        stall = [i for i, h in enumerate(hdr) if "smsp__average_warp" in h and "issue_stalled" in h and h.endswith(".ratio")
                 and "not_issued" not in h]
        for r in rows[2:]:
            cells = []
            for key, _ in WANT:
                if key in hdr:
                    i = hdr.index(key)
                    v = r[i]
                    try:
                        f = float(v)
                        v = f"{f:.1f}" if f < 1e4 else f"{f:.3g}"
                    except ValueError:
                        pass
                    u = units[i]
                    cells.append(v + (" " + u if u in ("Mbyte", "Gbyte", "Kbyte", "byte", "us", "ms", "usecond", "msecond", "ns") else ""))
                else:
                    cells.append("-")
            top = sorted(((float(r[i]), hdr[i].split("issue_stalled_")[1].split("_per")[0]) for i in stall if r[i]), reverse=True)[:4]
            tops = ", ".join(f"{n} {v:.2f}" for v, n in top if n != "selected")
            print(f"| {path.split('/')[-1]} | {short(r[hdr.index('Kernel Name')])} | " + " | ".join(cells) + f" | {tops} |")


if __name__ == "__main__":
    main(sys.argv[1:])
