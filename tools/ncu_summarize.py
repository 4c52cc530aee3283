#!/usr/bin/env python
"""Summarise an `ncu --csv` log (long format: one row per launch and metric) per kernel:
launches, total / share of gpu__time_duration, DRAM bytes per launch, tensor-pipe activity.

    python tools/ncu_summarize.py gpurun_out/launches.csv [--md profiles/x.md] [--traffic-json profiles/ncu_traffic.json]
"""
import argparse, collections, csv, io, json, re, sys

UNIT = {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def short(name):
    name = name.strip()
    if name.startswith("void "):
        name = name[5:]
    name = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    m = re.match(r"([^(<]*)(<[^(]*>)?", name)          # qualified name, optional template arguments
    base = m.group(1).split("::")[-1].strip() if m else name
    targs = (m.group(2) or "") if m else ""
    if base in ("operator", ""):
        return name[:40]
    return base + (targs if 0 < len(targs) <= 14 else "")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--md")
    ap.add_argument("--traffic-json")
    ap.add_argument("--title", default="ncu launch list")
    ap.add_argument("--source", default="")
    a = ap.parse_args()
    lines = [l for l in open(a.csv, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = collections.OrderedDict()          # launch id -> {kernel, metrics}
    for r in rows:
        d = per.setdefault(r["ID"], {"kernel": short(r["Kernel Name"]), "m": {}})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        name, unit = r["Metric Name"], r["Metric Unit"]
        if name == "gpu__time_duration.sum":
            v *= UNIT.get(unit, 1e-6)        # -> ms
        elif "bytes" in name:
            v *= BYTES.get(unit, 1.0)
        d["m"][name] = v
    agg = collections.OrderedDict()
    for d in per.values():
        g = agg.setdefault(d["kernel"], {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0, "tensor": 0.0, "nt": 0})
        g["n"] += 1
        g["ms"] += d["m"].get("gpu__time_duration.sum", 0.0)
        g["rd"] += d["m"].get("dram__bytes_read.sum", 0.0)
        g["wr"] += d["m"].get("dram__bytes_write.sum", 0.0)
        for k, v in d["m"].items():
            if k.startswith("sm__pipe_tensor"):
                g["tensor"] += v
                g["nt"] += 1
    tot = sum(g["ms"] for g in agg.values()) or 1.0
    out = [f"# {a.title}", "", a.source, "", "| kernel | launches | total ms | share | avg us | DRAM MB / launch (rd + wr) | tensor pipe % |",
           "|---|---|---|---|---|---|---|"]
    for k, g in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        dram = f"{g['rd'] / g['n'] / 1e6:.1f} + {g['wr'] / g['n'] / 1e6:.1f}" if g["rd"] + g["wr"] > 0 else "-"
        tp = f"{g['tensor'] / g['nt']:.1f}" if g["nt"] else "-"
        out.append(f"| {k} | {g['n']} | {g['ms']:.2f} | {100 * g['ms'] / tot:.1f}% | {1e3 * g['ms'] / g['n']:.1f} | {dram} | {tp} |")
    text = "\n".join(out) + "\n"
    print(text)
    if a.md:
        open(a.md, "w").write(text)
    if a.traffic_json:
        tj = {}
        for k, g in agg.items():
            if g["rd"] + g["wr"] > 0:
                tj[k] = {"dram_bytes_per_launch": round((g["rd"] + g["wr"]) / g["n"]), "launches": g["n"],
                         "avg_us": round(1e3 * g["ms"] / g["n"], 2), "source": a.source}
        json.dump(tj, open(a.traffic_json, "w"), indent=1)


if __name__ == "__main__":
    main()
