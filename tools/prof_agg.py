"""Aggregate a sefd_prof_dump CSV by launch label: python tools/prof_agg.py gpurun_out/fsn_profile.csv [top]"""
import collections
import csv
import sys
rows = list(csv.DictReader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = collections.OrderedDict()
for r in rows:
    k = r['label'] if r['label'] else 'cat' + r['category']
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += float(r['ms']); a[2] += float(r['gflop']); a[3] += float(r['gbyte'])
tot = sum(a[1] for a in agg.values())
print(f"total {tot:.3f} ms over {len(rows)} launches")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{a[1]:9.3f} ms n={a[0]:5d} avg {a[1] / a[0] * 1e3:8.1f} us {a[2] / max(a[1], 1e-9):7.1f} TF/s {a[3] * 1e3 / max(a[1], 1e-9):7.1f} GB/s  {k}")
