#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS instructions in libsefd.so (cuobjdump -sass): the evidence file
profiles/r2_sass_tc.txt.

    python tools/sass_evidence.py > profiles/r2_sass_tc.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dnn-based-speech-enhancement-in-the-frequency-domain_b200", "sefd", "libsefd.so")
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|LDTM(?:\.x\d+)?|UTCBAR(?:\.MULTICAST)?|UTMALDG\.\dD(?:\.MULTICAST)?|UTMASTG\.\dD|UBLKCP\.S\.G|"
                 r"UCGABAR_ARV|UCGABAR_WAIT|STAS\.\d+|HMMA\.\d+)\b")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur:
            for tok in PAT.findall(line):
                per[cur][tok] += 1
    names = demangle(list(per.keys()))
    print("SASS evidence of the Blackwell-native paths in libsefd.so (cuobjdump -sass, sm_100a), per kernel: instruction x count")
    print("  UTCHMMA = tcgen05.mma (kind::tf32 here), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit (.MULTICAST: to the cluster),")
    print("  UTMALDG = cp.async.bulk.tensor (TMA; .MULTICAST: into every CTA of the cluster), UBLKCP = cp.async.bulk, UCGABAR = barrier.cluster,")
    print("  STAS = st.async.shared::cluster; HMMA (legacy mma.sync) would be listed if any path used it\n")
    for k, c in per.items():
        if not c:
            continue
        n = names.get(k, k)
        n = re.sub(r"^void ", "", n)
        n = n.replace("(anonymous namespace)::", "").split("(")[0]
        print(f"{n:44s} " + ", ".join(f"{t} x{v}" for t, v in sorted(c.items())))
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print("\ntotal: " + ", ".join(f"{t} x{v}" for t, v in sorted(total.items())))


if __name__ == "__main__":
    main()
