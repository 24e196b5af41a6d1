#!/usr/bin/env python
"""Per-kernel SASS evidence that the hot kernels are Blackwell-native: counts of UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG
(TMA tensor load / store), LDTM / STTM (tcgen05.ld / st, TMEM) and UTCBAR (tcgen05.commit) in the built library.
    python tools/sass_counts.py > profiles/sass_counts.txt        (needs cuobjdump; no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mtvaf_b200", "_C", "libmtvaf_b200.so")
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
        elif cur:
            for k in KEYS:
                if k in line:
                    counts[cur][k] += 1
    rows, tot = [], collections.Counter()
    for fn, c in counts.items():
        if not any(c[k] for k in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM")):
            continue
        name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip() or fn
        name = name.replace("(anonymous namespace)::", "")
        rows.append((re.sub(r"\(.*", "", name)[:110], c))
        tot.update(c)
    rows.sort(key=lambda r: -r[1]["UTCHMMA"])
    print("# SASS instruction counts per kernel of mtvaf_b200/_C/libmtvaf_b200.so (cuobjdump -sass, sm_100a)")
    print("# UTCHMMA = tcgen05.mma (kind::f16), UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM = tcgen05.ld / st "
          "(TMEM), UTCBAR = tcgen05.commit")
    print("# regenerate: python tools/sass_counts.py > profiles/sass_counts.txt\n")
    print("%-112s %8s %8s %8s %6s %6s %7s" % ("kernel", *KEYS))
    for name, c in rows:
        print("%-112s %8d %8d %8d %6d %6d %7d" % (name, *(c[k] for k in KEYS)))
    print("\nTOTAL over %d tcgen05/TMA kernels: " % len(rows) + ", ".join("%s %d" % (k, tot[k]) for k in KEYS))


if __name__ == "__main__":
    sys.exit(main())
