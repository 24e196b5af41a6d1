"""Hot SASS instructions of one kernel from `ncu -i X.ncu-rep --page source --csv [--launch-skip i --launch-count 1]`:
top instructions by stall samples with their dominant stall reasons, plus totals per stall reason.

  python tools/ncu_sass_hot.py src.csv [top]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rd = csv.reader(open(path))
    hdr = None
    rows = []
    for r in rd:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 2 or not r[0].startswith("0x"):
            continue
        rows.append(dict(zip(hdr, r)))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = defaultdict(int)
    ts = ti = 0
    for i, d in enumerate(rows):
        d["_i"] = i
        d["_s"] = int(d["# Samples"] or 0)
        d["_n"] = int(d["Instructions Executed"] or 0)
        ts += d["_s"]
        ti += d["_n"]
        for c in stall_cols:
            tot[c] += int(d[c] or 0)
    print("instructions %d, samples %d, SASS lines %d" % (ti, ts, len(rows)))
    print("stall totals: " + "  ".join("%s=%.1f%%" % (k.replace("stall_", ""), 100.0 * v / max(1, ts))
                                       for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
    for d in sorted(rows, key=lambda d: -d["_s"])[:top]:
        st = sorted(((c, int(d[c] or 0)) for c in stall_cols), key=lambda kv: -kv[1])[:2]
        print("#%5d  samples %5.2f%%  inst %5.2f%%  %-28s %s" % (
            d["_i"], 100.0 * d["_s"] / max(1, ts), 100.0 * d["_n"] / max(1, ti),
            " ".join("%s=%d" % (a.replace("stall_", ""), b) for a, b in st if b), d["Source"].strip()[:90]))


if __name__ == "__main__":
    main()
