"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of total time).

  python tools/summarize_launches.py gpurun_out/launches.csv [-o profiles/xyz.txt] [--header "text"]
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    out = None
    header = ""
    if "-o" in sys.argv:
        out = sys.argv[sys.argv.index("-o") + 1]
    if "--header" in sys.argv:
        header = sys.argv[sys.argv.index("--header") + 1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        unit = d.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        rows.append((d["Kernel Name"], us))
    agg = OrderedDict()
    for name, us in rows:
        name = name.replace("<unnamed>::", "")
        name = re.sub(r"<.*", "", name)
        name = re.sub(r"\(.*", "", name)
        name = name.replace("void ", "")
        if name.startswith("at::native::") or name.startswith("at::"):
            name = "at:: (torch plumbing: fill/copy/cat)"
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    lines = []
    if header:
        lines += ["# " + l for l in header.split("\\n")]
    lines.append("# launches %d, total %.1f ms" % (len(rows), tot / 1e3))
    lines.append("%-48s %9s %12s %7s %10s" % ("kernel", "launches", "total_us", "share", "avg_us"))
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-48s %9d %12.1f %6.1f%% %10.1f" % (name[:48], n, us, 100 * us / tot, us / n))
    txt = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(txt)
    sys.stdout.write(txt)


if __name__ == "__main__":
    main()
