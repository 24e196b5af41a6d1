"""Per-source-line hot spots of an ncu report: instructions executed and stall samples.

  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; python tools/ncu_source_hot.py x.csv [file-substring] [top]
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    cur_file, hdr = None, None
    agg = defaultdict(lambda: [0, 0, ""])     # (file, line) -> [inst, samples, text]
    stalls = defaultdict(lambda: defaultdict(int))
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Name":
            cur_file = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] == "Kernel Name":
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
            continue
        d = dict(zip(hdr, r))
        # the CUDA-C view repeats "Source" twice (source text, then SASS); take per-line aggregates only
        try:
            inst = int(d.get("Instructions Executed", "0") or 0)
            samp = int(d.get("# Samples", "0") or 0)
        except ValueError:
            continue
        key = (cur_file, int(r[0]))
        a = agg[key]
        a[0] += inst
        a[1] += samp
        a[2] = r[1][:110]
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k:
                try:
                    stalls[key][k] += int(v or 0)
                except ValueError:
                    pass
    rows = [(k, v) for k, v in agg.items() if want in (k[0] or "")]
    tot_i = sum(v[0] for _, v in rows) or 1
    tot_s = sum(v[1] for _, v in rows) or 1
    print("total inst %d, samples %d" % (tot_i, tot_s))
    for k, v in sorted(rows, key=lambda kv: -kv[1][1])[:top]:
        st = sorted(stalls[k].items(), key=lambda kv: -kv[1])[:3]
        print("%-28s:%4d  inst %5.1f%%  samples %5.1f%%  %s | %s" % (
            (k[0] or "")[-28:], k[1], 100.0 * v[0] / tot_i, 100.0 * v[1] / tot_s,
            " ".join("%s=%d" % (a.replace("stall_", ""), b) for a, b in st), v[2].strip()))


if __name__ == "__main__":
    main()
