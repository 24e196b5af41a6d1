"""Per-variant duration and DRAM traffic of the tcgen05 GEMM launches of a training step, from

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:gemm_bf16 --csv --log-file X.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph

  python tools/summarize_gemm_traffic.py X.csv profiles/NAME.json

`mean_dram_bytes_per_launch` is what bench.py reports as roofline.traffic (the GEMM is the dominant kernel)."""
import collections
import csv
import json
import re
import sys

MULT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    lines = [l for l in open(src) if not l.startswith("==")]
    hdr, rows = None, collections.OrderedDict()
    for r in csv.reader(lines):
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        rec = rows.setdefault(d["ID"], {"name": d["Kernel Name"], "grid": d["Grid Size"]})
        rec[d["Metric Name"]] = float(d["Metric Value"].replace(",", "")) * MULT.get(d["Metric Unit"], 1.0)
    agg = collections.OrderedDict()
    for rec in rows.values():
        nm = re.search(r"(gemm_bf16_tc2?_kernel<[^>]*>)", rec["name"]).group(1)
        a = agg.setdefault((nm, rec["grid"]), [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += rec["gpu__time_duration.sum"]
        a[2] += rec["dram__bytes_read.sum"]
        a[3] += rec["dram__bytes_write.sum"]
    n = sum(a[0] for a in agg.values())
    tot_us = sum(a[1] for a in agg.values())
    tot_b = sum(a[2] + a[3] for a in agg.values())
    out = {"source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, --clock-control none, eager bench step "
                     "(B=256, L=128, P=16, bf16); template args = <BN, A_MN, B_MN, EPI>",
           "gemm_launches": n, "total_us": tot_us, "mean_dram_bytes_per_launch": tot_b / n,
           "variants": [{"kernel": k[0], "grid": k[1], "launches": a[0], "avg_us": a[1] / a[0],
                         "share_of_gemm_time": a[1] / tot_us, "dram_read_MB": a[2] / a[0] / 1e6,
                         "dram_write_MB": a[3] / a[0] / 1e6}
                        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "variants"}))


if __name__ == "__main__":
    main()
