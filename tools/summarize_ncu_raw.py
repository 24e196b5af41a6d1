#!/usr/bin/env python
"""Markdown table from `ncu -i X.ncu-rep --page raw --csv` of the hot-kernel capture (tools/profile_hot_kernels.py prints
the launch order = the variant names).

  python tools/summarize_ncu_raw.py raw.csv names.txt > profiles/NAME.md"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "M inst"),
        ("smsp__average_warp_latency_issue_stalled_barrier.pct", "stall barrier %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long scoreboard / issue")]
FLOPS = {"qkv_fwd_store": (2304, 768), "ffn1_fwd_gelu": (3072, 768), "attn_out_fwd_resid": (768, 768),
         "ffn2_fwd_resid": (768, 3072), "ffn2_dgrad_dgelu": (3072, 768), "ffn2_dgrad_dgelu_colsum": (3072, 768),
         "ffn1_dgrad_resid": (768, 3072), "ffn1_fwd_gelu_grad": (3072, 768), "ffn2_dgrad_mulaux_colsum": (3072, 768), "attn_out_dgrad_store": (768, 768), "ffn1_wgrad": (3072, 768)}


def num(s, unit_scale=None):
    s = s.replace(",", "")
    try:
        return float(s)
    except ValueError:
        return None


def main():
    raw, names = sys.argv[1], [l.strip() for l in open(sys.argv[2]) if l.strip()]
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
    rows = list(csv.reader(l for l in open(raw) if not l.startswith("==")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    head, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    idx = {h: i for i, h in enumerate(head)}
    print("| kernel (variant) | us | TFLOP/s | " + " | ".join(c[1] for c in COLS[1:]) + " |")
    print("|---" * (len(COLS) + 2) + "|")
    for name, r in zip(names, data):
        vals = []
        us = None
        for key, label in COLS:
            if key not in idx:
                vals.append("-")
                continue
            v, u = num(r[idx[key]]), units[idx[key]]
            if v is None:
                vals.append("-")
                continue
            if key == "gpu__time_duration.sum":
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
                us = v
            if key.startswith("dram__bytes"):
                v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            if key == "smsp__inst_executed.sum":
                v = v / 1e6
            vals.append("%.1f" % v)
        tf = "-"
        if name in FLOPS and us:
            n, k = FLOPS[name]
            tf = "%.0f" % (2.0 * T * n * k / us / 1e6)
        print("| %s | %s | %s | %s |" % (name, vals[0], tf, " | ".join(vals[1:])))


if __name__ == "__main__":
    main()
