#!/usr/bin/env python
"""Reads an .ncu-rep (ncu --set full) and writes (a) a markdown table of the per-launch metrics the docs quote and
(b) profiles/k_hist_traffic.json: mean dram__bytes_read.sum + dram__bytes_write.sum per k_hist launch, which bench.py
reports as roofline.traffic.   usage: tools/ncu_summary.py <report.ncu-rep> <workload key> [out.md]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    ("gpu__time_duration.sum", "ms", 1e-6), ("dram__bytes_read.sum", "DRAM read MB", 1e-6), ("dram__bytes_write.sum", "DRAM write MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % peak", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %", 1),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smem/LSU wavefronts % peak", 1),
    ("smsp__inst_executed_op_shared_atom.sum", "smem atomics", 1),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "atomic wavefronts", 1),
    ("launch__registers_per_thread", "regs", 1), ("launch__shared_mem_per_block_dynamic", "dyn smem KB", 1e-3),
    ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1),
]


def main():
    rep, key = sys.argv[1], sys.argv[2]
    out_md = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    name_i = col["Kernel Name"]

    def val(r, m, scale):
        x = float(r[col[m]].replace(",", "")) * scale
        u = units[col[m]]
        if m == "gpu__time_duration.sum":   # ncu prints ns / us / ms depending on the magnitude
            x = float(r[col[m]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        if m.startswith("dram__bytes"):
            x = float(r[col[m]].replace(",", "")) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
        if m == "launch__shared_mem_per_block_dynamic":
            x = float(r[col[m]].replace(",", "")) * (1.0 if u.startswith("Kbyte") else 1e3 if u.startswith("Mbyte") else 1e-3)
        return x

    lines = ["| # | kernel | " + " | ".join(t for _, t, _ in METRICS) + " |", "|---|---|" + "---|" * len(METRICS)]
    hist_bytes = []
    for i, r in enumerate(data):
        name = r[name_i].split("(")[0].replace("void ", "").replace("ygg::", "")
        vals = [val(r, m, s) for m, _, s in METRICS]
        lines.append(f"| {i} | `{name}` | " + " | ".join(f"{v:.4g}" for v in vals) + " |")
        if name.startswith("k_hist<"):
            hist_bytes.append((vals[1] + vals[2]) * 1e6)
    text = "\n".join(lines)
    print(text)
    if out_md:
        open(out_md, "a").write(text + "\n")
    if hist_bytes:
        p = os.path.join(ROOT, "profiles", "k_hist_traffic.json")
        table = json.load(open(p)) if os.path.exists(p) else {}
        table[key] = {"dram_bytes_per_launch": sum(hist_bytes) / len(hist_bytes), "launches": len(hist_bytes),
                      "per_launch": [round(b) for b in hist_bytes], "source": os.path.basename(rep),
                      "metric": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full --clock-control none"}
        json.dump(table, open(p, "w"), indent=1)


if __name__ == "__main__":
    main()
