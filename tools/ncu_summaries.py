#!/usr/bin/env python
"""Summaries of the ncu artefacts under gpurun_out/ for profiles/ (run here, in the build container).

    python tools/ncu_summaries.py launches gpurun_out/launches_r01g.csv gpurun_out/bench_r01g.json > profiles/launch_summary_r01g.md
    python tools/ncu_summaries.py kernels  gpurun_out/prof_r01g.ncu-rep                            > profiles/kernels_r01g.md
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict


def short(name):
    return name.split("(")[0].strip()


def launches(csv_path, bench_path):
    rows = list(csv.reader(l for l in open(csv_path) if l.startswith('"')))
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = OrderedDict()
    for r in rows[1:]:
        try:
            ns = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        n = short(r[kn])
        t = tot.setdefault(n, [0, 0.0])
        t[0] += 1
        t[1] += ns / 1e6
    bench = {}
    if bench_path:
        line = [l for l in open(bench_path) if l.startswith("{")][-1]
        bench = {k + "_kernel": v["share"] for k, v in json.loads(line)["kernels"].items()}
    total = sum(v[1] for v in tot.values())
    print("| kernel | launches | total ms | share (ncu) | share (bench events) |\n|---|---|---|---|---|")
    for n, (c, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        key = n.replace("void ", "").split("<")[0].replace("trmm2_kernel", "trmm_kernel")
        key = {"s3_pass_kernel": "sy2sb_pass_kernel", "s3_panel_kernel": "sy2sb_panel_kernel",
               "backtransform_wy_kernel": "backtransform_kernel"}.get(key, key)
        print(f"| {n[:60]} | {c} | {ms:.3f} | {ms / total:.4f} | {bench.get(key, '')} |")


METRICS = [("time", "gpu__time_duration.sum", 1e-6, "ms"),
           ("DRAM read", "dram__bytes_read.sum", 1e-6, "MB"), ("DRAM write", "dram__bytes_write.sum", 1e-6, "MB"),
           ("DRAM % peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1, ""),
           ("L2 hit %", "lts__t_sector_hit_rate.pct", 1, ""),
           ("FP64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1, ""),
           ("DMMA pipe %", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", 1, ""),
           ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1, ""),
           ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1, "")]


def kernels(rep_path):
    out = subprocess.run(["ncu", "-i", rep_path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    have = [(lab, m, sc, u) for lab, m, sc, u in METRICS if m in idx]
    print("| kernel | grid | block | regs | " + " | ".join(f"{lab}{' ' + u if u else ''}" for lab, _, _, u in have) + " |")
    print("|---|---|---|---|" + "---|" * len(have))
    seen = set()
    for r in rows[2:]:
        n = short(r[idx["Kernel Name"]])
        key = (n, r[idx["Grid Size"]])
        if key in seen:
            continue
        seen.add(key)
        vals = []
        for lab, m, sc, u in have:
            try:
                v = float(r[idx[m]].replace(",", ""))
                unit = units[idx[m]]
                if m.startswith("dram__bytes"):
                    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
                if m.startswith("gpu__time"):
                    v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
                vals.append(f"{v * sc:.4g}")
            except ValueError:
                vals.append(r[idx[m]])
        regs = r[idx["launch__registers_per_thread"]] if "launch__registers_per_thread" in idx else ""
        print(f"| {n[:48]} | {r[idx['Grid Size']]} | {r[idx['Block Size']]} | {regs} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        kernels(sys.argv[2])
