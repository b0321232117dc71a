#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) and/or a launch-list CSV into the small summaries kept under profiles/.

    python scripts/ncu_summary.py --rep gpurun_out/prof.ncu-rep --launches gpurun_out/launches.csv --out profiles/r01
"""
import argparse
import collections
import csv
import io
import json
import subprocess

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
]


def rep_summary(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]]}
        for k in KEYS:
            if k in idx:
                d[k] = f"{r[idx[k]]} {units[idx[k]]}".strip()
        res.append(d)
    return res


def launch_summary(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    return [{"kernel": k, "launches": a[0], "total_ms": round(a[1], 4), "share_pct": round(100 * a[1] / tot, 2)}
            for k, a in agg.items()]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    ap.add_argument("--out", required=True)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    doc = {"note": a.note}
    if a.launches:
        doc["launch_list"] = launch_summary(a.launches)
    if a.rep:
        doc["kernels"] = rep_summary(a.rep)
    with open(a.out + ".json", "w") as f:
        json.dump(doc, f, indent=1)
    with open(a.out + ".md", "w") as f:
        f.write(f"# ncu summary\n\n{a.note}\n\n")
        if a.launches:
            f.write("## Launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n\n")
            f.write("| kernel | launches | total ms | share % |\n|---|---:|---:|---:|\n")
            for r in doc["launch_list"]:
                f.write(f"| `{r['kernel']}` | {r['launches']} | {r['total_ms']} | {r['share_pct']} |\n")
        if a.rep:
            f.write("\n## Per-kernel metrics (ncu --set full)\n")
            for k in doc["kernels"]:
                f.write(f"\n### `{k['kernel'][:90]}`\n\n")
                for key in KEYS:
                    if key in k:
                        f.write(f"- {key}: {k[key]}\n")
    print("wrote", a.out + ".json", a.out + ".md")


if __name__ == "__main__":
    main()
