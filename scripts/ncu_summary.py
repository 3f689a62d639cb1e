#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` report of one kernel for profiles/: launch/occupancy/throughput
metrics, per-opcode instruction / shared-wavefront / global-request counts per unit of work, stall reasons, and the
hottest SASS lines.  Usage: python scripts/ncu_summary.py report.ncu-rep [--units N] [--top 20]
(--units: solved rows per launch, to print per-row numbers; default 1e6)."""
import argparse
import csv
import io
import subprocess
from collections import defaultdict

KEEP = ("Duration", "Executed Ipc Active", "Issue Slots Busy", "No Eligible", "Active Warps Per Scheduler",
        "Eligible Warps Per Scheduler", "Achieved Occupancy", "Theoretical Occupancy", "Executed Instructions",
        "Memory Throughput", "L2 Hit Rate", "L1/TEX Hit Rate", "Block Limit Registers", "Block Limit Shared Mem",
        "DRAM Throughput", "Registers Per Thread", "Dynamic Shared Memory Per Block", "Compute (SM) Throughput",
        "Grid Size", "Block Size")


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--units", type=float, default=1e6)
    ap.add_argument("--top", type=int, default=20)
    a = ap.parse_args()
    rows = page(a.report, "details")
    if rows:
        ix = {h: i for i, h in enumerate(rows[0])}
        print("== %s" % (rows[1][ix["Kernel Name"]] if len(rows) > 1 and "Kernel Name" in ix else a.report))
        for r in rows[1:]:
            if r[ix["Metric Name"]] in KEEP:
                print("  %-36s %s %s" % (r[ix["Metric Name"]], r[ix["Metric Value"]], r[ix["Metric Unit"]]))
    raw = page(a.report, "raw")
    if len(raw) > 2:
        d = dict(zip(raw[0], raw[-1]))
        for k in ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
                  "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                  "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
                  "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
                  "sm__pipe_tensor_subpipe_tcgen05_cycles_active.avg.pct_of_peak_sustained_active",
                  "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
                  "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"):
            if k in d:
                print("  %-78s %s" % (k, d[k]))
    src = page(a.report, "source")
    hi = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    if not hi:
        return
    hdr = src[hi[0]]
    ix = {h: i for i, h in enumerate(hdr)}
    data = src[hi[0] + 1:]
    f = lambda r, k: num(r[ix[k]]) if k in ix and ix[k] < len(r) else 0.0
    tot_s = sum(f(r, "# Samples") for r in data) or 1.0
    agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
    for r in data:
        sp = r[ix["Source"]].split()
        op = sp[0] if sp else "?"
        if op.startswith("@") and len(sp) > 1:
            op = sp[1]
        g = agg[op]
        g[0] += f(r, "Instructions Executed"); g[1] += f(r, "# Samples")
        g[2] += f(r, "L1 Wavefronts Shared"); g[3] += f(r, "L1 Tag Requests Global")
    u = a.units
    print("  per unit of work: %.0f warp-instructions, %.0f shared wavefronts, %.0f global L1 tag requests (%d SASS lines)" % (
        sum(g[0] for g in agg.values()) / u, sum(g[2] for g in agg.values()) / u, sum(g[3] for g in agg.values()) / u, len(data)))
    print("  opcode                  inst/unit  samples%  shared-wavefronts/unit  global-tags/unit")
    for op, g in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print("  %-22s %9.1f  %7.1f  %14.1f  %14.1f" % (op, g[0] / u, 100 * g[1] / tot_s, g[2] / u, g[3] / u))
    st = defaultdict(float)
    for r in data:
        for k in hdr:
            if k.startswith("stall_") and "Not Issued" not in k:
                st[k] += f(r, k)
    tot = sum(st.values()) or 1.0
    print("  stall reasons (%% of samples): %s" % ", ".join(
        "%s %.1f" % (k.replace("stall_", ""), 100 * v / tot) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    print("  hottest SASS lines:")
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:a.top]:
        top = sorted(((k, f(r, k)) for k in hdr if k.startswith("stall_") and "Not Issued" not in k), key=lambda kv: -kv[1])[:2]
        print("    %5.2f%%  x%-8.1f %-62s %s" % (100 * f(r, "# Samples") / tot_s, f(r, "Instructions Executed") / u,
                                               r[ix["Source"]][:62], " ".join("%s:%d" % (k.replace("stall_", ""), v) for k, v in top)))


if __name__ == "__main__":
    main()
