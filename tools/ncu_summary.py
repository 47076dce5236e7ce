#!/usr/bin/env python
"""Summarise ncu output for profiles/:  launch list (csv from --metrics gpu__time_duration.sum) and/or a
full-set report (.ncu-rep, read with `ncu -i ... --page raw --csv`).

    python tools/ncu_summary.py --launches gpurun_out/launches.csv --rep gpurun_out/prof.ncu-rep > profiles/rN_x.md
"""
import argparse
import collections
import csv
import re
import subprocess

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__t_bytes.sum", "L1 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (per SM)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__maximum_warps_per_active_cycle_pct", "theoretical occupancy %"),
    ("launch__registers_per_thread", "registers"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem / block"),
    ("launch__shared_mem_per_block_static", "static smem / block"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print("## launch list (cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| %s | %d | %.3f | %.1f%% | %.1f |" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot, v[1] / v[0]))
    print("| **total** | %d | %.3f | | |\n" % (sum(v[0] for v in agg.values()), tot / 1e3))


def report(path, pattern):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = collections.Counter()
    print("## full-set capture (%s)\n" % path.split("/")[-1])
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0]
        if pattern and not re.search(pattern, name):
            continue
        seen[name] += 1
        if seen[name] > 1:
            continue
        print("### %s\n\n| metric | value |\n|---|---:|" % name)
        for k, label in KEYS:
            if k in idx:
                print("| %s (`%s`) | %s %s |" % (label, k, r[idx[k]], units[idx[k]]))
        print()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--kernels", default="")
    ap.add_argument("--title", default="ncu summary")
    a = ap.parse_args()
    print("# %s\n" % a.title)
    if a.launches:
        launches(a.launches)
    if a.rep:
        report(a.rep, a.kernels)
