#!/usr/bin/env python
"""Compact text summary of one `ncu --set full` capture from its exported pages:
   python scripts/summarise_ncu.py NAME  (reads gpurun_out/r02_NAME_{details.txt,raw.csv,source.csv})"""
import collections
import csv
import re
import sys

name = sys.argv[1]
base = "gpurun_out/r02_%s" % name
out = []
det = open(base + "_details.txt").read()
head = [l for l in det.splitlines() if re.match(r"\s+\S.*\(\d+, \d+, \d+\)x\(\d+, \d+, \d+\)", l)]
out += ["# ncu --set full --clock-control none, one launch (cold caches, serialised): %s" % name] + [h.strip() for h in head[:1]]
keys = ("Duration", "Executed Ipc Active", "Issue Slots Busy", "Registers Per Thread", "Theoretical Occupancy", "Achieved Occupancy",
        "DRAM Throughput", "Memory Throughput", "L2 Cache Throughput", "Compute (SM) Throughput", "L2 Hit Rate", "L1/TEX Hit Rate",
        "Dynamic Shared Memory Per Block", "Block Limit Registers", "Block Limit Shared Mem", "No Eligible", "Eligible Warps Per Scheduler")
seen = set()
for l in det.splitlines():
    for k in keys:
        if l.strip().startswith(k) and (k, l.split()[-1]) not in seen:
            seen.add((k, l.split()[-1]))
            out.append("  " + " ".join(l.split()))
rows = list(csv.reader(open(base + "_raw.csv")))
want = ("dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "gpu__time_duration.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size")
out.append("raw metrics:")
for h, u, v in zip(rows[0], rows[1], rows[2]):
    if h in want:
        out.append("  %-82s %-12s %s" % (h, u, v))
for h, u, v in zip(rows[0], rows[1], rows[2]):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and float(v) >= 0.05:
        out.append("  stall %-40s %.3f per issue" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], float(v)))
src = list(csv.reader(open(base + "_source.csv")))
hdr = src[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.defaultdict(collections.Counter)
cnt, smp = collections.Counter(), collections.Counter()
for r in src[2:]:
    if len(r) < len(hdr):
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[idx["Source"]])
    op = m.group(2) if m else "?"
    op = ".".join(op.split(".")[:2]) if op.startswith(("MUFU", "LDS", "LDG", "DMMA", "F2F", "I2F", "UBLKCP", "SYNCS")) else op.split(".")[0]
    cnt[op] += float(r[idx["Instructions Executed"]])
    smp[op] += float(r[idx["# Samples"]])
    for s in stalls:
        tot[op][s] += float(r[idx[s]])
T, I = sum(smp.values()), sum(cnt.values())
out.append("SASS opcode histogram of the launch (executed warp instructions) and where the warp-stall samples fall:")
for k, v in cnt.most_common(24):
    top = ", ".join("%s %.0f%%" % (s[6:], 100 * c / max(smp[k], 1)) for s, c in tot[k].most_common(3))
    out.append("  %-14s instr %5.1f%%   samples %5.1f%%   | %s" % (k, 100 * v / I, 100 * smp[k] / T, top))
print("\n".join(out))
