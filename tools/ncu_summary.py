#!/usr/bin/env python3
"""Summarise an .ncu-rep (details + per-instruction source page) into text: python tools/ncu_summary.py rep [top_n]"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
det = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h = det[0]
idc, kn, mn, mv, mu = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
want = ["Duration", "Registers Per Thread", "Theoretical Occupancy", "Achieved Occupancy", "Executed Ipc Active", "Issue Slots Busy", "L1/TEX Hit Rate", "L2 Hit Rate",
        "DRAM Throughput", "Warp Cycles Per Issued Instruction", "Avg. Active Threads Per Warp", "Executed Instructions", "Memory Throughput", "Local Memory Spilling Requests"]
seen = {}
for r in det[1:]:
    if r[mn] in want:
        seen.setdefault((r[idc], r[kn].split("(")[0]), []).append(f"{r[mn]}={r[mv]}{r[mu]}")
for k, v in seen.items():
    print("##", k[0], k[1]); print("   " + "; ".join(v))
SRC_PLACEHOLDER = None
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
# the source page holds one table per launch, each starting with a header row
tables, cur = [], None
for r in src:
    if "Source" in r and "# Samples" in r:
        cur = {"h": r, "rows": []}; tables.append(cur)
    elif cur is not None and len(r) == len(cur["h"]):
        cur["rows"].append(r)
for ti, t in enumerate(tables):
    h = t["h"]
    isrc, isamp, iex, ith = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    stall = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot, EX, TH, data = collections.Counter(), 0, 0, []
    for r in t["rows"]:
        try:
            s, ex, th = int(r[isamp]), int(r[iex]), int(r[ith])
        except ValueError:
            continue
        EX += ex; TH += th
        for i, c in stall:
            try:
                tot[c] += int(r[i])
            except ValueError:
                pass
        data.append((s, ex, th, r[isrc]))
    T = sum(tot.values()) or 1
    print(f"== launch {ti}: warp instr {EX}, avg lanes {TH / max(EX, 1):.2f}; stalls: " + ", ".join(f"{c[6:]} {100 * v / T:.1f}%" for c, v in tot.most_common(7)))
    for s, ex, th, sc in sorted(data, reverse=True)[:topn]:
        print(f"   {s:7d} {ex:10d} {th / max(ex, 1):5.1f}  {sc[:100]}")
