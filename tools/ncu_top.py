#!/usr/bin/env python3
"""Summarise an .ncu-rep: headline counters, stall reasons per issue, and the top stall locations of the source page.
usage: python tools/ncu_top.py file.ncu-rep [ntop]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
keys = ("Duration", "Executed Ipc Active", "Issue Slots Busy", "Registers Per Thread", "Achieved Active Warps", "Theoretical Active Warps", "DRAM Throughput",
        "L1/TEX Hit", "Mem Busy", "Eligible Warps", "No Eligible", "Dynamic Shared Memory Per Block", "Waves Per SM", "Memory Throughput", "Max Bandwidth",
        "Compute (SM) Throughput", "L2 Cache Throughput", "L1/TEX Cache Throughput", "Mem Pipes Busy")
for l in det.splitlines():
    if any(k in l for k in keys) and "INF" not in l and "OPT" not in l:
        print(l.strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw"], capture_output=True, text=True).stdout
st = []
for l in raw.splitlines():
    if "smsp__average_warps_issue_stalled_" in l and "per_issue_active.ratio" in l or "smsp__average_warp_latency_issue_stalled" in l and ".ratio" in l:
        p = l.split()
        try:
            st.append((float(p[-1]), p[0].replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "")))
        except ValueError:
            pass
    for k in ("smsp__inst_executed.sum ", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
              "sm__inst_executed_pipe_fma.sum ", "sm__pipe_fma_cycles_active.avg.pct", "sm__inst_executed_pipe_lsu.sum ", "l1tex__throughput.avg.pct", "sm__pipe_fmaheavy", "smsp__inst_executed_pipe_fma"):
        if k in l:
            print(" ".join(l.split()))
print("stalls per issue:", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:10]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
for i, r in enumerate(rows):
    if "Address" in r and "# Samples" in r:
        h = i; break
if h is not None:
    hdr = rows[h]; data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
    ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    cols = [c for c in ("stall_long_sb", "stall_wait", "stall_mio", "stall_barrier", "stall_short_sb", "stall_math", "stall_not_selected", "stall_branch_resolving", "stall_lg") if c in hdr]
    tot = sum(int(r[isamp] or 0) for r in data) or 1
    print("samples", tot, "warp instructions", sum(int(r[iex] or 0) for r in data))
    for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:ntop]:
        extra = " ".join("%s=%s" % (c.replace("stall_", ""), r[hdr.index(c)]) for c in cols if r[hdr.index(c)] not in ("0", ""))
        print("%s %5.1f%% ex=%-8s %-60s %s" % (r[ia][-5:], 100 * int(r[isamp]) / tot, r[iex], r[isrc][:60], extra))
