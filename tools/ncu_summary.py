#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): headline metrics, stall breakdown, and the hottest source lines."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_uniform.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d["Kernel Name"][:80], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in keys:
        if k in d: print(f"  {k:70s} {d[k]}")
    st = [(float(v.replace(",", "")), k) for k, v in d.items() if "warps_issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k and v]
    for v, k in sorted(st, reverse=True)[:8]:
        print(f"  stall {k.split('stalled_')[1].split('_per_')[0]:30s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
ci, cs, ct = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
items = []; tot_i = tot_s = 0
for r in rows[hi + 1:]:
    if len(r) < len(h) or not r[0].strip().isdigit():
        continue
    try:
        ie, ss, ti = float(r[ci]), float(r[cs]), float(r[ct])
    except Exception:
        continue
    tot_i += ie; tot_s += ss
    items.append((ie, ss, ti, r[0], r[1][:120]))
print(f"-- source lines: total warp inst {tot_i:.4g}, samples {tot_s:.4g}")
by = 1 if len(sys.argv) > 3 and sys.argv[3] == "samples" else 0
for ie, ss, ti, l, t in sorted(items, key=lambda x: -x[by])[:topn]:
    print(f"  {100*ie/max(tot_i,1):5.1f}% inst {100*ss/max(tot_s,1):5.1f}% smp  thr/inst {ti/max(ie,1):4.1f}  L{l:>4s} {t}")
# phase breakdown by source-line ranges given as name:lo-hi,... in argv[4]
if len(sys.argv) > 4:
    print("-- phases")
    for spec in sys.argv[4].split(","):
        name, rng = spec.split(":")
        acc_i = acc_s = 0
        for part in rng.split("+"):
            lo, hi2 = (int(x) for x in part.split("-"))
            acc_i += sum(ie for ie, ss, ti, l, t in items if lo <= int(l) <= hi2)
            acc_s += sum(ss for ie, ss, ti, l, t in items if lo <= int(l) <= hi2)
        print(f"  {name:22s} {100*acc_i/max(tot_i,1):5.1f}% inst {100*acc_s/max(tot_s,1):5.1f}% smp   {acc_i/1e6:8.2f} M warp-inst")
