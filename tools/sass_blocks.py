#!/usr/bin/env python
"""Basic-block view of one kernel of an .ncu-rep (read here, no GPU): consecutive SASS instructions with the same execution count
are one block; prints the blocks by executed warp instructions (what an issue-bound kernel pays for), with their share, the
average active threads and the first instructions of each.  usage: sass_blocks.py REPORT [TOP]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1] if len(rows[0]) > 1 else "")
h = rows[1]
isrc, ie, ismp, it = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
ins = [(r[isrc].strip(), int(r[ie]), int(r[ismp]), int(r[it])) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
tot = sum(e for _, e, _, _ in ins)
warps = ins[0][1]
print(f"{len(ins)} SASS instructions, {tot/1e6:.1f} M executed warp instructions, {warps} warps launched ({tot/warps:.0f} per warp)")
blocks, cur = [], []
for i, x in enumerate(ins):
    if cur and x[1] != ins[i - 1][1]:
        blocks.append(cur); cur = []
    cur.append((i,) + x)
if cur:
    blocks.append(cur)
acc = 0
for b in sorted(blocks, key=lambda b: -sum(x[2] for x in b))[:top]:
    t = sum(x[2] for x in b); acc += t
    ops = " ".join(x[1].split()[0] for x in b[:10])
    print(f"@{b[0][0]:<5d} {len(b):3d} instr x {b[0][2]:7d} ({b[0][2]/warps:4.1f}/warp) = {t/1e6:6.2f} M {100*t/tot:5.1f}% cum {100*acc/tot:5.1f}%  "
          f"threads {sum(x[4] for x in b)/max(t,1):4.1f}  samples {sum(x[3] for x in b):5d} | {ops}")
