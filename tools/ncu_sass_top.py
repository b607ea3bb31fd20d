#!/usr/bin/env python3
"""Summarise `ncu --page source --csv` (SASS view) of one kernel: top instructions by stall samples
and totals by opcode.  usage: ncu_sass_top.py file.csv [N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
which = int(sys.argv[3]) if len(sys.argv) > 3 else len(hdrs) - 1      # several launches in one export: the last one by default
hdr = hdrs[which]
rows = rows[:hdrs[which + 1]] if which + 1 < len(hdrs) else rows
h = rows[hdr]; data = [r for r in rows[hdr + 1:] if len(r) == len(h) and r[0] != "Address" and r[h.index("Instructions Executed")].isdigit()]
S, I, W = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
T = h.index("Thread Instructions Executed")
tot_i = sum(int(r[I]) for r in data); tot_w = sum(int(r[W]) for r in data)
print(f"kernel: {rows[0][1][:90]}\ninstructions {tot_i}  samples {tot_w}  avg threads/inst {sum(int(r[T]) for r in data)/max(tot_i,1):.1f}")
by = collections.defaultdict(lambda: [0, 0])
for r in data:
    op = r[S].split()[0] if not r[S].strip().startswith("@") else r[S].split()[1]
    op = op.split(".")[0]
    by[op][0] += int(r[I]); by[op][1] += int(r[W])
print("by opcode (inst%, stall%):", ", ".join(f"{k} {v[0]/tot_i*100:.1f}/{v[1]/tot_w*100:.1f}" for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:16]))
idx = {id(r): i for i, r in enumerate(data)}
for r in sorted(data, key=lambda r: -int(r[W]))[:n]:
    print(f"  #{idx[id(r)]:5d} inst {int(r[I])/tot_i*100:5.2f}%  stall {int(r[W])/tot_w*100:5.2f}%  {r[S].strip()[:100]}")
