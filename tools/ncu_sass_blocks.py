#!/usr/bin/env python3
"""Basic-block view of `ncu --page source --csv --print-source sass`: contiguous SASS runs with the same
execution count, ranked by their share of all executed warp instructions.  usage: ncu_sass_blocks.py file.csv [N]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
which = int(sys.argv[3]) if len(sys.argv) > 3 else len(hdrs) - 1      # several launches in one export: the last one by default
hdr = hdrs[which]
rows = rows[:hdrs[which + 1]] if which + 1 < len(hdrs) else rows
h = rows[hdr]
data = [r for r in rows[hdr + 1:] if len(r) == len(h) and r[h.index("Instructions Executed")].isdigit()]
S, I, W = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[I]) for r in data)
totw = sum(int(r[W]) for r in data) or 1
blocks, cur = [], None
for i, r in enumerate(data):
    c = int(r[I])
    if cur and cur[2] == c:
        cur[1] = i
    else:
        cur = [i, i, c]
        blocks.append(cur)
print(f"{len(data)} SASS instructions, {tot} executed warp instructions")
for b in sorted(sorted(blocks, key=lambda b: -(b[1] - b[0] + 1) * b[2])[:top]):
    n = b[1] - b[0] + 1
    op = lambda t: (t.split()[1] if t.strip().startswith("@") else t.split()[0]).split(".")[0]
    ops = collections.Counter(op(data[k][S]) for k in range(b[0], b[1] + 1))
    st = sum(int(data[k][W]) for k in range(b[0], b[1] + 1))
    print(f"#{b[0]:5d}-{b[1]:5d} n={n:4d} x {b[2]:8d}  inst {n * b[2] / tot * 100:5.2f}%  stall {st / totw * 100:5.2f}%  "
          + " ".join(f"{k}:{v}" for k, v in ops.most_common(7)))
