#!/usr/bin/env python3
"""Executed warp instructions and stall samples per SOURCE LINE of one kernel: joins `ncu --page source --csv
--print-source sass` (instruction order = SASS order) with `nvdisasm -g` of the kernel's cubin (line info).
usage: ncu_lines.py sass.csv kernel.cubin mangled-name-substring [top N] [launch index]"""
import collections, csv, re, subprocess, sys
rows = list(csv.reader(open(sys.argv[1])))
cubin, sub = sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
which = int(sys.argv[5]) if len(sys.argv) > 5 else len(hdrs) - 1
hdr = hdrs[which]
rows = rows[:hdrs[which + 1]] if which + 1 < len(hdrs) else rows
h = rows[hdr]
data = [r for r in rows[hdr + 1:] if len(r) == len(h) and r[h.index("Instructions Executed")].isdigit()]
I, W = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# walk the function whose name contains `sub`
lines, cur, infn = [], ("?", 0), False
for l in dis:
    if l.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", l):
        infn = sub in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
if len(lines) != len(data):
    print(f"warning: {len(lines)} SASS instructions in the cubin vs {len(data)} in the profile", file=sys.stderr)
n = min(len(lines), len(data))
by = collections.defaultdict(lambda: [0, 0, 0])
for k in range(n):
    b = by[lines[k]]
    b[0] += int(data[k][I]); b[1] += int(data[k][W]); b[2] += 1
ti = sum(v[0] for v in by.values()) or 1; tw = sum(v[1] for v in by.values()) or 1
print(f"{ti} executed warp instructions, {tw} stall samples")
src = {}
for (f, ln), v in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src:
        try:
            src[f] = open(f"rav1d_b200/csrc/{f}").read().splitlines()
        except OSError:
            src[f] = []
    text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
    print(f"{f}:{ln:<5d} inst {v[0] / ti * 100:5.2f}%  stall {v[1] / tw * 100:5.2f}%  sass {v[2]:4d} | {text}")
