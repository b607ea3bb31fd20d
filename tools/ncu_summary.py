#!/usr/bin/env python3
"""One line per kernel launch from an `ncu --set full` report:
    ncu -i file.ncu-rep --page raw --csv > raw.csv;  python tools/ncu_summary.py raw.csv [last_n] > summary.csv
Columns are picked by metric name, so missing metrics print as empty fields; last_n keeps only the last n
launches (e.g. the second, warm submit of tools/one_frame.py)."""
import csv
import sys

COLS = [("time_us", "gpu__time_duration.sum", 1e-3), ("dram_rd_MB", "dram__bytes_read.sum", None), ("dram_wr_MB", "dram__bytes_write.sum", None),
        ("regs", "launch__registers_per_thread", 1), ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("pipe_alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
        ("pipe_fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
        ("pipe_lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1),
        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1), ("warp_inst", "smsp__inst_executed.sum", 1),
        ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1),
        ("threads_per_inst", "smsp__thread_inst_executed_per_inst_executed.ratio", 1), ("l2_hit_pct", "lts__t_sector_hit_rate.pct", 1)]

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units = rows[h], rows[h + 1]
ix = {n: i for i, n in enumerate(names)}
out = csv.writer(sys.stdout)
out.writerow(["kernel", "grid", "block"] + [c[0] for c in COLS])
body = [r for r in rows[h + 2:] if len(r) == len(names) and r[0].isdigit()]
if len(sys.argv) > 2:
    body = body[-int(sys.argv[2]):]
for r in body:
    line = [r[ix["Kernel Name"]].split("(")[0], r[ix["Grid Size"]].split(",")[0].strip("( "), r[ix["Block Size"]].split(",")[0].strip("( ")]
    for _, metric, scale in COLS:
        if metric not in ix:
            line.append(""); continue
        v = float(r[ix[metric]].replace(",", "")) if r[ix[metric]] not in ("", "n/a") else 0.0
        u = units[ix[metric]]
        if scale is None:      # bytes in whatever unit ncu chose -> MB
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
        elif metric == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1e-3)
        line.append(f"{v:.6g}")
    out.writerow(line)
