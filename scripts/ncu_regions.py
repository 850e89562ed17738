#!/usr/bin/env python
"""Summarise an ncu report of one kernel: key raw metrics + stall samples per barrier-delimited SASS region.
Usage: python scripts/ncu_regions.py gpurun_out/prof.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, v = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active"]
for h, x, u in zip(H, v, rows[1]):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and float(x or 0) > 0.2):
        print("%-90s %s %s" % (h, x, u))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
Hh = rows[hdr]; si = Hh.index("# Samples")
data = rows[hdr + 1:]
tot = sum(int(r[si]) for r in data)
print("total samples", tot)
acc, start = 0, 0
for i, r in enumerate(data):
    acc += int(r[si])
    ins = r[1].strip()
    if ins.startswith("BAR.SYNC") or "UTCBAR" in ins or i == len(data) - 1:
        if acc > tot * 0.004:
            print("  sass %5d-%5d  %5.1f%%  ends: %s" % (start, i, 100.0 * acc / tot, ins[:50]))
        acc, start = 0, i + 1
