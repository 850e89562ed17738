#!/usr/bin/env python
"""Summarise an ncu report holding one launch of several kernels: key raw metrics per kernel, and for the
kernel matching `pattern` the SASS instructions with the most stall samples.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [pattern]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg"]
ki = H.index("Kernel Name")
for r in rows[2:]:
    print("== %s" % r[ki].split("(")[0])
    for h, u, x in zip(H, U, r):
        if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")
                         and float(x or 0) > 0.5):
            print("   %-88s %s %s" % (h, x, u))
if pat:
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k",
                           "regex:" + pat], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(sass)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    Hh = rows[hdr]
    si, so = Hh.index("# Samples"), Hh.index("Source")
    data = [r for r in rows[hdr + 1:] if len(r) > si and r[si].isdigit()]
    tot = sum(int(r[si]) for r in data)
    print("== stall samples of %s: %d; instructions with >= 1 %%" % (pat, tot))
    for i, r in enumerate(data):
        if int(r[si]) >= 0.01 * tot:
            print("   sass %5d  %5.1f%%  %s" % (i, 100.0 * int(r[si]) / tot, r[so][:90]))
