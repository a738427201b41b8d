#!/usr/bin/env python3
"""Summarises an ncu report by source line: python tools/ncu_lines.py <report.ncu-rep> [top N]
(stall samples and executed warp instructions per CUDA source line, plus the headline counters)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for ri in range(2, len(rows)):
    print("== kernel:", rows[ri][rows[0].index("Kernel Name")] if "Kernel Name" in rows[0] else "")
    for h, u, v in zip(rows[0], rows[1], rows[ri]):
        if h in want:
            print(f"  {h} = {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, agg, stall_tot = None, {}, {}
hdr = None
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 3 and r[0] == "Line No":
        hdr = r
        continue
    if len(r) < 8 or r[0] in ("Function Name", ""):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    try:
        samples = int(r[4]) if r[4] not in ("-", "") else 0
        inst = int(r[7]) if r[7] not in ("-", "") else 0
    except ValueError:      # a source line whose text broke the CSV quoting (inline asm)
        continue
    a = agg.setdefault((cur, line), [0, 0, r[1][:100], {}])
    a[0] += samples
    a[1] += inst
    if hdr:
        for h, v in zip(hdr, r):
            if h.startswith("stall_") and "Not Issued" not in h and v.isdigit() and v != "0":
                a[3][h] = a[3].get(h, 0) + int(v)
                stall_tot[h] = stall_tot.get(h, 0) + int(v)
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print(f"total samples {ts}, warp instructions {ti}")
print("stall mix:", ", ".join(f"{k[6:]} {100*v/ts:.0f}%" for k, v in sorted(stall_tot.items(), key=lambda x: -x[1])[:8]))
print("--- top lines by stall samples")
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    st = ",".join(f"{h[6:]}:{v}" for h, v in sorted(a[3].items(), key=lambda x: -x[1])[:2])
    print(f"{k[0]}:{k[1]:4d} samp {100*a[0]/ts:5.1f}% inst {100*a[1]/ti:5.1f}% [{st}] {a[2]}")
print("--- top lines by instructions")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{k[0]}:{k[1]:4d} inst {100*a[1]/ti:5.1f}% {a[1]/1e6:8.1f}M samp {100*a[0]/ts:5.1f}% {a[2]}")
