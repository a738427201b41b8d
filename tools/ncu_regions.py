#!/usr/bin/env python3
"""Dynamic instruction / stall-sample share per code region of iamfb_fused.cuh (regions = the marker comments)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
quads = float(sys.argv[2]) if len(sys.argv) > 2 else 122880.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; agg = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 8: continue
    try: line = int(r[0]); inst = int(r[7]); samp = int(r[4])
    except ValueError: continue
    a = agg.setdefault((cur, line), [0, 0]); a[0] += inst; a[1] += samp
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
FILE = sys.argv[3] if len(sys.argv) > 3 else 'iamfb_fused.cuh'
lines = open('iac_b200/csrc/' + FILE).read().split('\n')
stream_keys = [('void stream_scan', 'scan'), ('int stream_q16', 'quantise'), ('Q4 stream_ld', 'input loads (stream_ld)'),
               ('void stream_gain', 'output gain'), ('Q4 stream_in', 'chain loads'), ('bool stream_derivable', 'derived select'),
               ('__global__ void', 'kernel prologue'), ('auto prefetch', 'prefetch'), ('auto render', 'render: setup'),
               ('// ---- phase A', 'phase A: derivation chain'), ('// ---- phase B', 'phase B: columns (recon, select)'),
               ('// element / output mix gains', 'finish: gains + peak + store'), ('auto wmax', 'wmax'),
               ('auto output', 'output'), ('// ---- scanner state', 'scanner state'), ('// ---- iteration t', 'tile loop'),
               ('// the last 240 instants', 'epilogue'), ('stream_div(const Q4', 'exact division'), ('void stream_mat_col', 'matrix columns'),
               ('float stream_slow_div', 'slow division'), ('void bar_stream_workers', 'barriers/helpers')]
keys = stream_keys if FILE == 'iamfb_stream.cuh' else [('__device__ __forceinline__ V4 lds4', 'helpers'), ('__device__ __noinline__ void slow_div4', 'exact_div'),
        ('void fused_reconstruct', 'reconstruct'), ('void fused_element', 'element setup'), ('// element mix gain / output mix gain', 'gains setup'),
        ('// render: out = 0; out += mat', 'csr+finalize'), ('void store_any', 'store_any'), ('void fused_scan', 'scan'),
        ('__global__ void', 'kernel prologue'), ('while (f < n_frames || flush_pending)', 'tile loop: render'),
        ('// the staged rows are consumed', 'advance/issue'), ('// ---------------------------------------------------------------- sliding maximum', 'wmax'),
        ('// ---------------------------------------------------------------- gain recurrence', 'scan call/EW'),
        ('// ------------------------------------------------------------------ output', 'output'),
        ('// ------------------------------------------------------------------ advance the ring', 'ring advance/realign'),
        ('// the last 240 instants, in time order', 'epilogue')]
marks = sorted((i, n) for i, l in enumerate(lines, 1) for k, n in keys if k in l)
def region(f, l):
    if f != FILE: return f
    name = 'top'
    for i, n in marks:
        if l >= i: name = n
    return name
reg = {}
for (f, l), a in agg.items():
    x = reg.setdefault(region(f, l), [0, 0]); x[0] += a[0]; x[1] += a[1]
print(f"total warp instructions {ti/1e6:.1f}M, samples {ts}")
for n, x in sorted(reg.items(), key=lambda x: -x[1][0]):
    print(f"{n:28s} inst {100*x[0]/ti:5.1f}% ({x[0]/1e6:6.1f}M = {x[0]/quads:6.0f}/quad-warp)  samples {100*x[1]/ts:5.1f}%")
