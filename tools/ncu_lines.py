#!/usr/bin/env python3
"""Per-source-line instruction / stall summary of an .ncu-rep (needs -lineinfo +
--import-source on).  usage: ncu_lines.py report.ncu-rep [top_n] [out.csv]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"],
                     capture_output=True, text=True).stdout
rows = []
fname = None
hdr = None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 8:
        continue
    try:
        line = int(r[0])
        inst = int(r[hdr.index("Instructions Executed")])
        samp = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    if r[2] != "-":  # per-SASS rows repeat the line's totals; keep only the line rows
        continue
    rows.append((fname, line, inst, samp, r[1].strip()))
ti = sum(x[2] for x in rows) or 1
ts = sum(x[3] for x in rows) or 1
print("total inst %d samples %d" % (ti, ts))
rows.sort(key=lambda x: -x[2])
w = None
if len(sys.argv) > 3:
    w = csv.writer(open(sys.argv[3], "w"))
    w.writerow(["file", "line", "inst_executed", "inst_pct", "stall_samples", "sample_pct", "source"])
for f, l, i, s, src in rows[:top]:
    print("%-18s %4d inst %10d %5.1f%%  samp %6d %5.1f%%  %s" % (f, l, i, 100 * i / ti, s, 100 * s / ts, src[:90]))
    if w:
        w.writerow([f, l, i, round(100 * i / ti, 2), s, round(100 * s / ts, 2), src])
