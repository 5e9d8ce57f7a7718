#!/usr/bin/env python
"""Per-CUDA-source-line warp-stall samples of an .ncu-rep (needs -lineinfo + --import-source on).
usage: scripts/ncu_lines.py rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fname = "?"; rows = []; hdr = None
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            rows.append((float(d["Warp Stall Sampling (All Samples)"] or 0), float(d["Instructions Executed"] or 0),
                         fname, int(r[0]), r[1].strip()[:105]))
        except ValueError:
            pass
tot = sum(x[0] for x in rows) or 1; toti = sum(x[1] for x in rows) or 1
print("total samples %d, warp instructions %d" % (tot, toti))
for s, ins, f, ln, src in sorted(rows, reverse=True)[:top]:
    print("%6.2f%% smp %6.2f%% inst  %s:%d  %s" % (100 * s / tot, 100 * ins / toti, f, ln, src))
