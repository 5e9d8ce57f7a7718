#!/usr/bin/env python
"""SASS instructions of an .ncu-rep with their samples and stall reasons.
usage: scripts/ncu_sass.py rep [regex of the region start] [lines] | or 'top N' for the N hottest instructions"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows if r and r[0].startswith("0x")]
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(float(r[ix["# Samples"]] or 0) for r in body) or 1.0
def fmt(r):
    s = float(r[ix["# Samples"]] or 0)
    st = sorted(((float(r[ix[n]] or 0), n[6:]) for n in stalls), reverse=True)[:3]
    return "%5.2f%% ex=%-10s %-64s %s" % (100 * s / tot, r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:64],
                                        " ".join("%s:%d" % (n, v) for v, n in st if v > 0))
if len(sys.argv) > 2 and sys.argv[2] == "top":
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    for r in sorted(body, key=lambda r: -float(r[ix["# Samples"]] or 0))[:n]:
        print(r[0][-5:], fmt(r))
else:
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    hits = [i for i, r in enumerate(body) if pat is None or pat.search(r[ix["Source"]])]
    start = hits[skip] if hits else 0
    for r in body[max(0, start - 10): start + n]:
        print(r[0][-5:], fmt(r))
print("total samples", int(tot))
