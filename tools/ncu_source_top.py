#!/usr/bin/env python
"""Top source lines of a kernel by warp-state samples, from `ncu -i X.ncu-rep --page source --csv` (view: source+SASS).
usage: ncu_source_top.py report.ncu-rep [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
if not out.strip():
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
for i, r in enumerate(rows):
    if any("Sampl" in c for c in r):
        hdr = r
        rows = rows[i + 1:]
        break
if hdr is None:
    print("no sampling columns found; header candidates:", rows[:3])
    sys.exit(0)
print("columns:", hdr)
scol = [i for i, c in enumerate(hdr) if "Sampl" in c and "Not" not in c][0]
src = [i for i, c in enumerate(hdr) if c.strip() in ("Source", "# Source")]
srci = src[0] if src else 1
recs = []
for r in rows:
    try:
        n = float(r[scol])
    except Exception:
        continue
    if n > 0:
        recs.append((n, r))
tot = sum(n for n, _ in recs)
recs.sort(key=lambda x: -x[0])
print("total samples", tot)
for n, r in recs[:topn]:
    print("%6.0f %5.1f%%  %s" % (n, 100 * n / tot, " | ".join(c[:110] for j, c in enumerate(r) if j != scol and c and j < 6)))
