#!/usr/bin/env python
"""Per-source-line executed warp instructions and stall samples of one kernel: ncu -i X --page source --csv (cuda view)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
for i, r in enumerate(rows):
    if "Instructions Executed" in r and "# Samples" in r:
        hdr, rows = r, rows[i + 1:]
        break
if hdr is None:
    print("no header", rows[:2]); sys.exit(0)
ie, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
ln = 0
src = hdr.index("Source")
recs = []
for r in rows:
    if len(r) <= max(ie, si) or not r[ln].strip().isdigit():   # keep the CUDA-source rows (first column = line number)
        continue
    try:
        recs.append((float(r[ie] or 0), float(r[si] or 0), r[ln], r[src]))
    except Exception:
        pass
tot_i, tot_s = sum(x[0] for x in recs), sum(x[1] for x in recs)
print("total warp instructions %.0f, samples %.0f" % (tot_i, tot_s))
for x in sorted(recs, key=lambda x: -x[0])[:topn]:
    print("%5.1f%% inst %5.1f%% smp  L%-5s %s" % (100 * x[0] / max(tot_i, 1), 100 * x[1] / max(tot_s, 1), x[2], x[3][:120]))
