#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
agg = defaultdict(lambda: [0.0, 0])
for r in rows:
    if r is hdr or r[ki] == "Kernel Name":
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void |<unnamed>::|at::native::|\(anonymous namespace\)::", "", name)[:90]
    agg[name][0] += v
    agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
print("launches %d, total kernel time %.2f ms (cold-cache, serialised under ncu: compare shares)" % (sum(v[1] for v in agg.values()), tot / 1e3))
for k, (us, n) in sorted(agg.items(), key=lambda t: -t[1][0])[:40]:
    print("%6.2f%%  %10.1f us  x%-5d %s" % (100 * us / tot, us, n, k))
