#!/usr/bin/env python
"""Turns the ncu launch list of ONE training step into profiles/<tag>_ncu_summary.json + a text table.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \\
        --clock-control none -s <skip> -c <count> --csv --log-file launches.csv python bench.py --steps 1 --warmup 3 --no-graph ...
    python tools/ncu_step_summary.py launches.csv profiles/r02_ncu_summary.json > profiles/r02_ncu_launch_shares_step.txt

One step = the launches between two consecutive opt_adamw_kernel launches (the last kernel of a step). Per kernel name: launches,
time and share (ncu times are cold-cache and serialised: compare SHARES with bench.py's, not absolutes); for the tcgen05 GEMM the
average DRAM bytes per launch (roofline.traffic of bench.py); tensor-pipe activity per family and, for attention, per head
dimension (launch order within the step: blocks enc0..dec2, each encoder / decoder / cross)."""
import csv
import json
import re
import sys
from collections import defaultdict

src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui, mi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name"), hdr.index("ID")
launches = {}
order = []
for r in rows:
    if r[ki] == "Kernel Name":
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    lid = int(r[ii])
    if lid not in launches:
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^void |<unnamed>::|at::native::|\(anonymous namespace\)::", "", name)[:90]
        launches[lid] = {"name": name}
        order.append(lid)
    m = r[mi]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        launches[lid]["us"] = v
    elif m.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1.0)
        launches[lid]["dram"] = launches[lid].get("dram", 0.0) + v
    elif m.startswith("sm__pipe_tensor"):
        launches[lid]["tensor"] = v
seq = [launches[i] for i in order]
ends = [i for i, l in enumerate(seq) if l["name"].startswith("opt_adamw_kernel")]
if len(ends) >= 2:
    seq = seq[ends[-2] + 1: ends[-1] + 1]
    note = "one step: the %d launches between two consecutive opt_adamw_kernel launches" % len(seq)
else:
    note = "no two optimizer launches in the capture: all %d launches" % len(seq)
agg = defaultdict(lambda: [0.0, 0, 0.0, 0.0])
for l in seq:
    a = agg[l["name"]]
    a[0] += l.get("us", 0.0); a[1] += 1; a[2] += l.get("dram", 0.0); a[3] += l.get("tensor", 0.0) * l.get("us", 0.0)
tot = sum(a[0] for a in agg.values())
print("%s; total kernel time %.2f ms (cold-cache, serialised under ncu: compare shares)" % (note, tot / 1e3))
shares = []
for k, (us, n, dram, tw) in sorted(agg.items(), key=lambda t: -t[1][0]):
    shares.append({"kernel": k, "launches": n, "us": round(us, 1), "share": round(us / tot, 4), "dram_MB_per_launch": round(dram / n / 1e6, 2),
                   "tensor_pipe_pct_time_weighted": round(tw / us, 2) if us else 0.0})
for s in shares[:45]:
    print("%6.2f%%  %10.1f us  x%-5d %8.2f MB/launch  tensor %5.1f%%  %s" % (100 * s["share"], s["us"], s["launches"], s["dram_MB_per_launch"],
                                                                          s["tensor_pipe_pct_time_weighted"], s["kernel"]))


def family(pred):
    ls = [l for l in seq if pred(l["name"])]
    us = sum(l.get("us", 0.0) for l in ls)
    return ls, us


gemm, gemm_us = family(lambda n: n.startswith("gemm_tcgen05_kernel"))
afwd, afwd_us = family(lambda n: n.startswith("attn_fwd"))
abwd, abwd_us = family(lambda n: n.startswith("attn_bwd"))
wavg = lambda ls, us: round(sum(l.get("tensor", 0.0) * l.get("us", 0.0) for l in ls) / us, 2) if us else None
HD_FWD = [56, 56, 44, 44, 44, 32, 32, 32, 20, 20, 20, 20, 20, 20, 20, 20, 20, 32, 32, 32, 44, 44, 44, 56]
by_hd = {}
for tag, ls, seqhd in (("fwd", afwd, HD_FWD), ("bwd", abwd, HD_FWD[::-1])):
    if len(ls) == 24:
        d = defaultdict(list)
        for l, hd in zip(ls, seqhd):
            d[hd].append(l)
        by_hd[tag] = {str(hd): {"tensor_pipe_pct": round(sum(x.get("tensor", 0.0) for x in v) / len(v), 2), "us_per_launch": round(sum(x.get("us", 0.0) for x in v) / len(v), 1),
                                "launches": len(v)} for hd, v in sorted(d.items())}
out = {"source": src, "note": note, "step_launches": len(seq), "kernel_time_ms": round(tot / 1e3, 3), "shares": shares[:40],
       "gemm_launches": len(gemm), "gemm_dram_bytes_per_launch": round(sum(l.get("dram", 0.0) for l in gemm) / max(len(gemm), 1), 1),
       "gemm_share": round(gemm_us / tot, 4),
       "tensor_pipe_pct": {"calm_gemm": wavg(gemm, gemm_us), "calm_attention_fwd": wavg(afwd, afwd_us), "calm_attention_bwd": wavg(abwd, abwd_us)},
       "attention_tensor_pipe_pct_by_head_dim": by_hd}
json.dump(out, open(dst, "w"), indent=1)
print("gemm: %d launches, %.1f MB dram per launch, tensor pipe %.1f%% (time-weighted) | attention fwd %.1f%% bwd %.1f%%" % (
    len(gemm), out["gemm_dram_bytes_per_launch"] / 1e6, out["tensor_pipe_pct"]["calm_gemm"] or 0, out["tensor_pipe_pct"]["calm_attention_fwd"] or 0,
    out["tensor_pipe_pct"]["calm_attention_bwd"] or 0))
