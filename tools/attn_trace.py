#!/usr/bin/env python
"""Timeline of the attention-backward pipeline on CTA 0 (globaltimer stamps, calm_debug_set_trace_buffer).
Event ids: 1000 ctrl TMA landed | 1010 tails zeroed | 102p MMA part p issue | 1030 all parts drained | 1040 final MMAs committed |
1050 tile epilogue done ; 2000 worker saw Q | 2010 arrived A | 202p wait S/dP part p | 203p got it | 204p math done | 205p arrived |
2060 final MMAs done | 2070 tile stores done."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "calm-vit-dte_b200"))
import torch  # noqa: E402
import calm_kernels as K  # noqa: E402
import calm_lib  # noqa: E402

dev = torch.device("cuda:0")
S, D, hd, B = int(sys.argv[1]) if len(sys.argv) > 1 else 224, None, None, 256
WHAT = sys.argv[2] if len(sys.argv) > 2 else "bwd"      # bwd | fwd (the forward is only instrumented in attention_long_sm100.cu)
D, hd = 3 * S, S // 4
B = {384: 64, 512: 32}.get(S, 256)
bf16 = torch.bfloat16
qkv = (torch.randn(B * S, 3 * D, device=dev)).to(bf16)
bias = (torch.randn(B, S, S, device=dev) * 0.5).to(bf16)
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
o, lse = K.attention_fwd(q, k, v, bias, B, S, 12, hd, 3 * D, 3 * D, 3 * D)
do = torch.randn_like(o)
for _ in range(2):
    K.attention_bwd(q, k, v, bias, o, do, lse, B, S, 12, hd, 3 * D, 3 * D, 3 * D, D)
cap = 4000                                  # pairs per region; 4 regions (controller / S-MMA lane, worker thread 0, PV-MMA lane, loader lane)
buf = torch.zeros(4 * 2 * cap, dtype=torch.int64, device=dev)
import ctypes  # noqa: E402
set_trace = calm_lib.load().calm_debug_set_trace_buffer      # only exported by -DCALM_BRINGUP builds (CALM_NVCC_FLAGS=-DCALM_BRINGUP)
set_trace.argtypes, set_trace.restype = [ctypes.c_void_p, ctypes.c_int32], None
set_trace(buf.data_ptr(), cap)
if WHAT == "fwd":
    K.attention_fwd(q, k, v, bias, B, S, 12, hd, 3 * D, 3 * D, 3 * D)
else:
    K.attention_bwd(q, k, v, bias, o, do, lse, B, S, 12, hd, 3 * D, 3 * D, 3 * D, D)
torch.cuda.synchronize()
set_trace(None, 0)
t = buf.cpu().tolist()
ev = []
for region in range(4):
    for i in range(cap):
        eid, ts = t[2 * (region * cap + i)], t[2 * (region * cap + i) + 1]
        if eid == 0:
            break
        ev.append((ts, eid))
ev.sort()
t0 = ev[0][0]
print("events", len(ev))
for ts, eid in ev[:int(os.environ.get("TRACE_PRINT", 120))]:
    print("%9.2f us  %d" % ((ts - t0) / 1000.0, eid))
# where the time goes: per role (1xxx controller lane, 2xxx worker thread 0), the average gap before each event id
for role in (1, 2):
    seq = [(ts, eid) for ts, eid in ev if eid // 1000 == role]
    gaps = {}
    for (t_a, a), (t_b, b_) in zip(seq, seq[1:]):
        if WHAT == "fwd":       # long forward ids carry the step number in the last two digits: fold pass A / pass B steps together
            nch = (S + 127) // 128
            fa = lambda e: e - e % 100 + (0 if e % 100 < nch else 50) if e % 1000 < 500 else e
            a, b_ = fa(a), fa(b_)
        g = gaps.setdefault((a, b_), [0, 0.0])
        g[0] += 1
        g[1] += (t_b - t_a) / 1000.0
    total = sum(g[1] for g in gaps.values())
    print("role %d: %.1f us traced" % (role, total))
    for (a, b_), g in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:18]:
        print("   %d -> %d   n %4d   avg %6.2f us   share %4.1f %%" % (a, b_, g[0], g[1] / g[0], 100 * g[1] / total))
