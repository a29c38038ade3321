"""Sweeps the GEMM N-tile width on the small / medium shapes of the step (tuning aid for the tile heuristic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "calm-vit-dte_b200"))
import torch, calm_lib, calm_kernels as K
lib = calm_lib.load()
dev = torch.device("cuda:0")
def timeit(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
shapes = [(20480, 240, 240), (20480, 480, 240), (20480, 240, 480), (20480, 160, 80), (32768, 384, 384), (32768, 768, 384), (32768, 384, 768),
          (45056, 528, 528), (45056, 1056, 528), (45056, 352, 176), (57344, 672, 672), (57344, 1344, 672), (57344, 2016, 672), (57344, 448, 224)]
for M, N, Kd in shapes:
    x = torch.randn(M, Kd, device=dev).bfloat16(); w = torch.randn(N, Kd, device=dev).bfloat16(); y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    res = []
    for bn in (0, 256, 240, 224, 192, 176, 160, 144, 128, 112, 96, 80, 64, 48):
        if bn and bn > (N + 15) // 16 * 16: continue
        us = timeit(lambda: K.gemm(x, w, y, M, N, Kd, lda=Kd, ldb=Kd, ldc=N, bn=bn))
        res.append((bn, us))
    best = min(res, key=lambda t: t[1])
    print("M%d N%d K%d  auto %.1fus  best bn=%d %.1fus  | %s" % (M, N, Kd, res[0][1], best[0], best[1], " ".join("%d:%.0f" % r for r in res[1:])), flush=True)
