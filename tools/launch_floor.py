#!/usr/bin/env python
"""GPU-side cost of one launch inside a CUDA graph: 200 back-to-back launches of a trivially small problem per kernel family."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "calm-vit-dte_b200"))
import torch
import calm_kernels as K
dev = torch.device("cuda:0")
bf16, f32 = torch.bfloat16, torch.float32

def graph_time(fn, n=200):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 5 / n * 1e3

for M, N, Kd in ((128, 16, 64), (128, 128, 64), (2048, 128, 64), (20480, 240, 240), (20480, 160, 80), (57344, 672, 672)):
    x = torch.randn(M, Kd, device=dev).to(bf16); w = torch.randn(N, Kd, device=dev).to(bf16); y = torch.empty(M, N, device=dev, dtype=bf16)
    print("gemm M%d N%d K%d: %.2f us/launch" % (M, N, Kd, graph_time(lambda: K.gemm(x, w, y, M, N, Kd, lda=Kd, ldb=Kd, ldc=N))), flush=True)
x = torch.randn(1024, device=dev)
print("cast_bf16 1024 elements: %.2f us/launch" % graph_time(lambda: K.cast_bf16(x)), flush=True)
xs = torch.randn(64, 672, device=dev); wln = torch.ones(672, device=dev)
print("layernorm_fwd 64 rows: %.2f us/launch" % graph_time(lambda: K.layernorm_fwd(xs, wln)), flush=True)
