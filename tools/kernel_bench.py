#!/usr/bin/env python
"""Per-kernel micro-benchmarks at the trainer shapes (B=256, stages (224,672) (176,528) (128,384) (80,240)), through the
C ABI. CUDA-event timing, 3 warm-ups, rotating operand sets (so that an operand is not L2-hot from the previous
iteration), median of the timed launches. Writes gpurun_out/kernel_bench.json and prints a table.

    python tools/kernel_bench.py [gemm] [attn] [cnn] [ln] [misc]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "calm-vit-dte_b200"))
import torch  # noqa: E402
import calm_kernels as K  # noqa: E402

bf16, f32 = torch.bfloat16, torch.float32
dev = torch.device("cuda:0")
B = int(os.environ.get("KB_BATCH", 256))
STAGES = [(224, 672, 56), (176, 528, 44), (128, 384, 32), (80, 240, 20)]
if os.environ.get("KB_RES") == "384":    # BASELINE configs[3] stage shapes (per-GPU batch 64 / 32)
    STAGES, B = [(384, 1152, 96), (336, 1008, 84), (288, 864, 72), (240, 720, 60)], int(os.environ.get("KB_BATCH", 64))
elif os.environ.get("KB_RES") == "512":
    STAGES, B = [(512, 1536, 128), (464, 1392, 116), (416, 1248, 104), (368, 1104, 92)], int(os.environ.get("KB_BATCH", 32))
if os.environ.get("KB_STAGES"):     # e.g. KB_STAGES=224 for an ncu capture of one shape
    STAGES = [s for s in STAGES if str(s[0]) in os.environ["KB_STAGES"].split(",")]
PEAK_TF = 1358.9
PEAK_GB = 6549.4
try:
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    PEAK_TF, PEAK_GB = pk["bf16_tflops_sustained"], pk["hbm_gbs"]
except Exception:
    pass
results = []


AB = None
last_ab = {}


def timeit(fn, nsets=3, iters=12):
    flags = [None]
    for i in range(3):
        for f in flags:
            fn(i % nsets)
    torch.cuda.synchronize()
    ts = {f: [] for f in flags}
    for i in range(iters):
        for f in flags:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(i % nsets)
            b.record()
            torch.cuda.synchronize()
            ts[f].append(a.elapsed_time(b))
    med = {f: sorted(v)[len(v) // 2] for f, v in ts.items()}
    last_ab.clear()
    last_ab.update(med)
    return med[flags[0]]


def report(family, name, ms, flops=None, bytes_=None):
    r = {"family": family, "name": name, "ms": ms}
    if flops:
        r["tflops"] = flops / ms / 1e9
        r["frac_tensor_peak"] = r["tflops"] / PEAK_TF
    if bytes_:
        r["gbs"] = bytes_ / ms / 1e6
        r["frac_hbm_peak"] = r["gbs"] / PEAK_GB
    results.append(r)
    if AB:
        print("%-8s %-62s %s" % (family, name, "  ".join("[%#x] %.3f ms (%+.1f%%)" % (f, t, 100 * (t - ms) / ms) for f, t in last_ab.items())), flush=True)
        return
    print("%-8s %-46s %8.3f ms %s %s" % (family, name, ms, ("%7.1f TF/s (%.2f)" % (r["tflops"], r["frac_tensor_peak"])) if flops else "",
                                          ("%7.0f GB/s (%.2f)" % (r["gbs"], r["frac_hbm_peak"])) if bytes_ else ""), flush=True)


def rnd(*shape, dtype=bf16, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(dtype)


def bench_gemm():
    for S, D, hd in STAGES:
        M = B * S
        for tag, N, Kd in (("qkv fwd", 3 * D, D), ("out_proj fwd (+f32 addend, f32 out)", D, D), ("mlp.0 fwd (GELU)", 2 * D, D),
                           ("mlp.3 fwd (+f32 addend, f32 out)", D, 2 * D)):
            xs = [rnd(M, Kd) for _ in range(3)]
            w = rnd(N, Kd, scale=0.05)
            f32out = "f32" in tag
            outs = [torch.empty(M, N, dtype=f32 if f32out else bf16, device=dev) for _ in range(3)]
            add = rnd(M, N, dtype=f32) if f32out else None
            aux = torch.empty(M, N, dtype=bf16, device=dev) if "GELU" in tag else None

            def fn(i):
                kw = {}
                if add is not None:
                    kw.update(addend=add, ld_addend=N)
                if aux is not None:
                    kw.update(epilogue=K.EPI_GELU, aux=aux, ld_aux=N)
                K.gemm(xs[i], w, outs[i], M, N, Kd, lda=Kd, ldb=Kd, ldc=N, **kw)
            report("gemm", "S%d %s M%d N%d K%d" % (S, tag, M, N, Kd), timeit(fn), flops=2.0 * M * N * Kd)
            del xs, outs
        # dgrad (B MN-major) and wgrad (split-K) of the qkv projection and of mlp.3
        for tag, Nout, Kin in (("qkv", 3 * D, D), ("mlp.3", D, 2 * D)):
            dys = [rnd(M, Nout) for _ in range(3)]
            w = rnd(Nout, Kin, scale=0.05)
            xs = [rnd(M, Kin) for _ in range(3)]
            dx = torch.empty(M, Kin, dtype=bf16, device=dev)
            report("gemm", "S%d %s dgrad M%d N%d K%d" % (S, tag, M, Kin, Nout),
                   timeit(lambda i: K.gemm(dys[i], w, dx, M, Kin, Nout, lda=Nout, ldb=Kin, ldc=Kin, b_major=K.MAJOR_MN)),
                   flops=2.0 * M * Nout * Kin)
            sp = K.gemm_default_splits(Nout, Kin, M)
            part = torch.empty(sp, Nout, Kin, dtype=f32, device=dev)
            report("gemm", "S%d %s wgrad M%d N%d K%d splits%d" % (S, tag, Nout, Kin, M, sp),
                   timeit(lambda i: K.gemm(dys[i], xs[i], part, Nout, Kin, M, lda=Nout, ldb=Kin, ldc=Kin, a_major=K.MAJOR_MN,
                                           b_major=K.MAJOR_MN, splits=sp, stride_split=Nout * Kin)), flops=2.0 * M * Nout * Kin)
            del dys, xs
        # mask logits (batched) and the mask MLP
        q, k = rnd(B, S, D), rnd(B, S, D)
        lg = torch.empty(B, S, S, dtype=bf16, device=dev)
        report("gemm", "S%d mask logits batch%d M%d N%d K%d" % (S, B, S, S, D),
               timeit(lambda i: K.gemm(q, k, lg, S, S, D, batch=B, lda=D, ldb=D, ldc=S, stride_a=S * D, stride_b=S * D, stride_c=S * S)),
               flops=2.0 * B * S * S * D)
        w1 = rnd(2 * S, S, scale=0.05)
        hid = torch.empty(M, 2 * S, dtype=bf16, device=dev)
        aux = torch.empty(M, 2 * S, dtype=bf16, device=dev)
        bias = rnd(2 * S, dtype=f32)
        report("gemm", "S%d linear_mask.0 (bias, GELU) M%d N%d K%d" % (S, M, 2 * S, S),
               timeit(lambda i: K.gemm(lg, w1, hid, M, 2 * S, S, lda=S, ldb=S, ldc=2 * S, bias=bias, epilogue=K.EPI_GELU, aux=aux, ld_aux=2 * S)),
               flops=2.0 * M * 2 * S * S)
    # seq-axis (left-multiply) GEMMs of the stage-changing blocks
    for S1, S2, D in ((224, 80, 672), (80, 176, 240), (224, 176, 672), (80, 224, 240)):
        w = rnd(S2, S1, scale=0.05)
        x = rnd(B, S1, D)
        y = torch.empty(B, S2, D, dtype=bf16, device=dev)
        report("gemm", "seq-axis W(%d,%d) . X_b(%d,%d) batch%d" % (S2, S1, S1, D, B),
               timeit(lambda i: K.gemm(w, x, y, S2, D, S1, batch=B, lda=S1, ldb=D, ldc=D, stride_a=0, stride_b=S1 * D, stride_c=S2 * D,
                                       b_major=K.MAJOR_MN)), flops=2.0 * B * S2 * D * S1)


def bench_attn():
    stages = STAGES
    if os.environ.get("KB_ATTN"):       # e.g. KB_ATTN=72x20,80x20 : extra (S x head_dim) shapes at 12 heads
        stages = [(int(a.split("x")[0]), 12 * int(a.split("x")[1]), int(a.split("x")[1])) for a in os.environ["KB_ATTN"].split(",")]
    for S, D, hd in stages:
        h = 12
        qkv = rnd(B * S, 3 * D)
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        bias = rnd(B, S, S)
        o, lse = K.attention_fwd(q, k, v, bias, B, S, h, hd, 3 * D, 3 * D, 3 * D)
        report("attn", "fwd S%d hd%d" % (S, hd), timeit(lambda i: K.attention_fwd(q, k, v, bias, B, S, h, hd, 3 * D, 3 * D, 3 * D)),
               flops=4.0 * B * h * S * S * hd)
        d_o = rnd(B * S, D)
        report("attn", "bwd S%d hd%d" % (S, hd),
               timeit(lambda i: K.attention_bwd(q, k, v, bias, o, d_o, lse, B, S, h, hd, 3 * D, 3 * D, 3 * D, D)),
               flops=10.0 * B * h * S * S * hd)


def bench_cnn():
    for S, _D, _hd in STAGES:
        x = rnd(B, S, S * 3, dtype=f32)
        ws = [rnd(32, 3, dtype=f32), rnd(32, dtype=f32), rnd(32, 9, dtype=f32, scale=0.3), rnd(32, dtype=f32),
              rnd(3, 32, dtype=f32, scale=0.3), rnd(3, dtype=f32)]
        px = B * S * S
        report("cnn", "fwd S%d" % S, timeit(lambda i: K.cnn_fwd(x, *ws, B, S)), bytes_=24.0 * px)
        dy = rnd(B, S, S * 3, dtype=f32)
        report("cnn", "bwd S%d" % S, timeit(lambda i: K.cnn_bwd(x, dy, *ws, B, S)), bytes_=36.0 * px)


def bench_ln():
    for S, D, hd in STAGES:
        rows = B * S
        x = rnd(rows, D, dtype=f32)
        w = rnd(D, dtype=f32) + 1
        y, mean, rstd = K.layernorm_fwd(x, w)
        report("ln", "fwd rows%d D%d" % (rows, D), timeit(lambda i: K.layernorm_fwd(x, w)), bytes_=6.0 * rows * D)
        dy = rnd(rows, D)
        dres = rnd(rows, D, dtype=f32)
        report("ln", "bwd rows%d D%d (+dres)" % (rows, D), timeit(lambda i: K.layernorm_bwd(dy, x, w, mean, rstd, dres)),
               bytes_=14.0 * rows * D)


def bench_misc():
    for S, D, hd in STAGES:
        x = rnd(B, S, D, dtype=f32)
        report("misc", "token_transpose S%d" % S, timeit(lambda i: K.token_transpose(x, B, S)), bytes_=8.0 * x.numel())
        report("misc", "cast_bf16 S%d" % S, timeit(lambda i: K.cast_bf16(x)), bytes_=6.0 * x.numel())
        inv = (1.0 / (10000.0 ** (torch.arange(0, hd, 2).float() / hd))).to(dev)
        cs = inv
        qk = rnd(B * S, 3 * D)
        out = K.rope_fwd(None, 0, qk, 3 * D, cs, B * S, S, 12, 0, hd)
        report("misc", "rope_fwd S%d hd%d" % (S, hd), timeit(lambda i: K.rope_fwd(None, 0, qk, 3 * D, cs, B * S, S, 12, 0, hd)),
               bytes_=4.0 * B * S * D)
        dout = rnd(B * S, D)
        report("misc", "rope_bwd S%d hd%d" % (S, hd), timeit(lambda i: K.rope_bwd(dout, D, out, cs, B * S, S, 12, 0, hd)),
               bytes_=6.0 * B * S * D)
        xb = rnd(B * S, 2 * S)
        report("misc", "colsum rows%d N%d" % (B * S, 2 * S), timeit(lambda i: K.colsum(xb, B * S, 2 * S, 2 * S)), bytes_=2.0 * xb.numel())


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "attn", "cnn", "ln", "misc"]
    for w in which:
        globals()["bench_" + w]()
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = os.environ.get("KB_TAG", "kernel_bench")
    json.dump({"batch": B, "peak_tflops": PEAK_TF, "peak_gbs": PEAK_GB, "results": results},
              open(os.path.join(ROOT, "gpurun_out", tag + ".json"), "w"), indent=1)
