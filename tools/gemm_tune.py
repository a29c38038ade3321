"""Schedule search for the GEMMs of one training step, on the live operands.

Runs ONE eager step of the trainer config with calm_kernels.gemm wrapped: every call is first timed under a set of per-call
schedule overrides (tile width `bn`, cluster-pair on / off — the `flags` / `bn_override` fields of calm_gemm_args) and then
executed normally. Prints, per shape class, the default time, the best variant and what the step would gain; writes
gpurun_out/gemm_tune.json. The result feeds the host-side heuristics of csrc/gemm_sm100.cu (it is a tuning aid, not a run-time
autotuner: the library picks schedules from the problem shape alone).

    python tools/gemm_tune.py [--res 224] [--task cls] [--batch N]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "calm-vit-dte_b200"))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=224)
    ap.add_argument("--task", default="cls")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--reps", type=int, default=8)
    args = ap.parse_args()
    import bench
    import calm_kernels as K
    import calm_lib as L
    import calm_ops
    import CALM_ViT_V2 as rvh
    calm_ops.PARALLEL = False
    dev = torch.device("cuda", 0)
    S, task = args.res, args.task
    B = args.batch or bench.DEFAULT_BATCH[S]
    torch.manual_seed(0)
    model = rvh.ViT(dev, type=8, **bench.vit_kwargs(S, task)).to(dev)
    model.train()
    tr = bench.Trainer(model, task, dev, B, S, False)
    xh, yh = bench.synth_batch(B, S, task, 2006)
    tr.x.copy_(xh)
    if yh is not None:
        tr.y.copy_(yh)
    for _ in range(2):
        tr._step()
    torch.cuda.synchronize()

    real = K.gemm
    results = {}

    def time_variant(a, kw, flags, bn):
        try:
            for _ in range(2):
                real(*a, **dict(kw, flags=flags, bn=bn))
        except Exception as e:       # a schedule the kernel refuses for this operand mix
            return None
        torch.cuda._sleep(400000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            real(*a, **dict(kw, flags=flags, bn=bn))
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) * 1e3 / args.reps

    def wrapped(*a, **kw):
        Aa, Bb, Cc, M, N, Kk = a[:6]
        batch = kw.get("batch", 1)
        key = "M%d N%d K%d b%d %s%s%s%s e%d%s s%d" % (
            M, N, Kk, batch, "AK" if kw.get("a_major", K.MAJOR_K) == K.MAJOR_K else "AM",
            "BK" if kw.get("b_major", K.MAJOR_K) == K.MAJOR_K else "BM", " f32" if Cc.dtype == torch.float32 else "",
            " add" if kw.get("addend") is not None else "", kw.get("epilogue", 0), " red" if kw.get("reduce_batch") else "", kw.get("splits", 1))
        if kw.get("flags", 0) == 0 and kw.get("bn", 0) == 0:
            ntn0 = -(-N // 256)
            bns = []
            for ntn in (ntn0, ntn0 + 1, ntn0 + 2, 2 * ntn0, 2 * ntn0 + 2):
                bn = -(-(-(-N // ntn)) // 16) * 16
                if 16 <= bn <= 256 and bn not in bns:
                    bns.append(bn)
            variants = [(0, 0)] + [(0, bn) for bn in bns] + [(L.GEMM_NO_CLUSTER, 0), (L.GEMM_FORCE_CLUSTER, 0)]   # bns[0] = the widest even tile
            variants += [(L.GEMM_FORCE_CLUSTER, bn) for bn in bns[1:2]] + [(L.GEMM_NO_CLUSTER, bn) for bn in bns[1:2]]
            rec = results.setdefault(key, {"n": 0, "t": {}})
            rec["n"] += 1
            for fl, bn in variants:
                t = time_variant(a, kw, fl, bn)
                if t is not None:
                    name = "f%d bn%d" % (fl, bn)
                    rec["t"][name] = rec["t"].get(name, 0.0) + t
        return real(*a, **kw)

    K.gemm = wrapped
    tr._step()
    torch.cuda.synchronize()
    K.gemm = real
    rows = []
    tot_def = tot_best = 0.0
    for key, rec in results.items():
        d = rec["t"].get("f0 bn0")
        if d is None:
            continue
        best = min(rec["t"], key=rec["t"].get)
        rows.append({"shape": key, "n": rec["n"], "default_us": d, "best": best, "best_us": rec["t"][best], "all": rec["t"]})
        tot_def += d
        tot_best += rec["t"][best]
    rows.sort(key=lambda r: r["best_us"] - r["default_us"])
    tot_wide = 0.0
    for r in rows:
        m = __import__("re").match(r"M(\d+) N(\d+)", r["shape"])
        N = int(m.group(2))
        ntn0 = -(-N // 256)
        bn0 = -(-(-(-N // ntn0)) // 16) * 16
        tot_wide += r["all"].get("f0 bn%d" % bn0, r["default_us"])
    print("GEMM time per step (back-to-back, warm L2): default (built-in heuristics) %.2f ms, widest-even-tile rule %.2f ms, "
          "best-of-variants %.2f ms" % (tot_def / 1e3, tot_wide / 1e3, tot_best / 1e3))
    for r in rows[:60]:
        print("%-52s n %2d default %7.1f us  best %-12s %7.1f us  gain %6.1f us | %s" % (
            r["shape"], r["n"], r["default_us"], r["best"], r["best_us"], r["default_us"] - r["best_us"],
            " ".join("%s=%.0f" % (k, v) for k, v in sorted(r["all"].items()))))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "gemm_tune_%d_%s.json" % (S, task)), "w"), indent=1)


if __name__ == "__main__":
    main()
