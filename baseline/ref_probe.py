"""Baseline harness: times the UNMODIFIED reference modules (no product code here).

Imports the reference's CALM_ViT_V2.ViT from baseline/_ref/ (a verbatim, gitignored copy of
/root/reference/CALM-ViT/{Vi_Tools_CNN_less_V2,CALM_ViT_V2}.py) or from /root/reference directly,
and reports, for the BASELINE.json configs:
  * CPU fp32 fwd+bwd img/s (config 1), with the core count
  * eager-PyTorch bf16-autocast training-step img/s on one GPU (configs 2-4)
  * a kernel-level profile of one GPU step (where the reference's time goes today)
Results go to gpurun_out/ref_probe.json and gpurun_out/ref_probe_profile.txt.
"""
import json, os, sys, time, types

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(HERE, "_ref"), "/root/reference/CALM-ViT"):
    if os.path.exists(os.path.join(p, "CALM_ViT_V2.py")):
        sys.path.insert(0, p)
        break
# CALM_ViT_V2.py:7 imports matplotlib (only used by save_samples); absent in this image.
_mpl = types.ModuleType("matplotlib"); _plt = types.ModuleType("matplotlib.pyplot"); _mpl.pyplot = _plt
sys.modules.setdefault("matplotlib", _mpl); sys.modules.setdefault("matplotlib.pyplot", _plt)

import torch  # noqa: E402
import CALM_ViT_V2 as rvh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "gpurun_out")
os.makedirs(OUT, exist_ok=True)
res = {"torch": torch.__version__, "cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads()}


def build(device, S=224, generate=False, sstep=16, R=80, M=240):
    # distributed_trainer_cls.py:148-151 / distributed_trainer_reg.py:140-143
    return rvh.ViT(device, type=8, heads=12, seq_length=S, in_features=3 * S, dim_step=3 * sstep,
                   mean_var_hidden=M, seq_len_step=sstep, seq_len_reduce=R,
                   out_features=1000 if not generate else 3 * S, force_reduce=False, generate=generate).to(device)


def make_step(model, device, B, S, generate, use_amp):
    from torch.amp import autocast, GradScaler
    opt = torch.optim.AdamW(model.parameters(), lr=3.1e-3, weight_decay=0.02, betas=(0.9, 0.98))
    scaler = GradScaler(enabled=use_amp)
    g = torch.Generator(device="cpu").manual_seed(2006)
    x = torch.randn(B, 3, S, S, generator=g).to(device)
    y = torch.softmax(torch.randn(B, 1000, generator=g) * 4, -1).to(device)  # soft labels like CutMix/MixUp
    ce = torch.nn.CrossEntropyLoss(); hub = torch.nn.HuberLoss(delta=1.0)

    def step():
        with autocast(device_type="cuda", enabled=use_amp, dtype=torch.bfloat16):
            y_hat, kl = model(x)
            if generate:
                img = y_hat.reshape(-1, S, S, 3).permute(0, 3, 1, 2)
                loss = hub(img, x) + kl * 0.1
            else:
                loss = ce(y_hat.squeeze(), y)
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1, error_if_nonfinite=False)
        scaler.step(opt); scaler.update(); opt.zero_grad()
        return loss
    return step


def time_gpu(tag, S, B, generate, warm=3, iters=8, profile=False):
    dev = torch.device("cuda")
    torch.manual_seed(0)
    try:
        model = build(dev, S=S, generate=generate); model.train()
        step = make_step(model, dev, B, S, generate, True)
        for _ in range(warm): loss = step()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): loss = step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res[tag] = {"S": S, "B": B, "generate": generate, "ms_per_step": ms, "img_per_s": B / ms * 1e3,
                    "loss": float(loss), "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30}
        print(tag, res[tag], flush=True)
        if profile:
            from torch.profiler import profile as prof_, ProfilerActivity
            with prof_(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
                step(); torch.cuda.synchronize()
            ka = p.key_averages()
            with open(os.path.join(OUT, "ref_probe_profile.txt"), "w") as f:
                f.write(ka.table(sort_by="self_cuda_time_total", row_limit=80, max_name_column_width=110))
            kern = [(k.key, k.self_device_time_total, k.count) for k in ka if k.self_device_time_total > 0 and getattr(k, "device_type", None) is not None]
            kern.sort(key=lambda t: -t[1])
            tot = sum(t[1] for t in kern)
            res[tag + "_profile"] = {"total_device_us": tot, "n_kernel_launches": sum(t[2] for t in kern),
                                     "top": [{"name": k[:120], "us": us, "count": c, "pct": 100 * us / max(tot, 1)} for k, us, c in kern[:45]]}
        del model, step
    except Exception as e:  # OOM etc.
        res[tag] = {"S": S, "B": B, "error": repr(e)[:300]}
        print(tag, "FAILED", repr(e)[:300], flush=True)
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()


def time_cpu(B=8, S=224):
    dev = torch.device("cpu"); torch.manual_seed(0)
    model = build(dev, S=S, generate=False); model.train()
    x = torch.randn(B, 3, S, S); y = torch.randint(0, 1000, (B,)); ce = torch.nn.CrossEntropyLoss()
    def step():
        yh, _ = model(x); loss = ce(yh.squeeze(), y); loss.backward(); model.zero_grad(set_to_none=True); return float(loss)
    step(); ts = []
    for _ in range(3):
        t = time.perf_counter(); step(); ts.append(time.perf_counter() - t)
    res["cpu_fp32_fwd_bwd"] = {"B": B, "S": S, "sec_per_step_best": min(ts), "img_per_s": B / min(ts), "threads": torch.get_num_threads()}
    print("cpu", res["cpu_fp32_fwd_bwd"], flush=True)


if __name__ == "__main__":
    if torch.cuda.is_available():
        res["gpu"] = torch.cuda.get_device_name(0); res["n_gpu"] = torch.cuda.device_count()
        time_gpu("ref_gpu_cls_224_b256", 224, 256, False, profile=True)
        time_gpu("ref_gpu_reg_224_b256", 224, 256, True)
        time_gpu("ref_gpu_cls_384_b64", 384, 64, False, warm=2, iters=4)
        time_gpu("ref_gpu_cls_512_b32", 512, 32, False, warm=2, iters=4)
    time_cpu()
    with open(os.path.join(OUT, "ref_probe.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: v for k, v in res.items() if not k.endswith("_profile")}, indent=1))
