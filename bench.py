#!/usr/bin/env python
"""bench.py — CALM-ViT training images/sec on N B200 GPUs (BASELINE.json's metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps 10 --warmup 3                     # this repo's CUDA path, cls trainer config at 224^2, B=256
    python bench.py --task reg                                         # configs[2]: regression / reconstruction trainer config
    python bench.py --res 384 [--batch 64] | --res 512 [--batch 32]    # configs[3]: longer rows / columns, scaled latent bank
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1     # the UNMODIFIED reference on the host CPU cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...              # data parallel, one rank per GPU

A "step" is one full training step of the trainer loop (distributed_trainer_cls.py:84-96): forward under autocast(bf16),
soft-target cross-entropy (or Huber + 0.1 kl), scaled backward, unscale, clip-grad-norm 1.0, AdamW, zero_grad — with synthetic
ImageNet-shaped data and randomly initialised weights.
  value     : images/sec, inputs resident in HBM, K steps bracketed by barrier + synchronize, CUDA events, max over ranks
  e2e       : same step driven from pinned HOST buffers (H2D of the batch + D2H of the loss inside the timed region)
  roofline  : the dominant kernel (tcgen05 GEMM, one kernel / ~780 launches per step): sum of algorithmic FLOPs / sum of CUDA-event
              launch durations of one step, vs the measured sustained bf16 peak of MEASURED_PEAKS.json; raw and bracket-corrected
  kernels   : the 8 kernel families with the largest share of the step: ms/step, achieved TFLOP/s or GB/s, fraction of peak
  reference_gpu_eager : the UNMODIFIED reference modules (baseline/_ref) run eagerly on the same GPU, same config, same process
  cpu_baseline : the unmodified reference (or the oracle port when it is absent) fwd+bwd on the host cores, a bounded sample
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "calm-vit-dte_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: the driver reads one JSON line
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

# (S, task) -> ViT kwargs; 224: the trainers' own (distributed_trainer_cls.py:148-151, distributed_trainer_reg.py:140-143);
# 384 / 512: SURVEY §8d config 4 with the scaled latent bank ("larger latent bank" of BASELINE.json configs[3])
LATENT = {224: (80, 240), 384: (144, 416), 512: (192, 544)}
DEFAULT_BATCH = {224: 256, 384: 64, 512: 32}
FLOP_PER_IMG = {(224, "cls"): 45.43e9, (224, "reg"): 45.56e9, (384, "cls"): 357.1e9, (512, "cls"): 1003.2e9}   # fwd+bwd


def vit_kwargs(S, task):
    R, M = LATENT[S]
    gen = task == "reg"
    return dict(heads=12, seq_length=S, in_features=3 * S, dim_step=48, mean_var_hidden=M, seq_len_step=16, seq_len_reduce=R,
                out_features=3 * S if gen else 1000, force_reduce=False, generate=gen)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def committed_profile(name):
    """Small JSON summaries of ncu runs committed under profiles/ (per-launch dram bytes, tensor-pipe %)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)
        return False

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 8 and r[4 + i] == "Active" for r in self.rows)]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def synth_batch(B, S, task, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(B, 3, S, S, generator=g)
    y = torch.softmax(torch.randn(B, 1000, generator=g) * 4, -1) if task == "cls" else None   # dense soft labels (CutMix/MixUp)
    return x, y


# ---------------------------------------------------------------------------------------------------------------------
# the unmodified reference (never the product): CPU arm and same-GPU eager arm
# ---------------------------------------------------------------------------------------------------------------------
def load_reference():
    """The reference's own CALM_ViT_V2 module from baseline/_ref (travels to the GPU box) or /root/reference, else None."""
    for d in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference/CALM-ViT"):
        if os.path.exists(os.path.join(d, "CALM_ViT_V2.py")) and os.path.exists(os.path.join(d, "Vi_Tools_CNN_less_V2.py")):
            break
    else:
        return None
    for n in ("matplotlib", "matplotlib.pyplot"):            # CALM_ViT_V2.py:7 — only save_samples uses it; absent in the image
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    names = ("CALM_ViT_V2", "Vi_Tools_CNN_less_V2")
    saved = {n: sys.modules.pop(n) for n in names if n in sys.modules}
    sys.path.insert(0, d)
    try:
        spec = importlib.util.spec_from_file_location("ref_CALM_ViT_V2", os.path.join(d, "CALM_ViT_V2.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(d)
        for n in names:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
    return mod


def reference_cpu_stepper(S, task, Bs):
    """One fwd+bwd of the reference path on the host cores (fp32, the trainers' loss): returns (step_fn, kind, n_threads)."""
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: use every host core anyway
    torch.manual_seed(0)
    kw = vit_kwargs(S, task)
    x, y = synth_batch(Bs, S, task, 2006)
    ref = load_reference()
    if ref is not None:
        model = ref.ViT(torch.device("cpu"), type=8, **kw)
        model.train()

        def one():
            model.zero_grad(set_to_none=True)
            out, kl = model(x)
            if task == "cls":
                loss = torch.nn.CrossEntropyLoss()(out.squeeze(), y)
            else:
                loss = torch.nn.HuberLoss(delta=1.0)(out.reshape(-1, S, S, 3).permute(0, 3, 1, 2), x) + kl * 0.1
            loss.backward()
        return one, "reference", torch.get_num_threads()
    from oracle import calm_oracle as O
    shapes = O.state_shapes(**{k: v for k, v in kw.items() if k != "force_reduce"})
    P = O.params_from_state_dict(O.random_state(shapes))

    def one():
        for p in P.values():
            p.grad = None
        (O.train_step_cls(P, kw["heads"], x, y, training=True) if task == "cls" else O.train_step_reg(P, kw["heads"], x, training=True))
    return one, "port", torch.get_num_threads()


def cpu_baseline(S, task, max_seconds=20.0):
    Bs = 8 if S == 224 else 2
    one, kind, nthr = reference_cpu_stepper(S, task, Bs)
    one()
    ts, t0 = [], time.perf_counter()
    while len(ts) < 2 or (time.perf_counter() - t0 < max_seconds and len(ts) < 8):
        t = time.perf_counter()
        one()
        ts.append(time.perf_counter() - t)
    best = min(ts)
    what = "unmodified reference modules (baseline/_ref)" if kind == "reference" else "oracle port of the reference path"
    return {"value": Bs / best, "unit": "images/sec", "cores": nthr, "kind": kind,
            "sample": "%s fwd+bwd, fp32, batch %d at %d^2 (BASELINE configs[0]), best of %d, %d torch threads" % (what, Bs, S, len(ts), nthr)}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on the host cores; rank 0 only."""
    if rank != 0:
        return
    S, task = args.res, args.task
    Bs = 8 if S == 224 else 2
    one, kind, nthr = reference_cpu_stepper(S, task, Bs)
    for _ in range(max(1, min(args.warmup, 2))):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = (time.perf_counter() - t0) / args.steps
    val = Bs / dt
    what = "unmodified reference modules (baseline/_ref)" if kind == "reference" else "oracle port of the reference path (baseline/_ref absent)"
    line = {"impl": "reference", "metric": "train images/sec at %d^2" % S, "value": val, "unit": "images/sec", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(S, task, args.batch or DEFAULT_BATCH[S]),
                       "sample": "bounded sample of that workload on the host CPU: %s, fwd+bwd in fp32, batch %d per step instead of %d "
                                 "(the per-image work is identical; a CPU step at the full batch would take minutes)" % (what, Bs, args.batch or DEFAULT_BATCH[S]),
                       "same_config": False},
            "cpu_baseline": {"value": val, "unit": "images/sec", "cores": nthr, "kind": kind,
                             "sample": "%s fwd+bwd fp32, batch %d/step, %d steps" % (what, Bs, args.steps)},
            "e2e": {"value": val, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def reference_gpu_eager(dev, S, task, B, steps=3, warm=2):
    """The unmodified reference modules, eager PyTorch under autocast(bf16) + GradScaler + clip + AdamW (the trainers' loop,
    distributed_trainer_cls.py:84-96) on this GPU at the same batch: the number the CUDA path has to beat."""
    ref = load_reference()
    if ref is None:
        return {"unavailable": "baseline/_ref not present"}
    from torch.amp import GradScaler
    torch.manual_seed(0)
    model = ref.ViT(dev, type=8, **vit_kwargs(S, task)).to(dev)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=3.1e-3, weight_decay=0.02, betas=(0.9, 0.98))
    scaler = GradScaler(enabled=True)
    x, y = synth_batch(B, S, task, 2006)
    x = x.to(dev)
    y = y.to(dev) if y is not None else None

    def step():
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            out, kl = model(x)
            if task == "cls":
                loss = torch.nn.CrossEntropyLoss()(out.squeeze(), y)
            else:
                loss = torch.nn.HuberLoss(delta=1.0)(out.reshape(-1, S, S, 3).permute(0, 3, 1, 2), x) + kl * 0.1
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1, error_if_nonfinite=False)
        scaler.step(opt)
        scaler.update()
        opt.zero_grad()
        return loss
    try:
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": B / ms * 1e3, "unit": "images/sec", "ms_per_step": ms, "steps": steps, "warmup": warm, "batch": B,
                "loss": float(loss), "what": "unmodified reference modules (baseline/_ref), eager PyTorch, autocast(bf16) + GradScaler + "
                                             "clip_grad_norm_ + AdamW, same GPU, same config and batch"}
    except Exception as e:      # e.g. out of memory next to the product's pools: report, never fail the bench line
        return {"unavailable": repr(e)[:200]}
    finally:
        del model, opt
        torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------------------------------
# the product arm
# ---------------------------------------------------------------------------------------------------------------------
def workload_name(S, task, B):
    src = "distributed_trainer_%s.py" % task if S == 224 else "SURVEY 8d config 4, scaled latent bank"
    return ("CALM-ViT %s trainer config (%s): %dx%dx3, heads 12, dim %d, latent (%d,%d), per-GPU batch %d, full training step "
            "(fwd+loss+bwd+unscale+clip+AdamW)" % (task, src, S, S, 3 * S, LATENT[S][0], LATENT[S][1], B))


def bracket_overhead_us(n=64):
    """What a (start event, launch, end event) bracket adds to a kernel compared with the same launch inside a CUDA graph,
    measured on a trivial kernel (calm_cast_bf16 of 8 elements): median bracketed time - per-launch time of a graph of n launches."""
    import calm_kernels as K
    x = torch.zeros(8, device="cuda")
    K.cast_bf16(x)
    torch.cuda.synchronize()
    torch.cuda._sleep(int(2e8))                       # same head start as the profiling step: the GPU never waits for the host
    ev = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); K.cast_bf16(x); b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    bracketed = t[len(t) // 2] * 1e3
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        K.cast_bf16(x)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            K.cast_bf16(x)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    in_graph = a.elapsed_time(b) / n * 1e3
    del g
    return max(0.0, bracketed - in_graph), bracketed, in_graph


class Trainer:
    """The per-rank training step of the reference loop, captured into one CUDA graph (static shapes)."""

    def __init__(self, model, task, dev, B, S, use_graph, torch_glue=False):
        self.model, self.task, self.dev, self.S = model, task, dev, S
        self.params = [p for p in model.parameters()]
        self.torch_glue = torch_glue
        if torch_glue:      # A/B arm: the reference loop's own torch objects (GradScaler, clip_grad_norm_, fused AdamW, F.* losses)
            from torch.amp import GradScaler
            self.opt = torch.optim.AdamW(self.params, lr=3.1e-3, weight_decay=0.02, betas=(0.9, 0.98), fused=True, capturable=True)
            self.scaler = GradScaler(enabled=True)
        else:               # product: calm_trainer (loss heads + unscale/clip/AdamW/scale-update kernels of libcalm_b200.so)
            import calm_trainer
            self.ct = calm_trainer
            self.glue = calm_trainer.TrainerStep(self.params, lr=3.1e-3, weight_decay=0.02, betas=(0.9, 0.98), max_norm=1.0)
        self.x = torch.zeros(B, 3, S, S, device=dev)
        self.y = torch.zeros(B, 1000, device=dev) if task == "cls" else None
        self.loss = torch.zeros((), device=dev)
        self.graphed = None
        self.use_graph = use_graph

    def _step(self):
        S = self.S
        if not self.torch_glue:
            with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
                y_hat, kl = self.model(self.x)
            if self.task == "cls":
                loss, _acc = self.ct.soft_target_cross_entropy(y_hat.squeeze(), self.y)
            else:
                loss, _h = self.ct.huber_kl_loss(y_hat, self.x, kl, 0.1, 1.0)
            self.glue.backward(loss)
            self.glue.step()
            self.glue.zero_grad()
            self.loss.copy_(loss.detach())
            return
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            y_hat, kl = self.model(self.x)
            if self.task == "cls":
                loss = torch.nn.functional.cross_entropy(y_hat.squeeze(), self.y)
            else:
                img = y_hat.reshape(-1, S, S, 3).permute(0, 3, 1, 2)
                loss = torch.nn.functional.huber_loss(img, self.x, delta=1.0) + kl * 0.1
        self.scaler.scale(loss).backward()
        self.scaler.unscale_(self.opt)
        torch.nn.utils.clip_grad_norm_(self.params, max_norm=1, error_if_nonfinite=False)
        self.scaler.step(self.opt)
        self.scaler.update()
        self.opt.zero_grad()
        self.loss.copy_(loss.detach())

    def capture(self):
        """3 eager steps on a side stream (allocator + lazy state warm-up), then one step captured into a CUDA graph: the
        product's own helper (calm_trainer.GraphedStep), which is what a user of the drop-in loop calls. A capture failure is
        fatal: a silently eager run would be reported as a slow step, not as the regression it is."""
        import calm_trainer
        self.graphed = calm_trainer.GraphedStep(self._step, warmup=3, capture=self.use_graph)

    def step(self):
        self.graphed()

    def release(self):
        if self.graphed is not None:
            self.graphed.release()


HBM_FAMILIES = ("calm_layernorm_fwd", "calm_layernorm_bwd", "calm_cnn_fwd", "calm_cnn_bwd", "calm_rope_fwd", "calm_rope_bwd",
                "calm_token_transpose", "calm_colsum", "calm_trainer_step", "calm_soft_ce_fwd", "calm_soft_ce_bwd")
TENSOR_FAMILIES = ("calm_gemm", "calm_attention_fwd", "calm_attention_bwd")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default 256 at 224^2, 64 at 384^2, 32 at 512^2)")
    ap.add_argument("--task", default="cls", choices=["cls", "reg"])
    ap.add_argument("--res", type=int, default=224, choices=[224, 384, 512])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--no-comm-split", action="store_true", help="N > 1: skip the second capture that times the step without all-reduces")
    ap.add_argument("--torch-glue", action="store_true", help="A/B: torch GradScaler/clip/AdamW/loss instead of calm_trainer")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.res != 224 and args.task != "cls":
        ap.error("--res 384/512 is the classification config (BASELINE configs[3])")
    if args.impl == "reference":
        run_reference(args, rank)
        return
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    S, task = args.res, args.task
    B = args.batch or DEFAULT_BATCH[S]
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner to stdout when the communicator is created; the driver reads ONE JSON line from
        # stdout, so fd 1 points at stderr until the communicator exists (first collective)
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group(backend="nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    import calm_lib
    import CALM_ViT_V2 as rvh
    torch.manual_seed(0)                      # identical initial weights on every rank (then broadcast from rank 0 anyway)
    model = rvh.ViT(dev, type=8, **vit_kwargs(S, task)).to(dev)
    model.train()
    # N > 1: the bucketed NCCL all-reduces (calm_ddp.DataParallel, side stream + events) are captured with the step
    use_graph = not args.no_graph
    wrapped = model
    if world > 1:
        from calm_ddp import DataParallel
        wrapped = DataParallel(model)
    tr = Trainer(wrapped, task, dev, B, S, use_graph, torch_glue=args.torch_glue)
    xh, yh = synth_batch(B, S, task, 2006 + rank)
    xh, yh = xh.pin_memory(), (yh.pin_memory() if yh is not None else None)
    tr.x.copy_(xh)
    if yh is not None:
        tr.y.copy_(yh)
    torch.manual_seed(1234 + rank)            # per-rank latent noise stream (the reference never seeds inside train())
    tr.capture()                              # raises on failure: no silent eager fallback
    for _ in range(max(args.warmup, 3)):
        tr.step()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------------------------
    n0 = calm_lib.launch_count
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for _ in range(args.steps):
            tr.step()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    launches_py = calm_lib.launch_count - n0
    # ---- end-to-end from pinned host buffers ----------------------------------------------------------------------
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # Input pipeline of the end-to-end loop: the pinned host batch of step i + 1 crosses PCIe on a copy stream while step i
    # computes (one H2D copy per step, all inside the timed region; the first one is exposed), then a device-to-device copy
    # hands it to the static input buffer of the captured step; the loss of every step is read back on the host.
    copy_stream = torch.cuda.Stream()
    x_stage = torch.empty_like(tr.x)
    y_stage = torch.empty_like(tr.y) if yh is not None else None

    def h2d_async():
        with torch.cuda.stream(copy_stream):
            x_stage.copy_(xh, non_blocking=True)
            if yh is not None:
                y_stage.copy_(yh, non_blocking=True)
            return copy_stream.record_event()
    torch.cuda.synchronize()
    f0.record()
    last = 0.0
    landed = h2d_async()
    for i in range(args.steps):
        main_stream = torch.cuda.current_stream()
        main_stream.wait_event(landed)
        tr.x.copy_(x_stage, non_blocking=True)
        if yh is not None:
            tr.y.copy_(y_stage, non_blocking=True)
        if i + 1 < args.steps:
            copy_stream.wait_event(main_stream.record_event())   # the staging buffers are free again
            landed = h2d_async()
        tr.step()
        last = tr.loss.item()                 # device -> host read of the step's result
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    ms_step = ms / args.steps
    value = B * world * args.steps / (ms / 1e3)
    e2e_value = B * world * args.steps / (ms_e2e / 1e3)
    pk = peaks()
    peak_tf = (pk or {}).get("bf16_tflops_sustained", 1400.0)
    peak_gb = (pk or {}).get("hbm_gbs", 6500.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if pk else "fallback (B200_PROFILING.md)"

    # ---- one extra eager step with CUDA events around every C-ABI launch: per-kernel-family time and work -----------
    roof, breakdown, kernels, launches_per_step = None, None, None, None
    if not args.no_profile:
        import calm_ops
        calm_ops.PARALLEL = False            # kernels are timed one at a time (the captured step overlaps independent ones)
        calm_lib.profile = []
        n1 = calm_lib.launch_count
        # The eager launch path (Python + ctypes) is slower than the small kernels: without a head start every interval between
        # two events would also contain the GPU waiting for the next launch. A ~0.3 s spin kernel lets the host enqueue the whole
        # step first, so the events bracket back-to-back kernels, as in the captured graph.
        torch.cuda._sleep(int(6e8))
        tr._step()
        torch.cuda.synchronize()
        launches_per_step = calm_lib.launch_count - n1
        prof, calm_lib.profile = calm_lib.profile, None
        calm_ops.PARALLEL = True
        fam, shapes, gemm_bytes = {}, {}, 0
        for name, work, a, b, *tag in prof:
            dt = a.elapsed_time(b)
            if tag and tag[0] and "bytes=" in tag[0]:
                gemm_bytes += int(tag[0].rsplit("bytes=", 1)[1])
                tag = [tag[0].rsplit(" bytes=", 1)[0]]
            if tag and tag[0]:
                d = shapes.setdefault(tag[0], {"ms": 0.0, "work": 0.0, "n": 0})
                d["ms"] += dt; d["work"] += work; d["n"] += 1
            d = fam.setdefault(name, {"ms": 0.0, "work": 0.0, "n": 0})
            d["ms"] += dt; d["work"] += work; d["n"] += 1
        tot = sum(d["ms"] for d in fam.values())
        top_shapes = sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])
        breakdown_shapes = [{"shape": k, "ms": round(v["ms"], 3), "n": v["n"], "tflops": round(v["work"] / (v["ms"] * 1e9), 1)} for k, v in top_shapes]
        breakdown = {k: {"ms": round(v["ms"], 3), "n": v["n"], "share": round(v["ms"] / tot, 4),
                         "rate": (v["work"] / (v["ms"] * 1e-3) if v["work"] and v["ms"] > 0 else None)} for k, v in fam.items()}
        try:
            ov, t_b, t_g = bracket_overhead_us()
        except Exception:
            ov, t_b, t_g = 0.0, 0.0, 0.0
        ncu = committed_profile("r02_ncu_summary.json") or {}
        kernels = []
        for name, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])[:8]:
            ms_corr = max(v["ms"] - v["n"] * ov * 1e-3, 0.25 * v["ms"])
            e = {"family": name, "launches": v["n"], "ms_per_step_event_bracketed": round(v["ms"], 3), "ms_per_step": round(ms_corr, 3),
                 "share_of_kernel_time": round(v["ms"] / tot, 4)}
            if v["work"] > 0:
                if name in TENSOR_FAMILIES:
                    r = v["work"] / (ms_corr * 1e-3) / 1e12
                    e.update({"bound": "tensor", "achieved": round(r, 1), "unit": "TFLOP/s", "frac": round(r / peak_tf, 4)})
                elif name in HBM_FAMILIES:
                    r = v["work"] / (ms_corr * 1e-3) / 1e9
                    e.update({"bound": "hbm", "achieved": round(r, 1), "unit": "GB/s (algorithmic bytes)", "frac": round(r / peak_gb, 4)})
            if name in ncu.get("tensor_pipe_pct", {}):
                e["ncu_tensor_pipe_pct"] = ncu["tensor_pipe_pct"][name]
            kernels.append(e)
        g = fam.get("calm_gemm")
        if g and g["ms"] > 0:
            raw = g["work"] / (g["ms"] * 1e-3) / 1e12
            ms_corr = max(g["ms"] - g["n"] * ov * 1e-3, 0.25 * g["ms"])
            ach = g["work"] / (ms_corr * 1e-3) / 1e12
            gt = ncu.get("gemm_dram_bytes_per_launch")
            roof = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (calm_gemm, %d launches/step, %.1f%% of kernel time)" % (g["n"], 100 * g["ms"] / tot),
                    "achieved": ach, "achieved_event_bracketed": raw, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                    "frac_event_bracketed": raw / peak_tf, "traffic": gt,
                    "algorithmic_bytes_per_launch": gemm_bytes / g["n"],
                    "traffic_note": ("traffic = average dram__bytes_read.sum + dram__bytes_write.sum per gemm_tcgen05_kernel launch over the launches of "
                                     "one training step under ncu (profiles/r02_ncu_summary.json, %s launches); algorithmic_bytes_per_launch = operands "
                                     "read once + results written once, same average, counted live" % ncu.get("gemm_launches")) if gt
                                    else "no committed ncu dram capture found (profiles/r02_ncu_summary.json)",
                    "flop_per_launch": g["work"] / g["n"], "bracket_overhead_us": round(ov, 2), "family_ms_per_step": round(ms_corr, 3),
                    "peak_source": peak_src + " bf16_tflops_sustained",
                    "timing": "CUDA events around every launch of one eager step (same stream), sum of FLOPs / sum of durations; `achieved` removes the "
                              "bracket overhead measured on a trivial kernel (bracketed %.1f us vs %.1f us per launch inside a CUDA graph), "
                              "`achieved_event_bracketed` is the raw figure" % (t_b, t_g)}
    flops_img = FLOP_PER_IMG[(S, task)]
    line = {
        "metric": "train images/sec at %d^2" % S, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(S, task, B), "parallelism": "dp%d" % world, "global_batch": B * world,
                   "cuda_graph": tr.graphed.graph is not None,
                   "l2": "inputs larger than L2 (batch %d MB, activations GBs per step); no explicit flush" % (B * 3 * S * S * 4 // 2 ** 20)},
        "model_flops_frac_of_bf16_peak": value * flops_img / world / (peak_tf * 1e12),
        "model_flops_frac_of_bf16_burst_peak": value * flops_img / world / ((pk or {}).get("bf16_tflops", 1650.0) * 1e12),
        "e2e": {"value": e2e_value, "unit": "images/sec", "h2d_bytes_per_step": xh.numel() * 4 + (yh.numel() * 4 if yh is not None else 0),
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": (launches_per_step or launches_py // max(args.steps, 1)) * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "clocks": clocks.summary(), "loss": last,
    }
    if roof:
        line["roofline"] = roof
    if kernels:
        line["kernels"] = kernels
        att = (committed_profile("r02_ncu_summary.json") or {}).get("attention_tensor_pipe_pct_by_head_dim")
        if att:
            line["attention_tensor_pipe_pct_by_head_dim"] = att
    if world > 1:
        line["buffers_in_sync"] = wrapped.check_buffers()
    tr.release()                               # destroy the captured graph (it references the NCCL communicator)
    if world > 1 and use_graph and not args.no_comm_split:
        # What the gradient exchange costs per step at this N: the same step captured again WITHOUT the all-reduces (every rank
        # trains on alone for these few steps), timed the same way; the difference is the exposed communication plus what the
        # NCCL kernels take from the compute they overlap with.
        wrapped.require_sync = False
        tr2 = Trainer(wrapped, task, dev, B, S, True, torch_glue=args.torch_glue)
        tr2.x.copy_(tr.x)
        if tr.y is not None:
            tr2.y.copy_(tr.y)
        del tr
        torch.cuda.empty_cache()
        tr2.capture()
        for _ in range(3):
            tr2.step()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            tr2.step()
        g1.record()
        barrier()
        t2 = torch.tensor([g0.elapsed_time(g1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms_alone = t2.item() / args.steps
        tr2.release()
        grad_bytes = 4 * sum(p.numel() for p in model.parameters())
        line["comm"] = {"ms_per_step_without_allreduce": ms_alone, "allreduce_cost_ms_per_step": ms_step - ms_alone,
                        "gradient_bytes_per_step": grad_bytes, "buckets": len(wrapped.buckets), "nccl_max_ctas": wrapped.comm_ctas or None,
                        "note": "same captured step with calm_ddp's all-reduces switched off, same process, same timing; the difference "
                                "is exposed communication + SM / HBM interference of the NCCL kernels"}
    if rank == 0:
        if breakdown is not None:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            json.dump(breakdown, open(os.path.join(ROOT, "gpurun_out", "bench_kernel_breakdown.json"), "w"), indent=1)
            json.dump(breakdown_shapes, open(os.path.join(ROOT, "gpurun_out", "bench_gemm_shapes.json"), "w"), indent=1)
        if world == 1 and not args.no_reference_gpu:
            tr = wrapped = model = None
            torch.cuda.empty_cache()
            line["reference_gpu_eager"] = reference_gpu_eager(dev, S, task, B)
            if "value" in line["reference_gpu_eager"]:
                line["speedup_vs_reference_gpu_eager"] = value / line["reference_gpu_eager"]["value"]
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(S, task)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
