"""Device-side training-step glue of the reference's per-rank loop.

The reference loop (distributed_trainer_cls.py:84-102, distributed_trainer_reg.py:76-98) is

    loss = criterion(y_hat.squeeze(), y)            |  loss = huber(img, x) + kl_loss * 0.1
    scaler.scale(loss).backward()
    scaler.unscale_(optimizer)
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1, error_if_nonfinite=False)
    scaler.step(optimizer); scaler.update(); optimizer.zero_grad()
    epoch_loss += loss.item(); ... accuracy via two torch.max + .item()

Here the same arithmetic runs as a handful of kernels of libcalm_b200.so with every scalar (loss scale, step count,
learning rate, gradient norm, skip flag, loss, accuracy) resident on the device:

    step = TrainerStep(model.parameters(), lr=3.1e-3, weight_decay=0.02, betas=(0.9, 0.98))
    loss, acc = soft_target_cross_entropy(y_hat.squeeze(), y)        # or huber_kl_loss(y_hat, x, kl_loss, 0.1)
    step.backward(loss)                                               # = scaler.scale(loss).backward()
    step.step()                                                       # unscale + clip + AdamW + scale update, 3 launches
    step.zero_grad()

No `.item()` is needed inside the loop; `float(loss)` when the loop wants to print. Plumbing only (tensors, pointer
tables); there is no PyTorch fallback for any of it.
"""
import ctypes as C
import math
import weakref

import torch
from torch.autograd import Function

import calm_lib as L
from calm_lib import ptr

f32 = torch.float32


# ---------------------------------------------------------------------------------------------------- loss heads
class _SoftCeFn(Function):
    @staticmethod
    def forward(ctx, logits, target):
        if logits.dim() != 2:
            raise L.CalmError("soft_target_cross_entropy expects (B, C) logits, got %s" % (tuple(logits.shape),))
        x = logits.float()
        if x.stride(1) != 1:
            x = x.contiguous()
        B, Cn = x.shape
        if target.dtype == torch.int64:
            if target.shape != (B,):
                raise L.CalmError("class-index target must have shape (B,)")
            t, lab, ld_t = None, target.contiguous(), 0
        else:
            if target.shape != x.shape:
                raise L.CalmError("probability target must have the logits' shape")
            t = target.float()
            if t.stride(1) != 1:
                t = t.contiguous()
            lab, ld_t = None, t.stride(0)
        stats = torch.empty(B, 4, dtype=f32, device=x.device)
        out = torch.empty(2, dtype=f32, device=x.device)
        L.call("calm_soft_ce_fwd", ptr(x), x.stride(0), ptr(t), ld_t, ptr(lab), ptr(stats), ptr(out), B, Cn,
               work=8.0 * B * Cn)
        ctx.save_for_backward(x, t if t is not None else lab, stats)
        ctx.index_target = t is None
        ctx.in_dtype = logits.dtype
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, dloss, _dout):
        x, tgt, stats = ctx.saved_tensors
        B, Cn = x.shape
        t, lab = (None, tgt) if ctx.index_target else (tgt, None)
        d = torch.empty(B, Cn, dtype=f32, device=x.device)
        g = dloss.float().contiguous()
        L.call("calm_soft_ce_bwd", ptr(x), x.stride(0), ptr(t), t.stride(0) if t is not None else 0, ptr(lab), ptr(stats),
               ptr(g), ptr(d), Cn, B, Cn, work=12.0 * B * Cn)
        return (d if ctx.in_dtype == f32 else d.to(ctx.in_dtype)), None


def soft_target_cross_entropy(logits, target):
    """`torch.nn.CrossEntropyLoss()(logits, target)` for probability targets (CutMix / MixUp labels) or int64 class indices
    (distributed_trainer_cls.py:63,86), fp32 like autocast's policy for cross_entropy. Returns (loss, accuracy): 0-dim
    tensors on the device; accuracy is the dominant-class accuracy the loop prints (:97-100)."""
    loss, out = _SoftCeFn.apply(logits, target)
    return loss, out[1]


class _HuberKlFn(Function):
    @staticmethod
    def forward(ctx, tokens, image, kl, kl_weight, delta):
        B = image.shape[0]
        S = image.shape[-1]
        if image.shape != (B, 3, S, S) or tokens.numel() != image.numel():
            raise L.CalmError("huber_kl_loss: tokens %s do not match the image %s" % (tuple(tokens.shape), tuple(image.shape)))
        t = tokens.float().contiguous()
        img = image.float().contiguous()
        nparts = L.load().calm_huber_parts(B, S)
        partial = torch.empty(nparts, dtype=f32, device=t.device)
        out = torch.empty(2, dtype=f32, device=t.device)
        klf = kl.float().contiguous() if kl is not None else None
        L.call("calm_huber_tokens_fwd", ptr(t), ptr(img), ptr(klf), float(kl_weight), float(delta), ptr(partial), nparts,
               ptr(out), B, S, work=8.0 * t.numel())
        ctx.save_for_backward(t, img)
        ctx.meta = (B, S, float(kl_weight), float(delta), tokens.shape, tokens.dtype, kl is not None, kl.dtype if kl is not None else None)
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, dloss, _dout):
        t, img = ctx.saved_tensors
        B, S, klw, delta, shape, dtype, has_kl, kl_dtype = ctx.meta
        d = torch.empty_like(t)
        dkl = torch.empty((), dtype=f32, device=t.device) if has_kl else None
        g = dloss.float().contiguous()
        L.call("calm_huber_tokens_bwd", ptr(t), ptr(img), ptr(g), klw, delta, ptr(d), ptr(dkl), B, S, work=12.0 * t.numel())
        d = d.view(shape)
        if dtype != f32:
            d = d.to(dtype)
        if has_kl and kl_dtype != f32:
            dkl = dkl.to(kl_dtype)
        return d, None, dkl, None, None


def huber_kl_loss(tokens, image, kl=None, kl_weight=0.1, delta=1.0):
    """`HuberLoss(delta)(y_hat.reshape(-1,S,S,3).permute(0,3,1,2), x) + kl_loss * kl_weight`
    (distributed_trainer_reg.py:76-88) without materialising the NCHW permute. Returns (loss, huber_term)."""
    loss, out = _HuberKlFn.apply(tokens, image, kl, kl_weight, delta)
    return loss, out[1]


# ---------------------------------------------------------------------------------------------------- whole-step CUDA graph
class GraphedStep:
    """One training step captured into a CUDA graph and replayed.

    The loop launches ~1,210 kernels per step through Python; run eagerly the drop-in path is bound by that launch
    path (55 - 90 ms/step on the trainer config, by host CPU) rather than by the GPU (~47 ms). All shapes of the path are static, so the whole
    step — forward, loss head, backward, `TrainerStep.step()`, `zero_grad()` — can be captured once:

        x_dev, y_dev = torch.empty(...), torch.empty(...)            # static input buffers the step function reads
        def one_step():
            with autocast("cuda", dtype=torch.bfloat16):
                y_hat, kl = model(x_dev)
            loss, acc = soft_target_cross_entropy(y_hat.squeeze(), y_dev)
            step.backward(loss); step.step(); step.zero_grad()
            loss_dev.copy_(loss.detach())
        graphed = GraphedStep(one_step)                               # 3 eager warm-up steps on a side stream, then capture
        for x, y in dataloader:
            x_dev.copy_(x, non_blocking=True); y_dev.copy_(y, non_blocking=True)
            graphed()                                                 # one graph launch

    Warm-up steps are real steps (they update the parameters), exactly like the first iterations of the loop. The latent noise
    (`torch.randn`) is drawn inside the graph from the CUDA generator torch registers with the capture, so every replay draws
    new noise. `bench.py` runs its timed region through this class.
    """

    def __init__(self, step_fn, warmup=3, capture=True):
        if not torch.cuda.is_available():
            raise L.CalmError("GraphedStep needs a CUDA device; there is no CPU fallback")
        self.step_fn = step_fn
        self.graph = None
        _live_graphs.add(self)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step_fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if capture:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step_fn()
            self.graph = g

    def __call__(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.step_fn()

    def release(self):
        """Destroys the captured graph (the step function stays callable eagerly). A graph that captured NCCL all-reduces
        (calm_ddp.DataParallel) keeps the communicator referenced: `dist.destroy_process_group()` waits for every such graph to
        be destroyed, so call this (or `release_graphs()`) first — otherwise the teardown never returns."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None


_live_graphs = weakref.WeakSet()


def release_graphs():
    """`GraphedStep.release()` on every live captured step of this process (call before `dist.destroy_process_group()`)."""
    for g in list(_live_graphs):
        g.release()


# ---------------------------------------------------------------------------------------------------- input side
class MixBatch:
    """`transforms.RandomChoice([CutMix(num_classes=1000, alpha=1.0), MixUp(num_classes=1000, alpha=0.8)])` of the reference's
    collate function (distributed_trainer_cls.py:58-61) applied to a batch that already lives on the device: one kernel
    mixes the images (no `roll` / `clone` copies), one writes the (B, num_classes) soft labels.

    `draw(H, W)` samples the parameters on the host with the torch CPU generator in torchvision's own order (multinomial for
    the choice, Beta for lam, then two randint for the CutMix box centre), so a seeded run picks what torchvision would pick.
    """

    def __init__(self, num_classes=1000, cutmix_alpha=1.0, mixup_alpha=0.8):
        self.num_classes = int(num_classes)
        beta = torch.distributions.Beta
        self._dist = [beta(torch.tensor([float(cutmix_alpha)]), torch.tensor([float(cutmix_alpha)])),
                      beta(torch.tensor([float(mixup_alpha)]), torch.tensor([float(mixup_alpha)]))]

    def draw(self, H, W):
        idx = int(torch.multinomial(torch.tensor([0.5, 0.5]), 1))            # RandomChoice.forward
        lam = float(self._dist[idx].sample(()))
        if idx == 1:                                                          # MixUp.make_params
            return {"mode": 0, "lam": lam, "box": (0, 0, 0, 0), "lam_labels": lam}
        r_x = torch.randint(W, size=(1,))                                     # CutMix.make_params
        r_y = torch.randint(H, size=(1,))
        r = 0.5 * math.sqrt(1.0 - lam)
        r_w_half, r_h_half = int(r * W), int(r * H)
        x1 = int(torch.clamp(r_x - r_w_half, min=0))
        y1 = int(torch.clamp(r_y - r_h_half, min=0))
        x2 = int(torch.clamp(r_x + r_w_half, max=W))
        y2 = int(torch.clamp(r_y + r_h_half, max=H))
        return {"mode": 1, "lam": lam, "box": (x1, y1, x2, y2), "lam_labels": float(1.0 - (x2 - x1) * (y2 - y1) / (W * H))}

    def __call__(self, images, labels, params=None):
        """images f32 (B, C, H, W) and int64 labels (B,) on the device -> (mixed images, soft labels (B, num_classes))."""
        if images.dim() != 4 or images.dtype != f32:
            raise L.CalmError("MixBatch expects a float32 (B, C, H, W) batch")
        if labels.dtype != torch.int64 or labels.shape != (images.shape[0],):
            raise L.CalmError("MixBatch expects int64 class indices of shape (B,)")
        B, Cc, H, W = images.shape
        p = params if params is not None else self.draw(H, W)
        x = images.contiguous()
        out = torch.empty_like(x)
        soft = torch.empty(B, self.num_classes, dtype=f32, device=x.device)
        x1, y1, x2, y2 = p["box"]
        L.call("calm_mix_batch", ptr(x), ptr(labels.contiguous()), ptr(out), ptr(soft), B, Cc, H, W, self.num_classes, int(p["mode"]),
               float(p["lam"]), 1.0 - float(p["lam"]), x1, y1, x2, y2, float(p["lam_labels"]), 1.0 - float(p["lam_labels"]),
               work=8.0 * x.numel())
        return out, soft


def cosine_annealing_lr(epoch, base_lr, T_max, eta_min=1e-6):
    """Learning rate after `epoch` calls of `CosineAnnealingLR(optimizer, T_max, eta_min).step()` (closed form of
    torch.optim.lr_scheduler.CosineAnnealingLR; the reference builds it with T_max=epochs, eta_min=1e-6 and steps it once per
    epoch, distributed_trainer_cls.py:52,108-109)."""
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * epoch / T_max)) / 2.0


# ---------------------------------------------------------------------------------------------------- optimizer step
class TrainerStep:
    """GradScaler + clip_grad_norm_ + AdamW of the reference loop as one object (distributed_trainer_cls.py:64,88-96,158).

    state (device, fp32): loss scale, growth tracker, step count, learning rate, last gradient norm, last skip flag.
    `step()` launches three kernels and never synchronises; it can be captured in a CUDA graph — the gradient pointer table
    is re-uploaded (one small async copy) only when a `.grad` tensor changed its address.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=1.0, use_scaler=True,
                 init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise L.CalmError("TrainerStep: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise L.CalmError("TrainerStep needs CUDA parameters (got %s); there is no CPU fallback" % dev)
        for p in self.params:
            if p.dtype != f32 or not p.is_contiguous() or p.device != dev:
                raise L.CalmError("TrainerStep: parameters must be contiguous fp32 tensors on one device")
        L.load()
        self.device = dev
        self.hyper = dict(beta1=float(betas[0]), beta2=float(betas[1]), eps=float(eps), weight_decay=float(weight_decay),
                          max_norm=float(max_norm) if max_norm else 0.0, growth_factor=float(growth_factor),
                          backoff_factor=float(backoff_factor), growth_interval=int(growth_interval), use_scaler=int(bool(use_scaler)))
        numels = [p.numel() for p in self.params]
        off = [0]
        for n in numels:
            off.append(off[-1] + n)
        chunk_tensor, chunk_start = [], []
        for i, n in enumerate(numels):
            for j in range(-(-n // L.OPT_CHUNK)):
                chunk_tensor.append(i)
                chunk_start.append(j)
        self.n_tensors, self.n_chunks, self.total = len(numels), len(chunk_tensor), off[-1]
        self._off = off
        i64, i32 = torch.int64, torch.int32
        self.elem_off = torch.tensor(off, dtype=i64, device=dev)
        self.chunk_tensor = torch.tensor(chunk_tensor, dtype=i32, device=dev)
        self.chunk_start = torch.tensor(chunk_start, dtype=i32, device=dev)
        self._param_ptr_list = [p.data_ptr() for p in self.params]
        self.param_ptrs = torch.tensor(self._param_ptr_list, dtype=i64, device=dev)
        self.grad_ptrs = torch.zeros(self.n_tensors, dtype=i64, device=dev)
        self._grad_ptr_list = None
        # pinned staging for the gradient pointer table: two alternate in eager mode (each guarded by an event), a third is
        # dedicated to CUDA-graph capture (the captured copy node re-reads it at every replay, so it is never rewritten)
        self._stage = [torch.zeros(self.n_tensors, dtype=i64).pin_memory() for _ in range(3)]
        self._stage_event = [None, None]
        self._stage_next = 0
        self._graph_stage_used = False
        self.exp_avg = torch.zeros(self.total, dtype=f32, device=dev)
        self.exp_avg_sq = torch.zeros(self.total, dtype=f32, device=dev)
        self.partial = torch.empty(self.n_chunks, dtype=f32, device=dev)
        st = torch.zeros(L.OPT_STATE_FLOATS, dtype=f32)
        st[L.OPT_SCALE] = float(init_scale) if use_scaler else 1.0
        st[L.OPT_LR] = float(lr)
        st[L.OPT_BIAS1] = 1.0
        st[L.OPT_BIAS2_SQRT] = 1.0
        self.state = st.to(dev)

    # -- GradScaler surface ----------------------------------------------------------------------------------------------
    def scale(self, loss):
        """`scaler.scale(loss)`: the loss times the current loss scale (read on the device)."""
        if not self.hyper["use_scaler"]:
            return loss
        return loss * self.state[L.OPT_SCALE]

    def backward(self, loss):
        """`scaler.scale(loss).backward()` without the multiply: the loss scale is handed to autograd as the upstream gradient."""
        if not self.hyper["use_scaler"]:
            loss.backward()
        else:
            loss.backward(gradient=self.state[L.OPT_SCALE].to(loss.dtype))

    def get_scale(self):
        return float(self.state[L.OPT_SCALE].item())

    # -- scheduler surface -----------------------------------------------------------------------------------------------
    def set_lr(self, lr):
        """What `scheduler.step()` (CosineAnnealingLR(T_max=epochs, eta_min=1e-6), distributed_trainer_cls.py:52,108-109) does to
        param_groups, once per epoch: one scalar written to the device state; a captured graph picks it up at the next replay.
        `cosine_annealing_lr` below gives the value."""
        self.state[L.OPT_LR:L.OPT_LR + 1].fill_(float(lr))

    def get_lr(self):
        return float(self.state[L.OPT_LR].item())

    # -- results of the last step (device scalars; reading them as floats synchronises) -------------------------------------
    @property
    def grad_norm(self):
        return self.state[L.OPT_GRAD_NORM]

    @property
    def found_inf(self):
        return self.state[L.OPT_FOUND_INF]

    @property
    def step_count(self):
        return self.state[L.OPT_STEP]

    def moments(self, i):
        """(exp_avg, exp_avg_sq) views of parameter i, shaped like it (torch.optim.AdamW's per-parameter state)."""
        a, b = self._off[i], self._off[i + 1]
        return self.exp_avg[a:b].view_as(self.params[i]), self.exp_avg_sq[a:b].view_as(self.params[i])

    # -- the step --------------------------------------------------------------------------------------------------------
    def _stage_grad_ptrs(self):
        """Returns the pinned host table to upload (or None when no .grad moved since the last upload)."""
        ptrs, pptrs = [], []
        for p in self.params:
            pptrs.append(p.data_ptr())
            g = p.grad
            if g is None:
                raise L.CalmError("TrainerStep.step: a parameter has no gradient (the reference runs DDP without "
                                  "find_unused_parameters: every parameter takes part in every step)")
            if g.dtype != f32 or not g.is_contiguous() or g.device != self.device:
                raise L.CalmError("TrainerStep.step: gradients must be contiguous fp32 tensors on the parameters' device")
            ptrs.append(g.data_ptr())
        if pptrs != self._param_ptr_list:
            # a parameter's storage was replaced after construction (model.to(), p.data = ..., load_state_dict(assign=True)):
            # re-upload the table — the AdamW kernel would otherwise update freed memory
            if torch.cuda.is_current_stream_capturing():
                raise L.CalmError("TrainerStep: a parameter was re-allocated since the last step; run one eager step before capturing")
            if any(p.dtype != f32 or not p.is_contiguous() or p.device != self.device for p in self.params):
                raise L.CalmError("TrainerStep: parameters must stay contiguous fp32 tensors on %s" % self.device)
            self.param_ptrs.copy_(torch.tensor(pptrs, dtype=torch.int64))
            self._param_ptr_list = pptrs
        if ptrs == self._grad_ptr_list:
            return None, None
        if torch.cuda.is_current_stream_capturing():
            if self._graph_stage_used:
                raise L.CalmError("TrainerStep: a second CUDA-graph capture needs a second TrainerStep (pointer staging is baked in)")
            self._graph_stage_used = True
            self._grad_ptr_list = None   # addresses inside the capture pool say nothing about the next eager step
            k = 2
        else:
            k = self._stage_next
            self._stage_next ^= 1
            if self._stage_event[k] is not None:
                self._stage_event[k].synchronize()   # the upload that last read this buffer has run (two steps ago)
            self._grad_ptr_list = ptrs
        self._stage[k].copy_(torch.tensor(ptrs, dtype=torch.int64))
        return self._stage[k], k

    def step(self):
        """unscale_ + clip_grad_norm_ + scaler.step(AdamW) + scaler.update(), no host synchronisation."""
        stage, k = self._stage_grad_ptrs()
        a = L.TrainerStepArgs()
        a.params, a.grads = ptr(self.param_ptrs), ptr(self.grad_ptrs)
        a.grads_host = stage.data_ptr() if stage is not None else None
        a.elem_off, a.chunk_tensor, a.chunk_start = ptr(self.elem_off), ptr(self.chunk_tensor), ptr(self.chunk_start)
        a.exp_avg, a.exp_avg_sq, a.partial, a.state = ptr(self.exp_avg), ptr(self.exp_avg_sq), ptr(self.partial), ptr(self.state)
        a.n_tensors, a.n_chunks = self.n_tensors, self.n_chunks
        for key, v in self.hyper.items():
            setattr(a, key, v)
        L.call("calm_trainer_step", C.byref(a), work=32.0 * self.total)
        if stage is not None and k < 2:
            ev = torch.cuda.Event()
            ev.record()
            self._stage_event[k] = ev

    def zero_grad(self, set_to_none=True):
        """`optimizer.zero_grad()` (default set_to_none=True, distributed_trainer_cls.py:96)."""
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    # -- checkpoint surface (torch.optim.AdamW-shaped, so reference-side tooling can read it) --------------------------------
    def state_dict(self):
        st = self.state.cpu()
        per = {}
        for i in range(self.n_tensors):
            m, v = self.moments(i)
            per[i] = {"step": st[L.OPT_STEP].clone(), "exp_avg": m.clone(), "exp_avg_sq": v.clone()}
        group = {"lr": float(st[L.OPT_LR]), "betas": (self.hyper["beta1"], self.hyper["beta2"]), "eps": self.hyper["eps"],
                 "weight_decay": self.hyper["weight_decay"], "params": list(range(self.n_tensors))}
        scaler = {"scale": float(st[L.OPT_SCALE]), "growth_factor": self.hyper["growth_factor"],
                  "backoff_factor": self.hyper["backoff_factor"], "growth_interval": self.hyper["growth_interval"],
                  "_growth_tracker": int(st[L.OPT_GROWTH_TRACKER])}
        return {"state": per, "param_groups": [group], "scaler": scaler}

    def load_state_dict(self, sd):
        st = self.state.cpu()
        for i, s in sd["state"].items():
            m, v = self.moments(int(i))
            m.copy_(s["exp_avg"])
            v.copy_(s["exp_avg_sq"])
            st[L.OPT_STEP] = float(s["step"])
        st[L.OPT_LR] = float(sd["param_groups"][0]["lr"])
        if "scaler" in sd:
            st[L.OPT_SCALE] = float(sd["scaler"]["scale"])
            st[L.OPT_GROWTH_TRACKER] = float(sd["scaler"]["_growth_tracker"])
        step = float(st[L.OPT_STEP])
        st[L.OPT_BIAS1] = 1.0 - math.pow(self.hyper["beta1"], step) if step > 0 else 1.0
        st[L.OPT_BIAS2_SQRT] = math.sqrt(1.0 - math.pow(self.hyper["beta2"], step)) if step > 0 else 1.0
        self.state.copy_(st)
