"""GPU parity at BASELINE.json's trainer configs against the UNMODIFIED reference modules (SURVEY §8c protocol).

For the classification (distributed_trainer_cls.py:148-151) and regression (distributed_trainer_reg.py:140-143) configs at
224^2, heads 12, latent (80,240):

    seed -> construct the reference's ViT with its default initialisation -> one warm-up training forward under CUDA autocast
    (moves u/v off their random init) -> state_dict -> load into (a) a second reference instance run in fp32 (TF32 off, math
    SDPA: the truth), (b) a third reference instance run under torch.autocast(bfloat16) exactly as the trainers do, (c) the
    drop-in product. Same input, same torch.manual_seed right before each forward.

Bars (north_star: 2e-2 in bf16 for activations, loss and gradients; bit-exact latent noise):
  * output, loss, kl: flat 2e-2 against the fp32 reference AND against the live bf16 reference;
  * the 12 latent-noise tensors of the forward: torch.equal to the reference's torch.randn_like draws;
  * every parameter gradient, per tensor: distance to the fp32 truth <= max(2e-2, 1.5 x the distance the reference's own
    bf16 run shows FOR THAT TENSOR) for every tensor of >= 256 elements; the tiny cancellation-dominated reductions (inv_freq,
    conv biases, ...) are held to pooled per-class statistics plus a per-tensor cap — tests/parity_util.py states the bar and
    why; the offenders are printed;
  * u/v after the step: 1e-4 against the fp32 reference (the product iterates in fp32; the reference's autocast run does its
    mat-vecs in bf16, so it is itself ~3e-3 from that — printed);
  * the fp32 oracle (oracle/calm_oracle.py) against the fp32 reference at this size: 1e-4 (pins the oracle at full width).
"""
import json
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import refload  # noqa: E402
import parity_util as pu  # noqa: E402
from parity_util import rel  # noqa: E402

TRAINER = {
    "cls": dict(heads=12, seq_length=224, in_features=672, dim_step=48, mean_var_hidden=240, seq_len_step=16, seq_len_reduce=80,
                out_features=1000, generate=False),
    "reg": dict(heads=12, seq_length=224, in_features=672, dim_step=48, mean_var_hidden=240, seq_len_step=16, seq_len_reduce=80,
                out_features=672, generate=True),
}
NPARAM = {"cls": (521, 42577130), "reg": (525, 40330509)}


def loss_of(task, out, kl, x, y):
    if task == "reg":                                          # distributed_trainer_reg.py:76-88
        img = out.reshape(-1, 224, 224, 3).permute(0, 3, 1, 2)
        return torch.nn.HuberLoss(delta=1.0)(img, x) + kl * 0.1
    return torch.nn.CrossEntropyLoss()(out.squeeze(), y)       # distributed_trainer_cls.py:84-86


class Recorder:
    """Wraps torch.randn / torch.randn_like for the duration of one forward and keeps what they returned."""

    def __init__(self, monkeypatch, name):
        self.draws = []
        real = getattr(torch, name)

        def wrapped(*a, **k):
            t = real(*a, **k)
            self.draws.append(t.detach().clone())
            return t
        monkeypatch.setattr(torch, name, wrapped)
        self.monkeypatch = monkeypatch

    def done(self):
        self.monkeypatch.undo()
        return self.draws


def run_model(model, task, x, y, seed, autocast, recorder=None):
    model.train()
    for p in model.parameters():
        p.grad = None
    torch.manual_seed(seed)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        out, kl = model(x)
        loss = loss_of(task, out, kl, x, y)
    loss.backward()
    torch.cuda.synchronize()
    draws = recorder.done() if recorder is not None else None
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    bufs = {k: v.detach().clone() for k, v in model.state_dict().items() if k.endswith(("weight_u", "weight_v"))}
    return dict(out=out.detach().float(), kl=kl.detach().float(), loss=loss.detach().float(), grads=grads, bufs=bufs, draws=draws)


@pytest.mark.parametrize("task,B", [("cls", 4), ("reg", 3)])
def test_trainer_config_matches_unmodified_reference(task, B, monkeypatch):
    ref_mod = refload.load_reference()
    if ref_mod is None:
        pytest.skip("unmodified reference not available (baseline/_ref or /root/reference)")
    import CALM_ViT_V2 as rvh
    from torch.nn.attention import SDPBackend, sdpa_kernel
    dev = torch.device("cuda:0")
    kw = TRAINER[task]
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    make_ref = lambda: ref_mod.ViT(dev, type=8, force_reduce=False, **kw).to(dev)
    torch.manual_seed(2006)
    ref = make_ref()                                            # the reference's own default initialisation
    g = torch.Generator().manual_seed(2006)
    x = torch.randn(B, 3, 224, 224, generator=g).to(dev)
    y = torch.softmax(torch.randn(B, 1000, generator=g) * 4, -1).to(dev) if task == "cls" else None
    ref.train()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ref(x)                                                  # warm-up training forward: one power iteration on every u/v
    sd0 = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    assert (len(list(ref.parameters())), sum(p.numel() for p in ref.parameters())) == NPARAM[task]
    seed = 77
    # (a) fp32 truth: the unmodified reference, no autocast, math SDPA, TF32 off
    ref32 = make_ref(); ref32.load_state_dict(sd0)
    with sdpa_kernel(SDPBackend.MATH):
        T = run_model(ref32, task, x, y, seed, autocast=False, recorder=Recorder(monkeypatch, "randn_like"))
    del ref32
    # (b) the reference exactly as the trainers run it
    refbf = ref; refbf.load_state_dict(sd0)
    R = run_model(refbf, task, x, y, seed, autocast=True, recorder=Recorder(monkeypatch, "randn_like"))
    del refbf, ref
    # (c) the product
    model = rvh.ViT(dev, type=8, force_reduce=False, **kw).to(dev)
    assert list(model.state_dict().keys()) == list(sd0.keys())
    model.load_state_dict(sd0)
    Pm = run_model(model, task, x, y, seed, autocast=True, recorder=Recorder(monkeypatch, "randn"))
    torch.cuda.empty_cache()

    report = {"task": task, "batch": B}
    # ---- latent noise: bit-exact, same order (zq then zkv per reduce block, Vi_Tools_CNN_less_V2.py:238-239)
    assert len(T["draws"]) == len(R["draws"]) == len(Pm["draws"]) == 12
    for i, (a, b, c) in enumerate(zip(T["draws"], R["draws"], Pm["draws"])):
        assert a.shape == c.shape == (B, 80, 240) and c.dtype == torch.float32
        assert torch.equal(a, b) and torch.equal(b, c), "latent noise draw %d differs from the reference's randn_like" % i
    # ---- activations / loss / kl: flat 2e-2
    d = {"out_vs_fp32": rel(Pm["out"], T["out"]), "out_vs_refbf16": rel(Pm["out"], R["out"]), "refbf16_out_vs_fp32": rel(R["out"], T["out"]),
         "loss_vs_fp32": rel(Pm["loss"], T["loss"]), "loss_vs_refbf16": rel(Pm["loss"], R["loss"]),
         "kl_vs_fp32": rel(Pm["kl"], T["kl"]), "kl_vs_refbf16": rel(Pm["kl"], R["kl"])}
    report["activations"] = d
    print("\n[%s B=%d] %s" % (task, B, {k: "%.3e" % v for k, v in d.items()}))
    # ---- gradients: per tensor for everything with >= 256 elements, pooled + capped for the tiny reductions (parity_util.py)
    for k, gp in Pm["grads"].items():
        assert torch.isfinite(gp).all(), k
    rows, offenders, summary = pu.gradient_report(T["grads"], R["grads"], Pm["grads"])
    report["gradients"] = summary
    report["per_tensor"] = [dict(key=k, numel=n, ours=a, ref=b, ours_vs_ref=c) for k, n, a, b, c in rows]
    pu.print_report(task, rows, offenders, summary)
    # ---- u/v after the step
    uv = max(rel(Pm["bufs"][k], T["bufs"][k]) for k in T["bufs"])
    uv_ref = max(rel(R["bufs"][k], T["bufs"][k]) for k in T["bufs"])
    report["uv"] = {"ours_vs_fp32_max": uv, "refbf16_vs_fp32_max": uv_ref, "n": len(T["bufs"])}
    print("[%s] u/v after the step vs fp32 reference: ours max %.3e, reference under autocast max %.3e (%d buffers)" % (task, uv, uv_ref, len(T["bufs"])))
    # ---- the oracle at full width against the fp32 reference (pins the checker used by the other tests at this size)
    from oracle import calm_oracle as O
    P32 = O.params_from_state_dict(sd0, device=dev)
    torch.manual_seed(seed)
    with sdpa_kernel(SDPBackend.MATH):
        o_out, o_kl = O.vit(P32, kw["heads"], x, True, None)
        o_loss = loss_of(task, o_out, o_kl, x, y)
        o_loss.backward()
    og = max(rel(P32[k].grad, T["grads"][k]) for k in T["grads"])
    report["oracle_fp32_vs_reference_fp32"] = {"out": rel(o_out, T["out"]), "loss": rel(o_loss, T["loss"]), "kl": rel(o_kl, T["kl"]), "grad_max": og}
    print("[%s] fp32 oracle vs fp32 reference: out %.2e loss %.2e kl %.2e worst gradient %.2e" % (
        task, rel(o_out, T["out"]), rel(o_loss, T["loss"]), rel(o_kl, T["kl"]), og))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump(report, open(os.path.join(out_dir, "parity_reference_%s.json" % task), "w"), indent=1)
    for k in ("out_vs_fp32", "out_vs_refbf16", "loss_vs_fp32", "loss_vs_refbf16", "kl_vs_fp32", "kl_vs_refbf16"):
        assert d[k] < 2e-2, (k, d[k])
    assert not offenders, "%d gradient tensors outside the bar of tests/parity_util.py: %s" % (len(offenders), offenders[:8])
    assert uv < 1e-4, uv
    assert rel(o_out, T["out"]) < 1e-4 and rel(o_loss, T["loss"]) < 1e-4 and og < 1e-3
