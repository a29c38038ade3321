"""GPU parity of the whole hot path: the drop-in modules (C-ABI kernels) against the oracle on identical weights, inputs
and latent noise. (The same comparison against the UNMODIFIED reference modules at the full trainer configs is
tests/test_reference_parity_gpu.py.)

Protocol: seed -> construct with the reference's default initialisation (the drop-in constructors draw the same
init as the reference's for the same seed) -> one warm-up training forward (moves u/v off their random start) -> state_dict ->
(a) fp32 oracle = the truth, (b) the oracle under torch.autocast(bfloat16) = what the reference computes in the trainers,
(c) the product; same input, same torch.manual_seed before each forward.

Bars, all norm-relative (||a-b|| / ||b||), north_star's bf16 tolerance: output / loss / kl flat 2e-2 against the fp32 truth and
against the bf16 policy; gradients per tensor as stated in tests/parity_util.py; u/v 1e-4; latent noise bit-exact.
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, HERE)
import synth  # noqa: E402
import parity_util as pu  # noqa: E402
from parity_util import rel  # noqa: E402
from oracle import calm_oracle as O  # noqa: E402


def load_fixture(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return z, json.loads(bytes(z["meta"]).decode())


def loss_fn(cfg, out, kl, x, y):
    if cfg["generate"]:
        S = cfg["seq_length"]
        return torch.nn.functional.huber_loss(out.reshape(-1, S, S, 3).permute(0, 3, 1, 2), x, delta=1.0) + kl * 0.1
    return torch.nn.functional.cross_entropy(out.squeeze(), y)


def run_product(name, monkeypatch, training=True):
    import CALM_ViT_V2 as rvh
    dev = torch.device("cuda:0")
    z, meta = load_fixture(name)
    cfg = meta["config"]
    kw = {k: v for k, v in cfg.items() if k != "batch"}
    model = rvh.ViT(dev, type=8, force_reduce=False, **kw).to(dev)
    sd = model.state_dict()
    assert list(sd.keys()) == list(meta["shapes"].keys())
    state = synth.synth_state(meta["shapes"])
    model.load_state_dict(state)
    x, y = synth.synth_input(cfg)
    x = x.to(dev)
    y = y.to(dev) if y is not None else None
    noise = synth.NoiseStream(cfg)
    monkeypatch.setattr(torch, "randn", lambda *a, **k: next(noise).to(dev))
    model.train(training)
    out, kl = model(x)
    loss = loss_fn(cfg, out, kl, x, y)
    loss.backward()
    torch.cuda.synchronize()
    monkeypatch.undo()
    return z, meta, cfg, model, state, x, y, out.detach(), kl.detach(), loss.detach()


def run_oracle(cfg, state, x, y, autocast):
    dev = x.device
    P = O.params_from_state_dict(state, device=dev)
    noise = (t.to(dev) for t in synth.NoiseStream(cfg))
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        out, kl = O.vit(P, cfg["heads"], x, True, noise)
        loss = loss_fn(cfg, out, kl, x, y)
    loss.backward()
    return P, out.detach(), kl.detach(), loss.detach()


def default_init_state(kw, x, seed=0):
    """The reference's default initialisation (same constructor calls, same seed) after one warm-up training forward."""
    import CALM_ViT_V2 as rvh
    dev = x.device
    torch.manual_seed(seed)
    model = rvh.ViT(dev, type=8, force_reduce=False, **kw).to(dev)
    model.train()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        model(x)
    return model, {k: v.detach().clone() for k, v in model.state_dict().items()}


def oracle_step(cfg, state, x, y, seed, autocast):
    from torch.nn.attention import SDPBackend, sdpa_kernel
    P = O.params_from_state_dict(state, device=x.device)
    torch.manual_seed(seed)
    with sdpa_kernel([SDPBackend.MATH] if not autocast else [SDPBackend.MATH, SDPBackend.EFFICIENT_ATTENTION, SDPBackend.FLASH_ATTENTION]):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out, kl = O.vit(P, cfg["heads"], x, True, None)
            loss = loss_fn(cfg, out, kl, x, y)
        loss.backward()
    grads = {k: P[k].grad for k in P if O.is_param(k)}
    return P, out.detach().float(), kl.detach().float(), loss.detach().float(), grads


@pytest.mark.parametrize("name", ["small_cls", "small_gen"])
def test_training_step_matches_oracle(name):
    """One training step at the small configs (S=160, D=480, head dims 40/28/16/4) with the reference's default initialisation:
    flat 2e-2 on output / loss / kl, the per-tensor gradient bar of tests/parity_util.py, u/v to 1e-4."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    cfg = dict(synth.CONFIGS[name], batch=4)
    kw = {k: v for k, v in cfg.items() if k != "batch"}
    x, y = synth.synth_input(cfg)
    x = x.to(dev)
    y = y.to(dev) if y is not None else None
    model, state = default_init_state(kw, x)
    seed = 21
    for p in model.parameters():
        p.grad = None
    torch.manual_seed(seed)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out, kl = model(x)
        loss = loss_fn(cfg, out, kl, x, y)
    loss.backward()
    torch.cuda.synchronize()
    P32, f_out, f_kl, f_loss, g32 = oracle_step(cfg, state, x, y, seed, autocast=False)
    Pbf, o_out, o_kl, o_loss, gbf = oracle_step(cfg, state, x, y, seed, autocast=True)
    d = {"out_vs_fp32": rel(out, f_out), "out_vs_bf16_policy": rel(out, o_out), "bf16_policy_vs_fp32": rel(o_out, f_out),
         "loss_vs_fp32": rel(loss, f_loss), "loss_vs_bf16_policy": rel(loss, o_loss), "kl_vs_fp32": rel(kl, f_kl), "kl_vs_bf16_policy": rel(kl, o_kl)}
    print("\n[%s] %s" % (name, {k: "%.3e" % v for k, v in d.items()}))
    grads = {k: p.grad for k, p in model.named_parameters()}
    assert all(torch.isfinite(g).all() for g in grads.values())
    rows, offenders, summary = pu.gradient_report(g32, gbf, grads)
    pu.print_report(name, rows, offenders, summary)
    for k in ("out_vs_fp32", "out_vs_bf16_policy", "loss_vs_fp32", "loss_vs_bf16_policy", "kl_vs_fp32", "kl_vs_bf16_policy"):
        assert d[k] < 2e-2, (k, d[k])
    assert not offenders, offenders[:8]
    sd = model.state_dict()
    for k in sd:                                           # power-iteration buffers advanced exactly once, in fp32
        if k.endswith(("weight_u", "weight_v")):
            assert rel(sd[k], P32[k]) < 1e-4, k


def test_golden_vector_anchor(monkeypatch):
    """Absolute anchor on the committed fp32 golden vectors of the UNMODIFIED reference (tests/golden/small_cls.npz, RNG-free
    synthetic state and injected noise): output / loss / kl flat 2e-2, late-layer gradients 2e-2."""
    z, meta, cfg, model, state, x, y, out, kl, loss = run_product("small_cls", monkeypatch)
    assert rel(out, torch.as_tensor(z["out_train"]).to(out.device)) < 2e-2
    assert abs(loss.item() - float(z["loss"])) < 2e-2 * abs(float(z["loss"]))
    assert abs(kl.item() - float(z["kl"])) < 2e-2 * abs(float(z["kl"]))
    P32, _, _, _ = run_oracle(cfg, state, x, y, autocast=False)
    params = dict(model.named_parameters())
    late = [k for k in params if k.startswith(("head.", "autoencoder.ln_final"))]
    assert late
    for k in late:
        assert rel(params[k].grad, P32[k].grad) < 2e-2, k
    sd = model.state_dict()
    for k in meta["buf_keys"]:
        assert rel(sd[k], P32[k]) < 1e-4, k


def test_eval_forward_and_mutation_contract(monkeypatch):
    z, meta, cfg, model, state, x, y, out, kl, loss = run_product("small_gen", monkeypatch)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    model.eval()
    with torch.no_grad():
        out_eval, kl_eval = model(x)
    assert rel(out_eval, torch.as_tensor(z["out_eval"]).to(out.device)) < 3e-2
    assert abs(kl_eval.item() - float(z["kl_eval"])) < 2e-2 * abs(float(z["kl_eval"]))
    assert all(torch.equal(before[k], v) for k, v in model.state_dict().items())   # eval mutates nothing
    # a second training forward advances u/v again (one power iteration per training forward)
    model.train()
    model(x)
    sd = model.state_dict()
    assert any(not torch.equal(before[k], sd[k]) for k in meta["buf_keys"])


def test_seeded_noise_matches_oracle_rng_order(monkeypatch):
    """Without injected noise: same torch CUDA seed before each forward -> the product's 12 torch.randn draws are BIT-IDENTICAL
    to the torch.randn_like draws of the reference algorithm (zq then zkv per reduce block, Vi_Tools_CNN_less_V2.py:238-239)."""
    dev = torch.device("cuda:0")
    cfg = synth.CONFIGS["small_cls"]
    kw = {k: v for k, v in cfg.items() if k != "batch"}
    x, y = synth.synth_input(cfg)
    x, y = x.to(dev), y.to(dev)
    model, state = default_init_state(kw, x)
    mine, theirs = [], []
    real_randn, real_randn_like = torch.randn, torch.randn_like
    monkeypatch.setattr(torch, "randn", lambda *a, **k: (mine.append(real_randn(*a, **k)), mine[-1])[1])
    torch.manual_seed(5)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out, kl = model(x)
    monkeypatch.undo()
    P = O.params_from_state_dict(state, device=dev, requires_grad=False)
    monkeypatch.setattr(torch, "randn_like", lambda *a, **k: (theirs.append(real_randn_like(*a, **k)), theirs[-1])[1])
    torch.manual_seed(5)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o_out, o_kl = O.vit(P, cfg["heads"], x, True, None)
    monkeypatch.undo()
    assert len(mine) == len(theirs) == 12
    for i, (a, b) in enumerate(zip(mine, theirs)):
        assert a.dtype == b.dtype == torch.float32 and torch.equal(a, b), "latent noise draw %d" % i
    assert rel(out, o_out) < 2e-2
    assert abs(kl.item() - o_kl.item()) < 2e-2 * abs(o_kl.item())
    torch.manual_seed(6)
    out2, _ = model(x)
    assert rel(out2, out) > 1e-4                       # different seed -> different latent noise


def test_second_forward_before_backward():
    """Two training forwards (and an eval pass in between) before any backward — gradient accumulation with a deferred backward,
    legal in the reference: the accumulated gradients equal those of the same two steps run one after the other."""
    import CALM_ViT_V2 as rvh
    dev = torch.device("cuda:0")
    _, meta = load_fixture("small_cls")
    cfg = meta["config"]
    kw = {k: v for k, v in cfg.items() if k != "batch"}
    state = synth.synth_state(meta["shapes"])
    x1, _ = synth.synth_input(cfg)
    x2, _ = synth.synth_input(cfg, seed=3)
    x1, x2 = x1.to(dev), x2.to(dev)

    def fresh():
        m = rvh.ViT(dev, type=8, force_reduce=False, **kw).to(dev)
        m.load_state_dict(state)
        m.train()
        return m

    def fwd(m, x, seed):
        torch.manual_seed(seed)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out, kl = m(x)
        return out.float().pow(2).mean() + 0.1 * kl
    seq = fresh()
    fwd(seq, x1, 0).backward()
    fwd(seq, x2, 1).backward()
    want = {k: p.grad.clone() for k, p in seq.named_parameters()}
    model = fresh()
    l1 = fwd(model, x1, 0)
    model.eval()
    with torch.no_grad():
        model(x1)
    model.train()
    l2 = fwd(model, x2, 1)
    l1.backward()
    l2.backward()
    torch.cuda.synchronize()
    for k, p in model.named_parameters():
        assert rel(p.grad, want[k]) < 1e-5, k
    for (k, a), b in zip(model.state_dict().items(), seq.state_dict().values()):
        assert torch.equal(a, b), k


@pytest.mark.parametrize("B", [1, 3])
def test_full_size_shapes_properties(B):
    """BASELINE.json's trainer config (224^2, heads 12, latent (80,240)): size-independent properties at full width —
    output shape, finite loss/gradients for every parameter, batch-independence of per-image results (row b of a batch
    equals the same image run alone, since nothing in the path mixes images except the shared weights)."""
    import CALM_ViT_V2 as rvh
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = rvh.ViT(dev, type=8, heads=12, seq_length=224, in_features=672, dim_step=48, mean_var_hidden=240,
                    seq_len_step=16, seq_len_reduce=80, out_features=1000, force_reduce=False, generate=False).to(dev)
    model.train()
    g = torch.Generator().manual_seed(2006)
    x = torch.randn(B, 3, 224, 224, generator=g).to(dev)
    with torch.no_grad():
        model(x)                                       # warm u/v (a fresh model's first sigma estimate is poor)
    model.eval()
    with torch.no_grad():
        out, kl = model(x)
        assert out.shape == (B, 1000) and torch.isfinite(out).all() and torch.isfinite(kl)
        single, _ = model(x[:1])
        assert rel(out[:1], single) < 1e-3
    model.train()
    out, kl = model(x)
    y = torch.softmax(torch.randn(B, 1000, generator=g) * 4, -1).to(dev)
    torch.nn.functional.cross_entropy(out, y).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    assert sum(p.numel() for p in model.parameters()) == 42577130      # SURVEY §6 parameter count, cls config


@pytest.mark.parametrize("S,B,R,M", [(384, 2, 80, 240), (512, 2, 192, 544)])
def test_highres_configs_match_oracle(S, B, R, M):
    """BASELINE configs[3] (384^2 / 512^2, original and scaled latent bank; head dims 60-128): one training step with the
    reference's default initialisation against the fp32 oracle and against the oracle under the trainers' autocast(bf16) policy
    (identical weights, input, latent noise): flat 2e-2 on output / loss / kl, the per-tensor gradient bar of parity_util.py."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    kw = dict(heads=12, seq_length=S, in_features=3 * S, dim_step=48, mean_var_hidden=M, seq_len_step=16, seq_len_reduce=R,
              out_features=1000, generate=False)
    cfg = dict(kw, batch=B)
    x, y = synth.synth_input(cfg)
    x, y = x.to(dev), y.to(dev)
    model, state = default_init_state(kw, x)
    seed = 11
    torch.manual_seed(seed)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out, kl = model(x)
        loss = torch.nn.functional.cross_entropy(out.squeeze(), y)
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    out, kl, loss = out.detach().float(), kl.detach().float(), loss.detach().float()
    del model
    torch.cuda.empty_cache()
    P32, f_out, f_kl, f_loss, g32 = oracle_step(cfg, state, x, y, seed, autocast=False)
    g32 = {k: v.clone() for k, v in g32.items()}
    del P32
    torch.cuda.empty_cache()
    Pbf, o_out, o_kl, o_loss, gbf = oracle_step(cfg, state, x, y, seed, autocast=True)
    d = {"out_vs_fp32": rel(out, f_out), "out_vs_bf16_policy": rel(out, o_out), "bf16_policy_vs_fp32": rel(o_out, f_out),
         "loss_vs_fp32": rel(loss, f_loss), "loss_vs_bf16_policy": rel(loss, o_loss), "kl_vs_fp32": rel(kl, f_kl), "kl_vs_bf16_policy": rel(kl, o_kl)}
    print("\n[S=%d latent (%d,%d)] %s" % (S, R, M, {k: "%.3e" % v for k, v in d.items()}))
    assert all(torch.isfinite(g).all() for g in grads.values())
    rows, offenders, summary = pu.gradient_report(g32, gbf, grads)
    pu.print_report("S=%d" % S, rows, offenders, summary)
    for k in ("out_vs_fp32", "out_vs_bf16_policy", "loss_vs_fp32", "loss_vs_bf16_policy", "kl_vs_fp32", "kl_vs_bf16_policy"):
        assert d[k] < 2e-2, (k, d[k])
    assert not offenders, offenders[:8]
    del Pbf
    torch.cuda.empty_cache()
