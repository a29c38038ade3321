"""Pins the oracle (oracle/calm_oracle.py) on the golden vectors produced by the unmodified reference
(tests/golden/gen_golden.py) and — when /root/reference is present, i.e. in the build container — on the live reference.
fp32 on CPU; tolerance 1e-4 relative (north_star's fp32 bar), measured ~1e-6.
"""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import synth  # noqa: E402
from oracle import calm_oracle as O  # noqa: E402

TOL = 1e-4


def load_fixture(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def relerr(a, b):
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def stats(t, key):
    t = t.detach().double().flatten()
    idx = synth.sample_indices(t.numel(), key)
    return np.concatenate([[t.sum().item(), t.norm().item()], t[idx].numpy()])


def run_oracle(meta, training=True):
    cfg = meta["config"]
    P = O.params_from_state_dict(synth.synth_state(meta["shapes"], seed=0))
    x, y = synth.synth_input(cfg)
    noise = synth.NoiseStream(cfg)
    if cfg["generate"]:
        loss, out = O.train_step_reg(P, cfg["heads"], x, training, noise)
    else:
        loss, out = O.train_step_cls(P, cfg["heads"], x, y, training, noise)
    return P, x, loss, out, noise


@pytest.mark.parametrize("name", ["tiny_cls", "tiny_gen", "small_cls", "small_gen"])
def test_oracle_matches_reference_golden(name):
    z, meta = load_fixture(name)
    cfg = meta["config"]
    P, x, loss, out, noise = run_oracle(meta)
    assert noise.k == 12                                         # 6 reduce blocks x (zq, zkv), reference draw order
    assert relerr(out, z["out_train"]) < TOL
    assert abs(float(loss) - float(z["loss"])) < TOL * abs(float(z["loss"]))
    # every parameter gradient: sum, norm and 16 sampled entries; norm-relative tolerance
    for k, ref in zip(meta["grad_keys"], z["grad_stats"]):
        got = stats(P[k].grad, k)
        scale = max(ref[1], 1e-12)
        assert abs(got[1] - ref[1]) < TOL * scale, k
        assert np.abs(got[2:] - ref[2:]).max() < 10 * TOL * scale / np.sqrt(max(P[k].numel(), 1)) + TOL * np.abs(ref[2:]).max(), k
    for k in z.files:
        if k.startswith("grad/"):
            assert relerr(P[k[5:]].grad, z[k]) < TOL, k
    # the power iteration mutated u/v exactly like the reference's forward pre-hooks
    for k, ref in zip(meta["buf_keys"], z["buf_stats"]):
        got = stats(P[k], k)
        assert np.abs(got - ref).max() < TOL, k
    # eval-mode forward with the warmed buffers: no noise, no buffer update
    before = {k: P[k].clone() for k in meta["buf_keys"]}
    with torch.no_grad():
        out_eval, kl_eval = O.vit(P, cfg["heads"], x, training=False)
    assert relerr(out_eval, z["out_eval"]) < TOL
    assert abs(float(kl_eval) - float(z["kl_eval"])) < TOL * abs(float(z["kl_eval"]))
    assert all(torch.equal(P[k], before[k]) for k in meta["buf_keys"])


@pytest.mark.parametrize("name", ["tiny_cls", "tiny_gen", "small_cls", "small_gen"])
def test_oracle_state_layout_matches_reference(name):
    """oracle.state_shapes re-derives the reference's state_dict keys, order and shapes from the ctor arguments alone."""
    _, meta = load_fixture(name)
    cfg = meta["config"]
    shapes = O.state_shapes(**cfg)
    assert list(shapes.keys()) == list(meta["shapes"].keys())
    assert all(list(shapes[k]) == meta["shapes"][k] for k in shapes)


def test_oracle_kl_and_shapes():
    z, meta = load_fixture("tiny_gen")
    cfg = meta["config"]
    P = O.params_from_state_dict(synth.synth_state(meta["shapes"]), requires_grad=False)
    x, _ = synth.synth_input(cfg)
    out, kl = O.vit(P, cfg["heads"], x, training=True, noise=synth.NoiseStream(cfg))
    assert out.shape == (cfg["batch"], cfg["seq_length"], cfg["in_features"])
    assert abs(float(kl) - float(z["kl"])) < TOL * abs(float(z["kl"]))


@pytest.mark.skipif(not os.path.exists("/root/reference/CALM-ViT/CALM_ViT_V2.py"), reason="reference only exists in the build container")
def test_oracle_matches_live_reference_rng_order():
    """Same torch seed right before each forward -> the oracle consumes torch.randn_like in the reference's order."""
    # the product modules carry the same names by design: import the reference in isolation, then restore sys.modules
    names = ("CALM_ViT_V2", "Vi_Tools_CNN_less_V2")
    saved = {n: sys.modules.pop(n) for n in names if n in sys.modules}
    sys.path.insert(0, "/root/reference/CALM-ViT")
    for n in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    try:
        import CALM_ViT_V2 as rvh
        assert rvh.__file__.startswith("/root/reference/")
    finally:
        sys.path.remove("/root/reference/CALM-ViT")
        for n in names:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
    cfg = synth.CONFIGS["tiny_cls"]
    kw = {k: v for k, v in cfg.items() if k != "batch"}
    torch.manual_seed(3)
    model = rvh.ViT(torch.device("cpu"), type=8, force_reduce=False, **kw)
    model.train()
    x, y = synth.synth_input(cfg)
    model(x)                                                      # warm u/v with the reference's own init
    P = O.params_from_state_dict(model.state_dict())
    torch.manual_seed(11)
    ref_out, ref_kl = model(x)
    torch.nn.functional.cross_entropy(ref_out.squeeze(), y).backward()
    torch.manual_seed(11)
    O.train_step_cls(P, cfg["heads"], x, y, training=True, noise=None)
    with torch.no_grad():
        pass
    worst = max(relerr(P[k].grad, p.grad) for k, p in model.named_parameters())
    assert worst < TOL, worst
    for k, b in model.state_dict().items():
        if k.endswith("weight_u") or k.endswith("weight_v"):
            assert relerr(P[k], b) < TOL, k
