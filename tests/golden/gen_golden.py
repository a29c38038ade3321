"""Generates tests/golden/<config>.npz by running the UNMODIFIED reference (imported from /root/reference/CALM-ViT, which
only exists in the build container) on the synthetic states / inputs / noise of synth.py.

    python tests/golden/gen_golden.py            # writes tiny_cls, tiny_gen, small_cls, small_gen

Per config the fixture holds: the state_dict key -> shape table, the training-forward output, kl, loss, summary
statistics (sum, L2 norm, 16 sampled entries) of every parameter gradient and of every post-forward power-iteration
vector, full gradients of the small parameters, and the eval-mode output computed afterwards with the warmed u/v.
The latent noise the reference draws with torch.randn_like (Vi_Tools_CNN_less_V2.py:238-239) is replaced, for this run
only, by synth.NoiseStream so that the fixture does not depend on torch's CPU RNG implementation.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import synth  # noqa: E402

REF = "/root/reference/CALM-ViT"
FULL_GRAD_MAX = 512


def import_reference():
    sys.path.insert(0, REF)
    for n in ("matplotlib", "matplotlib.pyplot"):                 # CALM_ViT_V2.py:7 — only save_samples uses it
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import CALM_ViT_V2 as rvh
    return rvh


def stats(t, key):
    t = t.detach().double().flatten()
    idx = synth.sample_indices(t.numel(), key)
    return np.concatenate([[t.sum().item(), t.norm().item()], t[idx].numpy()])


def run(rvh, name):
    cfg = synth.CONFIGS[name]
    kw = {k: v for k, v in cfg.items() if k != "batch"}
    torch.manual_seed(0)
    model = rvh.ViT(torch.device("cpu"), type=8, force_reduce=False, **kw)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synth.synth_state(shapes, seed=0))
    x, y = synth.synth_input(cfg)
    noise = synth.NoiseStream(cfg)
    real_randn_like = torch.randn_like
    torch.randn_like = lambda t, **kws: next(noise).to(t.dtype)
    try:
        model.train()
        out, kl = model(x)
        if cfg["generate"]:                                       # distributed_trainer_reg.py:76-88
            S = cfg["seq_length"]
            loss = torch.nn.HuberLoss(delta=1.0)(out.reshape(-1, S, S, 3).permute(0, 3, 1, 2), x) + kl * 0.1
        else:                                                     # distributed_trainer_cls.py:84-86
            loss = torch.nn.CrossEntropyLoss()(out.squeeze(), y)
        loss.backward()
    finally:
        torch.randn_like = real_randn_like
    assert noise.k == 12, noise.k
    fx = {"out_train": out.detach().numpy(), "kl": np.float64(float(kl.detach())), "loss": np.float64(float(loss.detach()))}
    gkeys, gstats, bkeys, bstats = [], [], [], []
    for k, p in model.named_parameters():
        gkeys.append(k); gstats.append(stats(p.grad, k))
        if p.numel() <= FULL_GRAD_MAX:
            fx["grad/" + k] = p.grad.numpy()
    for k, b in model.state_dict().items():
        if k.endswith("weight_u") or k.endswith("weight_v"):
            bkeys.append(k); bstats.append(stats(b, k))
    fx["grad_stats"] = np.stack(gstats); fx["buf_stats"] = np.stack(bstats)
    model.eval()
    with torch.no_grad():
        out_eval, kl_eval = model(x)
    fx["out_eval"] = out_eval.numpy(); fx["kl_eval"] = np.float64(float(kl_eval))
    meta = {"config": cfg, "shapes": {k: list(v) for k, v in shapes.items()}, "grad_keys": gkeys, "buf_keys": bkeys,
            "torch": torch.__version__, "n_params": sum(p.numel() for p in model.parameters())}
    fx["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **fx)
    print(name, "loss %.6f kl %.6f params %d -> %s (%.1f KB)" % (float(loss.detach()), float(kl.detach()), meta["n_params"], path,
                                                                 os.path.getsize(path) / 1024))


if __name__ == "__main__":
    rvh = import_reference()
    for name in (sys.argv[1:] or list(synth.CONFIGS)):
        run(rvh, name)
