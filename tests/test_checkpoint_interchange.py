"""CPU tests of the checkpoint tooling: reference-format round trips in both directions and the
remove_spectral_norm-style export. The live-reference cases only run where /root/reference exists (the build container)."""
import io
import os
import sys
import types

import pytest
import torch
from torch.nn.utils import remove_spectral_norm, spectral_norm

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import synth  # noqa: E402

import calm_checkpoint as ck  # noqa: E402

REF_DIR = "/root/reference/CALM-ViT"


def _kw(name):
    return {k: v for k, v in synth.CONFIGS[name].items() if k != "batch"}


def _reference_module():
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
    import importlib.util
    saved = {n: sys.modules.pop(n) for n in ("CALM_ViT_V2", "Vi_Tools_CNN_less_V2") if n in sys.modules}
    sys.path.insert(0, REF_DIR)
    try:
        spec = importlib.util.spec_from_file_location("ref_CALM_ViT_V2", os.path.join(REF_DIR, "CALM_ViT_V2.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(REF_DIR)
        for n in ("CALM_ViT_V2", "Vi_Tools_CNN_less_V2"):
            sys.modules.pop(n, None)
        sys.modules.update(saved)
    return mod


def test_fold_spectral_norm_matches_torch_remove_spectral_norm():
    torch.manual_seed(3)
    net = torch.nn.Sequential(spectral_norm(torch.nn.Linear(24, 40, bias=False)), torch.nn.GELU(),
                              spectral_norm(torch.nn.Conv2d(8, 8, 3, padding=1, groups=8)), spectral_norm(torch.nn.Linear(40, 6)))
    net.train()
    net[0](torch.randn(5, 24)); net[2](torch.randn(2, 8, 6, 6)); net[3](torch.randn(5, 40))   # move u / v off their init
    net.eval()
    sd = net.state_dict()
    folded = ck.fold_spectral_norm(sd)
    net[0](torch.randn(1, 24)); net[2](torch.randn(1, 8, 6, 6)); net[3](torch.randn(1, 40))    # eval forward: weight = W / sigma(u, v)
    for i in (0, 2, 3):
        remove_spectral_norm(net[i])
    want = net.state_dict()
    assert list(folded.keys()) == [k.replace("_orig", "") for k in sd if not k.endswith(("_u", "_v"))]
    assert set(folded) == set(want)
    for k in want:
        assert folded[k].shape == want[k].shape
        assert torch.allclose(folded[k], want[k], rtol=1e-6, atol=1e-7), k


def test_product_checkpoint_round_trip(tmp_path):
    import CALM_ViT_V2 as rvh
    torch.manual_seed(0)
    a = rvh.ViT(torch.device("cpu"), type=8, force_reduce=False, **_kw("tiny_cls"))
    path = ck.save_reference_checkpoint(a, str(tmp_path / "model_cls.pth"))
    torch.manual_seed(1)
    b = rvh.ViT(torch.device("cpu"), type=8, force_reduce=False, **_kw("tiny_cls"))
    res = ck.load_reference_checkpoint(b, path, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for (k1, v1), (k2, v2) in zip(a.state_dict().items(), b.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1
    folded = ck.fold_spectral_norm(a.state_dict())
    assert not any(k.endswith(("weight_orig", "weight_u", "weight_v")) for k in folded)
    assert sum(k.endswith(".weight") for k in folded) >= sum(k.endswith("weight_orig") for k in a.state_dict())


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "CALM_ViT_V2.py")), reason="reference only exists in the build container")
@pytest.mark.parametrize("name", ["tiny_cls", "tiny_gen"])
def test_checkpoints_interchange_with_live_reference(name, tmp_path):
    ref = _reference_module()
    import CALM_ViT_V2 as rvh
    torch.manual_seed(0)
    r = ref.ViT(torch.device("cpu"), type=8, force_reduce=False, **_kw(name))
    torch.manual_seed(5)
    p = rvh.ViT(torch.device("cpu"), type=8, force_reduce=False, **_kw(name))
    # reference -> product (what distributed_trainer_cls.py:106 writes)
    buf = io.BytesIO()
    torch.save(r.state_dict(), buf)
    buf.seek(0)
    res = ck.load_reference_checkpoint(p, buf, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for (k1, v1), (k2, v2) in zip(r.state_dict().items(), p.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1
    # product -> reference, through a file, strict
    with torch.no_grad():
        for t in p.parameters():
            t.mul_(1.25)
    path = ck.save_reference_checkpoint(p, str(tmp_path / "model.pth"))
    r.load_state_dict(torch.load(path), strict=True)
    for (k1, v1), (k2, v2) in zip(p.state_dict().items(), r.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1
