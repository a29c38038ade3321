"""Checkpoint interchange with the reference.

The drop-in modules keep the reference's state_dict layout (`weight_orig / weight_u / weight_v` per sn(...) layer plus
torch's spectral_norm version metadata), so `.pth` files written by `distributed_trainer_cls.py:106` load here and files
written here load there — `save_reference_checkpoint` / `load_reference_checkpoint` only wrap that with the checks a
maintainer wants (missing / unexpected keys, shapes).

`fold_spectral_norm` is the export `torch.nn.utils.remove_spectral_norm` would give (CALM_ViT_V2.py:9 imports it as `rsn`
and never uses it): every `<lin>.weight_orig/_u/_v` triple becomes one plain `<lin>.weight = weight_orig / sigma` with
`sigma = u^T W v` from the stored vectors — the weight an eval-mode forward uses (torch/nn/utils/spectral_norm.py:92-114
with `do_power_iteration=False`). Host-side tensor bookkeeping only; nothing here touches the device kernels.
"""
from collections import OrderedDict

import torch


def save_reference_checkpoint(model, path):
    """`torch.save(model.module.state_dict(), path)` of the reference loop (distributed_trainer_cls.py:106); DataParallel /
    DDP wrappers are unwrapped, tensors go to the CPU."""
    inner = getattr(model, "module", model)
    sd = inner.state_dict()
    out = OrderedDict((k, v.detach().cpu()) for k, v in sd.items())
    if hasattr(sd, "_metadata"):
        out._metadata = sd._metadata
    torch.save(out, path)
    return path


def load_reference_checkpoint(model, path_or_state, strict=True):
    """Loads a reference-format state_dict (a path or the dict). Returns torch's (missing, unexpected) result; with
    strict=True a shape or key mismatch raises exactly like the reference's `load_state_dict`."""
    sd = torch.load(path_or_state, map_location="cpu") if isinstance(path_or_state, (str, bytes)) or hasattr(path_or_state, "read") else path_or_state
    inner = getattr(model, "module", model)
    return inner.load_state_dict(sd, strict=strict)


def fold_spectral_norm(state_dict):
    """state_dict with every spectral-norm triple folded into a plain `weight` (see the module docstring). Conv weights keep
    their (out, in, kh, kw) shape; sigma is computed on the (out, -1) matrix like torch does (dim=0)."""
    out = OrderedDict()
    for k, v in state_dict.items():
        if k.endswith(".weight_u") or k.endswith(".weight_v"):
            continue
        if k.endswith(".weight_orig"):
            base = k[: -len("_orig")]
            w = v.detach().float()
            u = state_dict[base + "_u"].detach().float()
            vv = state_dict[base + "_v"].detach().float()
            sigma = torch.dot(u, torch.mv(w.reshape(w.shape[0], -1), vv))
            out[base] = (w / sigma).to(v.dtype)
        else:
            out[k] = v.detach().clone()
    return out
