"""Deterministic, RNG-free synthesis of model states, inputs and latent noise for the golden vectors.

Everything is a pure integer hash (numpy uint64 arithmetic) mapped to floats, so the generator script (which runs the
unmodified reference) and the tests (which run the oracle / the CUDA path) rebuild bit-identical tensors on any machine
without storing megabytes of weights. Test infrastructure only.
"""
import zlib

import numpy as np
import torch

CONFIGS = {
    # name: ViT kwargs (CALM_ViT_V2.py:22-25) + batch. "tiny" is CPU-only (its widths are not multiples of 8);
    # "small" satisfies every alignment rule of the CUDA path (D % 16 == 0 at every stage, S % 8 == 0).
    "tiny_cls": dict(heads=3, seq_length=48, in_features=144, dim_step=12, mean_var_hidden=12, seq_len_step=4,
                     seq_len_reduce=8, out_features=10, generate=False, batch=2),
    "tiny_gen": dict(heads=3, seq_length=48, in_features=144, dim_step=12, mean_var_hidden=12, seq_len_step=4,
                     seq_len_reduce=8, out_features=144, generate=True, batch=2),
    "small_cls": dict(heads=12, seq_length=160, in_features=480, dim_step=48, mean_var_hidden=16, seq_len_step=16,
                      seq_len_reduce=16, out_features=16, generate=False, batch=2),
    "small_gen": dict(heads=12, seq_length=160, in_features=480, dim_step=48, mean_var_hidden=16, seq_len_step=16,
                      seq_len_reduce=16, out_features=480, generate=True, batch=2),
}


def _hash_u01(n, salt):
    """n uniform floats in [0,1) from a splitmix64-style integer hash of (index, salt)."""
    with np.errstate(over="ignore"):
        x = np.arange(n, dtype=np.uint64) + np.uint64(salt) * np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return (x >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def uniform(shape, salt, lo=-1.0, hi=1.0):
    n = int(np.prod(shape)) if len(shape) else 1
    return torch.from_numpy((lo + (hi - lo) * _hash_u01(n, salt)).astype(np.float32)).reshape(shape)


def normalish(shape, salt):
    """Zero-mean, unit-variance noise (sum of 4 uniforms — the distribution is irrelevant for parity)."""
    n = int(np.prod(shape))
    s = sum(_hash_u01(n, salt * 7 + i) for i in range(4)) - 2.0
    return torch.from_numpy((s * np.sqrt(3.0)).astype(np.float32)).reshape(shape)


def key_salt(key, seed):
    return (zlib.crc32(key.encode()) ^ (seed * 0x5bd1e995)) & 0x7FFFFFFF


def synth_state(shapes, seed=0):
    """shapes: {state_dict key: shape}. Returns {key: fp32 tensor} with sensible magnitudes for every kind of entry."""
    out = {}
    for k, shp in shapes.items():
        shp = tuple(shp)
        s = key_salt(k, seed)
        if k.endswith("weight_orig"):
            fan_in = int(np.prod(shp[1:]))
            t = uniform(shp, s) * (1.5 / np.sqrt(fan_in))
        elif k.endswith("weight_u") or k.endswith("weight_v"):
            t = uniform(shp, s)
            t = t / t.norm()
        elif k.endswith("inv_freq"):
            half = shp[0]
            base = 1.0 / (10000.0 ** (torch.arange(0, 2 * half, 2).float() / (2 * half)))   # Vi_Tools…:62
            t = base * (1.0 + 0.05 * uniform(shp, s))
        elif k.endswith(".bias"):
            t = 0.1 * uniform(shp, s)
        else:  # ls_att / ls_mlp / LayerNorm weights
            t = 1.0 + 0.1 * uniform(shp, s)
        out[k] = t.contiguous()
    return out


def synth_input(cfg, seed=0):
    B, S = cfg["batch"], cfg["seq_length"]
    x = normalish((B, 3, S, S), 1000 + seed)
    if cfg["generate"]:
        return x, None
    y = torch.softmax(4.0 * normalish((B, cfg["out_features"]), 2000 + seed), dim=-1)  # dense soft labels (CutMix/MixUp)
    return x, y


class NoiseStream:
    """Iterator of latent noise tensors eps_k of shape (B, R, M): the 12 randn_like draws of one forward, in order."""

    def __init__(self, cfg, seed=0):
        self.shape = (cfg["batch"], cfg["seq_len_reduce"], cfg["mean_var_hidden"])
        self.seed = seed
        self.k = 0

    def __iter__(self):
        return self

    def __next__(self):
        t = normalish(self.shape, 3000 + 100 * self.seed + self.k)
        self.k += 1
        return t


def sample_indices(numel, key, n=16):
    """n reproducible flat indices into a tensor of `numel` elements."""
    u = _hash_u01(n, key_salt(key, 99))
    return torch.from_numpy(np.minimum((u * numel).astype(np.int64), numel - 1))
