"""ORACLE — test infrastructure only. Never imported by the product path (calm-vit-dte_b200/); only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker or
as the timed CPU baseline.

A functional, stateless restatement of the reference's hot path (CALM-ViT-DTE, CALM-ViT/Vi_Tools_CNN_less_V2.py and
CALM-ViT/CALM_ViT_V2.py) in plain PyTorch ops over a flat {state_dict key: tensor} mapping. It is written from the
reference's mathematics, not from its module code: there are no nn.Modules, no hooks and no hidden state here, which is
what makes it usable as an independent check of the CUDA path and of the drop-in modules.

Parity status: PINNED. tests/golden/*.npz were produced by importing the unmodified reference in the build container
(tests/golden/gen_golden.py); tests/test_oracle_golden.py checks this file against them (fp32, 1e-4 relative — measured
~1e-6), and, when /root/reference is present, against the live reference modules as well.

Numerics follow the arithmetic substrate the reference dispatches to (PyTorch, un-vendored and un-pinned in
requirements.txt:1 — oracle pinned on torch 2.11.0): because every contraction here is the same torch op the reference
calls (F.linear / matmul / F.scaled_dot_product_attention / F.layer_norm / F.softplus / torch.mv / torch.dot), running this
file under torch.autocast(bfloat16) on a CUDA device reproduces the reference's mixed-precision policy too.
"""
import math

import torch
import torch.nn.functional as F

SN_EPS = 1e-12   # torch/nn/utils/spectral_norm.py:30 (default eps)
LN_EPS = 1e-6    # Vi_Tools_CNN_less_V2.py:115 (norm_layer = LayerNorm(eps=1e-6))


# ---------------------------------------------------------------------------------------------------------------------
# Spectral normalisation — torch/nn/utils/spectral_norm.py:92-114 (one power iteration per training forward)
# ---------------------------------------------------------------------------------------------------------------------
def sn_weight(P, name, training):
    """Effective weight W_orig / sigma of the sn(...) layer `name`; updates P[name.weight_u/_v] in place when training."""
    w = P[name + ".weight_orig"]
    u = P[name + ".weight_u"]
    v = P[name + ".weight_v"]
    wm = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=SN_EPS))   # :103
            u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=SN_EPS))       # :106
        u, v = u.clone(), v.clone()                                        # :110-111
    sigma = torch.dot(u, torch.mv(wm, v))                                  # :113
    return w / sigma                                                       # :114


def sn_linear(P, name, x, training):
    """sn(Linear) applied on the last axis (bias only where the reference has one: linear_mask.{0,2})."""
    return F.linear(x, sn_weight(P, name, training), P.get(name + ".bias"))


def sn_seq_linear(P, name, x, training):
    """sn(Linear) applied along the SEQUENCE axis: permute(0,2,1) -> Linear -> permute back (Vi_Tools…:224-229,250-264)."""
    return sn_linear(P, name, x.transpose(1, 2), training).transpose(1, 2)


def layer_norm(P, name, x):
    """Weight-only LayerNorm, eps 1e-6 (Vi_Tools…:131-132,197,494)."""
    w = P[name + ".weight"]
    return F.layer_norm(x, (x.shape[-1],), w, None, LN_EPS)


# ---------------------------------------------------------------------------------------------------------------------
# RoPE with learned inverse frequencies (Vi_Tools…:80-95)
# ---------------------------------------------------------------------------------------------------------------------
def rope(x, inv_freq):
    """x (B, h, S, d): x*cos(emb) + rotate_half(x)*sin(emb), emb = [t (x) inv_freq | t (x) inv_freq]."""
    S, d = x.shape[2], x.shape[3]
    t = torch.arange(S, dtype=torch.float32, device=x.device)
    ang = torch.outer(t, inv_freq)
    emb = torch.cat((ang, ang), dim=-1)
    rot = torch.cat((-x[..., d // 2:], x[..., : d // 2]), dim=-1)
    return x * emb.cos() + rot * emb.sin()


# ---------------------------------------------------------------------------------------------------------------------
# Latent running state — ResidualStateManager(mode="sum") (Vi_Tools…:7-50)
# ---------------------------------------------------------------------------------------------------------------------
class LatentState:
    def __init__(self):
        self.zq = None
        self.zkv = None
        self.kl = 0.0
        self.count = 0

    def push(self, zq, zkv, mu_q, sd_q, mu_kv, sd_kv):
        kl = lambda mu, sd: -0.5 * torch.mean(1 + 2 * torch.log(sd) - mu.pow(2) - sd.pow(2))   # :24-25
        self.kl = kl(mu_q, sd_q) + kl(mu_kv, sd_kv) + self.kl                                    # :26
        if self.zq is None:
            self.zq, self.zkv = zq, zkv                                                          # :27-30
        else:
            self.zq, self.zkv = self.zq + zq, self.zkv + zkv                                     # :42-44
        self.count += 1
        return self.zq, self.zkv

    def kl_loss(self):
        return self.kl / self.count if self.count > 0 else 0.0                                   # :49-50


# ---------------------------------------------------------------------------------------------------------------------
# VMLA block (Vi_Tools…:207-315)
# ---------------------------------------------------------------------------------------------------------------------
def vmla(P, pre, heads, x_q, x_kv, state, training, noise=None, record=None):
    """One attention block. `pre` = state_dict prefix ending in '.', x_kv=None for self attention.
    noise: optional iterator yielding the eps tensors (zq first, then zkv) instead of torch.randn_like."""
    has = lambda n: (pre + n + ".weight_orig") in P
    reduce, t_reduce = has("encoder_q"), has("t_encoder_q")
    residual = x_q
    xq = layer_norm(P, pre + "ln_q", x_q)
    xkv = xq if x_kv is None else layer_norm(P, pre + "ln_kv", x_kv)
    qz = qr = xq
    kz = vz = kr = xkv
    if reduce:
        if t_reduce:                                                       # squeeze the sequence axis S1 -> R
            xq = sn_seq_linear(P, pre + "t_encoder_q", xq, training)
            xkv = sn_seq_linear(P, pre + "t_encoder_kv", xkv, training)
        mu_q, rho_q = sn_linear(P, pre + "encoder_q", xq, training).chunk(2, dim=-1)
        mu_kv, rho_kv = sn_linear(P, pre + "encoder_kv", xkv, training).chunk(2, dim=-1)
        sd_q = F.softplus(rho_q) + 1e-6                                    # :234-235
        sd_kv = F.softplus(rho_kv) + 1e-6
        if training:                                                       # :237-239 (zq drawn first)
            e_q = next(noise) if noise is not None else torch.randn_like(sd_q)
            e_kv = next(noise) if noise is not None else torch.randn_like(sd_kv)
            zq, zkv = mu_q + e_q * sd_q, mu_kv + e_kv * sd_kv
        else:
            zq, zkv = mu_q, mu_kv
        if state is not None:
            zq, zkv = state.push(zq, zkv, mu_q, sd_q, mu_kv, sd_kv)
        qz = qr = zq
        kz = vz = zkv
        if t_reduce:                                                       # expand R -> S2; kr comes from the UN-reduced xkv
            qz = sn_seq_linear(P, pre + "t_qz_upsample", qz, training)
            kz = sn_seq_linear(P, pre + "t_kz_upsample", kz, training)
            vz = sn_seq_linear(P, pre + "t_vz_upsample", vz, training)
            qr = sn_seq_linear(P, pre + "t_qr_proj", qr, training)
            kr = sn_seq_linear(P, pre + "t_kr_proj", kr, training)
    q = sn_linear(P, pre + "q_proj", qz, training)
    k = sn_linear(P, pre + "k_proj", kz, training)
    v = sn_linear(P, pre + "v_proj", vz, training)
    B, Sq, Skv = q.shape[0], q.shape[1], k.shape[1]
    split = lambda t, S: t.reshape(B, S, heads, t.shape[-1] // heads).transpose(1, 2)
    q, k, v = split(q, Sq), split(k, Skv), split(v, Skv)
    if reduce:                                                             # decoupled RoPE: [content | rope(qr/kr proj)]
        qr = split(sn_linear(P, pre + "qr_proj", qr, training), Sq)
        kr = split(sn_linear(P, pre + "kr_proj", kr, training), Skv)
        q = torch.cat((q, rope(qr, P[pre + "rope_q.inv_freq"])), dim=-1)
        k = torch.cat((k, rope(kr, P[pre + "rope_k.inv_freq"])), dim=-1)
    else:
        q = rope(q, P[pre + "rope_q.inv_freq"])
        k = rope(k, P[pre + "rope_k.inv_freq"])
    merge = lambda t, S: t.transpose(1, 2).reshape(B, S, -1)
    # learned additive mask: MLP over the key axis of the all-head, unscaled logits (:288-291)
    logits = merge(q, Sq) @ merge(k, Skv).transpose(1, 2)
    hid = F.gelu(sn_linear(P, pre + "linear_mask.0", logits, training))
    bias = sn_linear(P, pre + "linear_mask.2", hid, training).unsqueeze(1)
    a = F.scaled_dot_product_attention(q, k, v, attn_mask=bias, dropout_p=0.0, is_causal=False)   # :293-298
    x = sn_linear(P, pre + "out_proj", merge(a, Sq), training) * P[pre + "ls_att"]
    if residual.shape != x.shape:                                          # :302-308
        if has("input_t_proj"):
            residual = sn_seq_linear(P, pre + "input_t_proj", residual, training)
        if has("input_proj"):
            residual = sn_linear(P, pre + "input_proj", residual, training)
    x = x + residual
    y = layer_norm(P, pre + "ln_2", x)
    y = sn_linear(P, pre + "mlp.3", F.gelu(sn_linear(P, pre + "mlp.0", y, training)), training) * P[pre + "ls_mlp"]
    if record is not None:
        record[pre] = dict(bias=bias.detach(), attn=a.detach())
    return x + y


def cnn_residual(P, pre, x, training):
    """x + conv1x1(gelu(dwconv3x3(gelu(conv1x1(img))))) on the (B,S,S,3) pixel view of the tokens (Vi_Tools…:378-385,400-403)."""
    B, S = x.shape[0], x.shape[1]
    img = x.reshape(B, S, S, 3).permute(0, 3, 1, 2)
    h = F.gelu(F.conv2d(img, sn_weight(P, pre + "0", training), P[pre + "0.bias"]))
    h = F.gelu(F.conv2d(h, sn_weight(P, pre + "2", training), P[pre + "2.bias"], padding=1, groups=h.shape[1]))
    h = F.conv2d(h, sn_weight(P, pre + "4", training), P[pre + "4.bias"])
    return x + h.permute(0, 2, 3, 1).reshape(B, S, 3 * S)


def swap_axes_tokens(x):
    """Row tokens <-> column tokens: (B,S,S,3) with the two S axes exchanged (Vi_Tools…:394-395,397-398)."""
    B, S = x.shape[0], x.shape[1]
    return x.reshape(B, S, S, 3).permute(0, 2, 1, 3).reshape(B, S, 3 * S)


def block(P, pre, heads, x, state, training, noise=None, record=None):
    """Block.forward (Vi_Tools…:387-403): row attention, column attention, cross attention (stage change), CNN residual."""
    xq = vmla(P, pre + "encoder.", heads, x, None, None, training, noise, record)
    xkv = swap_axes_tokens(vmla(P, pre + "decoder.", heads, swap_axes_tokens(xq), None, None, training, noise, record))
    y = vmla(P, pre + "cross.", heads, xq, xkv, state, training, noise, record)
    return cnn_residual(P, pre + "proj.", y, training)


def encoder_decoder(P, pre, heads, img, training, noise=None, record=None):
    """EncoderDecoder_8.forward (Vi_Tools…:496-533). img (B,3,S,S) -> (tokens (B,S,3S), kl)."""
    state = LatentState()
    B, _, S, _ = img.shape
    x = img.permute(0, 2, 3, 1).reshape(B, S, 3 * S)                        # :389-391 (first block only)
    skips = []
    for i in range(3):
        x = block(P, pre + "encoder_blocks.%d." % i, heads, x, state, training, noise, record)
        skips.append(x)
    skip_1, skip_2, skip_b1 = skips
    x = block(P, pre + "block_bottle_neck_1.", heads, x, state, training, noise, record) + skip_b1       # :512-513
    skip_b2 = x
    x = block(P, pre + "block_bottle_neck_2.", heads, x, state, training, noise, record) + (skip_b2 + skip_b1)  # :515-516
    x = block(P, pre + "decoder_blocks.0.", heads, x, state, training, noise, record) + skip_2           # :519-520
    x = block(P, pre + "decoder_blocks.1.", heads, x, state, training, noise, record) + skip_1           # :521-522
    x = block(P, pre + "decoder_blocks.2.", heads, x, state, training, noise, record)
    return layer_norm(P, pre + "ln_final", x), state.kl_loss()


def vit(P, heads, img, training=True, noise=None, record=None):
    """ViT.forward (CALM_ViT_V2.py:70-84). Classification if the head exists in P, else the generate-mode CNN residual."""
    x, kl = encoder_decoder(P, "autoencoder.", heads, img, training, noise, record)
    if "head.0.weight_orig" in P:
        x = x.mean(dim=1)                                                   # permute + AdaptiveAvgPool1d(1) + squeeze (:73-75)
        x = sn_linear(P, "head.2", F.gelu(sn_linear(P, "head.0", x, training)), training)
    else:
        x = cnn_residual(P, "proj.", x, training)
    return x, kl


# ---------------------------------------------------------------------------------------------------------------------
# Helpers for tests / bench
# ---------------------------------------------------------------------------------------------------------------------
def is_param(key):
    """state_dict entries that are nn.Parameters in the reference (everything but the power-iteration vectors)."""
    return not (key.endswith(".weight_u") or key.endswith(".weight_v"))


def params_from_state_dict(sd, device="cpu", dtype=torch.float32, requires_grad=True):
    """Deep-copies a reference-layout state_dict into the flat mapping the functions above consume."""
    P = {}
    for k, t in sd.items():
        t = t.detach().to(device=device, dtype=dtype).clone()
        if requires_grad and is_param(k):
            t.requires_grad_(True)
        P[k] = t
    return P


def train_step_cls(P, heads, img, target, training=True, noise=None):
    """Forward + soft-target cross-entropy + backward (distributed_trainer_cls.py:84-87, without the GradScaler).
    Returns (loss, logits); gradients land in P[k].grad."""
    logits, _ = vit(P, heads, img, training, noise)
    loss = F.cross_entropy(logits.squeeze(), target)
    loss.backward()
    return loss.detach(), logits.detach()


def train_step_reg(P, heads, img, training=True, noise=None):
    """Forward + Huber(delta=1) reconstruction + 0.1*KL + backward (distributed_trainer_reg.py:76-88)."""
    y, kl = vit(P, heads, img, training, noise)
    S = img.shape[-1]
    rec = y.reshape(-1, S, S, 3).permute(0, 3, 1, 2)
    loss = F.huber_loss(rec, img, delta=1.0) + 0.1 * kl
    loss.backward()
    return loss.detach(), y.detach()
