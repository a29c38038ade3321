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


# ---------------------------------------------------------------------------------------------------------------------
# state_dict layout of the reference, derived from the constructor arguments alone
# ---------------------------------------------------------------------------------------------------------------------
def _sn(shapes, name, out_f, in_f, bias=False, conv_tail=None):
    """Entries an sn(Linear/Conv2d) contributes, in torch's registration order: [bias], weight_orig, weight_u, weight_v."""
    w = (out_f, in_f) if conv_tail is None else (out_f, in_f) + tuple(conv_tail)
    fan = in_f if conv_tail is None else in_f * conv_tail[0] * conv_tail[1]
    if bias:
        shapes[name + ".bias"] = (out_f,)
    shapes[name + ".weight_orig"] = w
    shapes[name + ".weight_u"] = (out_f,)
    shapes[name + ".weight_v"] = (fan,)


def _vmla_shapes(shapes, pre, heads, d1, d2, M, S1, R, S2, mlp_dim, is_cross):
    reduce, t_reduce = d1 != d2, S1 != S2                                    # force_reduce=False (Vi_Tools…:128-129)
    hd_half = d2 // heads // 2
    shapes[pre + "ls_att"] = (d2,)
    shapes[pre + "ls_mlp"] = (d2,)
    shapes[pre + "ln_q.weight"] = (d1,)
    if is_cross:
        shapes[pre + "ln_kv.weight"] = (d1,)
    if t_reduce:
        _sn(shapes, pre + "t_encoder_q", R, S1); _sn(shapes, pre + "t_encoder_kv", R, S1)
    if reduce:
        _sn(shapes, pre + "encoder_q", 2 * M, d1); _sn(shapes, pre + "encoder_kv", 2 * M, d1)
    if t_reduce:
        for n in ("t_qz_upsample", "t_kz_upsample", "t_vz_upsample", "t_qr_proj"):
            _sn(shapes, pre + n, S2, R)
        _sn(shapes, pre + "t_kr_proj", S2, S1)
    d_in = M if reduce else d2
    _sn(shapes, pre + "q_proj", heads * hd_half if reduce else heads * 2 * hd_half, d_in)
    _sn(shapes, pre + "k_proj", heads * hd_half if reduce else heads * 2 * hd_half, d_in)
    _sn(shapes, pre + "v_proj", d2, d_in)
    if reduce:
        _sn(shapes, pre + "qr_proj", heads * hd_half, M); _sn(shapes, pre + "kr_proj", heads * hd_half, d1)
    if t_reduce:
        _sn(shapes, pre + "input_t_proj", S2, S1)
    if reduce:
        _sn(shapes, pre + "input_proj", d2, d1)
    d_rope = hd_half if reduce else 2 * hd_half
    shapes[pre + "rope_q.inv_freq"] = (len(range(0, d_rope, 2)),)
    shapes[pre + "rope_k.inv_freq"] = (len(range(0, d_rope, 2)),)
    _sn(shapes, pre + "linear_mask.0", 2 * S2, S2, bias=True); _sn(shapes, pre + "linear_mask.2", S2, 2 * S2, bias=True)
    _sn(shapes, pre + "out_proj", d2, d2)
    shapes[pre + "ln_2.weight"] = (d2,)
    _sn(shapes, pre + "mlp.0", mlp_dim, d2); _sn(shapes, pre + "mlp.3", d2, mlp_dim)


def _cnn_shapes(shapes, pre, hidden=32):
    _sn(shapes, pre + "0", hidden, 3, bias=True, conv_tail=(1, 1))
    _sn(shapes, pre + "2", hidden, 1, bias=True, conv_tail=(3, 3))
    _sn(shapes, pre + "4", 3, hidden, bias=True, conv_tail=(1, 1))


def state_shapes(heads, seq_length, in_features, dim_step, mean_var_hidden, seq_len_step, seq_len_reduce, out_features,
                 generate, **_):
    """{key: shape} of ViT(type=8).state_dict(), in order (CALM_ViT_V2.py:22-67; Vi_Tools…:98-205,317-385,407-494)."""
    shapes = {}
    S, D, M, R = seq_length, in_features, mean_var_hidden, seq_len_reduce
    plan = [("encoder_blocks.%d." % i, -1) for i in range(3)] + [("block_bottle_neck_1.", 0), ("block_bottle_neck_2.", 0)] + \
           [("decoder_blocks.%d." % i, +1) for i in range(3)]
    for name, sign in plan:
        pre = "autoencoder." + name
        D2, S2 = D + sign * dim_step * 3, S + sign * seq_len_step * 3
        _vmla_shapes(shapes, pre + "encoder.", heads, D, D, M, S, R, S, 2 * D, False)
        _vmla_shapes(shapes, pre + "decoder.", heads, D, D, M, S, R, S, 2 * D, False)
        _vmla_shapes(shapes, pre + "cross.", heads, D, D2, M, S, R, S2, 2 * D2, True)
        _cnn_shapes(shapes, pre + "proj.")
        D, S = D2, S2
    shapes["autoencoder.ln_final.weight"] = (D,)
    if generate:
        _cnn_shapes(shapes, "proj.")
    else:
        _sn(shapes, "head.0", 2 * in_features, in_features); _sn(shapes, "head.2", out_features, 2 * in_features)
    return shapes


def random_state(shapes, seed=0):
    """Random reference-layout state (for timing the CPU baseline): U(-1,1)/sqrt(fan_in) weights, unit u/v, ones elsewhere."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        if k.endswith("weight_orig"):
            fan = 1
            for s in shp[1:]:
                fan *= s
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan)
        elif k.endswith("weight_u") or k.endswith("weight_v"):
            sd[k] = F.normalize(torch.randn(shp, generator=g), dim=0)
        elif k.endswith("inv_freq"):
            half = shp[0]
            sd[k] = 1.0 / (10000.0 ** (torch.arange(0, 2 * half, 2).float() / (2 * half)))
        elif k.endswith(".bias"):
            sd[k] = torch.zeros(shp)
        else:
            sd[k] = torch.ones(shp)
    return sd
