"""B200-native drop-in for the reference's architecture tools (CALM-ViT/Vi_Tools_CNN_less_V2.py).

Same module name, class names, constructor arguments, forward signatures, parameter/buffer registration order and
state_dict keys as the reference, so `import Vi_Tools_CNN_less_V2 as vt` (CALM_ViT_V2.py:3) and the trainers
(distributed_trainer_cls.py / distributed_trainer_reg.py) drive this path unchanged. The arithmetic does not run in
PyTorch: every forward/backward op is a hand-written sm_100a kernel of libcalm_b200.so reached through the C ABI
(include/calm_b200.h) — see calm_ops.py. The sn(...) wrappers are kept only as parameter containers (weight_orig /
weight_u / weight_v, checkpoint metadata); their Python power-iteration hooks never run because the wrapped layers are
never called — one batched kernel per Block does all power iterations (calm_sn_forward).

Numerics = the trainers' autocast(bfloat16) policy; the modules require CUDA tensors and raise otherwise.
"""
from functools import partial
from typing import Callable

import torch
from torch.nn.utils import spectral_norm as sn

import calm_ops as ops
from calm_ops import GroupSpec


class ResidualStateManager():
    """Latent running state shared by the cross blocks (reference :7-50). Only mode "sum" is exercised by the model
    (:499) and only that mode is executed on the device path; the other modes keep the reference's host-side formulas."""

    def __init__(self, smooth_factor: float = 2.0, momentum: float = 0.9, mode: str = "ema"):
        super().__init__()
        self.zq_sum = None
        self.zkv_sum = None
        self.tot_kl_loss = 0.0
        self.count = 0
        self.smooth_factor = smooth_factor
        self.mode = mode
        self.momentum = momentum

    def get_sums(self, zq, zkv, mean_q, var_q, mean_kv, var_kv):
        kl = lambda m, s: -0.5 * torch.mean(1 + 2 * torch.log(s) - m.pow(2) - s.pow(2))
        self.tot_kl_loss = kl(mean_q, var_q) + kl(mean_kv, var_kv) + self.tot_kl_loss
        if self.zq_sum is None:
            self.zq_sum, self.zkv_sum, self.count = zq, zkv, 1
            return self.zq_sum, self.zkv_sum
        self.count += 1
        if self.mode in ("sum", "sma"):
            self.zq_sum, self.zkv_sum = self.zq_sum + zq, self.zkv_sum + zkv
            if self.mode == "sma":
                return self.zq_sum / self.count, self.zkv_sum / self.count
        else:
            if self.mode == "ema":
                self.momentum = self.smooth_factor / (self.count + 1)
            elif self.mode == "lp":
                self.momentum = self.count / (self.count + 1)
            m = self.momentum
            self.zq_sum, self.zkv_sum = m * zq + (1 - m) * self.zq_sum, m * zkv + (1 - m) * self.zkv_sum
        return self.zq_sum, self.zkv_sum

    def get_kl_loss(self):
        return self.tot_kl_loss / self.count if self.count > 0 else 0.0


class RoPE(torch.nn.Module):
    """Rotary embedding container (reference :55-95). Inside VMLA_Block the rotation itself is fused into the attention
    core kernels (calm_rope_fwd/bwd read `inv_freq` directly); forward() is kept for stand-alone use of the class."""

    def __init__(self, seq: int, dim: int, theta: float = 10000.0, learned: bool = False, training: bool = True):
        super().__init__()
        self.seq, self.dim, self.theta, self.learned = seq, dim, theta, learned
        inv_freq = 1.0 / (self.theta ** (torch.arange(0, dim, 2).float() / self.dim))
        t = torch.arange(self.seq, dtype=torch.float)
        if learned:
            self.inv_freq = torch.nn.Parameter(inv_freq, requires_grad=True)
            self.register_buffer("t", t, persistent=False)
        else:
            self.register_buffer("inv_freq", inv_freq)
            self.register_buffer("t", t, persistent=False)
            emb = torch.outer(t, inv_freq).repeat(1, 2)
            self.register_buffer("cos_emb", emb.cos(), persistent=False)
            self.register_buffer("sin_emb", emb.sin(), persistent=False)

    def forward(self, x):
        emb = torch.outer(self.t[: x.shape[2]].to(x.device), self.inv_freq).repeat(1, 2)
        half = x.shape[-1] // 2
        return x * emb.cos() + torch.cat((-x[..., half:], x[..., :half]), dim=-1) * emb.sin()


def _sn_linear(i, o, bias=False):
    return sn(torch.nn.Linear(i, o, bias=bias))


class VMLA_Block(torch.nn.Module):
    """Multi-head latent distribution attention block (reference :98-315)."""

    def __init__(self, heads: int, dim1: int, dim2: int, mean_var_hidden: int, seq_length: int, seq_len_reduce: int,
                 seq_len_new: int, mlp_dim: int, force_reduce: bool = True, t_force_reduce: bool = False,
                 dropout: float = 0.0, use_mlp: bool = True, is_cross: bool = False, training: bool = True,
                 norm_layer: Callable[..., torch.nn.Module] = partial(torch.nn.LayerNorm, eps=1e-6)):
        super().__init__()
        self.ls_att = torch.nn.Parameter(torch.ones(dim2), requires_grad=True)
        self.ls_mlp = torch.nn.Parameter(torch.ones(dim2), requires_grad=True) if use_mlp else None
        self.training = training
        self.heads = heads
        self.head_dim_content = self.head_dim_rope = dim2 // heads // 2
        self.head_dim = 2 * self.head_dim_rope
        self.t_reduce = seq_len_new != seq_length or t_force_reduce
        self.reduce = dim1 != dim2 or force_reduce
        self.seq_length, self.seq_len_new, self.dim1, self.dim2 = seq_length, seq_len_new, dim1, dim2
        if dropout != 0.0:
            raise NotImplementedError("the B200 path implements the trainers' dropout=0.0 configuration")
        hc, hr, M = heads * self.head_dim_content, heads * self.head_dim_rope, mean_var_hidden
        # --- registration order below mirrors the reference so that state_dict() enumerates identically ---
        self.ln_q = ops.check_norm(norm_layer(dim1, bias=False))
        self.ln_kv = ops.check_norm(norm_layer(dim1, bias=False)) if is_cross else None
        self.t_encoder_q = self.t_encoder_kv = None
        if self.t_reduce:
            self.t_encoder_q = _sn_linear(seq_length, seq_len_reduce)
            self.t_encoder_kv = _sn_linear(seq_length, seq_len_reduce)
        self.encoder_q = self.encoder_kv = None
        if self.reduce:
            self.encoder_q = _sn_linear(dim1, 2 * M)
            self.encoder_kv = _sn_linear(dim1, 2 * M)
        self.t_qz_upsample = self.t_kz_upsample = self.t_vz_upsample = self.t_qr_proj = self.t_kr_proj = None
        if self.t_reduce:
            self.t_qz_upsample = _sn_linear(seq_len_reduce, seq_len_new)
            self.t_kz_upsample = _sn_linear(seq_len_reduce, seq_len_new)
            self.t_vz_upsample = _sn_linear(seq_len_reduce, seq_len_new)
            self.t_qr_proj = _sn_linear(seq_len_reduce, seq_len_new)
            self.t_kr_proj = _sn_linear(seq_length, seq_len_new)
        self.qz_upsample = self.kz_upsample = self.vz_upsample = None
        d_in = dim2 if (dim1 == dim2 and not force_reduce) else M
        self.q_proj = _sn_linear(d_in, hc if self.reduce else heads * self.head_dim)
        self.k_proj = _sn_linear(d_in, hc if self.reduce else heads * self.head_dim)
        self.v_proj = _sn_linear(d_in, dim2)
        self.qr_proj = self.kr_proj = None
        if self.reduce:
            self.qr_proj = _sn_linear(M, hr)
            self.kr_proj = _sn_linear(dim1, hr)
        self.input_t_proj = _sn_linear(seq_length, seq_len_new) if seq_len_new != seq_length else None
        self.input_proj = _sn_linear(dim1, dim2) if dim1 != dim2 else None
        d_rope = self.head_dim_rope if self.reduce else self.head_dim
        self.rope_q = RoPE(seq_len_new, d_rope, learned=True)
        self.rope_k = RoPE(seq_len_new, d_rope, learned=True)
        self.linear_mask = torch.nn.Sequential(
            _sn_linear(seq_len_new, seq_len_new * 2, bias=True),
            torch.nn.GELU(approximate='none'),
            _sn_linear(seq_len_new * 2, seq_len_new, bias=True),
        )
        self.out_proj = _sn_linear(dim2, dim2)
        self.dropout = torch.nn.Dropout(dropout)
        self.ln_2 = ops.check_norm(norm_layer(dim2, bias=False))
        self.mlp = None
        if use_mlp:
            self.mlp = torch.nn.Sequential(
                _sn_linear(dim2, mlp_dim),
                torch.nn.GELU(approximate='none'),
                torch.nn.Dropout(dropout, inplace=False),
                _sn_linear(mlp_dim, dim2),
            )

    # ------------------------------------------------------------------------------------------------------------
    def _sn_groups(self):
        """Bank groups of this block: layers sharing an input are stacked into one GEMM operand."""
        gs = []
        if self.reduce:
            singles = [self.t_encoder_q, self.t_encoder_kv, self.encoder_q, self.encoder_kv, self.t_qz_upsample,
                       self.t_kz_upsample, self.t_vz_upsample, self.t_qr_proj, self.t_kr_proj, self.q_proj, self.k_proj,
                       self.v_proj, self.qr_proj, self.kr_proj, self.input_t_proj, self.input_proj]
            gs += [GroupSpec([m]) for m in singles if m is not None]
        elif self.ln_kv is None:
            gs.append(GroupSpec([self.q_proj, self.k_proj, self.v_proj]))        # self attention: one (3D, D) GEMM
        else:
            gs.append(GroupSpec([self.q_proj]))
            gs.append(GroupSpec([self.k_proj, self.v_proj]))
        gs.append(GroupSpec([self.linear_mask[0]]))
        gs.append(GroupSpec([self.linear_mask[2]]))
        gs.append(GroupSpec([self.out_proj], rowscale=self.ls_att))              # LayerScale folded into W_eff rows
        if self.mlp is not None:
            gs.append(GroupSpec([self.mlp[0]]))
            gs.append(GroupSpec([self.mlp[3]], rowscale=self.ls_mlp))
        return gs

    def forward(self, input_q, input_kv=None, state_manager=None, mask=False):
        if not mask:
            # the reference dereferences mask_mat unconditionally (:290-291): mask=False raises there too
            raise AttributeError("'NoneType' object has no attribute 'unsqueeze' (VMLA_Block requires mask=True)")
        sc = ops.current_scope()
        if sc is not None and id(self.out_proj) in sc.bank.gid_of:
            return self._forward(sc, input_q, input_kv, state_manager)
        with ops.Scope(self, self._sn_groups, self.training) as sc:
            return self._forward(sc, input_q, input_kv, state_manager)

    def _forward(self, sc, input_q, input_kv, csm):
        bank, tok = sc.bank, sc.token
        gid = bank.gid
        B = input_q.shape[0]
        h, S2, D2 = self.heads, self.seq_len_new, self.dim2
        lin = lambda x, m, addend=None, out_f32=False: ops.LinearFn.apply(x, tok, addend, bank, gid(m), out_f32)
        seq = lambda x, *ms: ops.SeqLinearFn.apply(x, tok, bank, *[gid(m) for m in ms])
        if input_q.dim() == 4:      # the first Block hands the NCHW image over: tokenisation is fused into this LayerNorm (:389-391)
            xq, res = ops.ImageLayerNormFn.apply(input_q, self.ln_q.weight, self.ln_q.eps)
        else:
            input_q = input_q.float()
            xq, res = ops.LayerNormFn.apply(input_q, self.ln_q.weight, False, False, self.ln_q.eps)
        xkv = xq if input_kv is None else ops.LayerNormFn.apply(input_kv.float(), self.ln_kv.weight, False, False, self.ln_kv.eps)[0]
        if self.reduce:
            kr_t = xkv
            tq, tkv = xq, xkv
            if self.t_reduce:
                (tq,) = seq(xq, self.t_encoder_q)
                tkv, kr_t = seq(xkv, self.t_encoder_kv, self.t_kr_proj)
            mv_q, mv_kv = lin(tq, self.encoder_q), lin(tkv, self.encoder_kv)
            M = mv_q.shape[-1] // 2
            eps_q = eps_kv = None
            if self.training:   # same draw order/shape/dtype as the reference (:238-239): zq first, then zkv
                eps_q = torch.randn(mv_q.shape[0], mv_q.shape[1], M, dtype=torch.float32, device=mv_q.device)
                eps_kv = torch.randn(mv_kv.shape[0], mv_kv.shape[1], M, dtype=torch.float32, device=mv_kv.device)
            prev = (None, None, None)
            if csm is not None:
                if csm.mode != "sum":
                    raise NotImplementedError("device path implements ResidualStateManager(mode='sum') (reference :499)")
                prev = (csm.zq_sum, csm.zkv_sum, csm.tot_kl_loss if csm.count > 0 else None)
            zq32, zkv32, zq, zkv, kl = ops.LatentFn.apply(mv_q, mv_kv, eps_q, eps_kv, *prev)
            if csm is not None:
                csm.zq_sum, csm.zkv_sum, csm.tot_kl_loss, csm.count = zq32, zkv32, kl, csm.count + 1
            qz = qr = zq
            kz = vz = zkv
            if self.t_reduce:
                qz, qr = seq(zq, self.t_qz_upsample, self.t_qr_proj)
                kz, vz = seq(zkv, self.t_kz_upsample, self.t_vz_upsample)
            srcs = (lin(qz, self.q_proj), lin(qr, self.qr_proj), lin(kz, self.k_proj), lin(kr_t, self.kr_proj), lin(vz, self.v_proj))
            roles = dict(qc=(0, 0), qr=(1, 0), kc=(2, 0), kr=(3, 0), v=(4, 0))
            dc, dr = self.head_dim_content, self.head_dim_rope
        else:
            D = D2
            if input_kv is None:
                srcs = (lin(xq, self.q_proj),)
                roles = dict(qc=None, qr=(0, 0), kc=None, kr=(0, D), v=(0, 2 * D))
            else:
                srcs = (lin(xq, self.q_proj), lin(xkv, self.k_proj))
                roles = dict(qc=None, qr=(0, 0), kc=None, kr=(1, 0), v=(1, D))
            dc, dr = 0, self.head_dim
        lm0, lm2 = self.linear_mask[0], self.linear_mask[2]
        att = ops.AttnCoreFn.apply(tok, self.rope_q.inv_freq, self.rope_k.inv_freq, lm0.bias, lm2.bias, bank, gid(lm0), gid(lm2),
                                   roles, (B, S2, h, dc, dr), *srcs)
        # residual (resized when the stage changes, :302-308): bf16 like the reference's autocast Linear outputs
        if res.shape[1] != S2 or res.shape[2] != D2:
            r = ops.CastBf16Fn.apply(res)
            if self.input_t_proj is not None:
                (r,) = seq(r, self.input_t_proj)
            if self.input_proj is not None:
                r = lin(r, self.input_proj)
            res = r
        x = lin(att, self.out_proj, addend=res, out_f32=True)                    # (attn Wo^T) * ls_att + residual, fp32
        if self.mlp is None:
            return ops.LayerNormFn.apply(x, self.ln_2.weight, True, False, self.ln_2.eps)[0]
        y, res2 = ops.LayerNormFn.apply(x, self.ln_2.weight, False, True, self.ln_2.eps)   # d x feeds out_proj's backward GEMMs
        return ops.MlpFn.apply(y, tok, res2, bank, gid(self.mlp[0]), gid(self.mlp[3]), True)   # x + mlp(y) * ls_mlp


def _cnn(hidden_channels=32):
    return torch.nn.Sequential(
        sn(torch.nn.Conv2d(3, hidden_channels, kernel_size=1, groups=1, bias=True)),
        torch.nn.GELU(approximate='none'),
        sn(torch.nn.Conv2d(hidden_channels, hidden_channels, kernel_size=3, padding=1, bias=True, groups=hidden_channels,
                           padding_mode='zeros')),
        torch.nn.GELU(approximate='none'),
        sn(torch.nn.Conv2d(hidden_channels, 3, kernel_size=1, bias=True)),
    )


def _cnn_groups(proj):
    return [GroupSpec([proj[0]], conv=True), GroupSpec([proj[2]], conv=True), GroupSpec([proj[4]], conv=True)]


def _cnn_apply(sc, proj, x):
    gid = sc.bank.gid
    return ops.CnnFn.apply(x, sc.token, proj[0].bias, proj[2].bias, proj[4].bias, sc.bank, gid(proj[0]), gid(proj[2]), gid(proj[4]))


class Block(torch.nn.Module):
    """Row attention -> column attention -> cross attention (stage change) -> CNN residual (reference :317-403)."""

    def __init__(self, heads: int, dim1: int, dim_step: int, mean_var_hidden: int, seq_length: int, seq_len_step: int,
                 is_first_block: bool, is_last_block: bool, seq_len_reduce: int, force_reduce: bool = False,
                 training: bool = True, use_ape: bool = True,
                 norm_layer: Callable[..., torch.nn.Module] = partial(torch.nn.LayerNorm, eps=1e-6),
                 out_features_override: int = None):
        super().__init__()
        self.is_first_block = is_first_block
        common = dict(heads=heads, dim1=dim1, mean_var_hidden=mean_var_hidden, seq_length=seq_length,
                      seq_len_reduce=seq_len_reduce, force_reduce=force_reduce, training=training, use_mlp=True)
        self.encoder = VMLA_Block(dim2=dim1, seq_len_new=seq_length, mlp_dim=dim1 * 2, **common)
        self.decoder = VMLA_Block(dim2=dim1, seq_len_new=seq_length, mlp_dim=dim1 * 2, **common)
        dim_next = dim1 + dim_step * 3
        self.cross = VMLA_Block(dim2=dim_next if out_features_override is None else out_features_override,
                                seq_len_new=seq_length + seq_len_step * 3, mlp_dim=dim_next * 2, is_cross=True, **common)
        self.proj = _cnn(32)

    def _sn_groups(self):
        return self.encoder._sn_groups() + self.decoder._sn_groups() + self.cross._sn_groups() + _cnn_groups(self.proj)

    def forward(self, x, esm=None, dsm=None, csm=None, mask=True):
        with ops.Scope(self, self._sn_groups, self.training) as sc:
            # first block: the (B,3,S,S) image goes straight into the encoder, whose first LayerNorm also tokenises it
            xq = self.encoder(x, state_manager=esm, mask=mask)
            xkv, xq = ops.TokenSwapFn.apply(xq)                       # columns as tokens (+ alias of the row tokens)
            xkv = self.decoder(xkv, state_manager=dsm, mask=mask)
            xkv = ops.TokenSwapFn.apply(xkv)[0]                       # back to row tokens
            y = self.cross(xq, input_kv=xkv, state_manager=csm, mask=mask)
            return _cnn_apply(sc, self.proj, y)                       # y + CNN(y) on the (B,S,S,3) pixel view


def _stage_blocks(n, make, dim1, seq_length, dim_step, seq_len_step):
    blocks = torch.nn.ModuleList()
    for i in range(n):
        blocks.append(make(i, dim1, seq_length))
        dim1 += dim_step * 3
        seq_length += seq_len_step * 3
    return blocks, dim1, seq_length


class EncoderDecoder_8(torch.nn.Module):
    """3 down blocks, 2 bottleneck blocks, 3 up blocks with U-Net skips, final LayerNorm (reference :407-533)."""

    def __init__(self, heads: int = 12, dim1: int = 768, dim_step: int = 48, mean_var_hidden: int = 192, seq_length: int = 256,
                 seq_len_step: int = 16, seq_len_reduce: int = 128, out_features_override: int = None,
                 force_reduce: bool = False, training: bool = True,
                 norm_layer: Callable[..., torch.nn.Module] = partial(torch.nn.LayerNorm, eps=1e-6)):
        super().__init__()
        self.force_reduce = force_reduce
        if force_reduce:
            raise NotImplementedError("force_reduce=True crashes in the reference as well (SURVEY Appendix E #4)")
        mk = lambda ds, ss, first=False, last=False, ofo=None: (lambda i, d, s: Block(
            heads=heads, dim1=d, dim_step=ds, mean_var_hidden=mean_var_hidden, is_first_block=first and i == 0,
            is_last_block=last and i == 2, seq_length=s, seq_len_step=ss, seq_len_reduce=seq_len_reduce,
            out_features_override=ofo if (last and i == 2) else None, force_reduce=force_reduce, training=training))
        self.encoder_blocks, dim1, seq_length = _stage_blocks(3, mk(-dim_step, -seq_len_step, first=True), dim1, seq_length,
                                                              -dim_step, -seq_len_step)
        self.block_bottle_neck_1 = mk(0, 0)(1, dim1, seq_length)
        self.block_bottle_neck_2 = mk(0, 0)(1, dim1, seq_length)
        self.decoder_blocks, dim1, seq_length = _stage_blocks(3, mk(dim_step, seq_len_step, last=True, ofo=out_features_override),
                                                              dim1, seq_length, dim_step, seq_len_step)
        self.ln_final = ops.check_norm(norm_layer(dim1, bias=False))

    def forward(self, x):
        csm = ResidualStateManager(mode="sum")
        add = ops.Add3Fn.apply
        run = lambda blk, t: blk(t, esm=None, dsm=None, csm=csm, mask=True)
        skip_1 = run(self.encoder_blocks[0], x)
        skip_2 = run(self.encoder_blocks[1], skip_1)
        skip_b1 = run(self.encoder_blocks[2], skip_2)
        skip_b2 = add(run(self.block_bottle_neck_1, skip_b1), skip_b1, None)
        x = add(run(self.block_bottle_neck_2, skip_b2), skip_b2, skip_b1)
        x = add(run(self.decoder_blocks[0], x), skip_2, None)
        x = add(run(self.decoder_blocks[1], x), skip_1, None)
        x = run(self.decoder_blocks[2], x)
        x = ops.LayerNormFn.apply(x, self.ln_final.weight, True, False, self.ln_final.eps)[0]
        kl = csm.get_kl_loss()
        return x, (kl.reshape(()) if torch.is_tensor(kl) else kl)


class CALMLatentDiffusion(torch.nn.Module):
    """Constructible for API parity only — the reference class has no forward() (reference :535-595)."""

    def __init__(self, heads: int = 12, dim1: int = 672, dim_step: int = 48, mean_var_hidden: int = 204,
                 mean_var_hidden_diffusion: int = 96, seq_length: int = 224, seq_len_step: int = 16, seq_len_reduce: int = 80,
                 seq_len_reduce_diffusion: int = 32, out_features_override: int = None, force_reduce: bool = False,
                 training: bool = True, norm_layer: Callable[..., torch.nn.Module] = partial(torch.nn.LayerNorm, eps=1e-6)):
        super().__init__()
        self.force_reduce = force_reduce
        mk = lambda ds, ss, first=False, last=False: (lambda i, d, s: Block(
            heads=heads, dim1=d, dim_step=ds, mean_var_hidden=mean_var_hidden, is_first_block=first and i == 0,
            is_last_block=last and i == 2, seq_length=s, seq_len_step=ss, seq_len_reduce=seq_len_reduce,
            out_features_override=out_features_override if (last and i == 2) else None, force_reduce=force_reduce,
            training=training))
        self.encoder_blocks, dim1, seq_length = _stage_blocks(3, mk(-dim_step, -seq_len_step, first=True), dim1, seq_length,
                                                              -dim_step, -seq_len_step)
        self.decoder_blocks, dim1, seq_length = _stage_blocks(3, mk(dim_step, seq_len_step, last=True), dim1, seq_length,
                                                              dim_step, seq_len_step)
        self.ln_final = ops.check_norm(norm_layer(dim1, bias=False))


class Encoder_8(torch.nn.Module):
    """Encoder-only stack (reference :600-656). Dead code upstream (its defaults give an odd RoPE width and fail at the
    first forward); kept constructible with the same blocks, forward mirrors the skip rule."""

    def __init__(self, heads: int = 12, dim1: int = 672, dim_step: int = 24, mean_var_hidden: int = 192, seq_length: int = 224,
                 seq_len_step: int = 8, seq_len_reduce: int = 96, out_features_override: int = None, force_reduce: bool = False,
                 training: bool = True, norm_layer: Callable[..., torch.nn.Module] = partial(torch.nn.LayerNorm, eps=1e-6)):
        super().__init__()
        self.encoder_blocks = torch.nn.ModuleList()
        self.force_reduce = force_reduce
        for i in range(8):
            step = i in (2, 5)
            self.encoder_blocks.append(Block(
                heads=heads, dim1=dim1, dim_step=-dim_step if step else 0, mean_var_hidden=mean_var_hidden,
                is_first_block=(i == 0), is_last_block=(i == 7), seq_length=seq_length,
                seq_len_step=-seq_len_step if step else 0, seq_len_reduce=seq_len_reduce, out_features_override=None,
                force_reduce=force_reduce, training=training))
            if step:
                dim1 -= dim_step * 3
                seq_length -= seq_len_step * 3
        self.ln_final = ops.check_norm(norm_layer(dim1, bias=False))

    def forward(self, x):
        skip = None
        for block in self.encoder_blocks:
            x = block(x, esm=None, dsm=None, csm=None, mask=True)
            if skip is not None and x.shape == skip.shape:
                x = ops.Add3Fn.apply(x, skip, None)
            skip = x
        return ops.LayerNormFn.apply(x, self.ln_final.weight, True, False, self.ln_final.eps)[0]
