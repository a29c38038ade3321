"""Raw (non-autograd) Python wrappers over the C-ABI entry points of libcalm_b200.so.

One function per kernel family; each allocates its outputs with torch (device memory is the caller's — the library
allocates nothing), passes raw pointers / sizes through ctypes and launches on the current torch stream.
The autograd layer (calm_ops.py) and the GPU parity tests both go through these functions, i.e. through the C ABI.
"""
import ctypes as C
import os

import torch

import calm_lib as L
from calm_lib import BF16, F32, MAJOR_K, MAJOR_MN, EPI_NONE, EPI_GELU, EPI_DGELU, ptr

bf16 = torch.bfloat16
f32 = torch.float32
SN_ITEM_WEIGHTS = int(os.environ.get("CALM_SN_ITEM_WEIGHTS", "0"))   # tuning hook: work-item size of the spectral-norm kernels (0 = by table size)


def _dt(t):
    if t.dtype == bf16:
        return BF16
    if t.dtype == f32:
        return F32
    raise L.CalmError("unsupported dtype %s" % t.dtype)


def gemm(a, b, c, M, N, K, *, lda, ldb, ldc, batch=1, stride_a=0, stride_b=0, stride_c=0, a_major=MAJOR_K,
         b_major=MAJOR_K, alpha=1.0, bias=None, addend=None, ld_addend=0, stride_addend=0, epilogue=EPI_NONE, aux=None,
         ld_aux=0, stride_aux=0, reduce_batch=False, splits=1, stride_split=0, flags=0, bn=0):
    """C[b](M,N) = epi(alpha * A[b](M,K) . B[b](N,K)^T + bias + addend) — see calm_gemm in include/calm_b200.h."""
    assert a.dtype == bf16 and b.dtype == bf16
    g = L.GemmArgs()
    g.a, g.b, g.c = ptr(a), ptr(b), ptr(c)
    g.M, g.N, g.K, g.batch = M, N, K, batch
    g.lda, g.ldb, g.ldc = lda, ldb, ldc
    g.stride_a, g.stride_b, g.stride_c = stride_a, stride_b, stride_c
    g.a_major, g.b_major, g.c_dtype, g.epilogue = a_major, b_major, _dt(c), epilogue
    g.bias = ptr(bias)
    g.addend = ptr(addend)
    g.addend_dtype = _dt(addend) if addend is not None else F32
    g.ld_addend, g.stride_addend = ld_addend, stride_addend
    g.aux, g.ld_aux, g.stride_aux = ptr(aux), ld_aux, stride_aux
    g.reduce_batch, g.splits, g.stride_split = int(reduce_batch), splits, stride_split
    g.alpha = alpha
    g.flags, g.bn_override = flags, bn          # per-call schedule selectors (tests / tuning); 0 = the heuristics
    tag = None
    if L.profile is not None:
        # algorithmic bytes of this launch: every operand read once, the result (and the side outputs) written once
        nbytes = 2 * M * K * (batch if stride_a else 1) + 2 * N * K * (batch if stride_b else 1)
        nbytes += M * N * c.element_size() * (1 if reduce_batch else batch) * max(splits, 1)
        if addend is not None:
            nbytes += M * N * addend.element_size() * batch
        if aux is not None:
            nbytes += M * N * 2 * batch
        tag = "M%d N%d K%d b%d %s%s%s%s%s bytes=%d" % (M, N, K, batch, "AK" if a_major == MAJOR_K else "AM", "BK" if b_major == MAJOR_K else "BM",
                                               " f32" if c.dtype == f32 else "", " add" if addend is not None else "",
                                               (" epi%d" % epilogue if epilogue else "") + (" red" if reduce_batch else "") + (" s%d" % splits if splits > 1 else ""), nbytes)
    L.call("calm_gemm", C.byref(g), work=2.0 * M * N * K * batch, tag=tag)
    return c


def gemm_default_splits(M, N, K, batch=1, reduce_batch=False):
    return L.load().calm_gemm_default_splits(M, N, K, batch, int(reduce_batch))


# ------------------------------------------------------------------------------------------------ LayerNorm
def layernorm_fwd(x, w, eps=1e-6, out_dtype=bf16):
    """x f32 (..., D) -> y (bf16|f32), mean, rstd."""
    D = x.shape[-1]
    rows = x.numel() // D
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    mean = torch.empty(rows, dtype=f32, device=x.device)
    rstd = torch.empty(rows, dtype=f32, device=x.device)
    L.call("calm_layernorm_fwd", ptr(x), ptr(w), ptr(y), _dt(y), ptr(mean), ptr(rstd), rows, D, eps,
           work=float(rows) * D * (4 + y.element_size()))
    return y, mean, rstd


def layernorm_fwd_image(img, w, eps=1e-6):
    """img f32 (B,3,S,S) -> tokens f32 (B,S,3S) [= permute(0,2,3,1).reshape], y bf16 = LN(tokens) * w, mean, rstd: one pass."""
    B, _, S, _ = img.shape
    tokens = torch.empty(B, S, 3 * S, dtype=f32, device=img.device)
    y = torch.empty(B, S, 3 * S, dtype=bf16, device=img.device)
    mean = torch.empty(B * S, dtype=f32, device=img.device)
    rstd = torch.empty(B * S, dtype=f32, device=img.device)
    L.call("calm_layernorm_fwd_image", ptr(img), ptr(w), ptr(y), ptr(tokens), ptr(mean), ptr(rstd), B, S, eps, work=float(B) * 3 * S * S * 10)
    return y, tokens, mean, rstd


def layernorm_bwd(dy, x, w, mean, rstd, dres=None, want_bf16=False):
    """dx f32 = LN'(dy) (+ dres), dw f32 (D); with want_bf16 also the bf16 copy of dx (third result)."""
    D = x.shape[-1]
    rows = x.numel() // D
    nparts = L.load().calm_layernorm_bwd_parts(rows, D)
    dx = torch.empty(x.shape, dtype=f32, device=x.device)
    part = torch.empty(nparts * D, dtype=f32, device=x.device)
    dw = torch.empty(D, dtype=f32, device=x.device)
    dx16 = torch.empty(x.shape, dtype=bf16, device=x.device) if want_bf16 else None
    L.call("calm_layernorm_bwd", ptr(dy), _dt(dy), ptr(x), ptr(w), ptr(mean), ptr(rstd), ptr(dres), ptr(dx), ptr(dx16), ptr(part),
           nparts, ptr(dw), rows, D, work=float(rows) * D * (dy.element_size() + 4 + (4 if dres is not None else 0) + 4 + (2 if want_bf16 else 0)))
    return (dx, dw, dx16) if want_bf16 else (dx, dw)


# ------------------------------------------------------------------------------------------------ RoPE
def rope_fwd(content, ld_content, ropein, ld_rope, inv_freq, tokens, S, heads, dc, dr):
    """inv_freq f32 (dr/2): the learned frequencies; the kernels form cos / sin of position * inv_freq themselves."""
    out = torch.empty(tokens, heads * (dc + dr), dtype=bf16, device=ropein.device)
    L.call("calm_rope_fwd", ptr(content), ld_content, ptr(ropein), ld_rope, ptr(out), out.stride(0), ptr(inv_freq), tokens, S,
           heads, dc, dr, work=4.0 * tokens * heads * (dc + dr))
    return out


def rope_bwd(dout, ld_dout, out, inv_freq, tokens, S, heads, dc, dr, dcontent=None, ld_dcontent=0, dropein=None, ld_drope=0):
    """Returns (dcontent, dropein, dinv_freq); dcontent/dropein may be pre-allocated views (e.g. slices of one buffer)."""
    dev = dout.device
    if dc > 0 and dcontent is None:
        dcontent = torch.empty(tokens, heads * dc, dtype=bf16, device=dev)
        ld_dcontent = heads * dc
    if dropein is None:
        dropein = torch.empty(tokens, heads * dr, dtype=bf16, device=dev)
        ld_drope = heads * dr
    scratch = torch.empty(L.load().calm_rope_bwd_scratch_floats(S, dr), dtype=f32, device=dev)
    dinv = torch.empty(dr // 2, dtype=f32, device=dev)
    L.call("calm_rope_bwd", ptr(dout), ld_dout, ptr(out), out.stride(0), ptr(dcontent), ld_dcontent, ptr(dropein), ld_drope,
           ptr(inv_freq), ptr(scratch), ptr(dinv), tokens, S, heads, dc, dr, work=6.0 * tokens * heads * (dc + dr))
    return dcontent, dropein, dinv


# ------------------------------------------------------------------------------------------------ attention
def attention_fwd(q, k, v, bias, B, S, heads, hd, ld_q, ld_k, ld_v):
    """q/k/v bf16 token-major views; bias bf16 (B,S,S). Returns o bf16 (B*S, heads*hd), lse f32 (B, heads, S)."""
    o = torch.empty(B * S, heads * hd, dtype=bf16, device=q.device)
    lse = torch.empty(B, heads, S, dtype=f32, device=q.device)
    L.call("calm_attention_fwd", ptr(q), ptr(k), ptr(v), ptr(bias), ptr(o), ptr(lse), ld_q, ld_k, ld_v, o.stride(0), B, S,
           heads, hd, work=4.0 * B * heads * S * S * hd)
    return o, lse


def attention_bwd(q, k, v, bias, o, d_o, lse, B, S, heads, hd, ld_q, ld_k, ld_v, ld_do, dq=None, dk=None, dv=None,
                  ld_dq=0, ld_dk=0, ld_dv=0):
    dev = q.device
    D = heads * hd
    if dq is None:
        dq = torch.empty(B * S, D, dtype=bf16, device=dev); ld_dq = D
    if dk is None:
        dk = torch.empty(B * S, D, dtype=bf16, device=dev); ld_dk = D
    if dv is None:
        dv = torch.empty(B * S, D, dtype=bf16, device=dev); ld_dv = D
    dbias = torch.empty(B, S, S, dtype=bf16, device=dev)
    delta = torch.empty(B, heads, S, dtype=f32, device=dev)
    # per-head dS scratch of the tcgen05 path (summed over the heads by its second kernel)
    scratch = torch.empty(L.load().calm_attention_bwd_scratch_bytes(B, S, heads, hd), dtype=torch.uint8, device=dev)
    L.call("calm_attention_bwd", ptr(q), ptr(k), ptr(v), ptr(bias), ptr(o), ptr(d_o), ptr(lse), ptr(delta), ptr(dq), ptr(dk),
           ptr(dv), ptr(dbias), ptr(scratch), ld_q, ld_k, ld_v, o.stride(0), ld_do, ld_dq, ld_dk, ld_dv, B, S, heads, hd,
           work=10.0 * B * heads * S * S * hd)
    return dq, dk, dv, dbias


# ------------------------------------------------------------------------------------------------ latent
def latent_fwd(mv, eps, zsum_prev):
    """mv bf16 (rows, 2M) -> zsum f32 (rows, M), zsum bf16, kl_sum f32 scalar tensor (sum over elements of the KL integrand)."""
    rows, two_m = mv.shape[0], mv.shape[1]
    Mh = two_m // 2
    dev = mv.device
    nblocks = L.load().calm_latent_blocks(rows, Mh)
    zsum = torch.empty(rows, Mh, dtype=f32, device=dev)
    zb = torch.empty(rows, Mh, dtype=bf16, device=dev)
    part = torch.empty(nblocks, dtype=f32, device=dev)
    L.call("calm_latent_fwd", ptr(mv), ptr(eps), ptr(zsum_prev), ptr(zsum), ptr(zb), ptr(part), nblocks, rows, Mh)
    return zsum, zb, part


def latent_kl(part_q, part_kv, kl_prev, scale):
    out = torch.empty(1, dtype=f32, device=part_q.device)
    L.call("calm_latent_kl", ptr(part_q), ptr(part_kv), part_q.numel(), ptr(kl_prev), ptr(out), float(scale))
    return out


def latent_bwd(mv, eps, dz, kl_scale, dkl, dz_bf16=None, want_total=False):
    """Returns (dmv bf16, dz_total f32 | None): dz_total = dz + dz_bf16 is the gradient wrt the previous running sum."""
    rows, two_m = mv.shape[0], mv.shape[1]
    dmv = torch.empty_like(mv)
    tot = torch.empty(rows, two_m // 2, dtype=f32, device=mv.device) if want_total else None
    L.call("calm_latent_bwd", ptr(mv), ptr(eps), ptr(dz), ptr(dz_bf16), float(kl_scale), ptr(dkl), ptr(dmv), ptr(tot), rows,
           two_m // 2)
    return dmv, tot


# ------------------------------------------------------------------------------------------------ CNN residual
def cnn_fwd(x, w1, b1, w2, b2, w3, b3, B, S):
    y = torch.empty_like(x)
    L.call("calm_cnn_fwd", ptr(x), ptr(y), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(w3), ptr(b3), B, S, work=24.0 * B * S * S)
    return y


def cnn_bwd(x, dy, w1, b1, w2, b2, w3, b3, B, S, gp=None, want_bf16=False):
    """Returns dx f32 and gparams f32 (547): w1[96] b1[32] w2[288] b2[32] w3[96] b3[3] (+ the bf16 copy of dx with want_bf16)."""
    nblocks = L.load().calm_cnn_bwd_blocks(B, S)
    dx = torch.empty_like(x)
    part = torch.empty(nblocks * L.CNN_NPARAM, dtype=f32, device=x.device)
    if gp is None:
        gp = torch.empty(L.CNN_NPARAM, dtype=f32, device=x.device)
    dx16 = torch.empty(x.shape, dtype=bf16, device=x.device) if want_bf16 else None
    L.call("calm_cnn_bwd", ptr(x), ptr(dy), ptr(dx), ptr(dx16), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(w3), ptr(b3), ptr(part), nblocks,
           ptr(gp), B, S, work=(36.0 + (6.0 if want_bf16 else 0.0)) * B * S * S)
    return (dx, gp, dx16) if want_bf16 else (dx, gp)


# ------------------------------------------------------------------------------------------------ helpers
def token_transpose(x, B, S, addend=None, want_bf16=False):
    out = torch.empty_like(x)
    out16 = torch.empty(x.shape, dtype=bf16, device=x.device) if want_bf16 else None
    L.call("calm_token_transpose", ptr(x), ptr(addend), ptr(out), ptr(out16), B, S,
           work=3.0 * B * S * S * (8 + (4 if addend is not None else 0) + (2 if want_bf16 else 0)))
    return (out, out16) if want_bf16 else out


def nchw_to_tokens(x):
    B, _, S, _ = x.shape
    out = torch.empty(B, S, 3 * S, dtype=f32, device=x.device)
    L.call("calm_nchw_to_tokens", ptr(x), ptr(out), B, S)
    return out


def colsum(x, rows, N, ld):
    nparts = L.load().calm_colsum_parts(rows, N)
    part = torch.empty(nparts * N, dtype=f32, device=x.device)
    out = torch.empty(N, dtype=f32, device=x.device)
    L.call("calm_colsum", ptr(x), ld, ptr(part), nparts, ptr(out), rows, N, work=2.0 * rows * N)
    return out


def add3(a, b, c=None):
    out = torch.empty_like(a)
    L.call("calm_add3", ptr(a), ptr(b), ptr(c), ptr(out), a.numel())
    return out


def cast_bf16(x):
    out = torch.empty(x.shape, dtype=bf16, device=x.device)
    L.call("calm_cast_bf16", ptr(x), ptr(out), x.numel())
    return out


def cast_f32(x):
    out = torch.empty(x.shape, dtype=f32, device=x.device)
    L.call("calm_cast_f32", ptr(x), ptr(out), x.numel())
    return out


def seq_mean_fwd(x):
    B, S, D = x.shape
    out = torch.empty(B, D, dtype=bf16, device=x.device)
    L.call("calm_seq_mean_fwd", ptr(x), ptr(out), B, S, D)
    return out


def seq_mean_bwd(dout, B, S, D):
    dx = torch.empty(B, S, D, dtype=f32, device=dout.device)
    L.call("calm_seq_mean_bwd", ptr(dout), ptr(dx), B, S, D)
    return dx


# ------------------------------------------------------------------------------------------------ spectral norm
class SnTable:
    """Device-side description of a scope's sn(...) layers: the calm_sn_layer table, the (layer, row-chunk) work items and
    the scratch the kernels need. Holds references to every tensor whose pointer is baked into the table."""

    def __init__(self, entries, device):
        n = len(entries)
        arr = (L.SnLayer * n)()
        items = []
        # ~4 K weights per item at the 224^2 scopes (5 M weights: ~1.2 k CTAs; same-box sweep 16 K 59.3, 8 K 59.1, 4 K 58.9 ms/step); the
        # 384^2 / 512^2 scopes (20 - 45 M weights) take proportionally larger items: the per-layer reduction of the partials
        # (one CTA per layer) grows with the item count, not with the weights
        total = sum(e["rows"] * e["cols"] for e in entries)
        item_weights = SN_ITEM_WEIGHTS or max(4096, min(65536, total // 1184))
        self.keep = [entries]
        for i, e in enumerate(entries):
            rows, cols = e["rows"], e["cols"]
            chunk = max(4, min(rows, -(-item_weights // cols)))
            nit = -(-rows // chunk)
            for j in range(nit):
                items.append((i, j * chunk, min(rows, (j + 1) * chunk), j))
            tpart = torch.empty(nit * cols, dtype=f32, device=device)
            svec = torch.empty(rows, dtype=f32, device=device)
            self.keep += [tpart, svec]
            s = arr[i]
            for name in ("w", "u", "v", "rowscale", "w_eff", "grad_w", "grad_rowscale", "g_eff", "sigma"):
                t = e.get(name)
                setattr(s, name, ptr(t) if t is not None else None)
            s.tpart, s.svec = ptr(tpart), ptr(svec)
            s.rows, s.cols = rows, cols
            s.g_splits = e.get("g_splits", 1)
            s.eff_f32 = int(e.get("eff_f32", 0))
            s.g_split_stride = e.get("g_split_stride", rows * cols)
            s.item_count = nit
        it = (L.SnItem * len(items))()
        for j, (li, rb, re, loc) in enumerate(items):
            it[j].layer, it[j].row_begin, it[j].row_end, it[j].local_index = li, rb, re, loc
        L._tls.dev = None      # the pointers above were baked into a table, not handed to a launch
        self.n_layers, self.n_items = n, len(items)
        self.table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
        self.items = torch.frombuffer(bytearray(bytes(it)), dtype=torch.uint8).to(device)


def sn_table(entries, device):
    """entries: list of dicts with the calm_sn_layer fields (tensors or ints)."""
    return SnTable(entries, device)


def sn_forward(table, n_layers, max_rows, max_cols, training, eps=1e-12):
    L.call("calm_sn_forward", ptr(table.table), table.n_layers, ptr(table.items), table.n_items, int(training), eps)


def sn_backward(table, n_layers, max_rows, max_cols):
    L.call("calm_sn_backward", ptr(table.table), table.n_layers, ptr(table.items), table.n_items)
