"""TEST DOUBLE of calm_kernels — test infrastructure only, never imported by the product.

Re-states what each C-ABI kernel family computes (same argument meaning: leading dimensions, batch strides, majors,
epilogues, bf16 rounding points) with plain torch ops on CPU, so that the `-m "not gpu"` suite can exercise the
HOST-side logic of the drop-in modules (autograd wiring, spectral-norm bank bookkeeping, role/stride arithmetic,
state_dict layout) in a container without a GPU. The product path has no such fallback: calm_lib raises without CUDA.
"""
import math

import torch
import torch.nn.functional as F

bf16, f32 = torch.bfloat16, torch.float32
MAJOR_K, MAJOR_MN = 0, 1
EPI_NONE, EPI_GELU, EPI_DGELU = 0, 1, 2
launches = 0


def _view(t, size, stride):
    return torch.as_strided(t, size, stride, t.storage_offset())


def _operand(t, rows, Kd, ld, major, batch, bstride):
    st = (bstride, ld, 1) if major == MAJOR_K else (bstride, 1, ld)
    return _view(t, (batch, rows, Kd), st).float()


def gemm(a, b, c, M, N, K, *, lda, ldb, ldc, batch=1, stride_a=0, stride_b=0, stride_c=0, a_major=MAJOR_K,
         b_major=MAJOR_K, alpha=1.0, bias=None, addend=None, ld_addend=0, stride_addend=0, epilogue=EPI_NONE, aux=None,
         ld_aux=0, stride_aux=0, reduce_batch=False, splits=1, stride_split=0):
    global launches
    launches += 1
    assert N % 8 == 0 and lda % 8 == 0 and ldb % 8 == 0 and ldc % 8 == 0, "alignment rules of calm_gemm"
    A = _operand(a, M, K, lda, a_major, batch, stride_a)
    Bm = _operand(b, N, K, ldb, b_major, batch, stride_b)
    acc = torch.matmul(A, Bm.transpose(1, 2)) * alpha
    if reduce_batch or splits > 1:
        acc = acc.sum(0, keepdim=True) if reduce_batch else acc
        out = _view(c, (splits, M, N), (stride_split, ldc, 1))
        out.zero_()
        out[0].copy_(acc[0])
        return c
    if bias is not None:
        acc = acc + bias.float()
    if addend is not None:
        acc = acc + _view(addend, (batch, M, N), (stride_addend, ld_addend, 1)).float()
    if epilogue == EPI_GELU:   # aux <- gelu'(pre), C <- gelu(pre), pre rounded like the Linear output
        x = acc.to(bf16).float()
        dg = 0.5 * (1 + torch.erf(x / math.sqrt(2))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
        _view(aux, (batch, M, N), (stride_aux, ld_aux, 1)).copy_(dg.to(aux.dtype))
        acc = F.gelu(x)
    elif epilogue == EPI_DGELU:
        acc = acc * _view(aux, (batch, M, N), (stride_aux, ld_aux, 1)).float()
    _view(c, (batch, M, N), (stride_c, ldc, 1)).copy_(acc.to(c.dtype))
    return c


def gemm_default_splits(M, N, K, batch=1, reduce_batch=False):
    return 2 if K >= 64 else 1


def layernorm_fwd(x, w, eps=1e-6, out_dtype=bf16):
    D = x.shape[-1]
    mean = x.mean(-1)
    rstd = (x.var(-1, unbiased=False) + eps).rsqrt()
    y = ((x - mean[..., None]) * rstd[..., None] * w).to(out_dtype)
    return y, mean.reshape(-1), rstd.reshape(-1)


def layernorm_fwd_image(img, w, eps=1e-6):
    B, _, S, _ = img.shape
    tokens = img.permute(0, 2, 3, 1).reshape(B, S, 3 * S).contiguous()
    y, mean, rstd = layernorm_fwd(tokens, w, eps, bf16)     # `bf16` looked up at call time: the fp32-double test rebinds it
    return y, tokens, mean, rstd


def layernorm_bwd(dy, x, w, mean, rstd, dres=None, want_bf16=False):
    D = x.shape[-1]
    xh = (x - mean.view(*x.shape[:-1], 1)) * rstd.view(*x.shape[:-1], 1)
    g = dy.float() * w
    dx = rstd.view(*x.shape[:-1], 1) * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    if dres is not None:
        dx = dx + dres
    dw = (dy.float() * xh).reshape(-1, D).sum(0)
    return (dx, dw, dx.to(torch.bfloat16)) if want_bf16 else (dx, dw)


def _rope_table(inv_freq, S):
    ang = torch.outer(torch.arange(S, dtype=f32), inv_freq.detach())
    return torch.stack((ang.cos(), ang.sin()), -1)


def _rows(t, T, width, ld):
    return _view(t, (T, width), (ld, 1))


def rope_fwd(content, ld_content, ropein, ld_rope, inv_freq, tokens, S, heads, dc, dr):
    half = dr // 2
    cs = _rope_table(inv_freq, S)
    pos = torch.arange(tokens) % S
    c, s = cs[pos, :, 0][:, None, :], cs[pos, :, 1][:, None, :]
    r = _rows(ropein, tokens, heads * dr, ld_rope).float().view(tokens, heads, dr)
    x1, x2 = r[..., :half], r[..., half:]
    rot = torch.cat((x1 * c - x2 * s, x2 * c + x1 * s), -1)
    parts = [rot]
    if dc:
        parts = [_rows(content, tokens, heads * dc, ld_content).float().view(tokens, heads, dc), rot]
    return torch.cat(parts, -1).reshape(tokens, heads * (dc + dr)).to(bf16)


def rope_bwd(dout, ld_dout, out, inv_freq, tokens, S, heads, dc, dr, dcontent=None, ld_dcontent=0, dropein=None, ld_drope=0):
    half = dr // 2
    cs = _rope_table(inv_freq, S)
    pos = torch.arange(tokens) % S
    c, s = cs[pos, :, 0][:, None, :], cs[pos, :, 1][:, None, :]
    d = _rows(dout, tokens, heads * (dc + dr), ld_dout).float().view(tokens, heads, dc + dr)
    y = out.float().view(tokens, heads, dc + dr)
    dy1, dy2 = d[..., dc:dc + half], d[..., dc + half:]
    y1, y2 = y[..., dc:dc + half], y[..., dc + half:]
    dx = torch.cat((dy1 * c + dy2 * s, dy2 * c - dy1 * s), -1).reshape(tokens, heads * dr).to(bf16)
    if dropein is None:
        dropein = torch.empty(tokens, heads * dr, dtype=bf16); ld_drope = heads * dr
    _rows(dropein, tokens, heads * dr, ld_drope).copy_(dx)
    if dc:
        if dcontent is None:
            dcontent = torch.empty(tokens, heads * dc, dtype=bf16); ld_dcontent = heads * dc
        _rows(dcontent, tokens, heads * dc, ld_dcontent).copy_(d[..., :dc].reshape(tokens, heads * dc).to(bf16))
    dth = (y1 * dy2 - y2 * dy1).sum(1)                               # (tokens, half)
    dinv = (dth * pos[:, None].float()).sum(0)
    return dcontent, dropein, dinv


def _heads(t, B, S, h, hd, ld):
    return _rows(t, B * S, h * hd, ld).float().view(B, S, h, hd).transpose(1, 2)


def attention_fwd(q, k, v, bias, B, S, heads, hd, ld_q, ld_k, ld_v):
    Q, Kh, V = _heads(q, B, S, heads, hd, ld_q), _heads(k, B, S, heads, hd, ld_k), _heads(v, B, S, heads, hd, ld_v)
    s = Q @ Kh.transpose(-1, -2) / math.sqrt(hd) + bias.float().unsqueeze(1)
    o = (torch.softmax(s, -1) @ V).transpose(1, 2).reshape(B * S, heads * hd).to(bf16)
    return o, torch.logsumexp(s, -1)


def attention_bwd(q, k, v, bias, o, d_o, lse, B, S, heads, hd, ld_q, ld_k, ld_v, ld_do, dq=None, dk=None, dv=None,
                  ld_dq=0, ld_dk=0, ld_dv=0):
    D = heads * hd
    Q, Kh, V = [t.detach().requires_grad_(True) for t in
                (_heads(q, B, S, heads, hd, ld_q), _heads(k, B, S, heads, hd, ld_k), _heads(v, B, S, heads, hd, ld_v))]
    bz = bias.float().detach().requires_grad_(True)
    with torch.enable_grad():
        s = Q @ Kh.transpose(-1, -2) / math.sqrt(hd) + bz.unsqueeze(1)
        out = (torch.softmax(s, -1) @ V).transpose(1, 2).reshape(B * S, D)
        out.backward(_rows(d_o, B * S, D, ld_do).float())
    tok = lambda t: t.transpose(1, 2).reshape(B * S, D).to(bf16)
    res = []
    for g, buf, ld in ((Q.grad, dq, ld_dq), (Kh.grad, dk, ld_dk), (V.grad, dv, ld_dv)):
        if buf is None:
            buf = torch.empty(B * S, D, dtype=bf16); ld = D
        _rows(buf, B * S, D, ld).copy_(tok(g))
        res.append(buf)
    return res[0], res[1], res[2], bz.grad.to(bf16)


def latent_fwd(mv, eps, zsum_prev):
    Mh = mv.shape[1] // 2
    mu, rho = mv[:, :Mh].float(), mv[:, Mh:].float()
    sg = F.softplus(rho) + 1e-6
    z = mu.clone()
    if eps is not None:
        z = z + eps.reshape(-1, Mh) * sg
    if zsum_prev is not None:
        z = z + zsum_prev.reshape(-1, Mh)
    part = (1 + 2 * torch.log(sg) - mu * mu - sg * sg).sum().reshape(1)
    return z, z.to(bf16), part


def latent_kl(part_q, part_kv, kl_prev, scale):
    out = scale * (part_q.sum() + part_kv.sum())
    return (out + (kl_prev.reshape(()) if kl_prev is not None else 0.0)).reshape(1)


def latent_bwd(mv, eps, dz, kl_scale, dkl, dz_bf16=None, want_total=False):
    Mh = mv.shape[1] // 2
    mu, rho = mv[:, :Mh].float(), mv[:, Mh:].float()
    sg = F.softplus(rho) + 1e-6
    g = torch.zeros_like(mu)
    if dz is not None:
        g = g + dz.reshape(-1, Mh)
    if dz_bf16 is not None:
        g = g + dz_bf16.reshape(-1, Mh).float()
    gk = -2.0 * float(dkl) * kl_scale if dkl is not None else 0.0
    dmu = g + gk * mu
    dsg = gk * (sg - 1 / sg)
    if eps is not None:
        dsg = dsg + g * eps.reshape(-1, Mh)
    return torch.cat((dmu, dsg * torch.sigmoid(rho)), 1).to(bf16), (g.clone() if want_total else None)


def _cnn_ref(x, w1, b1, w2, b2, w3, b3, B, S):
    img = x.reshape(B, S, S, 3).permute(0, 3, 1, 2)
    h = F.gelu(F.conv2d(img, w1.view(32, 3, 1, 1), b1))
    h = F.gelu(F.conv2d(h, w2.view(32, 1, 3, 3), b2, padding=1, groups=32))
    return x + F.conv2d(h, w3.view(3, 32, 1, 1), b3).permute(0, 2, 3, 1).reshape(x.shape)


def cnn_fwd(x, w1, b1, w2, b2, w3, b3, B, S):
    return _cnn_ref(x, w1, b1, w2, b2, w3, b3, B, S)


def cnn_bwd(x, dy, w1, b1, w2, b2, w3, b3, B, S, gp=None, want_bf16=False):
    ps = [t.detach().clone().requires_grad_(True) for t in (x, w1, b1, w2, b2, w3, b3)]
    with torch.enable_grad():
        _cnn_ref(*ps, B, S).backward(dy)
    g = torch.cat([ps[1].grad.flatten(), ps[2].grad, ps[3].grad.flatten(), ps[4].grad, ps[5].grad.flatten(), ps[6].grad])
    if gp is None:
        gp = torch.empty(547)
    gp.copy_(g)
    return (ps[0].grad, gp, ps[0].grad.to(torch.bfloat16)) if want_bf16 else (ps[0].grad, gp)


def token_transpose(x, B, S, addend=None, want_bf16=False):
    out = x.reshape(B, S, S, 3).permute(0, 2, 1, 3).reshape(x.shape).contiguous()
    out = out + addend if addend is not None else out
    return (out, out.to(torch.bfloat16)) if want_bf16 else out


def nchw_to_tokens(x):
    B, _, S, _ = x.shape
    return x.permute(0, 2, 3, 1).reshape(B, S, 3 * S).contiguous()


def colsum(x, rows, N, ld):
    return _rows(x, rows, N, ld).float().sum(0)


def add3(a, b, c=None):
    return a + b + c if c is not None else a + b


def cast_bf16(x):
    return x.to(bf16)


def cast_f32(x):
    return x.float()


def seq_mean_fwd(x):
    return x.mean(1).to(bf16)


def seq_mean_bwd(dout, B, S, D):
    return (dout.float() / S).unsqueeze(1).expand(B, S, D).contiguous()


# ---- spectral norm bank: the "table" is just the python list of entries -----------------------------------------------
def sn_table(entries, device):
    return entries


def sn_forward(table, n_layers, max_rows, max_cols, training, eps=1e-12):
    for e in table:
        w = e["w"].detach().reshape(e["rows"], e["cols"])
        u, v = e["u"], e["v"]
        with torch.no_grad():
            if training:
                v.copy_(F.normalize(torch.mv(w.t(), u), dim=0, eps=eps))
                u.copy_(F.normalize(torch.mv(w, v), dim=0, eps=eps))
            sigma = torch.dot(u, torch.mv(w, v))
            e["sigma"].reshape(-1)[0] = sigma
            weff = w / sigma
            if e.get("rowscale") is not None:
                weff = weff * e["rowscale"].detach()[:, None]
            tgt = e["w_eff"].reshape(-1)[: w.numel()].view(e["rows"], e["cols"])
            tgt.copy_(weff.to(tgt.dtype))


def sn_backward(table, n_layers, max_rows, max_cols):
    for e in table:
        r, c = e["rows"], e["cols"]
        w = e["w"].detach().reshape(r, c)
        sigma = e["sigma"].reshape(-1)[0]
        G = torch.as_strided(e["g_eff"], (e["g_splits"], r, c), (e["g_split_stride"], c, 1), e["g_eff"].storage_offset()).sum(0)
        rs = e["rowscale"].detach() if e.get("rowscale") is not None else None
        if rs is not None:
            e["grad_rowscale"].reshape(-1)[:r].copy_((G * w).sum(1) / sigma)
            G = G * rs[:, None]
        dW = G / sigma - ((G * w).sum() / sigma ** 2) * torch.outer(e["u"], e["v"])
        e["grad_w"].reshape(-1)[: r * c].copy_(dW.reshape(-1))
