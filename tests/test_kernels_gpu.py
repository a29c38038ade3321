"""GPU parity tests of every kernel family, called through the C ABI (calm_kernels -> ctypes -> libcalm_b200.so) and
checked against plain PyTorch fp32 math on the same (bf16-rounded) inputs. Tolerances are written at each assert.
"""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

bf16, f32 = torch.bfloat16, torch.float32


def dev():
    return torch.device("cuda:0")


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def rnd(*shape, scale=1.0, dtype=bf16, seed=None):
    g = torch.Generator(device="cpu").manual_seed(seed if seed is not None else (hash(shape) & 0xffff))
    return (torch.randn(*shape, generator=g) * scale).to(dev()).to(dtype)


@pytest.fixture(scope="module")
def K():
    import calm_kernels
    return calm_kernels


# ----------------------------------------------------------------------------------------------------------- GEMM
GEMM_SHAPES = [
    # M, N, K
    (128, 64, 64), (300, 264, 240), (256, 672, 672), (1000, 1344, 672), (513, 240, 528), (96, 16, 80), (128, 1000, 1344),
    (384, 480, 240), (200, 528, 176), (130, 8, 64),   # BN=240 (tile width not a multiple of the 32-column epilogue chunk), tails
]


@pytest.mark.parametrize("M,N,Kd", GEMM_SHAPES)
@pytest.mark.parametrize("out_dtype", [bf16, f32])
def test_gemm_kmajor(K, M, N, Kd, out_dtype):
    a, b = rnd(M, Kd, seed=1), rnd(N, Kd, seed=2)
    c = torch.full((M, N), float("nan"), dtype=out_dtype, device=dev())
    K.gemm(a, b, c, M, N, Kd, lda=Kd, ldb=Kd, ldc=N, alpha=0.5)
    ref = 0.5 * (a.float() @ b.float().t())
    assert rel(c, ref) < (1e-5 if out_dtype == f32 else 4e-3)  # bf16 output rounding = 2^-9 relative


def test_gemm_epilogues(K):
    M, N, Kd = 384, 448, 224
    a, b = rnd(M, Kd, seed=3), rnd(N, Kd, scale=0.1, seed=4)
    bias = rnd(N, dtype=f32, seed=5)
    add = rnd(M, N, dtype=f32, seed=6)
    # bias + fp32 addend
    c = torch.empty(M, N, dtype=f32, device=dev())
    K.gemm(a, b, c, M, N, Kd, lda=Kd, ldb=Kd, ldc=N, bias=bias, addend=add, ld_addend=N)
    ref = a.float() @ b.float().t() + bias + add
    assert rel(c, ref) < 1e-5
    # bf16 addend
    addb = add.to(bf16)
    K.gemm(a, b, c, M, N, Kd, lda=Kd, ldb=Kd, ldc=N, addend=addb, ld_addend=N)
    assert rel(c, a.float() @ b.float().t() + addb.float()) < 1e-5
    # GELU: u = bf16(pre); C <- gelu(u), aux <- bf16(gelu'(u))   (the derivative is what the dgrad epilogue consumes)
    aux = torch.empty(M, N, dtype=bf16, device=dev())
    cg = torch.empty(M, N, dtype=bf16, device=dev())
    K.gemm(a, b, cg, M, N, Kd, lda=Kd, ldb=Kd, ldc=N, bias=bias, epilogue=K.EPI_GELU, aux=aux, ld_aux=N)
    u = (a.float() @ b.float().t() + bias).to(bf16).float().requires_grad_(True)
    act = torch.nn.functional.gelu(u)
    act.sum().backward()
    assert rel(cg, act.detach()) < 4e-3
    assert rel(aux, u.grad) < 4e-3
    # dGELU: C <- acc * aux
    g = rnd(M, Kd, seed=7)
    cd = torch.empty(M, N, dtype=bf16, device=dev())
    K.gemm(g, b, cd, M, N, Kd, lda=Kd, ldb=Kd, ldc=N, epilogue=K.EPI_DGELU, aux=aux, ld_aux=N)
    assert rel(cd, (g.float() @ b.float().t()) * aux.float()) < 5e-3
    assert rel(cd, (g.float() @ b.float().t()) * u.grad) < 8e-3   # end to end against the exact derivative


def test_gemm_dgrad_b_mn(K):
    # dX(M,Kin) = dY(M,Nout) . W(Nout,Kin): B read as MN-major straight from W
    M, Nout, Kin = 640, 528, 672
    dy, w = rnd(M, Nout, seed=8), rnd(Nout, Kin, scale=0.05, seed=9)
    dx = torch.empty(M, Kin, dtype=bf16, device=dev())
    K.gemm(dy, w, dx, M, Kin, Nout, lda=Nout, ldb=Kin, ldc=Kin, b_major=K.MAJOR_MN)
    assert rel(dx, dy.float() @ w.float()) < 4e-3


@pytest.mark.parametrize("M,Nout,Kin", [(4096, 240, 480), (5000, 672, 672), (777, 264, 240)])
def test_gemm_wgrad_splitk(K, M, Nout, Kin):
    # dW(Nout,Kin) = dY^T X : both operands MN-major, contraction over tokens, split-K fp32 partials
    dy, x = rnd(M, Nout, seed=10), rnd(M, Kin, seed=11)
    splits = K.gemm_default_splits(Nout, Kin, M)
    part = torch.empty(splits, Nout, Kin, dtype=f32, device=dev())
    K.gemm(dy, x, part, Nout, Kin, M, lda=Nout, ldb=Kin, ldc=Kin, a_major=K.MAJOR_MN, b_major=K.MAJOR_MN, splits=splits,
           stride_split=Nout * Kin)
    assert rel(part.sum(0), dy.float().t() @ x.float()) < 1e-5


def test_gemm_batched_mask_logits(K):
    Bn, S, D = 5, 224, 672
    q, k = rnd(Bn, S, D, seed=12), rnd(Bn, S, D, seed=13)
    c = torch.empty(Bn, S, S, dtype=bf16, device=dev())
    K.gemm(q, k, c, S, S, D, batch=Bn, lda=D, ldb=D, ldc=S, stride_a=S * D, stride_b=S * D, stride_c=S * S)
    assert rel(c, q.float() @ k.float().transpose(1, 2)) < 4e-3
    # backward pieces: dq = dL . k (B MN-major) + addend ; dk = dL^T . q (A MN-major, B MN-major)
    dl = rnd(Bn, S, S, seed=14)
    addq = rnd(Bn, S, D, seed=15)
    dq = torch.empty(Bn, S, D, dtype=bf16, device=dev())
    K.gemm(dl, k, dq, S, D, S, batch=Bn, lda=S, ldb=D, ldc=D, stride_a=S * S, stride_b=S * D, stride_c=S * D,
           b_major=K.MAJOR_MN, addend=addq, ld_addend=D, stride_addend=S * D)
    assert rel(dq, dl.float() @ k.float() + addq.float()) < 4e-3
    dk = torch.empty(Bn, S, D, dtype=bf16, device=dev())
    K.gemm(dl, q, dk, S, D, S, batch=Bn, lda=S, ldb=D, ldc=D, stride_a=S * S, stride_b=S * D, stride_c=S * D,
           a_major=K.MAJOR_MN, b_major=K.MAJOR_MN)
    assert rel(dk, dl.float().transpose(1, 2) @ q.float()) < 4e-3


@pytest.mark.parametrize("S1,S2,D,Bn", [(224, 80, 672, 3), (80, 176, 240, 4), (160, 112, 480, 2)])
def test_gemm_seq_axis(K, S1, S2, D, Bn):
    # Y_b(S2,D) = W(S2,S1) . X_b(S1,D): A broadcast K-major, B MN-major (no transposes materialised)
    w, x = rnd(S2, S1, scale=0.1, seed=16), rnd(Bn, S1, D, seed=17)
    y = torch.empty(Bn, S2, D, dtype=bf16, device=dev())
    K.gemm(w, x, y, S2, D, S1, batch=Bn, lda=S1, ldb=D, ldc=D, stride_a=0, stride_b=S1 * D, stride_c=S2 * D,
           b_major=K.MAJOR_MN)
    assert rel(y, torch.einsum("ts,bsd->btd", w.float(), x.float())) < 4e-3
    # dgrad: dX_b(S1,D) = W^T . dY_b : A MN-major broadcast
    dy = rnd(Bn, S2, D, seed=18)
    dx = torch.empty(Bn, S1, D, dtype=bf16, device=dev())
    K.gemm(w, dy, dx, S1, D, S2, batch=Bn, lda=S1, ldb=D, ldc=D, stride_a=0, stride_b=S2 * D, stride_c=S1 * D,
           a_major=K.MAJOR_MN, b_major=K.MAJOR_MN)
    assert rel(dx, torch.einsum("ts,btd->bsd", w.float(), dy.float())) < 4e-3
    # wgrad: dW(S2,S1) = sum_b dY_b . X_b^T : contraction over (batch, D)
    splits = K.gemm_default_splits(S2, S1, D, Bn, True)
    part = torch.empty(splits, S2, S1, dtype=f32, device=dev())
    K.gemm(dy, x, part, S2, S1, D, batch=Bn, lda=D, ldb=D, ldc=S1, stride_a=S2 * D, stride_b=S1 * D, reduce_batch=True,
           splits=splits, stride_split=S2 * S1)
    assert rel(part.sum(0), torch.einsum("btd,bsd->ts", dy.float(), x.float())) < 1e-5


def test_gemm_strided_views(K):
    # operands / outputs that are column slices of wider buffers (fused qkv layout)
    M, D = 512, 240
    x = rnd(M, D, seed=19)
    w = rnd(3 * D, D, scale=0.1, seed=20)
    qkv = torch.zeros(M, 3 * D, dtype=bf16, device=dev())
    K.gemm(x, w, qkv, M, 3 * D, D, lda=D, ldb=D, ldc=3 * D)
    assert rel(qkv, x.float() @ w.float().t()) < 4e-3
    # dgrad from the k-slice only
    dk = qkv[:, D:2 * D]
    dx = torch.empty(M, D, dtype=bf16, device=dev())
    K.gemm(dk, w[D:2 * D], dx, M, D, D, lda=3 * D, ldb=D, ldc=D, b_major=K.MAJOR_MN)
    assert rel(dx, dk.float() @ w[D:2 * D].float()) < 4e-3


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,Kd", [(20480, 240, 240), (19000, 528, 176), (18952, 96, 80)])
def test_gemm_pair_mode(K, M, N, Kd, a_mn, b_mn):
    """Large-M problems run as 2-CTA clusters: tcgen05.mma.cta_group::2 over a 256-row tile pair with half of B in each CTA
    (flag 8), or cta_group::1 with the B tile TMA-multicast (flags 8|16); odd tile counts included. Both must be bit-identical
    to the single-CTA schedule (same k order per output element)."""
    import calm_lib
    a = rnd(M, Kd, seed=51)
    b = rnd(N, Kd, scale=0.1, seed=52)
    A = a.t().contiguous() if a_mn else a            # MN-major: element (m, k) at [k * ld + m]
    Bm = b.t().contiguous() if b_mn else b
    bias = rnd(N, dtype=f32, seed=53)
    outs = []
    for flags in (8, 4, 8 | 16):                     # per-call calm_gemm_args.flags: FORCE_CLUSTER, NO_CLUSTER, FORCE_CLUSTER | PAIR_MULTICAST
        c = torch.full((M, N), float("nan"), dtype=bf16, device=dev())
        aux = torch.empty(M, N, dtype=bf16, device=dev())
        K.gemm(A, Bm, c, M, N, Kd, lda=M if a_mn else Kd, ldb=N if b_mn else Kd, ldc=N, a_major=a_mn, b_major=b_mn, bias=bias,
               epilogue=K.EPI_GELU, aux=aux, ld_aux=N, flags=flags)
        outs.append((c, aux))
    u = (a.float() @ b.float().t() + bias).to(bf16).float().requires_grad_(True)
    act = torch.nn.functional.gelu(u)
    act.sum().backward()
    assert rel(outs[0][1], u.grad) < 4e-3
    assert rel(outs[0][0], act.detach()) < 4e-3
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[2][0], outs[1][0]) and torch.equal(outs[2][1], outs[1][1])


@pytest.mark.parametrize("M,N,Kd,nb", [(300, 264, 240, 1), (1000, 1344, 128, 1), (130, 8, 64, 1), (224, 224, 96, 5), (96, 40, 80, 3)])
@pytest.mark.parametrize("form", ["bf16", "f32", "f32+add", "bf16+add_inplace", "gelu", "dgelu", "bf16<-f32add"])
def test_gemm_epilogue_forms(K, M, N, Kd, nb, form):
    """The TMA-staged epilogue (default) and the row-owner direct epilogue (flag 32) must agree bit for
    bit on every operand mix, ragged tiles and batches included; the last form has no staged variant and checks the fallback."""
    import calm_lib
    a, b = rnd(nb * M, Kd, seed=61), rnd(nb * N, Kd, scale=0.1, seed=62)
    bias = rnd(N, dtype=f32, seed=63)
    outs = []
    for flags in (0, 32):                           # default (TMA-staged), DIRECT_EPILOGUE
        if True:
            cdt = f32 if form in ("f32", "f32+add") else bf16
            c = torch.full((nb * M, N), float("nan"), dtype=cdt, device=dev())
            kw = dict(lda=Kd, ldb=Kd, ldc=N, batch=nb, stride_a=M * Kd, stride_b=N * Kd, stride_c=M * N, bias=bias, alpha=0.75)
            extra = None
            if form == "f32+add":
                kw.update(addend=rnd(nb * M, N, dtype=f32, seed=64), ld_addend=N, stride_addend=M * N)
            elif form == "bf16+add_inplace":
                c = rnd(nb * M, N, seed=65)
                kw.update(addend=c, ld_addend=N, stride_addend=M * N)
            elif form == "bf16<-f32add":
                kw.update(addend=rnd(nb * M, N, dtype=f32, seed=64), ld_addend=N, stride_addend=M * N)
            elif form == "gelu":
                extra = torch.full((nb * M, N), float("nan"), dtype=bf16, device=dev())
                kw.update(epilogue=K.EPI_GELU, aux=extra, ld_aux=N, stride_aux=M * N)
            elif form == "dgelu":
                extra = rnd(nb * M, N, seed=66)
                kw.update(epilogue=K.EPI_DGELU, aux=extra, ld_aux=N, stride_aux=M * N)
            K.gemm(a, b, c, M, N, Kd, flags=flags, **kw)
            outs.append((c, extra))
    assert not torch.isnan(outs[0][0].float()).any()
    assert torch.equal(outs[0][0], outs[1][0])
    if form == "gelu":
        assert torch.equal(outs[0][1], outs[1][1])
    if form in ("bf16", "f32"):
        ref = 0.75 * torch.bmm(a.view(nb, M, Kd).float(), b.view(nb, N, Kd).float().transpose(1, 2)).reshape(nb * M, N) + bias
        assert rel(outs[0][0], ref) < (1e-5 if form == "f32" else 4e-3)


def test_gemm_rejects_bad_args(K):
    import calm_lib
    a, b = rnd(64, 64), rnd(60, 64)
    c = torch.empty(64, 60, dtype=bf16, device=dev())
    with pytest.raises(calm_lib.CalmError):
        K.gemm(a, b, c, 64, 60, 64, lda=64, ldb=64, ldc=60)  # N % 8 != 0
    with pytest.raises(calm_lib.CalmError):
        K.gemm(a, b, c, 0, 64, 64, lda=64, ldb=64, ldc=64)   # empty


# ------------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("rows,D", [(1000, 672), (333, 240), (4096, 528), (7, 48), (777, 1152), (1030, 1536), (64, 1008), (515, 864)])
def test_layernorm(K, rows, D):
    x = rnd(rows, D, dtype=f32, seed=21) * 3 + 0.5
    w = rnd(D, dtype=f32, seed=22) + 1
    y, mean, rstd = K.layernorm_fwd(x, w)
    ref = torch.nn.functional.layer_norm(x, (D,), w, None, 1e-6)
    assert rel(y, ref) < 4e-3
    y32, _, _ = K.layernorm_fwd(x, w, out_dtype=f32)
    assert rel(y32, ref) < 1e-5
    dy = rnd(rows, D, seed=23)
    dres = rnd(rows, D, dtype=f32, seed=24)
    dx, dw = K.layernorm_bwd(dy, x, w, mean, rstd, dres)
    dx_b, dw_b, dx16 = K.layernorm_bwd(dy, x, w, mean, rstd, dres, want_bf16=True)     # optional bf16 copy for the next GEMM
    assert torch.equal(dx_b, dx) and torch.equal(dw_b, dw) and torch.equal(dx16, dx.to(bf16))
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (D,), wr, None, 1e-6).backward(dy.float())
    assert rel(dx, xr.grad + dres) < 1e-4
    assert rel(dw, wr.grad) < 1e-4


@pytest.mark.parametrize("B,S", [(3, 224), (2, 80), (1, 512), (2, 48)])
def test_layernorm_image(K, B, S):
    """First-block tokenisation fused into the LayerNorm: tokens bit-exact (a pure permutation), LN as the plain kernel."""
    img = rnd(B, 3, S, S, dtype=f32, seed=71) * 2 + 0.3
    w = rnd(3 * S, dtype=f32, seed=72) + 1
    y, tokens, mean, rstd = K.layernorm_fwd_image(img, w)
    ref_tok = img.permute(0, 2, 3, 1).reshape(B, S, 3 * S)
    assert torch.equal(tokens, ref_tok)
    y2, mean2, rstd2 = K.layernorm_fwd(ref_tok.contiguous(), w)
    assert rel(y, torch.nn.functional.layer_norm(ref_tok, (3 * S,), w, None, 1e-6)) < 4e-3
    assert rel(y, y2) < 1e-3 and rel(mean, mean2) < 1e-5 and rel(rstd, rstd2) < 1e-5


# ------------------------------------------------------------------------------------------------------ RoPE
def _rope_ref(x, inv_freq):
    # x (B, h, S, d) fp32; NeoX rotate-half with learned inv_freq
    S = x.shape[2]
    fr = torch.outer(torch.arange(S, device=x.device, dtype=f32), inv_freq)
    emb = torch.cat((fr, fr), -1)
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return x * emb.cos() + torch.cat((-x2, x1), -1) * emb.sin()


@pytest.mark.parametrize("B,S,h,dc,dr", [(3, 80, 12, 0, 20), (2, 176, 12, 22, 22), (2, 224, 12, 0, 56), (9, 176, 12, 0, 44),
                                         (5, 224, 12, 28, 28), (2, 128, 12, 0, 32), (3, 48, 3, 2, 6)])
def test_rope(K, B, S, h, dc, dr):
    tokens = B * S
    inv = (1.0 / (10000.0 ** (torch.arange(0, dr, 2).float() / dr))).to(dev())
    ropein = rnd(tokens, h * dr, seed=25)
    content = rnd(tokens, h * dc, seed=26) if dc else None
    out = K.rope_fwd(content, h * dc, ropein, h * dr, inv, tokens, S, h, dc, dr)
    xr = ropein.float().view(B, S, h, dr).transpose(1, 2).requires_grad_(True)
    invr = inv.clone().requires_grad_(True)
    r = _rope_ref(xr, invr)
    parts = [r] if not dc else [content.float().view(B, S, h, dc).transpose(1, 2), r]
    ref = torch.cat(parts, -1).transpose(1, 2).reshape(tokens, h * (dc + dr))
    assert rel(out, ref) < 4e-3
    dout = rnd(tokens, h * (dc + dr), seed=27)
    dcontent, dropein, dinv = K.rope_bwd(dout, h * (dc + dr), out, inv, tokens, S, h, dc, dr)
    ref.backward(dout.float())
    assert rel(dropein, xr.grad.transpose(1, 2).reshape(tokens, h * dr)) < 4e-3
    # d inv_freq is computed from the bf16-rounded forward output: 2e-2 (the bf16 activation tolerance)
    assert rel(dinv, invr.grad) < 2e-2
    if dc:
        assert torch.equal(dcontent, dout.view(tokens, h, dc + dr)[:, :, :dc].reshape(tokens, h * dc))


# ------------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,S,h,hd", [(2, 224, 12, 56), (2, 176, 12, 44), (3, 128, 12, 32), (2, 80, 12, 20), (2, 16, 12, 4),
                                       (5, 160, 12, 40), (3, 256, 4, 64),                 # tcgen05 / TMEM / TMA kernels (S <= 256, hd <= 64)
                                       (1, 384, 12, 96), (2, 512, 12, 128), (1, 288, 12, 72), (2, 336, 12, 84), (1, 464, 12, 116), (1, 128, 3, 72), (3, 320, 2, 64),   # long-row tcgen05 forward
                                       (2, 72, 12, 20), (2, 136, 4, 96), (1, 76, 3, 20),   # mma.sync kernels
                                       # more (image, head, tile) items than SMs: every persistent CTA walks several items (barrier phases,
                                       # buffer hand-over between items) — whole-row kernels with 2 tiles and 1 tile per item, long-row kernels
                                       (40, 224, 12, 56), (30, 176, 12, 44), (45, 128, 12, 32), (50, 80, 12, 20), (16, 240, 12, 60),
                                       (8, 384, 12, 96), (5, 512, 12, 128), (9, 336, 12, 84),
                                       # warp-per-item register kernels (S <= 128, hd <= 32): every S / 16, padded and unpadded head dims
                                       (2, 96, 3, 24), (1, 48, 5, 12), (2, 112, 12, 28), (1, 64, 2, 32), (2, 32, 1, 8), (700, 16, 12, 16),
                                       # pipelined whole-row forward at its corners: short rows with wide heads, one head, 64-wide heads at S = 224
                                       (2, 64, 3, 48), (1, 96, 2, 64), (1, 224, 2, 64), (150, 48, 1, 36)])
def test_attention(K, B, S, h, hd):
    """The implementation is selected by the shape alone (no process-wide switch): tcgen05 kernels where eligible (S <= 256,
    S % 16 == 0, hd <= 64), the mma.sync kernels otherwise (what the 384^2 / 512^2 configs use; S = 72 / 136 exercise their
    ragged-tile paths)."""
    _attention_case(K, B, S, h, hd)


def _attention_case(K, B, S, h, hd):
    D = h * hd
    qkv = rnd(B * S, 3 * D, seed=28)
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    bias = rnd(B, S, S, seed=29)
    o, lse = K.attention_fwd(q, k, v, bias, B, S, h, hd, 3 * D, 3 * D, 3 * D)
    qr, kr, vr = [t.float().reshape(B, S, h, hd).transpose(1, 2).detach().requires_grad_(True) for t in (q, k, v)]
    br = bias.float().requires_grad_(True)
    s = qr @ kr.transpose(-1, -2) / math.sqrt(hd) + br.unsqueeze(1)
    ref = (torch.softmax(s, -1) @ vr).transpose(1, 2).reshape(B * S, D)
    assert rel(o, ref) < 8e-3   # P is rounded to bf16 before P.V, output rounded to bf16
    assert rel(lse, torch.logsumexp(s, -1)) < 1e-4
    d_o = rnd(B * S, D, seed=30)
    dq, dk, dv, dbias = K.attention_bwd(q, k, v, bias, o, d_o, lse, B, S, h, hd, 3 * D, 3 * D, 3 * D, D)
    ref.backward(d_o.float())
    tok = lambda t: t.transpose(1, 2).reshape(B * S, D)
    assert rel(dv, tok(vr.grad)) < 1e-2
    assert rel(dq, tok(qr.grad)) < 1e-2
    assert rel(dk, tok(kr.grad)) < 1e-2
    assert rel(dbias, br.grad) < 1e-2


@pytest.mark.parametrize("B,S,h,hd", [(64, 224, 12, 56), (64, 176, 12, 44), (96, 128, 12, 32), (128, 80, 12, 20), (16, 384, 12, 96), (8, 512, 12, 128)])
def test_attention_bitwise_repeatable(K, B, S, h, hd):
    """No atomics, fixed summation orders: repeated launches at trainer-like sizes (several items per persistent CTA, so the timing of
    the producer / consumer warps differs from launch to launch) must give bit-identical results — a hand-over race between the
    loader, MMA and worker warps would show up here as a flipped bit long before it shows up as a tolerance failure."""
    D = h * hd
    reps = int(os.environ.get("CALM_TEST_REPEATS", "4"))
    qkv = rnd(B * S, 3 * D, seed=41)
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    bias = rnd(B, S, S, seed=42)
    d_o = rnd(B * S, D, seed=43)
    first = None
    for _ in range(reps):
        o, lse = K.attention_fwd(q, k, v, bias, B, S, h, hd, 3 * D, 3 * D, 3 * D)
        res = (o, lse) + tuple(K.attention_bwd(q, k, v, bias, o, d_o, lse, B, S, h, hd, 3 * D, 3 * D, 3 * D, D))
        res = [t.clone() for t in res]
        if first is None:
            first = res
        else:
            for name, a, b in zip(("o", "lse", "dq", "dk", "dv", "dbias"), first, res):
                assert torch.equal(a, b), "%s differs between two launches on the same inputs" % name


# ------------------------------------------------------------------------------------------------------ latent
@pytest.mark.parametrize("rows,Mh", [(2 * 80, 240), (48, 10), (33, 16)])   # 4-column kernels and the scalar fallback (Mh % 4 != 0)
def test_latent(K, rows, Mh):
    mv = rnd(rows, 2 * Mh, seed=31)
    eps = rnd(rows, Mh, dtype=f32, seed=32)
    prev = rnd(rows, Mh, dtype=f32, seed=33)
    zsum, zb, part = K.latent_fwd(mv, eps, prev)
    mu = mv[:, :Mh].float().requires_grad_(True)
    rho = mv[:, Mh:].float().requires_grad_(True)
    sg = torch.nn.functional.softplus(rho) + 1e-6
    z = mu + eps * sg
    ref = prev + z
    klsum = (1 + 2 * torch.log(sg) - mu.pow(2) - sg.pow(2)).sum()
    assert rel(zsum, ref) < 1e-5
    assert rel(zb, ref) < 4e-3
    assert abs(part.sum().item() - klsum.item()) < 1e-4 * abs(klsum.item()) + 1e-2
    dz = rnd(rows, Mh, dtype=f32, seed=34)
    dkl = torch.tensor([3.0e4], device=dev())   # large enough that the KL term matters next to dz
    kl_scale = -0.5 / (rows * Mh)
    dz2 = rnd(rows, Mh, seed=50)
    dmv, tot = K.latent_bwd(mv, eps, dz, kl_scale, dkl, dz_bf16=dz2, want_total=True)
    assert rel(tot, dz + dz2.float()) < 1e-6
    ((ref * (dz + dz2.float())).sum() + 3.0e4 * kl_scale * klsum).backward()
    assert rel(dmv[:, :Mh], mu.grad) < 4e-3
    assert rel(dmv[:, Mh:], rho.grad) < 4e-3
    klrun = K.latent_kl(part, part, torch.tensor([1.5], device=dev()), kl_scale)
    assert abs(klrun.item() - (1.5 + 2 * kl_scale * klsum.item())) < 1e-4
    # eval mode: no noise, first block: no previous sum
    z0, _, _ = K.latent_fwd(mv, None, None)
    assert rel(z0, mv[:, :Mh].float()) < 1e-6


# ------------------------------------------------------------------------------------------------------ CNN residual
@pytest.mark.parametrize("B,S", [(2, 80), (1, 224), (3, 37), (2, 176), (5, 16), (1, 36), (3, 128), (1, 384), (2, 112)])
def test_cnn(K, B, S):
    """Fused CNN residual against fp32 conv2d. The kernels hold the hidden activations in fp16x2 (11-bit mantissa) and the
    hidden gradients in bf16 with fp32 accumulation; the reference's autocast path stores both in bf16 (8-bit mantissa). The
    bar is therefore stated against that policy, measured in the same test: the kernel's distance to the fp32 result must be
    below HALF of the distance torch's own autocast(bfloat16) conv path shows (and below 4e-3 absolute for the branch)."""
    torch.backends.cudnn.allow_tf32 = False  # the fp32 conv reference must not run on TF32 tensor cores
    x = rnd(B, S, S, 3, dtype=f32, seed=35)
    w1, b1 = rnd(32, 3, dtype=f32, seed=36), rnd(32, dtype=f32, seed=37)
    w2, b2 = rnd(32, 9, dtype=f32, scale=0.3, seed=38), rnd(32, dtype=f32, seed=39)
    w3, b3 = rnd(3, 32, dtype=f32, scale=0.3, seed=40), rnd(3, dtype=f32, seed=41)
    F = torch.nn.functional

    def torch_path(autocast):
        ps = [t.clone().requires_grad_(True) for t in (x, w1, b1, w2, b2, w3, b3)]
        xr, w1r, b1r, w2r, b2r, w3r, b3r = ps
        with torch.autocast("cuda", dtype=bf16, enabled=autocast):
            img = xr.permute(0, 3, 1, 2)
            hmid = F.gelu(F.conv2d(img, w1r.view(32, 3, 1, 1), b1r))
            hmid = F.gelu(F.conv2d(hmid, w2r.view(32, 1, 3, 3), b2r, padding=1, groups=32))
            branch = F.conv2d(hmid, w3r.view(3, 32, 1, 1), b3r).permute(0, 2, 3, 1)
            out = xr + branch
        return ps, branch.float(), out

    y = K.cnn_fwd(x, w1, b1, w2, b2, w3, b3, B, S)
    ps, branch, ref = torch_path(False)
    _, branch_bf, _ = torch_path(True)
    e_fwd, e_pol = rel(y - x, branch), rel(branch_bf, branch)
    print("\n[cnn B=%d S=%d] forward branch: ours %.2e, torch autocast(bf16) %.2e" % (B, S, e_fwd, e_pol))
    assert e_fwd < 4e-3 and e_fwd < 0.5 * e_pol
    assert rel(y, ref) < 2e-3
    dy = rnd(B, S, S, 3, dtype=f32, seed=42) * 300.0      # GradScaler-scaled gradients are large: nothing on that side may be fp16
    dx, gp = K.cnn_bwd(x, dy, w1, b1, w2, b2, w3, b3, B, S)
    dx_b, gp_b, dx16 = K.cnn_bwd(x, dy, w1, b1, w2, b2, w3, b3, B, S, want_bf16=True)
    assert torch.equal(dx_b, dx) and torch.equal(gp_b, gp) and torch.equal(dx16, dx.to(bf16))     # deterministic, bf16 copy exact
    ref.backward(dy)
    grads = lambda ps: torch.cat([ps[1].grad.flatten(), ps[2].grad, ps[3].grad.flatten(), ps[4].grad, ps[5].grad.flatten(), ps[6].grad])
    refg, refdx = grads(ps), ps[0].grad
    ps_bf, _, out_bf = torch_path(True)
    out_bf.backward(dy)
    e_dx, e_dx_pol = rel(dx - dy, refdx - dy), rel(ps_bf[0].grad - dy, refdx - dy)
    names = ["w1", "b1", "w2", "b2", "w3", "b3"]
    offs = [0, 96, 128, 416, 448, 544, 547]
    print("[cnn B=%d S=%d] dx branch: ours %.2e, torch autocast(bf16) %.2e" % (B, S, e_dx, e_dx_pol))
    assert e_dx < 6e-3 and e_dx < 0.75 * e_dx_pol + 1e-3
    assert rel(dx, refdx) < 3e-3
    gbf = grads(ps_bf)
    for n, lo, hi in zip(names, offs[:-1], offs[1:]):
        eo, ep = rel(gp[lo:hi], refg[lo:hi]), rel(gbf[lo:hi], refg[lo:hi])
        print("[cnn B=%d S=%d] d%s: ours %.2e, torch autocast(bf16) %.2e" % (B, S, n, eo, ep))
        assert eo < 5e-3, (n, eo)
    assert rel(gp, refg) < 3e-3


# ------------------------------------------------------------------------------------------------------ helpers
def test_token_helpers(K):
    B, S = 3, 80
    x = rnd(B, S, S * 3, dtype=f32, seed=43)
    t = K.token_transpose(x, B, S)
    assert torch.equal(t, x.view(B, S, S, 3).permute(0, 2, 1, 3).reshape(B, S, 3 * S))  # pure permutation: bit-exact
    img = rnd(B, 3, S, S, dtype=f32, seed=44)
    assert torch.equal(K.nchw_to_tokens(img), img.permute(0, 2, 3, 1).reshape(B, S, 3 * S))
    a, b, c = (rnd(B, S, 3 * S, dtype=f32, seed=45 + i) for i in range(3))
    assert torch.equal(K.add3(a, b), a + b)
    assert rel(K.add3(a, b, c), a + b + c) < 1e-6
    assert torch.equal(K.cast_bf16(a), a.to(bf16))
    assert torch.equal(K.cast_f32(a.to(bf16)), a.to(bf16).float())
    assert rel(K.token_transpose(x, B, S, addend=a), t + a) < 1e-7
    t_b, t16 = K.token_transpose(x, B, S, addend=a, want_bf16=True)
    assert torch.equal(t_b, K.token_transpose(x, B, S, addend=a)) and torch.equal(t16, t_b.to(bf16))
    m = K.seq_mean_fwd(x)
    assert rel(m, x.mean(1)) < 4e-3
    dm = rnd(B, 3 * S, seed=48)
    assert rel(K.seq_mean_bwd(dm, B, S, 3 * S), (dm.float() / S).unsqueeze(1).expand(B, S, 3 * S)) < 1e-6
    xb = rnd(1000, 448, seed=49)
    assert rel(K.colsum(xb, 1000, 448, 448), xb.float().sum(0)) < 1e-5


@pytest.mark.parametrize("rows,N,ld", [(20480, 160, 160), (4099, 80, 240), (57344, 448, 448), (300, 352, 360), (777, 6, 10), (5, 224, 224)])
def test_colsum_shapes(K, rows, N, ld):
    # linear_mask bias gradients: the 128-bit kernel (N, ld multiples of 8) and the 4-byte fallback, strided views included
    buf = rnd(rows, ld, seed=rows % 97)
    got = K.colsum(buf, rows, N, ld)
    assert rel(got, buf[:, :N].float().sum(0)) < 2e-5


# ------------------------------------------------------------------------------------------------------ spectral norm
def _sn_ref(w, u, v, eps=1e-12):
    # torch/nn/utils/spectral_norm.py:97-112, one power iteration
    F = torch.nn.functional
    v2 = F.normalize(torch.mv(w.t(), u), dim=0, eps=eps)
    u2 = F.normalize(torch.mv(w, v2), dim=0, eps=eps)
    sigma = torch.dot(u2, torch.mv(w, v2))
    return u2, v2, sigma


def test_spectral_norm_bank(K):
    shapes = [(672, 672), (1344, 672), (240, 480), (80, 224), (32, 3), (32, 9), (3, 32), (448, 224)]
    F = torch.nn.functional
    ents, refs = [], []
    for i, (r, c) in enumerate(shapes):
        w = rnd(r, c, dtype=f32, seed=60 + i)
        u = F.normalize(rnd(r, dtype=f32, seed=80 + i), dim=0)
        v = F.normalize(rnd(c, dtype=f32, seed=100 + i), dim=0)
        rs = (rnd(r, dtype=f32, seed=120 + i) + 1.5) if i in (0, 1) else None
        small = r * c < 1024
        g = rnd(2, r, c, dtype=f32, seed=140 + i)
        e = dict(w=w, u=u.clone(), v=v.clone(), rowscale=rs, rows=r, cols=c, eff_f32=small,
                 w_eff=torch.empty(r, c, dtype=f32 if small else bf16, device=dev()),
                 grad_w=torch.empty(r, c, dtype=f32, device=dev()),
                 grad_rowscale=torch.empty(r, dtype=f32, device=dev()) if rs is not None else None,
                 g_eff=g, g_splits=2, g_split_stride=r * c,
                 sigma=torch.empty(1, dtype=f32, device=dev()))
        ents.append(e)
        refs.append((w, u, v, rs, g))
    table = K.sn_table(ents, dev())
    mr, mc = max(s[0] for s in shapes), max(s[1] for s in shapes)
    K.sn_forward(table, len(ents), mr, mc, True)
    K.sn_backward(table, len(ents), mr, mc)
    for e, (w, u, v, rs, g) in zip(ents, refs):
        u2, v2, sigma = _sn_ref(w, u, v)
        assert rel(e["u"], u2) < 1e-5 and rel(e["v"], v2) < 1e-5
        assert abs(e["sigma"].item() - sigma.item()) < 1e-5 * abs(sigma.item())
        wo = w.clone().requires_grad_(True)
        rsr = rs.clone().requires_grad_(True) if rs is not None else None
        weff = wo / torch.dot(u2, torch.mv(wo, v2))
        if rsr is not None:
            weff = weff * rsr[:, None]
        tol = 1e-5 if e["eff_f32"] else 4e-3
        assert rel(e["w_eff"], weff) < tol
        weff.backward(g.sum(0))
        assert rel(e["grad_w"], wo.grad) < 1e-4
        if rsr is not None:
            assert rel(e["grad_rowscale"], rsr.grad) < 1e-4
    # eval mode: buffers untouched, sigma from the stored u, v
    before = [(e["u"].clone(), e["v"].clone()) for e in ents]
    K.sn_forward(table, len(ents), mr, mc, False)
    for e, (u0, v0) in zip(ents, before):
        assert torch.equal(e["u"], u0) and torch.equal(e["v"], v0)
        assert abs(e["sigma"].item() - torch.dot(u0, torch.mv(e["w"], v0)).item()) < 1e-4 * abs(e["sigma"].item())
