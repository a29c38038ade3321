"""Autograd layer of the CALM-ViT hot path: torch.autograd.Functions whose forward/backward launch the sm_100a kernels
of libcalm_b200.so through the C ABI (calm_kernels -> calm_lib -> ctypes). No arithmetic is done by PyTorch here except
the autograd engine's own gradient accumulation where one tensor has several consumers.

Numerics are those of the trainers' `autocast(bfloat16)` region (distributed_trainer_cls.py:84): bf16 GEMM/attention
operands with fp32 accumulation, fp32 LayerNorm statistics / residual stream / latent sampling, fp32 master weights.

Spectral norm: all sn(...) layers of one scope (a Block, a stand-alone VMLA_Block, the ViT head) live in an SNBank —
ONE batched power-iteration launch per forward writes every bf16 effective weight W/sigma (LayerScale folded in), the
weight-gradient GEMMs deposit split-K fp32 partials of dL/dW_eff into bank-owned buffers, and ONE batched launch pair
per backward turns them into dL/dW_orig. The bank's autograd node is ordered after all of its
consumers by a 1-element `token` tensor every consumer takes as an input.
"""
import functools
import os
import threading
import weakref

import torch

import calm_kernels as K
import calm_lib as L
from calm_lib import MAJOR_MN, EPI_NONE, EPI_GELU, EPI_DGELU

bf16, f32 = torch.bfloat16, torch.float32
Function = torch.autograd.Function


@functools.lru_cache(maxsize=None)
def _splits(M, N, Kd, batch=1, reduce_batch=False):
    return K.gemm_default_splits(M, N, Kd, batch, reduce_batch)


def _is_sn(m):
    return hasattr(m, "weight_orig") and hasattr(m, "weight_u") and hasattr(m, "weight_v")


# =====================================================================================================================
# Spectral-norm bank
# =====================================================================================================================
def _require_cuda(dev):
    if dev.type != "cuda":
        raise L.CalmError("calm_b200 modules run on CUDA only (parameters are on %s); there is no CPU fallback" % dev)


class GroupSpec:
    """A set of sn(...) layers with equal fan-in whose effective weights are stacked row-wise into one GEMM operand."""

    def __init__(self, modules, rowscale=None, conv=False):
        self.modules = list(modules)
        self.rowscale = rowscale          # LayerScale Parameter folded into the rows (single-layer groups only)
        self.conv = conv                  # fp32 effective weight for the fused CNN stencil kernel
        assert rowscale is None or len(self.modules) == 1


class WeightSet:
    """Everything one FORWARD of a bank produces and its backward consumes: the bf16 effective weights W/sigma of every group,
    sigma, the split-K partials of dL/dW_eff, the flat dL/dW_orig buffer and the device table that points at all of them.
    A bank normally owns one set and reuses it every step. When a new forward starts while the previous forward's autograd graph
    is still alive (gradient accumulation with a deferred backward, an evaluation pass between forward and backward) that set is
    FROZEN — its table is repointed to a snapshot of u / v, which the new forward is about to advance — and the new forward gets
    another set, so both backward passes see the weights, sigma and u / v of their own forward, as in the reference where every
    forward's `weight` tensor lives in its own graph (torch/nn/utils/spectral_norm.py:92-114)."""

    def __init__(self, bank):
        self.bank = bank
        dev = bank.device
        self.w_eff = [torch.empty(g["rows"], g["cols"], dtype=f32 if g["conv"] else bf16, device=dev) for g in bank.groups]
        self.g_eff = [None] * len(bank.groups)
        self.splits = [1] * len(bank.groups)
        self.sigma = torch.empty(len(bank.layers), dtype=f32, device=dev)
        self.flat_grad = torch.zeros(bank.flat_numel, dtype=f32, device=dev)
        self.cnn_gp = {}
        self.table = None
        self.dirty = True
        self.uv_snap = None        # [(u copy, v copy)] per layer once frozen
        self.lease = None          # weak reference to the autograd node of the forward that owns this set
        self.busy = False          # True from that forward until its bank node has run backward
        self.version = 0

    groups = property(lambda self: self.bank.groups)

    def in_use(self):
        return self.busy and self.lease is not None and self.lease() is not None

    def weight(self, gid):
        return self.w_eff[gid]

    def wgrad_buffer(self, gid, splits):
        if self.g_eff[gid] is None or self.splits[gid] != splits:
            g = self.bank.groups[gid]
            self.g_eff[gid] = torch.empty(splits, g["rows"], g["cols"], dtype=f32, device=self.bank.device)
            self.splits[gid] = splits
            self.dirty = True
        return self.g_eff[gid]

    def cnn_grad_buffer(self, g1, g2, g3):
        """Persistent 547-float parameter-gradient block of one fused CNN; the conv layers' dW_eff are slices of it."""
        buf = self.cnn_gp.get(g1)
        if buf is None:
            buf = self.cnn_gp[g1] = torch.empty(L.CNN_NPARAM, dtype=f32, device=self.bank.device)
            for gid, lo in ((g1, 0), (g2, 128), (g3, 448)):
                g = self.bank.groups[gid]
                self.g_eff[gid] = buf[lo: lo + g["rows"] * g["cols"]].view(1, g["rows"], g["cols"])
                self.splits[gid] = 1
            self.dirty = True
        return buf

    def freeze(self):
        """Called when another forward of the bank is about to advance u / v while this set's backward is still pending."""
        if self.uv_snap is None:
            mods = [l["module"] for l in self.bank.layers]
            self.uv_snap = [(m.weight_u.detach().clone(), m.weight_v.detach().clone()) for m in mods]
            self.dirty = True

    def thaw(self):
        if self.uv_snap is not None:
            self.uv_snap = None
            self.dirty = True

    def upload(self):
        bank = self.bank
        ents = []
        for i, l in enumerate(bank.layers):
            g = bank.groups[l["gid"]]
            m = l["module"]
            lo = l["row_off"] * l["cols"]
            u, v = (m.weight_u, m.weight_v) if self.uv_snap is None else self.uv_snap[i]
            e = dict(w=m.weight_orig, u=u, v=v, rows=l["rows"], cols=l["cols"], eff_f32=g["conv"],
                     w_eff=self.w_eff[l["gid"]].view(-1)[lo:], grad_w=self.flat_grad[l["grad_off"]:], sigma=self.sigma[i:],
                     g_splits=self.splits[l["gid"]], g_split_stride=g["rows"] * g["cols"])
            if self.g_eff[l["gid"]] is not None:
                e["g_eff"] = self.g_eff[l["gid"]].view(-1)[lo:]
            if l["rowscale"] is not None:
                e["rowscale"] = l["rowscale"]
                e["grad_rowscale"] = self.flat_grad[l["rs_off"]:]
            ents.append(e)
        self.table = K.sn_table(ents, bank.device)
        self.dirty = False


class SNBank:
    def __init__(self, specs):
        assert specs
        self.specs = specs
        self.counter = 0
        self.fingerprint = None
        self.groups = []
        self.gid_of = {}
        self.sets = []
        self.active = None
        self._build()

    # ---- construction -------------------------------------------------------------------------------------------
    def _build(self):
        dev = self.specs[0].modules[0].weight_orig.device
        _require_cuda(dev)
        self.device = dev
        if dev.type == "cuda":
            # the tcgen05 GEMM moves 16-byte rows with TMA: every Linear's fan-in and fan-out must be a multiple of 8 (bf16).
            # Said here, with the layer's shape, instead of as a failed launch in the middle of the first forward.
            for sp in self.specs:
                if sp.conv:
                    continue
                for m in sp.modules:
                    o, i = m.weight_orig.shape[0], m.weight_orig[0].numel()
                    if o % 8 or i % 8:
                        raise L.CalmError("calm_b200: sn(Linear(%d -> %d)) is not supported: in/out features must be multiples of 8 (e.g. "
                                          "ViT(out_features=...) and every stage width dim +- k*3*dim_step, seq +- k*3*seq_len_step)" % (i, o))
        self.groups, self.gid_of, self.layers = [], {}, []
        off = 0
        for gid, sp in enumerate(self.specs):
            rows = [m.weight_orig.shape[0] for m in sp.modules]
            cols = sp.modules[0].weight_orig[0].numel()
            assert all(m.weight_orig[0].numel() == cols for m in sp.modules), "fused layers need equal fan-in"
            g = dict(gid=gid, rows=sum(rows), cols=cols, conv=sp.conv, layer_ids=[])
            r0 = 0
            for m, r in zip(sp.modules, rows):
                g["layer_ids"].append(len(self.layers))
                self.layers.append(dict(module=m, gid=gid, row_off=r0, rows=r, cols=cols, grad_off=off, rowscale=sp.rowscale))
                self.gid_of[id(m)] = gid
                off += (r * cols + 3) & ~3          # keep every layer's gradient block 16-byte aligned
                r0 += r
            self.groups.append(g)
        self.rs_layers = [l for l in self.layers if l["rowscale"] is not None]
        for l in self.rs_layers:
            l["rs_off"] = off
            off += (l["rows"] + 3) & ~3
        self.flat_numel = off
        self.max_rows = max(l["rows"] for l in self.layers)
        self.max_cols = max(l["cols"] for l in self.layers)
        self.params = [l["module"].weight_orig for l in self.layers] + [l["rowscale"] for l in self.rs_layers]
        self.sets, self.active = [], None
        self.fingerprint = self._fingerprint()

    def _fingerprint(self):
        """Addresses of EVERY tensor whose pointer is baked into the device table (weight_orig / u / v of each layer, the
        LayerScale vectors): `.to()`, `p.data = ...`, `load_state_dict(assign=True)` or `remove_spectral_norm` on any single layer
        changes it, and the bank is rebuilt before the kernels could touch freed memory."""
        fp = []
        for l in self.layers:
            m = l["module"]
            fp += [m.weight_orig.data_ptr(), m.weight_u.data_ptr(), m.weight_v.data_ptr()]
            if l["rowscale"] is not None:
                fp.append(l["rowscale"].data_ptr())
        return tuple(fp)

    # ---- per-step API -------------------------------------------------------------------------------------------
    def gid(self, module):
        return self.gid_of[id(module)]

    def _acquire(self):
        cur = self.active
        if cur is not None and not cur.in_use():
            ws = cur
        else:
            if cur is not None:
                cur.freeze()                     # its backward is still to come: keep the u / v it was computed from
            ws = next((w for w in self.sets if not w.in_use()), None)
            if ws is None:
                ws = WeightSet(self)
                self.sets.append(ws)
        ws.thaw()
        self.active = ws
        return ws

    def begin(self, training):
        """Runs the batched power iteration into a free weight set; returns (set, autograd token or None)."""
        if self._fingerprint() != self.fingerprint:
            self._build()                      # parameters were moved (.to()) or replaced
        ws = self._acquire()
        if ws.dirty:
            ws.upload()
        self.counter += 1
        ws.version = self.counter
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.params):
            return ws, _SNBankFn.apply(self, ws, bool(training), *self.params)
        ws.busy = False
        K.sn_forward(ws.table, len(self.layers), self.max_rows, self.max_cols, training)
        return ws, None


class _SNBankFn(Function):
    @staticmethod
    def forward(ctx, bank, ws, training, *params):
        K.sn_forward(ws.table, len(bank.layers), bank.max_rows, bank.max_cols, training)
        ctx.bank, ctx.ws, ctx.version = bank, ws, ws.version
        ws.lease, ws.busy = weakref.ref(ctx), True
        return torch.zeros(1, dtype=f32, device=bank.device)

    @staticmethod
    def backward(ctx, _g):
        bank, ws = ctx.bank, ctx.ws
        _check_version(ws, ctx.version)
        for g_eff in ws.g_eff:
            if g_eff is None:
                raise L.CalmError("an sn(...) layer of this scope received no weight gradient (unused in forward?)")
        if ws.dirty:
            ws.upload()
        K.sn_backward(ws.table, len(bank.layers), bank.max_rows, bank.max_cols)
        flat = ws.flat_grad.clone()            # persistent scratch -> tensors autograd may keep as .grad
        ws.busy = False                        # every consumer of this forward has run its backward: the set may be reused
        grads = [flat[l["grad_off"]: l["grad_off"] + l["rows"] * l["cols"]].view_as(l["module"].weight_orig) for l in bank.layers]
        grads += [flat[l["rs_off"]: l["rs_off"] + l["rows"]] for l in bank.rs_layers]
        return (None, None, None, *grads)


def _check_version(ws, version):
    if ws.version != version:
        raise L.CalmError("this forward's effective weights have been released (its backward already ran once — retain_graph is not "
                          "supported — or the spectral-norm bank was rebuilt after the parameters moved)")


_banks = weakref.WeakKeyDictionary()   # owner module -> SNBank (kept out of module __dict__: pickle/deepcopy/state_dict safe)
_tls = threading.local()


class Scope:
    """Context manager giving the sn(...) layers under `owner` their bank for one forward call."""

    def __init__(self, owner, spec_fn, training):
        self.owner, self.spec_fn, self.training = owner, spec_fn, training

    def __enter__(self):
        self.prev = getattr(_tls, "scope", None)
        bank = _banks.get(self.owner)
        if bank is None:
            bank = SNBank(self.spec_fn())
            _banks[self.owner] = bank
        self.bank = bank
        self.ws, self.token = bank.begin(self.training)
        _tls.scope = self
        return self

    def __exit__(self, *exc):
        _tls.scope = self.prev
        return False


def current_scope():
    return getattr(_tls, "scope", None)


# =====================================================================================================================
# raw GEMM helpers on bank weights (feature-axis Linear: y = x W^T)
# =====================================================================================================================
def _as2d(x):
    """(…, K) tensor with contiguous last dim and uniformly strided rows -> (rows, K) view + row stride."""
    if x.dim() == 2:
        assert x.stride(1) == 1
        return x, x.shape[0], x.stride(0)
    x = x.contiguous() if not x.is_contiguous() else x
    x2 = x.view(-1, x.shape[-1])
    return x2, x2.shape[0], x2.stride(0)


def _lin_fwd(ws, gid, x2, M, lda, out, ldc, **kw):
    g = ws.groups[gid]
    K.gemm(x2, ws.w_eff[gid], out, M, g["rows"], g["cols"], lda=lda, ldb=g["cols"], ldc=ldc, **kw)


def _lin_dgrad(ws, gid, dy2, M, ld_dy, out, ld_out, **kw):
    """dX(M, cols) = dY(M, rows) . W(rows, cols) — W read MN-major, no transposed copy."""
    g = ws.groups[gid]
    K.gemm(dy2, ws.w_eff[gid], out, M, g["cols"], g["rows"], lda=ld_dy, ldb=g["cols"], ldc=ld_out, b_major=MAJOR_MN, **kw)


def _lin_wgrad(ws, gid, dy2, ld_dy, x2, ld_x, M):
    """dW_eff(rows, cols) = dY^T X — contraction over the M tokens, split-K fp32 partials into the forward's weight set."""
    g = ws.groups[gid]
    s = _splits(g["rows"], g["cols"], M)
    buf = ws.wgrad_buffer(gid, s)
    K.gemm(dy2, x2, buf, g["rows"], g["cols"], M, lda=ld_dy, ldb=ld_x, ldc=g["cols"], a_major=MAJOR_MN, b_major=MAJOR_MN,
           splits=s, stride_split=g["rows"] * g["cols"])


# ---------------------------------------------------------------------------------------------------------------------
# fork / join of independent launches
# ---------------------------------------------------------------------------------------------------------------------
# Many kernels of a step are small next to the GPU (a GEMM of 160 tiles on 148 SMs leaves the second wave 8 % full) and
# independent of each other (dgrad / wgrad of one Linear, the three GEMMs that follow dpre in an MLP backward, the dq / dk mask
# GEMMs): issued on forked streams they share the SMs — the CTAs of the second kernel start on the SMs the first one's last
# wave leaves idle. Inside the captured training step the fork / join events become parallel branches of the CUDA graph.
# Rules that keep it safe with the caching allocator: branches only LAUNCH (every output is allocated before the fork, on the
# forking stream) and the join happens before the autograd Function returns.
PARALLEL = os.environ.get("CALM_FORK", "1") != "0"      # bench.py's per-kernel timing pass switches it off (kernels timed alone)
_side_streams = {}


def _side(dev, i):
    key = (dev.index, i)
    st = _side_streams.get(key)
    if st is None:
        st = _side_streams[key] = torch.cuda.Stream(device=dev)
    return st


def fork_join(dev, *branches):
    """Runs the callables concurrently: branches[0] on the current stream, the others on side streams forked from / joined to it."""
    branches = [b for b in branches if b is not None]
    if not PARALLEL or len(branches) < 2 or dev.type != "cuda":
        for b in branches:
            b()
        return
    cur = torch.cuda.current_stream(dev)
    start = cur.record_event()
    done = []
    for i, b in enumerate(branches[1:]):
        st = _side(dev, i)
        st.wait_event(start)
        with torch.cuda.stream(st):
            b()
            done.append(st.record_event())
    branches[0]()
    for e in done:
        cur.wait_event(e)


# bf16 copies of fp32 gradients that their producer (LayerNorm backward) already wrote: id(tensor) -> (weakref, bf16 tensor).
# A hit requires the very same tensor object autograd hands on (identity through the weak reference), so a gradient that was
# accumulated, copied or whose memory was recycled can never pick up a stale copy; entries die with their fp32 tensor.
_BF16_SIDE = {}
_SIDE_ON = os.environ.get("CALM_BF16_SIDE", "1") != "0"     # A/B switch: 0 = every fp32 gradient is cast by its consumer


def _offer_bf16(g, g16):
    key = id(g)
    _BF16_SIDE[key] = (weakref.ref(g, lambda _r, k=key: _BF16_SIDE.pop(k, None)), g16)


def _bf16_grad(g):
    if g.dtype == bf16:
        return g
    e = _BF16_SIDE.get(id(g))
    if e is not None and e[0]() is g and e[1].shape == g.shape:
        return e[1]
    return K.cast_bf16(g.contiguous())


# =====================================================================================================================
# Functions
# =====================================================================================================================
def check_norm(m):
    """The path implements the trainers' norm_layer: weight-only torch.nn.LayerNorm (`partial(LayerNorm, eps=1e-6)` with
    bias=False, Vi_Tools_CNN_less_V2.py:115,131). Anything else is refused at construction instead of being silently ignored."""
    if not isinstance(m, torch.nn.LayerNorm) or m.bias is not None or m.weight is None or len(m.normalized_shape) != 1:
        raise NotImplementedError("calm_b200 implements weight-only torch.nn.LayerNorm over the feature axis (norm_layer(dim, bias=False)); "
                                  "got %r" % (m,))
    return m


class LayerNormFn(Function):
    """y = LN(x)*w in bf16 (or fp32), plus an alias of x so that the residual consumer's gradient is fused into dx."""

    @staticmethod
    def forward(ctx, x, w, out_f32, grad_feeds_gemm=False, eps=1e-6):
        """grad_feeds_gemm: the gradient of x goes straight into a Linear's backward (ln_2 -> out_proj): LayerNorm backward then
        also writes it as bf16, the form that GEMM reads."""
        x = x.contiguous()
        y, mean, rstd = K.layernorm_fwd(x, w, float(eps), f32 if out_f32 else bf16)
        ctx.save_for_backward(x, w, mean, rstd)
        ctx.set_materialize_grads(False)
        ctx.side = bool(grad_feeds_gemm) and _SIDE_ON
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dres):
        x, w, mean, rstd = ctx.saved_tensors
        if dy is None:
            return dres, None, None, None, None
        dres = dres.contiguous() if dres is not None else None
        if not ctx.side:
            dx, dw = K.layernorm_bwd(dy.contiguous(), x, w, mean, rstd, dres)
            return dx, dw, None, None, None
        dx, dw, dx16 = K.layernorm_bwd(dy.contiguous(), x, w, mean, rstd, dres, want_bf16=True)
        _offer_bf16(dx, dx16)       # the producing Linear's backward takes this gradient as a bf16 GEMM operand
        return dx, dw, None, None, None


class ImageLayerNormFn(Function):
    """First block: (B,3,S,S) image -> (LN(tokens) * w in bf16, the fp32 row tokens) in one kernel (Vi_Tools…:389-391 + :211)."""

    @staticmethod
    def forward(ctx, img, w, eps):
        if img.dim() != 4 or img.shape[1] != 3 or img.shape[2] != img.shape[3]:
            raise L.CalmError("expected a (B, 3, S, S) image, got %s" % (tuple(img.shape),))
        y, tokens, mean, rstd = K.layernorm_fwd_image(img.contiguous().float(), w, float(eps))
        ctx.save_for_backward(tokens, w, mean, rstd)
        ctx.set_materialize_grads(False)
        return y, tokens

    @staticmethod
    def backward(ctx, dy, dres):
        tokens, w, mean, rstd = ctx.saved_tensors
        if dy is None:
            dx, dw = dres, None
        else:
            dx, dw = K.layernorm_bwd(dy.contiguous(), tokens, w, mean, rstd, dres.contiguous() if dres is not None else None)
        dimg = None
        if ctx.needs_input_grad[0] and dx is not None:      # cold path: only if the input image needs a gradient
            B, S, _ = tokens.shape
            dimg = dx.view(B, S, S, 3).permute(0, 3, 1, 2).contiguous()
        return dimg, dw, None


class LinearFn(Function):
    """y = x W_eff^T (+ addend) for one bank group; x bf16 (…, K); out bf16 or fp32."""

    @staticmethod
    def forward(ctx, x, token, addend, bank, gid, out_f32):
        bank = bank.active          # the weight set of THIS forward (effective weights, sigma, wgrad partials)
        g = bank.groups[gid]
        x2, M, lda = _as2d(x)
        out = torch.empty(*x.shape[:-1], g["rows"], dtype=f32 if out_f32 else bf16, device=x.device)
        kw = {}
        if addend is not None:
            a2, _, ld_a = _as2d(addend)
            kw = dict(addend=a2, ld_addend=ld_a)
        _lin_fwd(bank, gid, x2, M, lda, out, g["rows"], **kw)
        ctx.save_for_backward(x)
        ctx.bank, ctx.gid, ctx.version = bank, gid, bank.version
        ctx.addend_dtype = addend.dtype if addend is not None else None
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        bank, gid = ctx.bank, ctx.gid
        _check_version(bank, ctx.version)
        g = bank.groups[gid]
        dyb = _bf16_grad(dy)
        dy2, M, ld_dy = _as2d(dyb)
        x2, _, ld_x = _as2d(x)
        dx = torch.empty(x.shape, dtype=bf16, device=x.device) if ctx.needs_input_grad[0] else None
        fork_join(x.device,
                  (lambda: _lin_dgrad(bank, gid, dy2, M, ld_dy, dx, g["cols"])) if dx is not None else None,
                  (lambda: _lin_wgrad(bank, gid, dy2, ld_dy, x2, ld_x, M)) if ctx.needs_input_grad[1] else None)
        dadd = None
        if ctx.needs_input_grad[2]:
            dadd = dy if ctx.addend_dtype == dy.dtype else (dyb if ctx.addend_dtype == bf16 else K.cast_f32(dy))
        return dx, None, dadd, None, None, None


def _mlp_forward(bank, g1, g2, x2, M, lda, b1, b2, addend2, ld_add, out):
    """hidden = gelu(x W1^T + b1) ; out = hidden W2^T + b2 + addend. Returns (pre, hidden) for backward, where `pre` holds
    gelu'(pre-activation): the GELU epilogue saves the derivative, the dgrad epilogue multiplies by it."""
    H = bank.groups[g1]["rows"]
    pre = torch.empty(M, H, dtype=bf16, device=x2.device)
    hid = torch.empty(M, H, dtype=bf16, device=x2.device)
    _lin_fwd(bank, g1, x2, M, lda, hid, H, bias=b1, epilogue=EPI_GELU, aux=pre, ld_aux=H)
    kw = dict(addend=addend2, ld_addend=ld_add) if addend2 is not None else {}
    _lin_fwd(bank, g2, hid, M, H, out, bank.groups[g2]["rows"], bias=b2, **kw)
    return pre, hid


def _mlp_backward(bank, g1, g2, dy2, M, ld_dy, x2, ld_x, pre, hid, need_dx, need_w, need_bias):
    """Returns (dx bf16 | None, db1, db2)."""
    H, N2, K1 = bank.groups[g1]["rows"], bank.groups[g2]["rows"], bank.groups[g1]["cols"]
    dpre = torch.empty(M, H, dtype=bf16, device=dy2.device)
    dx = torch.empty(M, K1, dtype=bf16, device=dy2.device) if need_dx else None
    # the weight gradient of the second layer needs dy and the hidden activation only: it runs beside the dgrad that produces dpre
    fork_join(dy2.device,
              lambda: _lin_dgrad(bank, g2, dy2, M, ld_dy, dpre, H, epilogue=EPI_DGELU, aux=pre, ld_aux=H),
              (lambda: _lin_wgrad(bank, g2, dy2, ld_dy, hid, H, M)) if need_w else None)
    db1 = db2 = None

    def _bias():
        nonlocal db1, db2
        db2 = K.colsum(dy2, M, N2, ld_dy)
        db1 = K.colsum(dpre, M, H, H)
    # first branch = forking stream: the only one that allocates (colsum partials)
    fork_join(dy2.device,
              _bias if need_bias else None,
              (lambda: _lin_dgrad(bank, g1, dpre, M, H, dx, K1)) if need_dx else None,
              (lambda: _lin_wgrad(bank, g1, dpre, H, x2, ld_x, M)) if need_w else None)
    return dx, db1, db2


class MlpFn(Function):
    """Two sn(Linear)s with an exact GELU between them (mlp: Vi_Tools…:199-205,311-314; cls head: CALM_ViT_V2.py:49-53)."""

    @staticmethod
    def forward(ctx, x, token, addend, bank, g1, g2, out_f32):
        bank = bank.active          # the weight set of THIS forward (effective weights, sigma, wgrad partials)
        x2, M, lda = _as2d(x)
        N2 = bank.groups[g2]["rows"]
        out = torch.empty(*x.shape[:-1], N2, dtype=f32 if out_f32 else bf16, device=x.device)
        a2 = ld_a = None
        if addend is not None:
            a2, _, ld_a = _as2d(addend)
        pre, hid = _mlp_forward(bank, g1, g2, x2, M, lda, None, None, a2, ld_a or 0, out)
        ctx.save_for_backward(x)
        ctx.pre, ctx.hid = pre, hid
        ctx.bank, ctx.g1, ctx.g2, ctx.version = bank, g1, g2, bank.version
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        bank = ctx.bank
        _check_version(bank, ctx.version)
        dyb = _bf16_grad(dy)
        dy2, M, ld_dy = _as2d(dyb)
        x2, _, ld_x = _as2d(x)
        dx, _, _ = _mlp_backward(bank, ctx.g1, ctx.g2, dy2, M, ld_dy, x2, ld_x, ctx.pre, ctx.hid, ctx.needs_input_grad[0],
                                 ctx.needs_input_grad[1], False)
        ctx.pre = ctx.hid = None
        return (dx.view(x.shape) if dx is not None else None), None, (dy if ctx.needs_input_grad[2] else None), None, None, None, None


class SeqLinearFn(Function):
    """sn(Linear)s applied along the SEQUENCE axis of x (B, S1, D): Y_i[b] = W_i (S2_i x S1) . X[b] — batched left-multiply
    GEMMs, no permute copies (Vi_Tools…:224-229,250-264,304-306). Several layers may share the input (one dX)."""

    @staticmethod
    def forward(ctx, x, token, bank, *gids):
        bank = bank.active          # the weight set of THIS forward (effective weights, sigma, wgrad partials)
        x = x.contiguous()
        B, S1, D = x.shape
        outs = []
        for gid in gids:
            g = bank.groups[gid]
            assert g["cols"] == S1
            outs.append(torch.empty(B, g["rows"], D, dtype=bf16, device=x.device))

        def _one(gid, y):
            g = bank.groups[gid]
            return lambda: K.gemm(bank.w_eff[gid], x, y, g["rows"], D, S1, batch=B, lda=S1, ldb=D, ldc=D, stride_a=0, stride_b=S1 * D,
                                  stride_c=g["rows"] * D, b_major=MAJOR_MN)
        fork_join(x.device, *[_one(gid, y) for gid, y in zip(gids, outs)])
        ctx.save_for_backward(x)
        ctx.bank, ctx.gids, ctx.version = bank, gids, bank.version
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *dys):
        (x,) = ctx.saved_tensors
        bank = ctx.bank
        _check_version(bank, ctx.version)
        B, S1, D = x.shape
        dys = [_bf16_grad(dy).contiguous() if dy is not None else torch.zeros(B, bank.groups[gid]["rows"], D, dtype=bf16, device=x.device)
               for gid, dy in zip(ctx.gids, dys)]
        dxs = [torch.empty(B, S1, D, dtype=bf16, device=x.device) for _ in ctx.gids] if ctx.needs_input_grad[0] else []

        def _dgrad_chain():                      # dX = sum_i W_i^T dY_i, accumulated through the epilogue addend
            prev = None
            for gid, dy, nx in zip(ctx.gids, dys, dxs):
                g = bank.groups[gid]
                S2 = g["rows"]
                kw = dict(addend=prev, ld_addend=D, stride_addend=S1 * D) if prev is not None else {}
                K.gemm(bank.w_eff[gid], dy, nx, S1, D, S2, batch=B, lda=S1, ldb=D, ldc=D, stride_a=0, stride_b=S2 * D,
                       stride_c=S1 * D, a_major=MAJOR_MN, b_major=MAJOR_MN, **kw)
                prev = nx

        def _wgrad(gid, dy):
            S2 = bank.groups[gid]["rows"]
            s = _splits(S2, S1, D, B, True)
            buf = bank.wgrad_buffer(gid, s)
            return lambda: K.gemm(dy, x, buf, S2, S1, D, batch=B, lda=D, ldb=D, ldc=S1, stride_a=S2 * D, stride_b=S1 * D,
                                  reduce_batch=True, splits=s, stride_split=S2 * S1)
        fork_join(x.device, _dgrad_chain if dxs else None,
                  *([_wgrad(gid, dy) for gid, dy in zip(ctx.gids, dys)] if ctx.needs_input_grad[1] else []))
        return ((dxs[-1] if dxs else None), None, None) + (None,) * len(ctx.gids)


class AttnCoreFn(Function):
    """RoPE (+ content|rope concat) -> all-head mask logits -> mask MLP over the key axis -> softmax(QK^T/sqrt(hd)+bias)V.
    (Vi_Tools…:271-299). `roles` maps qc/qr/kc/kr/v to (source index, column offset) inside the bf16 source matrices
    (fused qkv GEMM outputs or separate projections); gradients are written straight into per-source buffers."""

    @staticmethod
    def forward(ctx, token, inv_q, inv_k, b1, b2, bank, g1, g2, roles, dims, *srcs):
        bank = bank.active          # the weight set of THIS forward (effective weights, sigma, wgrad partials)
        B, S, heads, dc, dr = dims
        hd = dc + dr
        D = heads * hd
        T = B * S
        src2 = [_as2d(s) for s in srcs]

        def role(name):
            if roles[name] is None:
                return None, 0
            i, off = roles[name]
            t2, _, ld = src2[i]
            return t2[:, off:], ld

        qc, ld_qc = role("qc"); qr, ld_qr = role("qr")
        kc, ld_kc = role("kc"); kr, ld_kr = role("kr")
        v, ld_v = role("v")
        roped = {}

        def _rope(name, c_, ldc_, r_, ldr_, inv):
            def run():
                roped[name] = K.rope_fwd(c_, ldc_, r_, ldr_, inv, T, S, heads, dc, dr)
            return run
        # the two rotations are independent 10 - 30 us kernels: side by side, each hides the other's launch / tail latency
        fork_join(v.device, _rope("q", qc, ld_qc, qr, ld_qr, inv_q), _rope("k", kc, ld_kc, kr, ld_kr, inv_k))
        q, k = roped["q"], roped["k"]
        logits = torch.empty(B, S, S, dtype=bf16, device=q.device)
        K.gemm(q, k, logits, S, S, D, batch=B, lda=D, ldb=D, ldc=S, stride_a=S * D, stride_b=S * D, stride_c=S * S)
        bias = torch.empty(B, S, S, dtype=bf16, device=q.device)
        pre, hid = _mlp_forward(bank, g1, g2, logits.view(T, S), T, S, b1, b2, None, 0, bias)
        o, lse = K.attention_fwd(q, k, v, bias, B, S, heads, hd, D, D, ld_v)
        ctx.save_for_backward(inv_q, inv_k, *srcs)
        ctx.saved = (q, k, logits, pre, hid, bias, o, lse)
        ctx.bank, ctx.g1, ctx.g2, ctx.version = bank, g1, g2, bank.version
        ctx.roles, ctx.dims = roles, dims
        return o.view(B, S, D)

    @staticmethod
    def backward(ctx, d_o):
        inv_q, inv_k, *srcs = ctx.saved_tensors
        q, k, logits, pre, hid, bias, o, lse = ctx.saved
        bank = ctx.bank
        _check_version(bank, ctx.version)
        B, S, heads, dc, dr = ctx.dims
        hd = dc + dr
        D = heads * hd
        T = B * S
        roles = ctx.roles
        src2 = [_as2d(s) for s in srcs]
        dsrc = [torch.empty(s.shape, dtype=bf16, device=s.device) for s in srcs]
        dsrc2 = [_as2d(d) for d in dsrc]

        def role(name, bufs):
            if roles[name] is None:
                return None, 0
            i, off = roles[name]
            t2, _, ld = bufs[i]
            return t2[:, off:], ld

        v, ld_v = role("v", src2)
        dv, ld_dv = role("v", dsrc2)
        d_o2, _, ld_do = _as2d(_bf16_grad(d_o))
        dq = torch.empty(T, D, dtype=bf16, device=q.device)
        dk = torch.empty(T, D, dtype=bf16, device=q.device)
        _, _, _, dbias = K.attention_bwd(q, k, v, bias, o, d_o2, lse, B, S, heads, hd, D, D, ld_v, ld_do, dq=dq, dk=dk, dv=dv,
                                         ld_dq=D, ld_dk=D, ld_dv=ld_dv)
        need_w = ctx.needs_input_grad[0]
        dlog, db1, db2 = _mlp_backward(bank, ctx.g1, ctx.g2, dbias.view(T, S), T, S, logits.view(T, S), S, pre, hid, True, need_w, True)
        # logits = q k^T (all heads, unscaled): dq += dL k ; dk += dL^T q   (accumulated in place through the addend)
        dqc, ld_dqc = role("qc", dsrc2); dqr, ld_dqr = role("qr", dsrc2)
        dkc, ld_dkc = role("kc", dsrc2); dkr, ld_dkr = role("kr", dsrc2)
        dinv = {}

        def _q_side():      # the q chain and the k chain are independent from here on: mask-logit dgrad, then the rotation's backward
            K.gemm(dlog, k, dq, S, D, S, batch=B, lda=S, ldb=D, ldc=D, stride_a=S * S, stride_b=S * D, stride_c=S * D,
                   b_major=MAJOR_MN, addend=dq, ld_addend=D, stride_addend=S * D)
            dinv["q"] = K.rope_bwd(dq, D, q, inv_q, T, S, heads, dc, dr, dcontent=dqc, ld_dcontent=ld_dqc, dropein=dqr, ld_drope=ld_dqr)[2]

        def _k_side():
            K.gemm(dlog, q, dk, S, D, S, batch=B, lda=S, ldb=D, ldc=D, stride_a=S * S, stride_b=S * D, stride_c=S * D,
                   a_major=MAJOR_MN, b_major=MAJOR_MN, addend=dk, ld_addend=D, stride_addend=S * D)
            dinv["k"] = K.rope_bwd(dk, D, k, inv_k, T, S, heads, dc, dr, dcontent=dkc, ld_dcontent=ld_dkc, dropein=dkr, ld_drope=ld_dkr)[2]
        fork_join(q.device, _q_side, _k_side)
        ctx.saved = None
        return (None, dinv["q"], dinv["k"], db1, db2, None, None, None, None, None, *dsrc)


class LatentFn(Function):
    """Latent bottleneck sampling + ResidualStateManager('sum') running sums and KL (Vi_Tools…:232-244, 23-30)."""

    @staticmethod
    def forward(ctx, mv_q, mv_kv, eps_q, eps_kv, prev_q, prev_kv, prev_kl):
        B, R, M2 = mv_q.shape
        Mh = M2 // 2
        rows = B * R
        zq, zq16, pq = K.latent_fwd(mv_q.view(rows, M2), eps_q, prev_q)
        zkv, zkv16, pkv = K.latent_fwd(mv_kv.view(rows, M2), eps_kv, prev_kv)
        scale = -0.5 / (rows * Mh)
        kl = K.latent_kl(pq, pkv, prev_kl, scale)
        ctx.save_for_backward(mv_q, mv_kv, eps_q, eps_kv)
        ctx.scale = scale
        ctx.has_prev = (prev_q is not None, prev_kv is not None, prev_kl is not None)
        ctx.set_materialize_grads(False)
        shp = (B, R, Mh)
        return zq.view(shp), zkv.view(shp), zq16.view(shp), zkv16.view(shp), kl

    @staticmethod
    def backward(ctx, dzq, dzkv, dzq16, dzkv16, dkl):
        mv_q, mv_kv, eps_q, eps_kv = ctx.saved_tensors
        B, R, M2 = mv_q.shape
        rows = B * R
        c = lambda t: t.contiguous() if t is not None else None
        hq, hkv, hkl = ctx.has_prev
        dmv_q, tq = K.latent_bwd(mv_q.view(rows, M2), eps_q, c(dzq), ctx.scale, dkl, c(dzq16), want_total=hq)
        dmv_kv, tkv = K.latent_bwd(mv_kv.view(rows, M2), eps_kv, c(dzkv), ctx.scale, dkl, c(dzkv16), want_total=hkv)
        shp = (B, R, M2 // 2)
        return (dmv_q.view(B, R, M2), dmv_kv.view(B, R, M2), None, None, (tq.view(shp) if hq else None),
                (tkv.view(shp) if hkv else None), (dkl if hkl else None))


class CnnFn(Function):
    """Per-Block CNN residual, one fused stencil kernel each way (Vi_Tools…:378-385,400-403; CALM_ViT_V2.py:60-67,80-83)."""

    @staticmethod
    def forward(ctx, x, token, b1, b2, b3, bank, g1, g2, g3):
        bank = bank.active          # the weight set of THIS forward (effective weights, sigma, wgrad partials)
        x = x.contiguous()
        B, S = x.shape[0], x.shape[1]
        w1, w2, w3 = bank.weight(g1), bank.weight(g2), bank.weight(g3)
        y = K.cnn_fwd(x, w1, b1, w2, b2, w3, b3, B, S)
        ctx.save_for_backward(x, b1, b2, b3)
        ctx.bank, ctx.gs, ctx.version = bank, (g1, g2, g3), bank.version
        return y

    @staticmethod
    def backward(ctx, dy):
        x, b1, b2, b3 = ctx.saved_tensors
        bank = ctx.bank
        _check_version(bank, ctx.version)
        g1, g2, g3 = ctx.gs
        B, S = x.shape[0], x.shape[1]
        # weight gradients wrt the effective conv weights land in the bank: w1[0:96] w2[128:416] w3[448:544]
        gp = bank.cnn_grad_buffer(g1, g2, g3)
        if _SIDE_ON:
            dx, gp, dx16 = K.cnn_bwd(x, dy.contiguous(), bank.weight(g1), b1, bank.weight(g2), b2, bank.weight(g3), b3, B, S, gp=gp,
                                     want_bf16=True)
            _offer_bf16(dx, dx16)   # next consumer: the cross block's MLP backward (bf16 GEMM operands)
        else:
            dx, gp = K.cnn_bwd(x, dy.contiguous(), bank.weight(g1), b1, bank.weight(g2), b2, bank.weight(g3), b3, B, S, gp=gp)
        return dx, None, gp[96:128].clone(), gp[416:448].clone(), gp[544:547].clone(), None, None, None, None


class TokenSwapFn(Function):
    """Row tokens <-> column tokens of the (B,S,S,3) pixel grid (Vi_Tools…:394-395,397-398). Also returns an alias of the
    input so that the gradient of the other consumer (the cross block's query path) is fused into the backward transpose."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.set_materialize_grads(False)
        return K.token_transpose(x, x.shape[0], x.shape[1]), x.view_as(x)

    @staticmethod
    def backward(ctx, dt, dalias):
        if dt is None:
            return dalias
        dt = dt.contiguous()
        add = dalias.contiguous() if dalias is not None else None
        if not _SIDE_ON:
            return K.token_transpose(dt, dt.shape[0], dt.shape[1], addend=add)
        d, d16 = K.token_transpose(dt, dt.shape[0], dt.shape[1], addend=add, want_bf16=True)
        _offer_bf16(d, d16)         # next consumer: the previous attention block's MLP backward (bf16 GEMM operands)
        return d


class ImageToTokensFn(Function):
    """(B,3,S,S) image -> S row tokens of width 3S (Vi_Tools…:389-391)."""

    @staticmethod
    def forward(ctx, img):
        return K.nchw_to_tokens(img.contiguous().float())

    @staticmethod
    def backward(ctx, d):
        B, S, _ = d.shape
        return d.view(B, S, S, 3).permute(0, 3, 1, 2).contiguous()   # cold path: only if the input image needs a gradient


class Add3Fn(Function):
    """U-Net skip adds (Vi_Tools…:513,516,520,522), out of place."""

    @staticmethod
    def forward(ctx, a, b, c):
        ctx.has_c = c is not None
        return K.add3(a.contiguous(), b.contiguous(), c.contiguous() if c is not None else None)

    @staticmethod
    def backward(ctx, g):
        return g, g, (g if ctx.has_c else None)


class SeqMeanFn(Function):
    """Mean over the sequence axis feeding the classifier head (CALM_ViT_V2.py:73-75): fp32 (B,S,D) -> bf16 (B,D)."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = x.shape
        return K.seq_mean_fwd(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        B, S, D = ctx.shape
        return K.seq_mean_bwd(_bf16_grad(g).contiguous(), B, S, D)


class CastBf16Fn(Function):
    @staticmethod
    def forward(ctx, x):
        return K.cast_bf16(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return K.cast_f32(g.contiguous())
