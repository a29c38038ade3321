"""ctypes binding of libcalm_b200.so — the C-ABI kernel library declared in include/calm_b200.h.

Only plumbing lives here: locate/load the shared library, declare every prototype, translate torch tensors into raw
device pointers and raise on a non-zero return code. There is NO fallback: if the library is missing (or there is no
CUDA device) every call raises, so a silent PyTorch/CPU path can never stand in for the kernels.
"""
import ctypes as C
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcalm_b200.so")

BF16, F32 = 0, 1
MAJOR_K, MAJOR_MN = 0, 1
EPI_NONE, EPI_GELU, EPI_DGELU = 0, 1, 2
GEMM_SIMT, GEMM_NO_CLUSTER, GEMM_FORCE_CLUSTER, GEMM_PAIR_MULTICAST, GEMM_DIRECT_EPILOGUE = 1, 4, 8, 16, 32   # calm_gemm_args.flags (per call)
CNN_NPARAM = 547
OPT_CHUNK = 8192
OPT_SCALE, OPT_GROWTH_TRACKER, OPT_STEP, OPT_LR, OPT_GRAD_NORM, OPT_FOUND_INF, OPT_MULT, OPT_BIAS1, OPT_BIAS2_SQRT = range(9)
OPT_STATE_FLOATS = 16

i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class GemmArgs(C.Structure):
    """Mirror of calm_gemm_args (include/calm_b200.h)."""
    _fields_ = [
        ("a", vp), ("b", vp), ("c", vp),
        ("M", i32), ("N", i32), ("K", i32), ("batch", i32),
        ("lda", i64), ("ldb", i64), ("ldc", i64),
        ("stride_a", i64), ("stride_b", i64), ("stride_c", i64),
        ("a_major", i32), ("b_major", i32), ("c_dtype", i32), ("epilogue", i32),
        ("bias", vp),
        ("addend", vp), ("addend_dtype", i32), ("bn_override", i32), ("ld_addend", i64), ("stride_addend", i64),
        ("aux", vp), ("ld_aux", i64), ("stride_aux", i64),
        ("reduce_batch", i32), ("splits", i32), ("stride_split", i64),
        ("alpha", f32), ("flags", i32),
    ]


class SnLayer(C.Structure):
    """Mirror of calm_sn_layer (include/calm_b200.h)."""
    _fields_ = [
        ("w", vp), ("u", vp), ("v", vp), ("rowscale", vp), ("w_eff", vp),
        ("grad_w", vp), ("grad_rowscale", vp), ("g_eff", vp), ("tpart", vp), ("svec", vp), ("sigma", vp),
        ("g_split_stride", i64),
        ("rows", i32), ("cols", i32), ("g_splits", i32), ("eff_f32", i32),
        ("item_count", i32), ("_pad", i32),
    ]


class TrainerStepArgs(C.Structure):
    """Mirror of calm_trainer_step_args."""
    _fields_ = [
        ("params", vp), ("grads", vp), ("grads_host", vp), ("elem_off", vp), ("chunk_tensor", vp), ("chunk_start", vp),
        ("exp_avg", vp), ("exp_avg_sq", vp), ("partial", vp), ("state", vp),
        ("n_tensors", i32), ("n_chunks", i32),
        ("beta1", f32), ("beta2", f32), ("eps", f32), ("weight_decay", f32), ("max_norm", f32),
        ("growth_factor", f32), ("backoff_factor", f32), ("growth_interval", i32), ("use_scaler", i32), ("_pad", i32),
    ]


class SnItem(C.Structure):
    """Mirror of calm_sn_item."""
    _fields_ = [("layer", i32), ("row_begin", i32), ("row_end", i32), ("local_index", i32)]


# name -> (restype, argtypes); every symbol include/calm_b200.h declares
PROTOTYPES = {
    "calm_abi_version": (i32, []),
    "calm_last_error": (C.c_char_p, []),
    "calm_gemm": (i32, [C.POINTER(GemmArgs), vp]),
    "calm_gemm_default_splits": (i32, [i32, i32, i32, i32, i32]),
    "calm_sn_forward": (i32, [vp, i32, vp, i32, i32, f32, vp]),
    "calm_sn_backward": (i32, [vp, i32, vp, i32, vp]),
    "calm_layernorm_fwd": (i32, [vp, vp, vp, i32, vp, vp, i64, i32, f32, vp]),
    "calm_layernorm_fwd_image": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, f32, vp]),
    "calm_layernorm_bwd": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, i64, i32, vp]),
    "calm_layernorm_bwd_parts": (i32, [i64, i32]),
    "calm_rope_fwd": (i32, [vp, i64, vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, vp]),
    "calm_rope_bwd_scratch_floats": (i32, [i32, i32]),
    "calm_rope_bwd": (i32, [vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, i64, i32, i32, i32, i32, vp]),
    "calm_attention_fwd": (i32, [vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, i32, i32, i32, i32, vp]),
    "calm_attention_bwd_scratch_bytes": (i64, [i32, i32, i32, i32]),
    "calm_attention_bwd": (i32, [vp] * 13 + [i64] * 8 + [i32] * 4 + [vp]),
    "calm_latent_fwd": (i32, [vp, vp, vp, vp, vp, vp, i32, i64, i32, vp]),
    "calm_latent_kl": (i32, [vp, vp, i32, vp, vp, f32, vp]),
    "calm_latent_bwd": (i32, [vp, vp, vp, vp, f32, vp, vp, vp, i64, i32, vp]),
    "calm_latent_blocks": (i32, [i64, i32]),
    "calm_cnn_fwd": (i32, [vp] * 8 + [i32, i32, vp]),
    "calm_cnn_bwd": (i32, [vp] * 11 + [i32, vp, i32, i32, vp]),
    "calm_cnn_bwd_blocks": (i32, [i32, i32]),
    "calm_token_transpose": (i32, [vp, vp, vp, vp, i32, i32, vp]),
    "calm_nchw_to_tokens": (i32, [vp, vp, i32, i32, vp]),
    "calm_colsum": (i32, [vp, i64, vp, i32, vp, i64, i32, vp]),
    "calm_colsum_parts": (i32, [i64, i32]),
    "calm_add3": (i32, [vp, vp, vp, vp, i64, vp]),
    "calm_cast_bf16": (i32, [vp, vp, i64, vp]),
    "calm_cast_f32": (i32, [vp, vp, i64, vp]),
    "calm_seq_mean_fwd": (i32, [vp, vp, i32, i32, i32, vp]),
    "calm_seq_mean_bwd": (i32, [vp, vp, i32, i32, i32, vp]),
    "calm_soft_ce_fwd": (i32, [vp, i64, vp, i64, vp, vp, vp, i32, i32, vp]),
    "calm_soft_ce_bwd": (i32, [vp, i64, vp, i64, vp, vp, vp, vp, i64, i32, i32, vp]),
    "calm_huber_parts": (i32, [i32, i32]),
    "calm_huber_tokens_fwd": (i32, [vp, vp, vp, f32, f32, vp, i32, vp, i32, i32, vp]),
    "calm_huber_tokens_bwd": (i32, [vp, vp, vp, f32, f32, vp, vp, i32, i32, vp]),
    "calm_trainer_step": (i32, [C.POINTER(TrainerStepArgs), vp]),
    "calm_mix_batch": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, f32, i32, i32, i32, i32, f32, f32, vp]),
}

_lib = None
_lock = threading.Lock()
launch_count = 0  # kernels-launching C-ABI calls made through this binding (bench.py reports it)


class CalmError(RuntimeError):
    pass


def load(build_if_missing=True):
    """Load libcalm_b200.so (building it in-tree with nvcc if it is absent and nvcc exists). Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH) and build_if_missing:
            import importlib.util
            spec = importlib.util.spec_from_file_location("calm_build_ext", os.path.join(HERE, "build_ext.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        if not os.path.exists(LIB_PATH):
            raise CalmError("libcalm_b200.so not found at %s — run `python __graft_entry__.py` (build) first" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError = the library does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        if lib.calm_abi_version() != 1:
            raise CalmError("libcalm_b200.so ABI version %d, expected 1" % lib.calm_abi_version())
        _lib = lib
    return _lib


def last_error():
    return load().calm_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise CalmError("%s failed (rc=%d): %s" % (what, rc, last_error()))


_tls = threading.local()   # device ordinal of the tensors handed to ptr() since the last call() on this thread


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL). The tensor must live on a CUDA device; all tensors of one call must
    live on the same device (call() launches on that device's current stream, whatever the thread's current device is)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise CalmError("calm_b200 kernels need CUDA tensors (got %s); there is no CPU fallback" % t.device)
    idx = t.device.index
    seen = getattr(_tls, "dev", None)
    if seen is None:
        _tls.dev = idx
    elif seen != idx:
        _tls.dev = None
        raise CalmError("calm_b200: tensors of one kernel call live on different devices (cuda:%d and cuda:%d)" % (seen, idx))
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


profile = None   # bench.py's per-kernel timing pass sets this to a list; entries: (name, work, start_event, end_event)


def _launch(fn, name, args, work, tag):
    if profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args, stream())
        e1.record()
        profile.append((name, work, e0, e1, tag))
        return rc
    return fn(*args, stream())


def call(name, *args, work=0.0, tag=None):
    """Invoke an int32-returning entry point on the current torch stream OF THE TENSORS' DEVICE (appended as last argument).
    The reference trainers only do `model.to(cuda:local_rank)` and never call torch.cuda.set_device, and autograd's worker
    threads pick their own device: the launch is therefore wrapped in a device guard whenever the tensors' device (recorded by
    ptr()) is not the thread's current one, so forward and backward always run on the device that owns the memory.
    `work` = algorithmic FLOPs (contractions) or bytes (memory-bound kernels) of this launch, used only by the
    optional CUDA-event profiling pass."""
    global launch_count
    lib = load()
    fn = getattr(lib, name)
    dev = getattr(_tls, "dev", None)
    _tls.dev = None
    if dev is not None and dev != torch.cuda.current_device():
        with torch.cuda.device(dev):
            rc = _launch(fn, name, args, work, tag)
    else:
        rc = _launch(fn, name, args, work, tag)
    launch_count += 1
    if rc != 0:
        raise CalmError("%s failed (rc=%d): %s" % (name, rc, last_error()))
