"""Test infrastructure: imports the UNMODIFIED reference modules (CALM-ViT/CALM_ViT_V2.py + Vi_Tools_CNN_less_V2.py) under
private names, so that they can live in one process next to the drop-in modules of the same names.

Search order: baseline/_ref/ (a verbatim, git-ignored copy that travels to the GPU box) and /root/reference/CALM-ViT (the build
container only). Returns None when neither exists — callers skip. Nothing in the product imports this file.
"""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference/CALM-ViT")
_cache = {}


def reference_dir():
    for d in CANDIDATES:
        if os.path.exists(os.path.join(d, "CALM_ViT_V2.py")) and os.path.exists(os.path.join(d, "Vi_Tools_CNN_less_V2.py")):
            return d
    return None


def load_reference():
    """-> the reference's CALM_ViT_V2 module object (its `vt` attribute is the reference's Vi_Tools_CNN_less_V2), or None."""
    d = reference_dir()
    if d is None:
        return None
    if d in _cache:
        return _cache[d]
    for n in ("matplotlib", "matplotlib.pyplot"):            # CALM_ViT_V2.py:7 — only save_samples uses it; absent in the image
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    names = ("CALM_ViT_V2", "Vi_Tools_CNN_less_V2")
    saved = {n: sys.modules.pop(n) for n in names if n in sys.modules}
    sys.path.insert(0, d)
    try:
        spec = importlib.util.spec_from_file_location("ref_CALM_ViT_V2", os.path.join(d, "CALM_ViT_V2.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)                         # its `import Vi_Tools_CNN_less_V2 as vt` resolves inside d
    finally:
        sys.path.remove(d)
        for n in names:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
    _cache[d] = mod
    return mod
