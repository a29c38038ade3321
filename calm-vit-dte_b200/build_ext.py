"""Builds libcalm_b200.so (the C-ABI kernel library, include/calm_b200.h) in-tree with nvcc for sm_100a.

No torch headers are involved: the library is plain CUDA C++ behind an extern "C" boundary, loaded with ctypes by
calm_ops.py. nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with the tree.
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "libcalm_b200.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["core.cu", "gemm_sm100.cu", "attention.cu", "attention_sm100.cu", "attention_long_sm100.cu", "attention_small.cu", "spectral.cu", "layernorm.cu", "rope.cu", "latent.cu", "cnn.cu", "misc.cu", "trainer.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
NVCC_FLAGS += os.environ.get("CALM_NVCC_FLAGS", "").split()   # bring-up builds, e.g. -DCALM_MBAR_TIMEOUT_CYCLES=2000000000 -DCALM_BRINGUP


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libcalm_b200.so")


def _digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "calm_b200.h")]
    for p in files:
        with open(p, "rb") as f:
            h.update(p.encode()); h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library. Returns the .so path."""
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp.txt")
    digest = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == digest:
        return OUT
    nvcc = _nvcc()
    flags = list(NVCC_FLAGS)

    def one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose and r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(digest)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
