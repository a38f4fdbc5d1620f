"""TEST INFRASTRUCTURE: builders + ctypes bindings of the two CPU checkers of the bilateral-filter row.

  * `oracle/libbilateral_oracle.so`   — oracle/bilateral_oracle.c, the plain-C restatement (always buildable: gcc only);
  * `oracle/_ref/libbilateralfilter_ref.so` — the REFERENCE ITSELF: its two source files
    SCD-AAAI2023/wrapper/bilateralfilter/{bilateralfilter,permutohedral}.cpp compiled where they lie under /root/reference
    (g++ -O2 -fopenmp, the flags of the reference's own setup.py:20-27 minus the SWIG wrapper), output only into oracle/_ref/
    (git-ignored, travels to the GPU box).  No reference source is copied into the repo.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libbilateral_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_SO = os.path.join(REF_DIR, "libbilateralfilter_ref.so")
REF_SRC = "/root/reference/SCD-AAAI2023/wrapper/bilateralfilter"
# Itanium mangling of `void bilateralfilter_batch(float*, int, float*, int, float*, int, int, int, int, int, float, float)`
# (bilateralfilter.hpp:12): the reference exports no extern "C" symbol, SWIG binds the C++ one
REF_SYMBOL = "_Z21bilateralfilter_batchPfiS_iS_iiiiiff"
_FP = ctypes.POINTER(ctypes.c_float)


def _newer(target, sources):
    return os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in sources)


def build(verbose=False):
    """compile the restatement; compile the reference when its sources are present (authoring container)"""
    src = os.path.join(HERE, "bilateral_oracle.c")
    if not _newer(ORACLE_SO, [src]):
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", src, "-o", ORACLE_SO, "-lm"], check=True)
    srcs = [os.path.join(REF_SRC, f) for f in ("bilateralfilter.cpp", "permutohedral.cpp")]
    if all(os.path.exists(s) for s in srcs) and not _newer(REF_SO, srcs):
        os.makedirs(REF_DIR, exist_ok=True)
        subprocess.run(["g++", "-O2", "-fopenmp", "-shared", "-fPIC", "-w", "-I", REF_SRC] + srcs + ["-o", REF_SO], check=True)
    if verbose:
        print("bilateral oracle:", ORACLE_SO, "| reference:", REF_SO if os.path.exists(REF_SO) else "absent")


def have_reference():
    return os.path.exists(REF_SO)


def _flat(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_FP)


def reference_filter(images, ins, sigma_rgb, sigma_xy):
    """the compiled reference: images (N,3,H,W), ins (N,K,H,W) fp32 -> outs (N,K,H,W)"""
    lib = ctypes.CDLL(REF_SO)
    fn = getattr(lib, REF_SYMBOL)
    fn.restype = None
    fn.argtypes = [_FP, ctypes.c_int, _FP, ctypes.c_int, _FP, ctypes.c_int] + [ctypes.c_int] * 4 + [ctypes.c_float] * 2
    N, K, H, W = ins.shape
    images, pi = _flat(images)
    ins, pn = _flat(ins)
    outs = np.zeros_like(ins)
    fn(pi, images.size, pn, ins.size, outs.ctypes.data_as(_FP), outs.size, N, K, H, W, sigma_rgb, sigma_xy)
    return outs


def oracle_filter(images, ins, sigma_rgb, sigma_xy, want_lattice=False):
    """the C restatement; optionally also the number of lattice points per image"""
    if not os.path.exists(ORACLE_SO):
        build()
    lib = ctypes.CDLL(ORACLE_SO)
    fn = lib.bilateral_oracle_batch
    fn.restype = ctypes.c_int
    fn.argtypes = [_FP, _FP, _FP] + [ctypes.c_int] * 4 + [ctypes.c_float] * 2 + [ctypes.POINTER(ctypes.c_int)]
    N, K, H, W = ins.shape
    images, pi = _flat(images)
    ins, pn = _flat(ins)
    outs = np.zeros_like(ins)
    m = (ctypes.c_int * N)()
    rc = fn(pi, pn, outs.ctypes.data_as(_FP), N, K, H, W, sigma_rgb, sigma_xy, m)
    if rc:
        raise MemoryError("bilateral_oracle_batch")
    return (outs, np.array(list(m))) if want_lattice else outs


def synth(N, K, H, W, seed=0, kind="natural"):
    """seeded inputs of the shape DenseEnergyLoss feeds the filter (utils/losses.py:41-45,66-70: de-normalised RGB in 0..255,
    softmax probabilities times the ROI mask).  kind: natural (smooth colour field + noise), noise (iid colours: a huge lattice),
    flat (one colour: a tiny lattice, very long splat lists)"""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    if kind == "noise":
        img = rng.uniform(0, 255, (N, 3, H, W))
    elif kind == "flat":
        img = np.full((N, 3, H, W), 117.0)
    else:
        ph = rng.uniform(0, 6.28, (N, 3, 1, 1))
        fr = rng.uniform(0.01, 0.08, (N, 3, 1, 1))
        img = 127.5 + 100.0 * np.sin(fr * xx + ph) * np.cos(fr * 0.7 * yy + ph) + rng.normal(0, 6.0, (N, 3, H, W))
        img = np.clip(img, 0, 255)
    logits = rng.normal(0, 2.0, (N, K, H, W)).astype(np.float32)
    e = np.exp(logits - logits.max(1, keepdims=True))
    seg = e / e.sum(1, keepdims=True)
    roi = (rng.uniform(0, 1, (N, 1, H, W)) > 0.1).astype(np.float32)
    return img.astype(np.float32), (seg * roi).astype(np.float32)


if __name__ == "__main__":
    build(verbose=True)
