"""Shared helpers for the parity tests: run the CPU oracle (test infrastructure) on the same
flattened scene the GPU path receives, and compare images / hit identities / ray counts."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_render")


def build_oracle():
    """Compile oracle/rt_oracle.cpp if the .so is missing or stale (gcc is on every box)."""
    src = os.path.join(ROOT, "oracle", "rt_oracle.cpp")
    hdr = os.path.join(ROOT, "include", "rt_b200.h")
    if (not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-o", ORACLE_SO, src, "-lpthread"])
    return ORACLE_SO


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        import raytrace_b200 as R
        lib = C.CDLL(build_oracle())
        lib.rto_render.restype = C.c_int
        lib.rto_render.argtypes = [C.POINTER(R.SceneDesc), C.POINTER(R.RenderParams), C.c_void_p, C.c_void_p,
                                   C.POINTER(R.Counters), C.c_int]
        _oracle = lib
    return _oracle


def oracle_render(scene, max_level, type=0x80, want_ids=True, threads=None, rank=0, world=1, tile_rows=64, flags=0):
    """-> (image HxWx3 u8, ids HxW structured or None, Counters)"""
    import raytrace_b200 as R
    lib = oracle_lib()
    d = scene.flatten()
    w, h = scene.width, scene.height
    out = np.empty((h, w, 3), np.uint8)
    ids = np.zeros(w * h, R.HIT_DTYPE) if want_ids else None
    cnt = R.Counters()
    p = R.RenderParams(type, max_level, rank, world, flags, tile_rows)
    rc = lib.rto_render(d, C.byref(p), out.ctypes.data, ids.ctypes.data if want_ids else None, C.byref(cnt),
                        threads or min(32, os.cpu_count() or 1))
    assert rc == 0, rc
    return out, (ids.reshape(h, w) if want_ids else None), cnt


def compare_images(a, b):
    """-> dict(n_diff, n_gt1, max_abs, psnr, frac_within1)"""
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    px = d.max(axis=2)
    mse = float((d.astype(np.float64) ** 2).mean())
    psnr = float("inf") if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    return {"n_diff": int((px > 0).sum()), "n_gt1": int((px > 1).sum()), "max_abs": int(px.max()),
            "psnr": psnr, "frac_within1": float((px <= 1).mean())}


def compare_ids(a, b):
    ne = np.zeros(a.shape, bool)
    for f in ("object", "sub", "index", "octant"):
        ne |= a[f] != b[f]
    return int(ne.sum()), int((a["distance"] != b["distance"]).sum())
