"""Python face of the oracle (TEST INFRASTRUCTURE - see sangnom_oracle.c header).

Two checkers:
  * `oracle_frame(...)`      - our plain-C restatement (liboracle.so), always available.
  * `reference_plugin_path()`- the UNMODIFIED reference compiled into oracle/_ref/ (built in the
                               container that has /root/reference; the prebuilt .so travels to the
                               GPU box). Driven through the fake AviSynth host like a real host would.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(HERE, "liboracle.so")
REF_PLUGIN = os.path.join(HERE, "_ref", "libsangnom2_ref.so")

_lib = None


def build(quiet=True):
    """Compile liboracle.so and, when the reference tree is present, oracle/_ref."""
    subprocess.run(["make", "-f", os.path.join(HERE, "Makefile")], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_LIB):
            build()
        L = C.CDLL(ORACLE_LIB)
        L.sn_oracle_pool_stride.restype = C.c_int
        L.sn_oracle_pool_rows.restype = C.c_int
        L.sn_oracle_pool_bytes.restype = C.c_size_t
        L.sn_oracle_pool_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
        L.sn_oracle_threshold.restype = C.c_float
        L.sn_oracle_threshold.argtypes = [C.c_int, C.c_int, C.c_int]
        L.sn_oracle_process_plane.restype = C.c_int
        L.sn_oracle_process_plane.argtypes = [C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_float,
                                              C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.sn_oracle_frame.restype = C.c_int
        L.sn_oracle_frame.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_ssize_t), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_ssize_t), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float),
                                      C.c_void_p]
        L.sn_oracle_set_saturating.restype = None
        L.sn_oracle_set_saturating.argtypes = [C.c_int]
        _lib = L
    return _lib


def reference_plugin_path():
    """Path of the compiled unmodified reference plugin, or None if it was never built."""
    return REF_PLUGIN if os.path.exists(REF_PLUGIN) else None


def threshold(aa, bits, sample_bytes):
    return float(_load().sn_oracle_threshold(int(aa), int(bits), int(sample_bytes)))


def pool_geometry(luma_w, luma_h_out):
    L = _load()
    return L.sn_oracle_pool_stride(luma_w), L.sn_oracle_pool_rows(luma_h_out)


def new_pool(luma_w, luma_h_out, sample_bytes):
    return np.zeros(_load().sn_oracle_pool_bytes(luma_w, luma_h_out, sample_bytes), dtype=np.uint8)


def resolve_offset(order, parity):
    """order 0: by field parity of the frame (top-first -> keep top), 1: keep top, 2: keep bottom."""
    if order == 0:
        return 0 if parity else 1
    return 0 if order == 1 else 1


def oracle_frame(planes, bits, order=1, aa=48, aac=0, dh=False, luma=True, chroma=True, parity=True, pool=None, saturate=False):
    """Run one frame through the C restatement. `planes`: list of 1..4 2-D numpy arrays (Y[,U,V[,A]]).

    Returns the output planes (alpha, if present, is returned as a copy of the source scaled to the
    output height by row duplication for dh - the reference leaves it unwritten).
    pool=None => fresh zero-filled pool (the parity contract); pass new_pool(...) to chain frames.
    saturate=True => the arithmetic of the reference's opt=1 (SSE2) path instead of opt=0.
    """
    L = _load()
    L.sn_oracle_set_saturating(int(bool(saturate)))
    n = min(len(planes), 3)
    dt = planes[0].dtype
    sb = dt.itemsize
    hy, wy = planes[0].shape
    out_h = hy * 2 if dh else hy
    off = resolve_offset(order, parity)
    srcs = [np.ascontiguousarray(p) for p in planes[:n]]
    dsts = [np.zeros((p.shape[0] * (2 if dh else 1), p.shape[1]), dtype=dt) for p in srcs]
    proc = [luma, chroma, chroma][:n]
    thr = [threshold(a, bits, sb) for a in [aa, aac, aac][:n]]
    vp = lambda arrs: (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    ss = lambda arrs: (C.c_ssize_t * n)(*[a.strides[0] for a in arrs])
    rc = L.sn_oracle_frame(vp(srcs), ss(srcs), vp(dsts), ss(dsts),
                           (C.c_int * n)(*[p.shape[1] for p in srcs]), (C.c_int * n)(*[p.shape[0] for p in srcs]),
                           n, sb, wy, out_h, int(dh), off, (C.c_int * n)(*[int(b) for b in proc]),
                           (C.c_float * n)(*thr), pool.ctypes.data if pool is not None else None)
    L.sn_oracle_set_saturating(0)
    if rc != 0:
        raise RuntimeError(f"oracle failed rc={rc}")
    if len(planes) == 4:
        a = planes[3]
        dsts.append(np.repeat(a, 2, axis=0) if dh else a.copy())
    return dsts
