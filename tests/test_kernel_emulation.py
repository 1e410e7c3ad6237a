"""The device kernels' source, compiled for the host (tests/emul/) and run block by block on CPU threads,
against the oracle. Checks the packed-lane arithmetic, the shared-memory halo exchange and the cost-state
hand-over between the planes of a frame without a GPU; the GPU tests then only have to confirm that the
hardware does what the source says."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import assert_planes_equal
from kat import KAT_IN, KATS
from oracle import oracle as O
from pysangnom import cuda
from pysangnom.clips import make_frame
from pysangnom.fakehost import FORMATS

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL = os.path.join(HERE, "emul", "libkernel_emul.so")


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-f", os.path.join(HERE, "emul", "Makefile")], check=True, stdout=subprocess.DEVNULL)
    L = C.CDLL(EMUL)
    L.emul_frame.restype = C.c_int
    L.emul_frame.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.POINTER(C.c_int),
                             C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int, C.c_int]
    return L


def emulate(L, planes, bits, order=1, aa=48, aac=0, dh=False, luma=True, chroma=True, parity=True):
    """Frame-level host logic in numpy (field placement, plane skipping), plane passes through the emulated kernel."""
    off = cuda.resolve_offset(order, parity)
    sb = planes[0].dtype.itemsize
    outs, proc = [], []
    for p, a in enumerate(planes[:3]):
        enabled = dh or (luma if p == 0 else chroma)
        if dh:
            d = np.full((a.shape[0] * 2, a.shape[1]), 0xEE, dtype=a.dtype)
            d[off::2] = a
        elif enabled:
            d = np.full_like(a, 0xEE)
            d[off::2] = a[off::2]
        else:
            d = a.copy()
        outs.append(d)
        if enabled:
            proc.append(p)
    n = len(proc)
    if n:
        rc = L.emul_frame(sb, n, (C.c_void_p * n)(*[outs[p].ctypes.data for p in proc]),
                          (C.c_longlong * n)(*[outs[p].strides[0] for p in proc]), (C.c_int * n)(*[outs[p].shape[1] for p in proc]),
                          (C.c_int * n)(*[outs[p].shape[0] for p in proc]), (C.c_int * n)(*[off] * n),
                          (C.c_float * n)(*[cuda.threshold(aa if p == 0 else aac, bits, sb) for p in proc]),
                          outs[0].shape[1], outs[0].shape[0])
        assert rc == 0
    return outs


@pytest.mark.parametrize("kw,expected", KATS, ids=["order1_aa48", "order2_aa48", "order1_aa0"])
def test_known_answers(emul, kw, expected):
    assert np.array_equal(emulate(emul, [KAT_IN], 8, **kw)[0], expected)


EMUL_CASES = [
    ("Y8", 8, 4, dict(order=1), "noise"), ("Y8", 40, 2, dict(order=2), "noise"), ("Y8", 100, 30, dict(order=1, aa=10), "edges"),
    ("Y8", 33, 16, dict(order=2, aa=128), "noise"), ("YV12", 96, 64, dict(order=0, aa=48, aac=48), "noise"),
    ("YV12", 100, 48, dict(order=1, aa=48, aac=20), "edges"), ("YV12", 72, 40, dict(luma=False, aa=48, aac=48), "noise"),
    ("YUV422P8", 68, 30, dict(order=2, aa=30, aac=90), "noise"), ("YV411", 64, 32, dict(order=1, aa=48, aac=30), "edges"),
    ("YV24", 44, 20, dict(dh=True, aa=48, aac=48), "noise"), ("YV12", 352, 64, dict(order=0, aa=48, aac=48), "edges"),
    ("Y8", 16, 8, dict(order=1, aa=0), "edges"),
]


@pytest.mark.parametrize("fmtname,w,h,kw,kind", EMUL_CASES, ids=[f"{c[0]}_{c[1]}x{c[2]}_{i}" for i, c in enumerate(EMUL_CASES)])
def test_emulated_kernel_matches_oracle(emul, fmtname, w, h, kw, kind):
    fmt = FORMATS[fmtname]
    for i in range(2):
        fr = make_frame(31, w, h, fmt, kind, i)
        got = emulate(emul, fr, fmt.bits, parity=(i == 0), **kw)
        exp = O.oracle_frame(fr, fmt.bits, parity=(i == 0), **kw)
        assert_planes_equal(got, exp[:3], f"{fmtname} {w}x{h} {kw} frame {i}")
