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
from pysangnom.formats import FORMATS

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL = os.path.join(HERE, "emul", "libkernel_emul.so")


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-f", os.path.join(HERE, "emul", "Makefile")], check=True, stdout=subprocess.DEVNULL)
    L = C.CDLL(EMUL)
    L.emul_frame.restype = C.c_int
    L.emul_frame.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.POINTER(C.c_void_p),
                             C.POINTER(C.c_longlong), C.POINTER(C.c_int),
                             C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.emul_carry_bytes.restype = C.c_size_t
    L.emul_carry_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    return L


def emulate(L, planes, bits, order=1, aa=48, aac=0, dh=False, luma=True, chroma=True, parity=True, cluster=1, out_of_place=False, carry=None, saturate=False):
    """Frame-level host logic in numpy (field placement, plane skipping), plane passes through the emulated kernel.
    out_of_place: the kept rows are handed to the kernel as a separate packed field buffer (what the host path
    uploads) and the kernel writes them into the dst plane itself.
    carry: (in, out) numpy byte buffers of emul_carry_bytes() - persistent-pool mode."""
    off = cuda.resolve_offset(order, parity)
    sb = planes[0].dtype.itemsize
    outs, proc, fields = [], [], {}
    for p, a in enumerate(planes[:3]):
        enabled = dh or (luma if p == 0 else chroma)
        if dh:
            d = np.full((a.shape[0] * 2, a.shape[1]), 0xEE, dtype=a.dtype)
            kept = a
        elif enabled:
            d = np.full_like(a, 0xEE)
            kept = a[off::2]
        else:
            d = a.copy()
        if enabled:
            if out_of_place:
                fields[p] = np.ascontiguousarray(kept)
            else:
                d[off::2] = kept
            proc.append(p)
        outs.append(d)
    n = len(proc)
    if n:
        srcs = (C.c_void_p * n)(*[fields[p].ctypes.data if out_of_place else None for p in proc])
        spitch = (C.c_longlong * n)(*[fields[p].strides[0] if out_of_place else 0 for p in proc])
        rc = L.emul_frame(sb, n, (C.c_void_p * n)(*[outs[p].ctypes.data for p in proc]),
                          (C.c_longlong * n)(*[outs[p].strides[0] for p in proc]), srcs, spitch, (C.c_int * n)(*[outs[p].shape[1] for p in proc]),
                          (C.c_int * n)(*[outs[p].shape[0] for p in proc]), (C.c_int * n)(*[off] * n),
                          (C.c_float * n)(*[cuda.threshold(aa if p == 0 else aac, bits, sb) for p in proc]),
                          outs[0].shape[1], outs[0].shape[0], cluster,
                          carry[0].ctypes.data if carry else None, carry[1].ctypes.data if carry else None, int(saturate))
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
    # 16-bit and fp32 flavours
    ("Y16", 40, 12, dict(order=1), "noise"), ("Y10", 33, 10, dict(order=2, aa=100), "edges"), ("YUV420P16", 96, 48, dict(order=0, aa=48, aac=48), "noise"),
    ("YUV420P10", 100, 40, dict(order=1, aa=48, aac=20), "edges"), ("YUV422P10", 68, 30, dict(order=2, aa=30, aac=90), "noise"),
    ("YUV444P16", 44, 20, dict(dh=True, aa=48, aac=48), "noise"), ("YUV420P16", 72, 40, dict(luma=False, aa=48, aac=48), "noise"),
    ("Y32", 40, 12, dict(order=1), "noise"), ("YUV420PS", 96, 48, dict(order=2, aa=48, aac=24), "noise"),
    ("YUV420PS", 100, 40, dict(order=0, aa=48, aac=24), "edges"), ("YUV444PS", 44, 20, dict(dh=True, aa=128, aac=128), "edges"),
    ("YUV422PS", 68, 30, dict(order=1, aa=10, aac=0), "noise"),
]


@pytest.mark.parametrize("fmtname,w,h,kw,kind", EMUL_CASES, ids=[f"{c[0]}_{c[1]}x{c[2]}_{i}" for i, c in enumerate(EMUL_CASES)])
def test_emulated_kernel_matches_oracle(emul, fmtname, w, h, kw, kind):
    fmt = FORMATS[fmtname]
    for i in range(2):
        fr = make_frame(31, w, h, fmt, kind, i)
        got = emulate(emul, fr, fmt.bits, parity=(i == 0), out_of_place=(i == 1), **kw)
        exp = O.oracle_frame(fr, fmt.bits, parity=(i == 0), **kw)
        assert_planes_equal(got, exp[:3], f"{fmtname} {w}x{h} {kw} frame {i}")


CLUSTER_CASES = [("Y8", 128, 20, dict(order=1), 2), ("YV12", 256, 36, dict(order=0, aa=48, aac=48), 4), ("YV12", 512, 24, dict(order=2, aa=48, aac=48), 8),
                 ("Y8", 120, 20, dict(order=1), 2), ("YV12", 250, 36, dict(order=0, aa=48, aac=48), 4), ("Y8", 64, 12, dict(order=2, aa=20), 8),
                 ("Y16", 120, 20, dict(order=1), 2), ("YUV420P10", 250, 36, dict(order=0, aa=48, aac=48), 4), ("YUV420PS", 120, 24, dict(order=2, aa=48, aac=24), 2),
                 ("Y32", 64, 12, dict(order=1), 8), ("YUV422P16", 68, 30, dict(order=2, aa=30, aac=90), 2)]


@pytest.mark.parametrize("fmtname,w,h,kw,cluster", CLUSTER_CASES, ids=[f"{c[0]}_{c[1]}x{c[2]}_G{c[4]}" for c in CLUSTER_CASES])
def test_cluster_split_matches_oracle(emul, fmtname, w, h, kw, cluster):
    """A plane split into column segments over the blocks of a cluster (DSMEM halo + cluster barrier) gives the
    same bytes as the unsplit plane - i.e. as the oracle."""
    fmt = FORMATS[fmtname]
    fr = make_frame(41, w, h, fmt, "noise", 0)
    got = emulate(emul, fr, fmt.bits, cluster=cluster, **kw)
    exp = O.oracle_frame(fr, fmt.bits, **kw)
    assert_planes_equal(got, exp[:3], f"{fmtname} {w}x{h} {kw} G={cluster}")


# The picture's right edge one to three columns into the second block of a cluster (the first block's last thread
# then needs the replicated edge in ITS staged rows, and for 4-column threads two threads share it), and next to the
# end of the pool; every sample width.
SEGMENT_EDGE_CASES = [(f, w, 12, 2) for f in ("Y8", "Y16", "Y32") for w in (33, 34, 35, 61, 62, 63)] + [("YV411", 4 * w, 16, 2) for w in (33, 34, 35)]


@pytest.mark.parametrize("fmtname,w,h,cluster", SEGMENT_EDGE_CASES, ids=[f"{c[0]}_{c[1]}x{c[2]}_G{c[3]}" for c in SEGMENT_EDGE_CASES])
def test_picture_edge_near_a_segment_boundary(emul, fmtname, w, h, cluster):
    fmt = FORMATS[fmtname]
    for kind in ("noise", "edges"):
        fr = make_frame(43, w, h, fmt, kind, 0)
        got = emulate(emul, fr, fmt.bits, cluster=cluster, order=1, aa=48, aac=48)
        exp = O.oracle_frame(fr, fmt.bits, order=1, aa=48, aac=48)
        assert_planes_equal(got, exp[:3], f"{fmtname} {w}x{h} {kind} G={cluster}")


PERSISTENT_CASES = [("Y8", 40, 12, dict(order=1), 1), ("YV12", 100, 48, dict(order=0, aa=48, aac=48), 1), ("YV12", 72, 40, dict(luma=False, aa=48, aac=48), 1),
                    ("YV12", 96, 64, dict(chroma=False), 1), ("YV24", 44, 20, dict(dh=True, aa=48, aac=48), 1), ("YUV420P16", 100, 40, dict(order=2, aa=48, aac=20), 1),
                    ("Y32", 40, 12, dict(order=1), 1), ("YUV422PS", 68, 30, dict(order=1, aa=10, aac=30), 1), ("YV12", 250, 36, dict(order=0, aa=48, aac=48), 4),
                    ("YV411", 72, 32, dict(order=1, aa=48, aac=30), 1)]


@pytest.mark.parametrize("fmtname,w,h,kw,cluster", PERSISTENT_CASES, ids=[f"{c[0]}_{c[1]}x{c[2]}_{i}" for i, c in enumerate(PERSISTENT_CASES)])
def test_persistent_pool_chain_matches_oracle(emul, fmtname, w, h, kw, cluster):
    """Persistent-pool mode: four frames in sequence, each starting from the pool state the previous one left, against
    the oracle run with one pool for the whole clip (= one long-lived reference instance, tests/test_oracle.py)."""
    fmt = FORMATS[fmtname]
    out_h = h * 2 if kw.get("dh") else h
    pool = O.new_pool(w, out_h, fmt.sample_bytes)
    nbytes = emul.emul_carry_bytes(fmt.sample_bytes, w, out_h)
    carry = [np.zeros(max(nbytes, 1), np.uint8), np.zeros(max(nbytes, 1), np.uint8)]
    differs_from_fresh = False
    for i in range(4):
        fr = make_frame(57, w, h, fmt, "noise" if i != 2 else "edges", i)
        got = emulate(emul, fr, fmt.bits, parity=(i % 2 == 0), cluster=cluster, carry=(carry[i & 1], carry[(i & 1) ^ 1]), **kw)
        exp = O.oracle_frame(fr, fmt.bits, parity=(i % 2 == 0), pool=pool, **kw)
        assert_planes_equal(got, exp[:3], f"persistent {fmtname} {w}x{h} {kw} frame {i}")
        fresh = O.oracle_frame(fr, fmt.bits, parity=(i % 2 == 0), **kw)
        differs_from_fresh |= any(not np.array_equal(a, b) for a, b in zip(exp[:3], fresh[:3]))
    if w % 32 != 0 and fmt.sample_bytes < 4:
        assert differs_from_fresh, "case does not exercise the carried state"


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32], ids=["u8", "u16", "f32"])
@pytest.mark.parametrize("w,h", [(128, 128), (256, 128), (130, 70), (64, 200), (7, 5)])
@pytest.mark.parametrize("kind", [0, 1, 2], ids=["transpose", "right", "left"])
def test_turn_kernel(emul, dtype, w, h, kind):
    """sangnom_turn.cuh (AA chain): transpose / TurnRight / TurnLeft of a plane, whole tiles (word path with in-register
    cell transposition) and ragged edges (sample path)."""
    emul.emul_turn.restype = C.c_int
    emul.emul_turn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int]
    rng = np.random.default_rng(w * 1000 + h)
    a = rng.integers(0, 250, size=(h, w)).astype(dtype)
    out = np.zeros((w, h), dtype=dtype)
    assert emul.emul_turn(a.itemsize, kind, a.ctypes.data, a.strides[0], out.ctypes.data, out.strides[0], w, h) == 0
    exp = [a.T, np.rot90(a, -1), np.rot90(a, 1)][kind]
    assert np.array_equal(out, exp)


SAT_CASES = [("Y8", 100, 30, dict(order=1, aa=48), "noise", 1), ("YV12", 96, 64, dict(order=0, aa=48, aac=48), "noise", 1), ("YV12", 256, 36, dict(order=2, aa=128, aac=128), "noise", 4),
             ("YUV422P8", 68, 30, dict(order=2, aa=30, aac=90), "edges", 1), ("Y16", 40, 12, dict(order=1, aa=128), "noise", 1), ("YUV420P16", 96, 48, dict(order=0, aa=48, aac=48), "noise", 2),
             ("YUV420P10", 100, 40, dict(order=1, aa=48, aac=20), "noise", 1), ("YUV444P16", 44, 20, dict(dh=True, aa=48, aac=48), "noise", 1), ("YUV420PS", 96, 48, dict(order=2, aa=48, aac=24), "noise", 1)]


@pytest.mark.parametrize("fmtname,w,h,kw,kind,cluster", SAT_CASES, ids=[f"{c[0]}_{c[1]}x{c[2]}_{i}" for i, c in enumerate(SAT_CASES)])
def test_saturating_flavour_matches_oracle(emul, fmtname, w, h, kw, kind, cluster):
    """SN_FLAG_SATURATE kernels (the reference's SSE2 arithmetic) against the oracle's saturating flavour, which
    tests/test_oracle.py pins to the compiled reference run with opt=1."""
    fmt = FORMATS[fmtname]
    differs = False
    for i in range(2):
        fr = make_frame(131, w, h, fmt, kind, i)
        got = emulate(emul, fr, fmt.bits, parity=(i == 0), cluster=cluster, saturate=True, **kw)
        exp = O.oracle_frame(fr, fmt.bits, parity=(i == 0), saturate=True, **kw)
        assert_planes_equal(got, exp[:3], f"saturating {fmtname} {w}x{h} {kw} frame {i}")
        wrap = O.oracle_frame(fr, fmt.bits, parity=(i == 0), **kw)
        differs |= any(not np.array_equal(a, b) for a, b in zip(exp[:3], wrap[:3]))
    if fmtname in ("Y8", "YV12") and kind == "noise":
        assert differs, "case does not exercise the saturation"


HELPER_CASES = [("YV12", 960, 24, dict(order=0, aa=48, aac=48), "noise"), ("YV12", 960, 16, dict(order=2, aa=48, aac=48), "edges"),
                ("YUV422P8", 960, 12, dict(order=1, aa=30, aac=90), "noise"), ("YV411", 1280, 12, dict(order=1, aa=48, aac=30), "noise"),
                ("YV12", 944, 20, dict(order=1, aa=128, aac=128), "noise"), ("YV12", 960, 8, dict(luma=False, aa=48, aac=48), "noise")]


@pytest.mark.parametrize("fmtname,w,h,kw,kind", HELPER_CASES, ids=[f"{c[0]}_{c[1]}x{c[2]}_{i}" for i, c in enumerate(HELPER_CASES)])
@pytest.mark.parametrize("saturate", [False, True], ids=["wrap", "sat"])
def test_wide_subsampled_planes(emul, fmtname, w, h, kw, kind, saturate):
    """8-bit planes wide enough that the chroma pixel threads end inside a 32-thread group (a "mixed" warp on the
    device: pixel lanes and state-only lanes in one warp, the general variant of a row), run by the unclustered
    instantiation as the launcher would; in place and out of place, wrapping and saturating flavour."""
    fmt = FORMATS[fmtname]
    for i in range(2):
        fr = make_frame(73, w, h, fmt, kind, i)
        got = emulate(emul, fr, fmt.bits, parity=(i == 0), out_of_place=(i == 1), saturate=saturate, **kw)
        exp = O.oracle_frame(fr, fmt.bits, parity=(i == 0), saturate=saturate, **kw)
        assert_planes_equal(got, exp[:3], f"wide {fmtname} {w}x{h} {kw} frame {i}")


# Planes tall and wide enough that warps (and, split over a cluster, whole blocks) right of the chroma rectangle retire
# part-way down the sweep when their columns leave the dependency cone (sangnom_plan.h). The emulation poisons the
# hand-over scratch and the shared memory, so a cone that is too tight shows up as wrong samples.
CONE_CASES = [("YV12", 1600, 400, dict(order=0, aa=48, aac=48), 1), ("YV12", 1024, 400, dict(order=1, aa=48, aac=48), 4),
              ("YV411", 1536, 120, dict(order=2, aa=48, aac=30), 1), ("YUV422P8", 1024, 200, dict(order=1, aa=20, aac=90), 2),
              ("YUV420P16", 800, 400, dict(order=0, aa=48, aac=48), 1), ("YUV420PS", 512, 400, dict(order=2, aa=48, aac=24), 4),
              ("YUV420P10", 1024, 240, dict(order=1, aa=48, aac=48), 8)]


@pytest.mark.parametrize("fmtname,w,h,kw,cluster", CONE_CASES, ids=[f"{c[0]}_{c[1]}x{c[2]}_G{c[4]}" for c in CONE_CASES])
def test_dependency_cone_trimming(emul, fmtname, w, h, kw, cluster):
    fmt = FORMATS[fmtname]
    fr = make_frame(211, w, h, fmt, "noise", 0)
    got = emulate(emul, fr, fmt.bits, parity=False, cluster=cluster, **kw)
    exp = O.oracle_frame(fr, fmt.bits, parity=False, **kw)
    assert_planes_equal(got, exp[:3], f"cone {fmtname} {w}x{h} {kw} G={cluster}")
