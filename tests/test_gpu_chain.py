"""Anti-aliasing chain on the device (sangnom_cuda_chain_*): SangNom2(dh=true) -> turn ->
SangNom2(dh=true) -> turn back must equal the same four steps done one by one with the oracle and numpy turns,
i.e. what a script builds from the reference filter and the host's turn filters."""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_planes_equal
from oracle import oracle as O
from pysangnom.clips import make_frame
from pysangnom.formats import FORMATS

pytestmark = pytest.mark.gpu

TURNS = {0: (lambda a: a.T, lambda a: a.T), 1: (lambda a: np.rot90(a, -1), lambda a: np.rot90(a, 1)),
         2: (lambda a: np.rot90(a, 1), lambda a: np.rot90(a, -1))}


@pytest.fixture(scope="module")
def cuda():
    from pysangnom import cuda as c
    c.load()
    return c


def oracle_chain(planes, bits, aa, aac, turn, off1, off2):
    fwd, back = TURNS[turn]
    order1, order2 = (1 if off1 == 0 else 2), (1 if off2 == 0 else 2)
    a = O.oracle_frame(planes[:3], bits, order=order1, aa=aa, aac=aac, dh=True)
    t = [np.ascontiguousarray(fwd(p)) for p in a[:3]]
    b = O.oracle_frame(t, bits, order=order2, aa=aa, aac=aac, dh=True)
    return [np.ascontiguousarray(back(p)) for p in b[:3]]


CHAIN_CASES = [("YUV444P16", 96, 64, 0, 0, 0), ("YUV420P8", 64, 48, 1, 0, 1), ("YUV420PS", 40, 24, 2, 1, 0), ("Y8", 100, 50, 1, 0, 0),
               ("YUV422P10", 68, 30, 0, 1, 1), ("YUV444P8", 256, 128, 2, 0, 0), ("Y16", 130, 70, 1, 1, 0)]


@pytest.mark.parametrize("fmtname,w,h,turn,off1,off2", CHAIN_CASES, ids=[f"{c[0]}_{c[1]}x{c[2]}_t{c[3]}" for c in CHAIN_CASES])
def test_chain_matches_stepwise_oracle(cuda, fmtname, w, h, turn, off1, off2):
    fmt = FORMATS[fmtname]
    frames = [make_frame(61, w, h, fmt, "edges" if i else "noise", i) for i in range(3)]
    with cuda.Chain(fmt.sample_bytes, w, h, turn=turn, max_frames_in_flight=3) as ch:        # one frame per chunk: the slots rotate
        got = ch.process_frames(frames, fmt.bits, aa=48, aac=30, offset1=off1, offset2=off2)
        st = ch.stats()
    for i, fr in enumerate(frames):
        exp = oracle_chain(fr, fmt.bits, 48, 30, turn, off1, off2)
        assert got[i][0].shape == (2 * h, 2 * w)
        assert_planes_equal(got[i], exp, f"chain {fmtname} turn {turn} frame {i}")
    assert st["frames"] == 3 and st["kernel_launches"] == 6 and st["pass_kernel_launches"] >= 6
    nbytes = sum(p.nbytes for p in frames[0][:3]) * 3
    assert st["h2d_bytes"] == nbytes and st["d2h_bytes"] == 4 * nbytes          # one crossing each way


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32], ids=["u8", "u16", "f32"])
@pytest.mark.parametrize("kind", [0, 1, 2], ids=["transpose", "right", "left"])
def test_turn_planes_device(cuda, dtype, kind):
    import torch
    lib = cuda.load()
    rng = np.random.default_rng(kind)
    shapes = [(1080, 1920), (270, 480), (77, 130)]
    srcs = [rng.integers(0, 255, size=s).astype(dtype) for s in shapes]
    d_src = [torch.from_numpy(a.view(np.uint8).reshape(a.shape[0], -1)).cuda() for a in srcs]
    d_dst = [torch.zeros((a.shape[1], a.shape[0] * a.itemsize), dtype=torch.uint8, device="cuda") for a in srcs]
    planes = (cuda.SnTurnPlane * len(srcs))(*[cuda.SnTurnPlane(s.data_ptr(), s.shape[1], d.data_ptr(), d.shape[1], a.shape[1], a.shape[0])
                                              for a, s, d in zip(srcs, d_src, d_dst)])
    torch.cuda.synchronize()
    assert lib.sangnom_cuda_turn_planes_device(srcs[0].itemsize, kind, planes, len(srcs), C.c_void_p(0)) == 0
    torch.cuda.synchronize()
    for a, d in zip(srcs, d_dst):
        got = d.cpu().numpy().view(dtype)
        assert np.array_equal(got, [a.T, np.rot90(a, -1), np.rot90(a, 1)][kind])


@pytest.mark.parametrize("pad_writable", [False, True], ids=["exact", "padding_writable"])
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32], ids=["u8", "u16", "f32"])
@pytest.mark.parametrize("kind", [0, 1, 2], ids=["transpose", "right", "left"])
def test_turn_planes_ragged_sizes_through_tma(cuda, dtype, kind, pad_writable):
    """Planes whose base and pitch are multiples of 16 bytes take the TMA kernel (sangnom_turn_tma.cuh) whatever their
    size: tiles that stick out of the plane are zero-filled by the tensor-map load and clipped by the store, for the
    flipped turns at NEGATIVE tile coordinates. Sizes here are deliberately not multiples of anything; the padding of
    every row and the rows behind the plane must stay untouched."""
    import torch
    lib = cuda.load()
    rng = np.random.default_rng(10 + kind)
    sb = np.dtype(dtype).itemsize
    # (rows, columns). The TMA path takes planes whose destination rows (= source rows, in bytes) are a multiple of 16
    # and, for TurnLeft, whose source rows are too; the others exercise the plain kernel next to it in the same call.
    shapes = [(80, 130), (208, 333), (1080, 1920), (64, 31), (16, 500), (304, 1), (144, 129), (77, 130), (1, 500)]
    srcs, d_src, d_dst, planes = [], [], [], []
    for (h, w) in shapes:
        a = rng.integers(0, 255, size=(h, w)).astype(dtype)
        sp = (w * sb + 15) // 16 * 16 + 16
        dp = (h * sb + 15) // 16 * 16 + 32
        s = torch.full((h, sp), 0x11, dtype=torch.uint8, device="cuda")
        s[:, :w * sb] = torch.from_numpy(a.view(np.uint8).reshape(h, -1)).cuda()
        d = torch.full((w + 2, dp), 0xEE, dtype=torch.uint8, device="cuda")       # two guard rows behind the plane
        srcs.append(a); d_src.append(s); d_dst.append(d)
        planes.append(cuda.SnTurnPlane(s.data_ptr(), sp, d.data_ptr(), dp, w, h, cuda.TURN_DST_PADDING_WRITABLE if pad_writable else 0))
    arr = (cuda.SnTurnPlane * len(planes))(*planes)
    torch.cuda.synchronize()
    assert lib.sangnom_cuda_turn_planes_device(sb, kind, arr, len(planes), C.c_void_p(0)) == 0
    torch.cuda.synchronize()
    for a, d in zip(srcs, d_dst):
        h, w = a.shape
        full = d.cpu().numpy()
        got = full[:w, :h * sb].copy().view(dtype)
        assert np.array_equal(got, [a.T, np.rot90(a, -1), np.rot90(a, 1)][kind]), f"{a.shape}"
        # SN_TURN_DST_PADDING_WRITABLE: the bytes up to the next multiple of 16 of every row are the library's to overwrite
        keep_from = (h * sb + 15) // 16 * 16 if pad_writable else h * sb
        assert (full[:w, keep_from:] == 0xEE).all() and (full[w:] == 0xEE).all(), f"{a.shape}: wrote outside the plane"
