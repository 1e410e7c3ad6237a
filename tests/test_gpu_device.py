"""Device-resident entry (sangnom_cuda_process_planes_device): planes live in torch tensors, the kernel
works in place on the kept field exactly like the reference's `process(dstp, ...)` seam. Also the
size-independent properties at BASELINE.json's full sizes, where the CPU oracle would take too long."""
import numpy as np
import pytest
import torch

from helpers import assert_planes_equal, parity_of
from oracle import oracle as O
from pysangnom.clips import make_frame
from pysangnom.fakehost import FORMATS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    from pysangnom import cuda as c
    c.load()
    return c


def to_dev(a, pitch_align=256):
    h, w = a.shape
    pitch = (w * a.itemsize + pitch_align - 1) // pitch_align * pitch_align
    t = torch.zeros((h, pitch), dtype=torch.uint8, device="cuda")
    t[:, :w * a.itemsize] = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(h, -1)).cuda()
    return t, pitch


def from_dev(t, w, dtype):
    return t[:, :w * np.dtype(dtype).itemsize].cpu().numpy().view(dtype).copy()


def device_frames(cuda, fmt, w, h, frames, order=1, aa=48, aac=0, mode="inplace", stream=None):
    sb = fmt.sample_bytes
    keep, jobs = [], []
    with cuda.Context(sb, w, h) as ctx:
        for k, planes in enumerate(frames):
            off = cuda.resolve_offset(order, parity_of(k))
            for p, a in enumerate(planes[:3]):
                thr = cuda.threshold(aa if p == 0 else aac, fmt.bits, sb)
                if mode == "inplace":
                    t, pitch = to_dev(a)      # the discarded field is still there: it must be overwritten
                    jobs.append(cuda.make_job(0, 0, t.data_ptr(), pitch, a.shape[1], a.shape[0], off, cuda.MODE_INPLACE, thr, p, k))
                    keep.append((t, a.shape[1], a.dtype))
                else:
                    s, sp = to_dev(a)
                    d = torch.full_like(s, 0xEE)
                    jobs.append(cuda.make_job(s.data_ptr(), sp, d.data_ptr(), sp, a.shape[1], a.shape[0], off, cuda.MODE_FIELD, thr, p, k))
                    keep.append((d, a.shape[1], a.dtype, s))
        ctx.process_jobs_device(jobs, stream)
        torch.cuda.synchronize()
        st = ctx.stats()
    outs, i = [], 0
    for planes in frames:
        outs.append([from_dev(keep[i + p][0], keep[i + p][1], keep[i + p][2]) for p in range(len(planes[:3]))])
        i += len(planes[:3])
    return outs, st


@pytest.mark.parametrize("fmtname,w,h,mode", [("YUV420P8", 352, 288, "inplace"), ("YUV420P8", 352, 288, "field"),
                                              ("YUV444P16", 200, 120, "inplace"), ("YUV420PS", 320, 240, "field"),
                                              ("Y8", 1000, 300, "inplace")])
def test_device_path_matches_oracle(cuda, fmtname, w, h, mode):
    fmt = FORMATS[fmtname]
    frames = [make_frame(13, w, h, fmt, "edges" if mode == "field" else "noise", i) for i in range(3)]
    got, st = device_frames(cuda, fmt, w, h, frames, order=0, aa=48, aac=33, mode=mode,
                            stream=torch.cuda.current_stream().cuda_stream)
    for i, fr in enumerate(frames):
        exp = O.oracle_frame(fr, fmt.bits, order=0, aa=48, aac=33, parity=parity_of(i))
        assert_planes_equal(got[i], exp[:3], f"{fmtname} {mode} frame {i}")
    assert st["kernel_launches"] == min(3, fmt.components) and st["h2d_bytes"] == 0 and st["d2h_bytes"] == 0


FULL = [("YUV420P8", 1920, 1080, dict(order=0, aa=48, aac=48)), ("YUV420PS", 3840, 2160, dict(order=2, aa=48, aac=24)),
        ("YUV420P10", 3840, 2160, dict(order=1, aa=48, aac=48)), ("YUV444P16", 1920, 2160, dict(order=1, aa=48)),
        ("Y8", 7680, 4320, dict(order=1, aa=48))]


@pytest.mark.parametrize("fmtname,w,h,kw", FULL, ids=[f[0] + f"_{f[1]}x{f[2]}" for f in FULL])
def test_full_size_properties(cuda, fmtname, w, h, kw):
    """BASELINE.json sizes. Properties that need no oracle run:
    (1) kept rows untouched, border row = its neighbour;
    (2) batch-independence: a frame gives the same bytes alone and inside a batch;
    (3) a clip that is constant along x interpolates to the plain vertical mean (all nine costs tie, so
        buffer 4 wins, reference SangNom2.cpp:214-217)."""
    fmt = FORMATS[fmtname]
    np_ = min(fmt.components, 3)
    frames = [make_frame(17, w, h, fmt, "noise", i) for i in range(2)]
    got, _ = device_frames(cuda, fmt, w, h, frames, **kw)
    alone, _ = device_frames(cuda, fmt, w, h, frames[1:], **kw) if kw["order"] != 0 else (None, None)
    for i, fr in enumerate(frames):
        off = cuda.resolve_offset(kw["order"], parity_of(i))
        for p in range(np_):
            src, out = fr[p], got[i][p]
            assert np.array_equal(out[off::2], src[off::2])
            if off == 0:
                assert np.array_equal(out[-1], out[-2])
            else:
                assert np.array_equal(out[0], out[1])
    if alone is not None:
        for p in range(np_):
            assert np.array_equal(alone[0][p].view(np.uint8), got[1][p].view(np.uint8))
    # (3)
    col = make_frame(19, 8, h, fmt, "noise", 0)
    flat = [np.repeat(pl[:, :1], fmt.plane_shape(w, h, p)[1], axis=1) for p, pl in enumerate(col)]
    out, _ = device_frames(cuda, fmt, w, h, [flat], **kw)
    off = cuda.resolve_offset(kw["order"], True)
    for p in range(np_):
        kept = out[0][p][off::2]
        mid = out[0][p][off + 1::2][:len(kept) - 1]
        if fmt.bits == 32:
            # fp32: the 3-tap (4c+5c-c)*0.125 is not exactly c, so a 3-tap direction may win by an ulp;
            # its mean then differs from the vertical mean by rounding only
            exp = (kept[:-1] + kept[1:]) * np.float32(0.5)
            assert np.max(np.abs(mid - exp)) <= 1e-6
        else:
            exp = ((kept[:-1].astype(np.int64) + kept[1:] + 1) >> 1).astype(mid.dtype)
            assert np.array_equal(mid, exp)


@pytest.mark.parametrize("fmtname,w,h,kw", [("Y8", 7680, 4320, dict(order=2, aa=48)), ("YUV420P8", 1920, 1080, dict(order=0, aa=48, aac=48)),
                                            ("Y16", 8192, 256, dict(order=1, aa=48)), ("Y32", 8192, 128, dict(order=1, aa=48))],
                         ids=["cfg5a_8k_y8", "cfg2_1080p", "widest_u16", "widest_f32"])
def test_full_size_against_oracle(cuda, fmtname, w, h, kw):
    """One frame at BASELINE.json's largest sizes and at the widest planes the kernels accept (8-block clusters),
    bit-exact against the oracle."""
    fmt = FORMATS[fmtname]
    fr = make_frame(23, w, h, fmt, "noise", 0)
    got, _ = device_frames(cuda, fmt, w, h, [fr], mode="field", **kw)
    exp = O.oracle_frame(fr, fmt.bits, parity=True, **kw)
    assert_planes_equal(got[0], exp[:len(got[0])], f"{fmtname} {w}x{h}")
