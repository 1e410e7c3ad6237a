"""Device-resident entry (sangnom_cuda_process_planes_device): planes live in torch tensors, the kernel
works in place on the kept field exactly like the reference's `process(dstp, ...)` seam. Also the
size-independent properties at BASELINE.json's full sizes, where the CPU oracle would take too long."""
import numpy as np
import pytest
import torch

from helpers import assert_planes_equal, parity_of
from oracle import oracle as O
from pysangnom.clips import make_frame
from pysangnom.formats import FORMATS

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


# Every BASELINE.json configuration at its full size (plus the widest planes the kernels accept), one or two frames,
# bit-exact against the oracle through BOTH entries. The rows after cfg2 are the launches whose planes are split over
# a thread-block cluster AND are narrower than the pool (subsampled chroma: the hand-over regions cross segment
# boundaries), the pad-column case of the AA chain's second stage (S = 2176), and the widest planes.
FULL_ORACLE = [
    ("cfg2_1080p_420p8", "YUV420P8", 1920, 1080, dict(order=0, aa=48, aac=48), 2),
    ("cfg3_stage1_444p16_dh", "YUV444P16", 1920, 1080, dict(dh=True, aa=48), 1),
    ("cfg3_stage2_444p16_dh_T", "YUV444P16", 2160, 1920, dict(dh=True, aa=48), 1),
    ("cfg4_2160p_420ps", "YUV420PS", 3840, 2160, dict(order=2, aa=48, aac=24), 1),
    ("cfg5a_8k_y8", "Y8", 7680, 4320, dict(order=2, aa=48), 1),
    ("cfg5b_2160p_420p10", "YUV420P10", 3840, 2160, dict(order=1, aa=48, aac=48), 1),
    ("uhd_420p8_cluster_chroma", "YUV420P8", 3840, 2160, dict(order=0, aa=48, aac=48), 2),
    ("uhd_422p8_cluster_chroma", "YUV422P8", 3840, 1080, dict(order=2, aa=48, aac=30), 1),
    ("uhd_411_cluster_chroma", "YV411", 4096, 540, dict(order=1, aa=48, aac=48), 1),
    ("widest_u16", "Y16", 8192, 256, dict(order=1, aa=48), 1),
    ("widest_f32", "Y32", 8192, 128, dict(order=1, aa=48), 1),
    ("u16_420_odd_segments", "YUV420P16", 2080, 540, dict(order=0, aa=48, aac=48), 2),
]


def device_frames_any(cuda, fmt, w, h, frames, kw):
    """Device entry for any script arguments: dh (separated rows in, double-height plane out), disabled planes."""
    sb = fmt.sample_bytes
    dh = kw.get("dh", False)
    order, aa, aac = kw.get("order", 1), kw.get("aa", 48), kw.get("aac", 0)
    keep, jobs = [], []
    with cuda.Context(sb, w, h * 2 if dh else h) as ctx:
        for k, planes in enumerate(frames):
            off = cuda.resolve_offset(order, parity_of(k))
            for p, a in enumerate(planes[:3]):
                enabled = dh or (kw.get("luma", True) if p == 0 else kw.get("chroma", True))
                thr = cuda.threshold(aa if p == 0 else aac, fmt.bits, sb)
                s, sp = to_dev(a)
                d = torch.full((a.shape[0] * (2 if dh else 1), sp), 0xEE, dtype=torch.uint8, device="cuda")
                mode = cuda.MODE_DH if dh else (cuda.MODE_FIELD if enabled else cuda.MODE_COPY)
                jobs.append(cuda.make_job(s.data_ptr(), sp, d.data_ptr(), sp, a.shape[1], d.shape[0], off, mode, thr, p, k))
                keep.append((d, a.shape[1], a.dtype, s))
        ctx.process_jobs_device(jobs)
        ctx.synchronize()
    outs, i = [], 0
    for planes in frames:
        np_ = len(planes[:3])
        outs.append([from_dev(keep[i + p][0], keep[i + p][1], keep[i + p][2]) for p in range(np_)])
        i += np_
    return outs


@pytest.mark.parametrize("entry", ["device", "host"])
@pytest.mark.parametrize("case", FULL_ORACLE, ids=[c[0] for c in FULL_ORACLE])
def test_full_size_against_oracle(cuda, case, entry):
    name, fmtname, w, h, kw, nframes = case
    fmt = FORMATS[fmtname]
    frames = [make_frame(23, w, h, fmt, "noise" if i == 0 else "edges", i) for i in range(nframes)]
    args = dict(order=kw.get("order", 1), aa=kw.get("aa", 48), aac=kw.get("aac", 0), dh=kw.get("dh", False))
    if entry == "device":
        got = device_frames_any(cuda, fmt, w, h, frames, kw)
    else:
        with cuda.Context(fmt.sample_bytes, w, h * 2 if args["dh"] else h) as ctx:
            got = ctx.process_frames(frames, fmt.bits, parities=[parity_of(i) for i in range(nframes)], **args)
    for i, fr in enumerate(frames):
        exp = O.oracle_frame(fr, fmt.bits, parity=parity_of(i), **args)
        assert_planes_equal(got[i][:3], exp[:3], f"{name} {entry} frame {i}")
