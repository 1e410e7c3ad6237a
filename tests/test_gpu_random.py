"""Randomised GPU parity (hypothesis): small geometries x formats x script arguments through the device entry with
deliberately unaligned pointers and pitches (the kernels' non-bulk / non-vector paths), and with the column-segment
knobs lowered so that planes of a few hundred columns are split over thread-block clusters - including subsampled
chroma planes narrower than the pool, whose hand-over regions then cross segment boundaries. Checker: the oracle
(pinned to the compiled reference by tests/test_oracle.py). Reference semantics under test:
/root/reference/src/SangNom2.cpp:133-136,269-270 (every plane's recursion sweeps the luma-sized pool), :287-288."""
import os

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from helpers import assert_planes_equal, parity_of
from oracle import oracle as O
from pysangnom.clips import make_frame
from pysangnom.formats import FORMATS

pytestmark = pytest.mark.gpu

FMTS = ["Y8", "YV12", "YV16", "YV24", "YV411", "Y10", "YUV420P10", "YUV422P16", "YUV444P16", "Y32", "YUV420PS", "YUV444PS"]
# SANGNOM_RANDOM_EXAMPLES=N: a longer, differently seeded soak (the default run is derandomised so that it repeats)
_N = int(os.environ.get("SANGNOM_RANDOM_EXAMPLES", "0"))


@pytest.fixture(scope="module")
def cuda():
    from pysangnom import cuda as c
    c.load()
    return c


def dev_plane(a, skew, extra, fill=None):
    """Plane `a` in device memory at an address `skew` samples past an aligned base, pitch = row + `extra` samples."""
    h, w = a.shape
    sb = a.itemsize
    pitch = (w + extra) * sb
    buf = torch.full((h * pitch + skew * sb + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    view = buf[skew * sb: skew * sb + h * pitch].view(h, pitch)
    if fill is None:
        view[:, :w * sb] = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(h, -1)).cuda()
    return buf, view, pitch


@settings(max_examples=_N or 70, deadline=None, derandomize=not _N, suppress_health_check=list(HealthCheck))
@given(fmtname=st.sampled_from(FMTS), wq=st.integers(2, 180), h2=st.integers(1, 14), order=st.integers(0, 2),
       aa=st.integers(0, 128), aac=st.integers(0, 128), dh=st.booleans(), luma=st.booleans(), chroma=st.booleans(),
       seed=st.integers(0, 2 ** 16), kind=st.sampled_from(["noise", "edges"]), skew=st.sampled_from([0, 0, 1, 3, 8]),
       extra=st.sampled_from([0, 0, 1, 5, 16]), seg=st.sampled_from([None, "small", "small"]), mode=st.sampled_from(["field", "inplace"]),
       saturate=st.booleans())
def test_random_geometry_device_entry(cuda, fmtname, wq, h2, order, aa, aac, dh, luma, chroma, seed, kind, skew, extra, seg, mode, saturate):
    fmt = FORMATS[fmtname]
    sb = fmt.sample_bytes
    w = wq * 4
    h = h2 * 4
    if dh:
        mode = "field"
    for name in ("SANGNOM_U8_SEG", "SANGNOM_WIDE_SEG"):
        os.environ.pop(name, None)
    if seg == "small":
        os.environ["SANGNOM_U8_SEG"] = "256"
        os.environ["SANGNOM_WIDE_SEG"] = "128"
    try:
        frames = [make_frame(seed, w, h, fmt, kind, i) for i in range(2)]
        args = dict(order=order, aa=aa, aac=aac, dh=dh, luma=luma, chroma=chroma)
        exp = [O.oracle_frame(fr, fmt.bits, parity=parity_of(i), saturate=saturate, **args) for i, fr in enumerate(frames)]
        keep, jobs, outs = [], [], []
        with cuda.Context(sb, w, h * 2 if dh else h, flags=cuda.FLAG_SATURATE if saturate else 0) as ctx:
            for k, planes in enumerate(frames):
                off = cuda.resolve_offset(order, parity_of(k))
                row = []
                for p, a in enumerate(planes[:3]):
                    enabled = dh or (luma if p == 0 else chroma)
                    thr = cuda.threshold(aa if p == 0 else aac, fmt.bits, sb)
                    sbuf, sview, sp = dev_plane(a, skew, extra)
                    if mode == "inplace" and enabled:
                        jobs.append(cuda.make_job(0, 0, sview.data_ptr(), sp, a.shape[1], a.shape[0], off, cuda.MODE_INPLACE, thr, p, k))
                        keep.append(sbuf)
                        row.append((sview, a))
                        continue
                    dbuf, dview, dp = dev_plane(np.empty((a.shape[0] * (2 if dh else 1), a.shape[1]), a.dtype), (skew * 3) % 5, extra // 2, fill=0xEE)
                    m = cuda.MODE_DH if dh else (cuda.MODE_FIELD if enabled else cuda.MODE_COPY)
                    jobs.append(cuda.make_job(sview.data_ptr(), sp, dview.data_ptr(), dp, a.shape[1], dview.shape[0], off, m, thr, p, k))
                    keep += [sbuf, dbuf]
                    row.append((dview, a))
                outs.append(row)
            ctx.process_jobs_device(jobs)
            ctx.synchronize()
        for i, row in enumerate(outs):
            got = [v[:, :a.shape[1] * sb].cpu().numpy().view(a.dtype).copy() for v, a in row]
            assert_planes_equal(got, exp[i][:3], f"random {fmtname} {w}x{h} {args} skew={skew} extra={extra} seg={seg} {mode} sat={saturate} frame {i}")
    finally:
        for name in ("SANGNOM_U8_SEG", "SANGNOM_WIDE_SEG"):
            os.environ.pop(name, None)


@settings(max_examples=(_N // 3) or 25, deadline=None, derandomize=not _N, suppress_health_check=list(HealthCheck))
@given(fmtname=st.sampled_from(["YV12", "YV411", "YUV420P10", "YUV420PS", "YV24", "Y8"]), wq=st.integers(66, 300), h2=st.integers(2, 10),
       order=st.integers(0, 2), seed=st.integers(0, 2 ** 16), pinned=st.booleans(), pad=st.sampled_from([0, 0, 3, 16]))
def test_random_geometry_host_entry_clustered(cuda, fmtname, wq, h2, order, seed, pinned, pad):
    """Host entry (pinned and pageable planes, odd host pitches) with forced cluster splits."""
    fmt = FORMATS[fmtname]
    sb = fmt.sample_bytes
    w, h = wq * 4, h2 * 4
    os.environ["SANGNOM_U8_SEG"] = "256"
    os.environ["SANGNOM_WIDE_SEG"] = "128"
    try:
        frames = [make_frame(seed, w, h, fmt, "noise", i) for i in range(3)]
        exp = [O.oracle_frame(fr, fmt.bits, order=order, aa=48, aac=40, parity=parity_of(i)) for i, fr in enumerate(frames)]
        with cuda.Context(sb, w, h, max_frames_in_flight=4) as ctx:
            jobs, outs, keep = [], [], []
            for k, planes in enumerate(frames):
                srcs, dsts = [], []
                for a in planes[:3]:
                    alloc = (lambda shape: cuda.pinned_empty(shape, a.dtype)) if pinned else (lambda shape: np.empty(shape, a.dtype))
                    s_ = alloc((a.shape[0], a.shape[1] + pad))[:, :a.shape[1]]
                    s_[...] = a
                    d = alloc((a.shape[0], a.shape[1] + pad))[:, :a.shape[1]]
                    d[...] = 0
                    srcs.append(s_)
                    dsts.append(d)
                keep.append(srcs)
                jobs += ctx.frame_jobs(srcs, dsts, fmt.bits, order=order, aa=48, aac=40, parity=parity_of(k), frame_key=k)
                outs.append(dsts)
            ctx.process_jobs(jobs)
        for i in range(len(frames)):
            assert_planes_equal([np.ascontiguousarray(d) for d in outs[i]], exp[i][:3], f"host random {fmtname} {w}x{h} order={order} pinned={pinned} pad={pad} frame {i}")
    finally:
        for name in ("SANGNOM_U8_SEG", "SANGNOM_WIDE_SEG"):
            os.environ.pop(name, None)
