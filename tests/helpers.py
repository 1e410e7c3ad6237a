"""Shared helpers for the parity tests: run the same seeded frames through the oracle and the GPU."""
import numpy as np

from oracle import oracle as O
from pysangnom.clips import make_frame
from pysangnom.formats import FORMATS

# (name, format, width, height, script arguments, content kind, frames)
# The BASELINE.json configs at sizes the CPU oracle finishes in seconds, plus the edge cases the
# survey lists (non-mod-32 widths, 4:2:2, 4:1:1, Y-only dh, chroma-only, tiny planes).
CASES = [
    ("cfg1_yv12_luma_noise", "YV12", 720, 480, dict(order=1, aa=48, chroma=False), "noise", 3),
    ("cfg1_yv12_luma_edges", "YV12", 720, 480, dict(order=1, aa=48, chroma=False), "edges", 2),
    ("cfg2_420p8_dfr_noise", "YUV420P8", 1920, 1080, dict(order=0, aa=48, aac=48), "noise", 2),
    ("cfg2_420p8_dfr_edges", "YUV420P8", 1920, 1080, dict(order=0, aa=48, aac=48), "edges", 2),
    ("cfg3_444p16_dh", "YUV444P16", 640, 360, dict(dh=True, aa=48), "noise", 2),
    ("cfg3_444p16_dh_T", "YUV444P16", 360, 640, dict(dh=True, aa=48), "edges", 2),
    ("cfg4_420ps", "YUV420PS", 960, 540, dict(order=2, aa=48, aac=24), "noise", 2),
    ("cfg4_420ps_edges", "YUV420PS", 960, 540, dict(order=2, aa=48, aac=24), "edges", 2),
    ("cfg5_420p10", "YUV420P10", 960, 540, dict(order=1, aa=48, aac=48), "noise", 2),
    ("cfg5_y8", "Y8", 1280, 720, dict(order=1, aa=48), "edges", 2),
    ("422p8_kb", "YUV422P8", 640, 482, dict(order=2, aa=48, aac=48), "noise", 2),
    ("y12_dh_dfr", "Y12", 644, 482, dict(dh=True, order=0), "noise", 2),
    ("yv411", "YV411", 640, 480, dict(order=1, aa=48, aac=30), "noise", 1),
    ("chroma_only", "YUV420P8", 720, 480, dict(luma=False, aa=48, aac=48), "noise", 2),
    ("444ps_small", "YUV444PS", 100, 50, dict(order=1, aa=128, aac=128), "edges", 1),
    ("420p16_aa0", "YUV420P16", 332, 244, dict(order=2, aa=0, aac=0), "edges", 1),
    ("tiny_y8", "Y8", 8, 4, dict(order=1), "noise", 1),
    ("two_rows", "Y8", 40, 2, dict(order=2), "noise", 1),
    ("422p10_w2", "YUV422P10", 34, 16, dict(order=1, aa=20, aac=90), "noise", 1),
]


def case_frames(case):
    name, fmtname, w, h, kw, kind, nframes = case
    fmt = FORMATS[fmtname]
    seed = sum(map(ord, name))
    return fmt, [make_frame(seed, w, h, fmt, kind, i) for i in range(nframes)]


def parity_of(i):
    return i % 2 == 0      # DoubleWeave-style: even frames are top-field-first


def oracle_outputs(fmt, frames, kw):
    return [O.oracle_frame(fr, fmt.bits, order=kw.get("order", 1), aa=kw.get("aa", 48), aac=kw.get("aac", 0),
                           dh=kw.get("dh", False), luma=kw.get("luma", True), chroma=kw.get("chroma", True), parity=parity_of(i))
            for i, fr in enumerate(frames)]


def assert_planes_equal(got, exp, what):
    assert len(got) == len(exp), what
    for p, (g, e) in enumerate(zip(got, exp)):
        assert g.shape == e.shape and g.dtype == e.dtype, f"{what} plane {p}: {g.shape}/{g.dtype} vs {e.shape}/{e.dtype}"
        if not np.array_equal(g.view(np.uint8), e.view(np.uint8)):     # bit-exact, fp32 included
            bad = np.argwhere(g != e)
            raise AssertionError(f"{what} plane {p}: {len(bad)} samples differ, first at {bad[:4].tolist()} "
                                 f"got {g[tuple(bad[0])]} expected {e[tuple(bad[0])]}")
