"""Plugin-level parity on the GPU: our AviSynth plugin, driven by the fake host exactly like the
reference plugin, against the oracle / golden fixtures / (when built) the live reference."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import CASES, assert_planes_equal, case_frames, oracle_outputs, parity_of
from kat import KAT_IN, KATS
from oracle import oracle as O
from pysangnom.clips import make_frame
from pysangnom.formats import FORMATS
from fakehost import MT_NICE_FILTER, FakeHost

pytestmark = pytest.mark.gpu
PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "avisynth-sangnom2_b200")
OURS = os.path.join(PKG, "libsangnom2_b200.so")
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_plugin(plugin, fmt, w, h, frames, kw, func="SangNom2", fresh=False, **host_kw):
    """Default host: one that reports no SSE2, i.e. opt=-1 resolves to the opt=0 C++ arithmetic (the parity contract)
    on both plugins; tests of the SSE2 flavour pass cpu_flags=CPUF_SSE2."""
    host_kw.setdefault("cpu_flags", 0)
    outs = []
    with FakeHost(**host_kw) as host:
        host.load_plugin(plugin)
        src = host.source(w, h, fmt, len(frames), parity_mode=2)
        for i, fr in enumerate(frames):
            src.set_frame(i, fr)
        flt = None
        for i in range(len(frames)):
            if flt is None or fresh:
                flt = host.invoke(func, src, **kw)
            outs.append(flt.get_frame(i))
    return outs


@pytest.mark.parametrize("kw,expected", KATS, ids=["order1_aa48", "order2_aa48", "order1_aa0"])
def test_known_answers(kw, expected):
    out = run_plugin(OURS, FORMATS["Y8"], 16, 8, [[KAT_IN]], kw)
    assert np.array_equal(out[0][0], expected)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_plugin_matches_golden_and_oracle(case):
    name, fmtname, w, h, kw, kind, nframes = case
    fmt, frames = case_frames(case)
    got = run_plugin(OURS, fmt, w, h, frames, kw)      # ONE long-lived instance, batched internally
    g = GOLDEN["cases"][name]
    assert [[sha(p) for p in fr[:3]] for fr in got] == g["output_sha256"], "differs from the reference-generated fixture"
    exp = oracle_outputs(fmt, frames, kw)
    for i in range(nframes):
        assert_planes_equal(got[i][:3], exp[i][:3], f"{name} frame {i}")


@pytest.mark.skipif(O.reference_plugin_path() is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("func,kw", [("SangNom2", dict(order=0, aa=30, aac=70)), ("SangNom", dict(order=0, aa=48)),
                                     ("SangNom", dict(order=2, aa=48, opt=1)), ("SangNom", dict(order=1, aa=10, opt=0))])
@pytest.mark.parametrize("fmtname", ["YV12", "YUV420P16", "YUV420PS"])
def test_same_calls_on_both_plugins(func, kw, fmtname):
    """Identical host call sequence on the reference plugin (C++ path: host reports no SSE2) and on ours.
    Covers the legacy SangNom() entry: order remap and its aac := opt quirk (SangNom2.cpp:437-472)."""
    fmt = FORMATS[fmtname]
    w, h = 176, 144
    frames = [make_frame(21, w, h, fmt, "edges", i) for i in range(3)]
    ref = run_plugin(O.reference_plugin_path(), fmt, w, h, frames, kw, func=func, fresh=True, cpu_flags=0)
    got = run_plugin(OURS, fmt, w, h, frames, kw, func=func, cpu_flags=0)
    for i in range(3):
        assert_planes_equal(got[i][:3], ref[i][:3], f"{func} {kw} {fmtname} frame {i}")


@pytest.mark.skipif(O.reference_plugin_path() is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("fmtname,w,h,kw", [("YV12", 720, 480, dict(order=1, aa=48, chroma=False)), ("YV12", 176, 144, dict(luma=False, aa=48, aac=48)),
                                            ("YUV420P10", 100, 48, dict(order=0, aa=48, aac=48)), ("YUV444PS", 100, 50, dict(dh=True, aa=48, aac=48))])
def test_persistent_mode_equals_one_long_lived_reference_instance(monkeypatch, fmtname, w, h, kw):
    """SANGNOM_B200_PERSISTENT=1: our plugin pulled sequentially == ONE reference instance (opt=0) pulled sequentially,
    including the frames where that differs from a fresh instance (pad columns / chroma-only leftovers)."""
    monkeypatch.setenv("SANGNOM_B200_PERSISTENT", "1")
    monkeypatch.setenv("SANGNOM_B200_BATCH", "3")
    fmt = FORMATS[fmtname]
    frames = [make_frame(77, w, h, fmt, "noise", i) for i in range(7)]
    ref = run_plugin(O.reference_plugin_path(), fmt, w, h, frames, dict(opt=0, **kw), fresh=False)
    got = run_plugin(OURS, fmt, w, h, frames, kw, fresh=False)
    for i in range(len(frames)):
        assert_planes_equal(got[i][:3], ref[i][:3], f"persistent {fmtname} {kw} frame {i}")


@pytest.mark.skipif(O.reference_plugin_path() is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("fmtname,w,h,kw", [("YV12", 352, 288, dict(order=0, aa=48, aac=48)), ("YUV420P16", 176, 144, dict(order=2, aa=48, aac=48)),
                                            ("YUV444P10", 200, 120, dict(dh=True, aa=30)), ("YUV420PS", 176, 144, dict(order=1, aa=48, aac=24))])
def test_opt1_selects_the_sse2_arithmetic(fmtname, w, h, kw):
    """SangNom2(opt=1) on our plugin == the reference's SSE2 path (opt=1 on a host that reports SSE2)."""
    from fakehost import CPUF_SSE2
    fmt = FORMATS[fmtname]
    frames = [make_frame(19, w, h, fmt, "noise", i) for i in range(3)]
    ref = run_plugin(O.reference_plugin_path(), fmt, w, h, frames, dict(opt=1, **kw), fresh=True, cpu_flags=CPUF_SSE2)
    got = run_plugin(OURS, fmt, w, h, frames, dict(opt=1, **kw), cpu_flags=CPUF_SSE2)
    for i in range(len(frames)):
        assert_planes_equal(got[i][:3], ref[i][:3], f"opt=1 {fmtname} {kw} frame {i}")


@pytest.mark.skipif(O.reference_plugin_path() is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("func,kw", [("SangNom2", dict()), ("SangNom2", dict(order=0, aac=48)), ("SangNom", dict()), ("SangNom", dict(order=2, aa=64))])
@pytest.mark.parametrize("fmtname", ["YV12", "YUV420P10", "YUV444P16", "YUV420PS"])
def test_default_arguments_equal_the_stock_reference(func, kw, fmtname):
    """A script that does not pass opt gets, on an SSE2 host (every x86-64 host), the reference's SSE2 arithmetic
    (SangNom2.cpp:312: `opt < 0 && CPUF_SSE2`) - from our plugin too, including the legacy SangNom() entry, which can
    never pass opt at all. Noise content, where the two flavours differ."""
    from fakehost import CPUF_SSE2
    fmt = FORMATS[fmtname]
    w, h = 176, 144
    frames = [make_frame(31, w, h, fmt, "noise", i) for i in range(3)]
    ref = run_plugin(O.reference_plugin_path(), fmt, w, h, frames, kw, func=func, fresh=True, cpu_flags=CPUF_SSE2)
    got = run_plugin(OURS, fmt, w, h, frames, kw, func=func, cpu_flags=CPUF_SSE2)
    for i in range(len(frames)):
        assert_planes_equal(got[i][:3], ref[i][:3], f"{func} {kw} {fmtname} default-opt frame {i}")
    if fmt.sample_bytes == 1:
        wrap = run_plugin(OURS, fmt, w, h, frames, kw, func=func, cpu_flags=0)
        assert any(not np.array_equal(a, b) for fa, fb in zip(got, wrap) for a, b in zip(fa[:3], fb[:3])), "flavours should differ on noise"


def test_legacy_semantics_without_reference():
    """SangNom(order=0) keeps the bottom field, (order=2) is double-rate; aac takes the script's opt."""
    fmt = FORMATS["YV12"]
    w, h = 96, 64
    frames = [make_frame(4, w, h, fmt, "noise", i) for i in range(2)]
    got = run_plugin(OURS, fmt, w, h, frames, dict(order=0, aa=48, opt=1), func="SangNom")
    for i, fr in enumerate(frames):
        exp = O.oracle_frame(fr, 8, order=2, aa=48, aac=1)
        assert_planes_equal(got[i][:3], exp[:3], f"legacy frame {i}")
    got = run_plugin(OURS, fmt, w, h, frames, dict(order=2), func="SangNom")
    for i, fr in enumerate(frames):
        exp = O.oracle_frame(fr, 8, order=0, aa=48, aac=0, parity=parity_of(i))
        assert_planes_equal(got[i][:3], exp[:3], f"legacy dfr frame {i}")


def test_filter_object_properties():
    fmt = FORMATS["YUVA420P8"]
    w, h, n = 64, 32, 5
    with FakeHost(cpu_flags=0) as host:
        host.load_plugin(OURS)
        src = host.source(w, h, fmt, n)
        frames = [make_frame(2, w, h, fmt, "noise", i) for i in range(n)]
        for i, fr in enumerate(frames):
            src.set_frame(i, fr)
            src.set_prop(i, "_FieldBased", i)
        flt = host.invoke("SangNom2", src, dh=True, aa=20)
        info = flt.info()
        assert (info["width"], info["height"], info["num_frames"]) == (w, 2 * h, n)      # dh doubles vi.height
        assert flt.mt_mode() == MT_NICE_FILTER
        planes, props = flt.get_frame(3, with_props=["_FieldBased"])
        assert props == {"_FieldBased": 3}                                                 # NewVideoFrameP copies props
        exp = O.oracle_frame(frames[3], 8, order=1, aa=20, dh=True)
        assert_planes_equal(planes[:3], exp[:3], "dh frame 3")
        assert np.array_equal(planes[3], np.repeat(frames[3][3], 2, axis=0))              # alpha: copied (row-doubled)
    with FakeHost(cpu_flags=0, has_v8=False) as host:
        host.load_plugin(OURS)
        src = host.source(w, h, fmt, 1)
        src.set_frame(0, frames[0])
        src.set_prop(0, "_FieldBased", 7)
        planes, props = host.invoke("SangNom2", src).get_frame(0, with_props=["_FieldBased"])
        assert props == {}                                                                 # pre-v8 host: NewVideoFrame


def test_batched_prefetch_and_seek(monkeypatch):
    monkeypatch.setenv("SANGNOM_B200_BATCH", "4")
    monkeypatch.setenv("SANGNOM_B200_PREFETCH", "0")
    fmt = FORMATS["Y8"]
    w, h, n = 64, 32, 10
    frames = [make_frame(6, w, h, fmt, "edges", i) for i in range(n)]
    exp = [O.oracle_frame(fr, 8, order=0, parity=parity_of(i))[0] for i, fr in enumerate(frames)]
    with FakeHost(cpu_flags=0) as host:
        host.load_plugin(OURS)
        src = host.source(w, h, fmt, n, parity_mode=2)
        for i, fr in enumerate(frames):
            src.set_frame(i, fr)
        flt = host.invoke("SangNom2", src, order=0)
        assert np.array_equal(flt.get_frame(0)[0], exp[0])
        assert src.requests() == [0, 1, 2, 3]              # first sequential miss pulls a whole batch
        for i in (1, 2, 3):
            assert np.array_equal(flt.get_frame(i)[0], exp[i])
        assert src.requests() == [0, 1, 2, 3]              # served from the finished batch
        assert np.array_equal(flt.get_frame(4)[0], exp[4])
        assert src.requests()[4:] == [4, 5, 6, 7]
        assert np.array_equal(flt.get_frame(9)[0], exp[9])   # seek: single frame, clip end respected
        assert src.requests()[8:] == [9]
        assert np.array_equal(flt.get_frame(8)[0], exp[8])   # backwards
        assert np.array_equal(flt.get_frame(2)[0], exp[2])


def test_next_batch_is_prefetched_asynchronously(monkeypatch):
    """One batch ahead: after a sequential miss the following batch is submitted at once (sangnom_cuda_submit) and
    waited for only when one of its frames is pulled; seeks, the clip end and the filter's destruction with a batch in
    flight all keep the output exact."""
    monkeypatch.setenv("SANGNOM_B200_BATCH", "4")
    monkeypatch.setenv("SANGNOM_B200_PREFETCH", "1")
    fmt = FORMATS["YUV420P8"]
    w, h, n = 96, 64, 19
    frames = [make_frame(8, w, h, fmt, "noise", i) for i in range(n)]
    exp = [O.oracle_frame(fr, 8, order=0, aa=48, aac=48, parity=parity_of(i)) for i, fr in enumerate(frames)]
    with FakeHost(cpu_flags=0) as host:
        host.load_plugin(OURS)
        src = host.source(w, h, fmt, n, parity_mode=2)
        for i, fr in enumerate(frames):
            src.set_frame(i, fr)
        flt = host.invoke("SangNom2", src, order=0, aa=48, aac=48)
        assert_planes_equal(flt.get_frame(0)[:3], exp[0][:3], "frame 0")
        assert src.requests() == [0, 1, 2, 3, 4, 5, 6, 7]          # batch 0 finished, batch 1 already in flight
        for i in range(1, 9):
            assert_planes_equal(flt.get_frame(i)[:3], exp[i][:3], f"frame {i}")
        assert sorted(set(src.requests())) == list(range(16))      # pulling frame 4 waited for batch 1 and started batch 2, frame 8 batch 3
        assert_planes_equal(flt.get_frame(17)[:3], exp[17][:3], "seek forward")          # with a batch pending
        assert_planes_equal(flt.get_frame(18)[:3], exp[18][:3], "last frame")
        assert_planes_equal(flt.get_frame(3)[:3], exp[3][:3], "seek back")
        for i in range(4, 12):
            assert_planes_equal(flt.get_frame(i)[:3], exp[i][:3], f"frame {i} again")
        # leave with a batch in flight: the destructor must wait for it


def test_two_batches_ahead_by_default(monkeypatch):
    """Default mode: two batches are kept in flight behind the one being served (one new batch per GetFrame call), in
    frame order; seeks across them, reading to the clip end and leaving with two batches in flight stay exact."""
    monkeypatch.setenv("SANGNOM_B200_BATCH", "4")
    fmt = FORMATS["YUV420P8"]
    w, h, n = 96, 64, 31
    frames = [make_frame(12, w, h, fmt, "noise", i) for i in range(n)]
    exp = [O.oracle_frame(fr, 8, order=0, aa=48, aac=48, parity=parity_of(i)) for i, fr in enumerate(frames)]
    with FakeHost(cpu_flags=0) as host:
        host.load_plugin(OURS)
        src = host.source(w, h, fmt, n, parity_mode=2)
        for i, fr in enumerate(frames):
            src.set_frame(i, fr)
        flt = host.invoke("SangNom2", src, order=0, aa=48, aac=48)
        assert_planes_equal(flt.get_frame(0)[:3], exp[0][:3], "frame 0")
        assert src.requests() == list(range(8))                     # batch 0 finished, batch 1 in flight
        assert_planes_equal(flt.get_frame(1)[:3], exp[1][:3], "frame 1")
        assert src.requests() == list(range(12))                    # ... and batch 2 behind it
        for i in range(2, 4):
            assert_planes_equal(flt.get_frame(i)[:3], exp[i][:3], f"frame {i}")
        assert src.requests() == list(range(12))                    # never more than two ahead
        for i in range(4, 14):
            assert_planes_equal(flt.get_frame(i)[:3], exp[i][:3], f"frame {i}")
        assert src.requests() == sorted(src.requests()) and len(set(src.requests())) == len(src.requests())   # in order, nothing fetched twice
        assert_planes_equal(flt.get_frame(29)[:3], exp[29][:3], "seek forward over two batches in flight")
        assert_planes_equal(flt.get_frame(30)[:3], exp[30][:3], "last frame")
        assert_planes_equal(flt.get_frame(16)[:3], exp[16][:3], "a frame of a batch that was in flight during the seek")
        assert_planes_equal(flt.get_frame(2)[:3], exp[2][:3], "seek back")
        for i in range(3, 12):
            assert_planes_equal(flt.get_frame(i)[:3], exp[i][:3], f"frame {i} again")
        # leave with two batches in flight: the destructor must wait for both


def test_prefetch_failure_does_not_fail_a_finished_frame(monkeypatch):
    """An error while fetching LATER frames for the speculative next batch must not fail the GetFrame whose own frame
    is finished; it surfaces when one of those frames is requested, and the filter keeps working afterwards."""
    from fakehost import AvisynthError
    monkeypatch.setenv("SANGNOM_B200_BATCH", "4")
    fmt = FORMATS["Y8"]
    w, h, n = 64, 32, 12
    frames = [make_frame(9, w, h, fmt, "edges", i) for i in range(n)]
    exp = [O.oracle_frame(fr, 8, order=1)[0] for fr in frames]
    with FakeHost(cpu_flags=0) as host:
        host.load_plugin(OURS)
        src = host.source(w, h, fmt, n)
        for i, fr in enumerate(frames):
            src.set_frame(i, fr)
        src.fail_at(5)
        flt = host.invoke("SangNom2", src)
        for i in range(4):                                  # batch 0; its prefetch of 4..7 hits the failing frame
            assert np.array_equal(flt.get_frame(i)[0], exp[i])
        with pytest.raises(AvisynthError, match="injected failure at frame 5"):
            flt.get_frame(4)
        src.fail_at(-1)
        for i in range(4, n):
            assert np.array_equal(flt.get_frame(i)[0], exp[i])


def test_plugin_on_all_devices(monkeypatch):
    """SANGNOM_B200_DEVICES=all: one filter instance, one pipeline per GPU. Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    monkeypatch.setenv("SANGNOM_B200_DEVICES", "all")
    monkeypatch.setenv("SANGNOM_B200_BATCH", "8")
    fmt = FORMATS["YUV420P8"]
    w, h, n = 176, 144, 40
    frames = [make_frame(15, w, h, fmt, "noise", i) for i in range(n)]
    got = run_plugin(OURS, fmt, w, h, frames, dict(order=0, aa=48, aac=48))
    for i, fr in enumerate(frames):
        exp = O.oracle_frame(fr, 8, order=0, aa=48, aac=48, parity=parity_of(i))
        assert_planes_equal(got[i][:3], exp[:3], f"all-devices frame {i}")
