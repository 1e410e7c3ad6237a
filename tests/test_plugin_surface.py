"""Script surface of our AviSynth plugin vs the reference's: registered names and parameter strings,
defaults, every validation message (SangNom2.cpp:407-422, :446-459) in the reference's check order."""
import os

import pytest

from oracle import oracle as O
from pysangnom.formats import FORMATS
from fakehost import CPUF_SSE2, AvisynthError, FakeHost

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "avisynth-sangnom2_b200")
OURS = os.path.join(PKG, "libsangnom2_b200.so")

# (format, width, height, kwargs, expected message tail) - ordered so each one trips exactly the check named
BAD = [
    ("RGBP8", 64, 32, {}, "clip must be in Y/YUV planar format."),
    ("YUY2", 64, 32, {}, "clip must be in Y/YUV planar format."),
    ("Y8", 64, 33, {}, "height must be even."),
    ("YV12", 64, 34, {}, "height must be mod4."),
    ("YV12", 64, 32, dict(order=3), "order must be between 0..2."),
    ("YV12", 64, 32, dict(order=-1), "order must be between 0..2."),
    ("YV12", 64, 32, dict(aa=129), "aa must be between 0..128."),
    ("YV12", 64, 32, dict(aa=-1), "aa must be between 0..128."),
    ("YV12", 64, 32, dict(aac=200), "aac must be between 0..128."),
    ("YV12", 64, 32, dict(opt=2), "opt must be between -1..2."),
    ("YV12", 64, 32, dict(opt=-2), "opt must be between -1..2."),
    # several wrong at once: the first check in the reference's order wins
    ("YV12", 64, 34, dict(order=7, aa=500), "height must be mod4."),
    ("Y8", 64, 33, dict(opt=9), "height must be even."),
]


def _error(plugin, func, fmtname, w, h, kw, cpu_flags=CPUF_SSE2):
    with FakeHost(cpu_flags=cpu_flags) as host:
        host.load_plugin(plugin)
        src = host.source(w, h, FORMATS[fmtname], 1)
        with pytest.raises(AvisynthError) as e:
            host.invoke(func, src, **kw)
        return str(e.value)


def test_registration_matches_reference():
    with FakeHost() as host:
        assert host.load_plugin(OURS) == "SangNom2"
        ours = host.functions()
    assert ours == {"SangNom2": "c[order]i[aa]i[aac]i[threads]i[dh]b[luma]b[chroma]b[opt]i", "SangNom": "c[order]i[aa]i[opt]i"}
    if O.reference_plugin_path():
        with FakeHost() as host:
            assert host.load_plugin(O.reference_plugin_path()) == "SangNom2"
            assert host.functions() == ours


@pytest.mark.parametrize("fmtname,w,h,kw,tail", BAD)
def test_sangnom2_validation_messages(fmtname, w, h, kw, tail):
    msg = _error(OURS, "SangNom2", fmtname, w, h, kw)
    assert msg == "SangNom2: " + tail
    if O.reference_plugin_path():
        assert _error(O.reference_plugin_path(), "SangNom2", fmtname, w, h, kw) == msg


@pytest.mark.parametrize("fmtname,w,h,kw,tail", [b for b in BAD if "aac" not in b[3] and "opt" not in b[3]])
def test_legacy_validation_messages(fmtname, w, h, kw, tail):
    msg = _error(OURS, "SangNom", fmtname, w, h, kw)
    assert msg == "SangNom: " + tail
    if O.reference_plugin_path():
        assert _error(O.reference_plugin_path(), "SangNom", fmtname, w, h, kw) == msg


def test_opt1_requires_sse2():
    assert _error(OURS, "SangNom2", "YV12", 64, 32, dict(opt=1), cpu_flags=0) == "SangNom2: opt=1 requires SSE2."
    if O.reference_plugin_path():
        assert _error(O.reference_plugin_path(), "SangNom2", "YV12", 64, 32, dict(opt=1), cpu_flags=0) == "SangNom2: opt=1 requires SSE2."


def test_unknown_named_argument_is_a_script_error():
    # the legacy function has no aac/dh parameters (signature "c[order]i[aa]i[opt]i")
    assert "does not have a named argument" in _error(OURS, "SangNom", "YV12", 64, 32, dict(aac=3))


def test_construction_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    msg = _error(OURS, "SangNom2", "YV12", 64, 32, {})
    assert msg.startswith("SangNom2: no CUDA device") and "no CPU path" in msg
