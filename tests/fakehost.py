"""ctypes driver for the fake AviSynth host (tests/fakehost_src/fake_host.cpp -> tests/fakehost_src/libfakeavs.so).

Test infrastructure: loads an AviSynth plugin (.so exporting AvisynthPluginInit3) the way a
frameserver would, builds source clips from numpy planes, calls the registered script functions
with named arguments and pulls frames back as numpy arrays. The same driver is pointed at our
plugin (libsangnom2_b200.so) and, in the oracle tests, at the unmodified reference plugin.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(os.path.dirname(_HERE), "avisynth-sangnom2_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
FAKEHOST_LIB = os.path.join(_HERE, "fakehost_src", "libfakeavs.so")
CPUF_SSE2 = 0x20
CACHE_GET_MTMODE = 509
MT_NICE_FILTER, MT_MULTI_INSTANCE, MT_SERIALIZED = 1, 2, 3
from pysangnom.formats import FORMATS, ClipFormat  # noqa: E402,F401


class AvisynthError(RuntimeError):
    """What env->ThrowError raised inside the plugin."""


_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(FAKEHOST_LIB):
        raise FileNotFoundError(f"{FAKEHOST_LIB} missing - run __graft_entry__.build() or make -f tests/fakehost_src/Makefile")
    L = C.CDLL(FAKEHOST_LIB)
    vp, ci, cp = C.c_void_p, C.c_int, C.c_char_p
    L.fh_env_create.restype, L.fh_env_create.argtypes = vp, [ci, ci, ci]
    L.fh_env_destroy.restype, L.fh_env_destroy.argtypes = None, [vp]
    L.fh_load_plugin.restype, L.fh_load_plugin.argtypes = cp, [vp, cp, cp, ci]
    L.fh_function_count.restype, L.fh_function_count.argtypes = ci, [vp]
    L.fh_function_name.restype, L.fh_function_name.argtypes = cp, [vp, ci]
    L.fh_function_params.restype, L.fh_function_params.argtypes = cp, [vp, ci]
    L.fh_frames_allocated.restype, L.fh_frames_allocated.argtypes = C.c_long, [vp]
    L.fh_source_create.restype, L.fh_source_create.argtypes = vp, [ci] * 10
    L.fh_source_set_plane.restype, L.fh_source_set_plane.argtypes = ci, [vp, ci, ci, vp, ci]
    L.fh_source_set_prop.restype, L.fh_source_set_prop.argtypes = ci, [vp, ci, cp, C.c_longlong]
    L.fh_source_request_count.restype, L.fh_source_request_count.argtypes = ci, [vp]
    L.fh_source_request_at.restype, L.fh_source_request_at.argtypes = ci, [vp, ci]
    L.fh_source_clear_requests.restype, L.fh_source_clear_requests.argtypes = None, [vp]
    L.fh_source_fail_at.restype, L.fh_source_fail_at.argtypes = None, [vp, ci]
    L.fh_clip_release.restype, L.fh_clip_release.argtypes = None, [vp]
    L.fh_invoke.restype = vp
    L.fh_invoke.argtypes = [vp, cp, vp, ci, C.POINTER(cp), C.POINTER(ci), cp, ci]
    L.fh_clip_info.restype, L.fh_clip_info.argtypes = ci, [vp, C.POINTER(ci)]
    L.fh_clip_cache_hints.restype, L.fh_clip_cache_hints.argtypes = ci, [vp, ci, ci]
    L.fh_clip_parity.restype, L.fh_clip_parity.argtypes = ci, [vp, ci]
    L.fh_get_frame.restype, L.fh_get_frame.argtypes = vp, [vp, vp, ci, cp, ci]
    L.fh_frame_plane.restype, L.fh_frame_plane.argtypes = vp, [vp, ci, C.POINTER(ci)]
    L.fh_frame_get_prop.restype, L.fh_frame_get_prop.argtypes = ci, [vp, cp, C.POINTER(C.c_longlong)]
    L.fh_frame_release.restype, L.fh_frame_release.argtypes = None, [vp]
    _lib = L
    return L


class Clip:
    def __init__(self, host, handle, fmt=None, is_source=False):
        self.host, self.handle, self._fmt, self.is_source = host, handle, fmt, is_source

    def info(self):
        out = (C.c_int * 7)()
        _load().fh_clip_info(self.handle, out)
        return dict(width=out[0], height=out[1], num_frames=out[2], components=out[3], sub_w=out[4], sub_h=out[5], bits=out[6])

    @property
    def fmt(self):
        i = self.info()
        return ClipFormat(i["components"], i["sub_w"], i["sub_h"], i["bits"])

    def set_frame(self, n, planes):
        L = _load()
        for p, a in enumerate(planes):
            a = np.ascontiguousarray(a)
            rc = L.fh_source_set_plane(self.handle, n, p, a.ctypes.data_as(C.c_void_p), a.strides[0])
            if rc != 0:
                raise IndexError(f"frame {n} plane {p}")

    def set_prop(self, n, key, value):
        _load().fh_source_set_prop(self.handle, n, key.encode(), int(value))

    def requests(self):
        L = _load()
        return [L.fh_source_request_at(self.handle, i) for i in range(L.fh_source_request_count(self.handle))]

    def clear_requests(self):
        _load().fh_source_clear_requests(self.handle)

    def fail_at(self, n):
        """Fault injection: the source raises a script error when frame n is requested (-1: never)."""
        _load().fh_source_fail_at(self.handle, int(n))

    def mt_mode(self):
        return _load().fh_clip_cache_hints(self.handle, CACHE_GET_MTMODE, 0)

    def parity(self, n):
        return bool(_load().fh_clip_parity(self.handle, n))

    def get_frame(self, n, with_props=()):
        """Returns list of numpy planes (copies). with_props: keys to read back -> (planes, props)."""
        L = _load()
        err = C.create_string_buffer(1024)
        fh = L.fh_get_frame(self.host.env, self.handle, n, err, len(err))
        if not fh:
            raise AvisynthError(err.value.decode())
        try:
            info = self.info()
            fmt = ClipFormat(info["components"], info["sub_w"], info["sub_h"], info["bits"])
            planes = []
            for p in range(info["components"]):
                out = (C.c_int * 3)()
                ptr = L.fh_frame_plane(fh, p, out)
                pitch, row_size, height = out[0], out[1], out[2]
                raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(height, pitch))
                planes.append(raw[:, :row_size].copy().view(fmt.dtype))
            if with_props:
                props = {}
                for k in with_props:
                    v = C.c_longlong()
                    if L.fh_frame_get_prop(fh, k.encode(), C.byref(v)):
                        props[k] = v.value
                return planes, props
            return planes
        finally:
            L.fh_frame_release(fh)

    def release(self):
        if self.handle:
            _load().fh_clip_release(self.handle)
            self.handle = None


class FakeHost:
    """One IScriptEnvironment. cpu_flags/has_v8 model the host capabilities a plugin probes."""

    def __init__(self, cpu_flags=CPUF_SSE2, has_v8=True, poison_new_frames=True):
        self.env = _load().fh_env_create(cpu_flags, int(has_v8), int(poison_new_frames))
        self._clips = []

    def load_plugin(self, path):
        err = C.create_string_buffer(1024)
        name = _load().fh_load_plugin(self.env, os.fsencode(path), err, len(err))
        if name is None:
            raise OSError(err.value.decode())
        return name.decode()

    def functions(self):
        L = _load()
        return {L.fh_function_name(self.env, i).decode(): L.fh_function_params(self.env, i).decode()
                for i in range(L.fh_function_count(self.env))}

    def frames_allocated(self):
        return _load().fh_frames_allocated(self.env)

    def source(self, width, height, fmt: ClipFormat, num_frames=1, parity_mode=2):
        h = _load().fh_source_create(width, height, fmt.components, fmt.sub_w, fmt.sub_h, fmt.bits,
                                     int(fmt.rgb), int(fmt.planar), num_frames, parity_mode)
        c = Clip(self, h, fmt, True)
        self._clips.append(c)
        return c

    def looped_source(self, width, height, fmt: ClipFormat, stored, num_frames, parity_mode=2):
        """A long clip that repeats its `stored` frames (set_frame(0..stored-1)); for throughput runs."""
        L = _load()
        L.fh_source_create_looped.restype, L.fh_source_create_looped.argtypes = C.c_void_p, [C.c_int] * 11
        h = L.fh_source_create_looped(width, height, fmt.components, fmt.sub_w, fmt.sub_h, fmt.bits,
                                      int(fmt.rgb), int(fmt.planar), stored, num_frames, parity_mode)
        c = Clip(self, h, fmt, True)
        self._clips.append(c)
        return c

    def invoke(self, func, clip, **kwargs):
        names = (C.c_char_p * max(1, len(kwargs)))(*[k.encode() for k in kwargs])
        vals = (C.c_int * max(1, len(kwargs)))(*[int(v) for v in kwargs.values()])
        err = C.create_string_buffer(1024)
        h = _load().fh_invoke(self.env, func.encode(), clip.handle, len(kwargs), names, vals, err, len(err))
        if not h:
            raise AvisynthError(err.value.decode())
        c = Clip(self, h)
        self._clips.append(c)
        return c

    def close(self):
        for c in reversed(self._clips):
            c.release()
        self._clips.clear()
        if self.env:
            _load().fh_env_destroy(self.env)
            self.env = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
