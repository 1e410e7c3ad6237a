"""Seeded synthetic clips: identical bytes for the oracle and the GPU.

Two content kinds: "noise" (uniform full range - hits the integer wrap paths and the threshold
fallback) and "edges" (drifting sinusoidal diagonals + 2% noise - hits the directional branches).
"""
from __future__ import annotations

import numpy as np

from .formats import ClipFormat


def _max_code(fmt: ClipFormat):
    return (1 << fmt.bits) - 1


def plane_noise(rng, shape, fmt: ClipFormat, chroma=False):
    if fmt.bits == 32:
        a = rng.random(shape, dtype=np.float32)
        return (a - np.float32(0.5)) if chroma else a
    return rng.integers(0, _max_code(fmt) + 1, size=shape, dtype=np.int64).astype(fmt.dtype)


def plane_edges(rng, shape, fmt: ClipFormat, chroma=False, phase=0.0):
    h, w = shape
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    v = 0.5 + 0.25 * np.sin((x * 0.9 + y * 0.35) * 0.21 + phase) + 0.2 * np.sin((x * 0.3 - y * 0.8) * 0.13 + 2 * phase)
    v += (rng.random(shape) - 0.5) * 0.04
    v = np.clip(v, 0.0, 1.0)
    if fmt.bits == 32:
        a = v.astype(np.float32)
        return (a - np.float32(0.5)) if chroma else a
    return np.round(v * _max_code(fmt)).astype(fmt.dtype)


def make_frame(seed, width, height, fmt: ClipFormat, kind="noise", frame_index=0):
    """Planes [Y, (U, V), (A)] for one frame; deterministic in (seed, frame_index, kind, geometry)."""
    rng = np.random.default_rng([seed, frame_index, 0 if kind == "noise" else 1])
    gen = plane_noise if kind == "noise" else (lambda r, s, f, chroma=False: plane_edges(r, s, f, chroma, 0.37 * frame_index))
    planes = []
    for p in range(fmt.components):
        planes.append(gen(rng, fmt.plane_shape(width, height, p), fmt, chroma=p in (1, 2)))
    return planes
