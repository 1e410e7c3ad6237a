"""Planar clip formats the filter accepts (reference README.md:21-22: Y/YUV(A) 8..32-bit planar) under the
names BASELINE.json uses. Pure description: plane shapes, bit depth, sample type."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class ClipFormat:
    """Planar format: components 1 (Y), 3 (YUV) or 4 (YUVA); log2 chroma subsampling; bit depth."""
    components: int = 3
    sub_w: int = 1
    sub_h: int = 1
    bits: int = 8
    rgb: bool = False
    planar: bool = True

    @property
    def dtype(self):
        return np.uint8 if self.bits <= 8 else (np.uint16 if self.bits <= 16 else np.float32)

    @property
    def sample_bytes(self):
        return np.dtype(self.dtype).itemsize

    def plane_shape(self, width, height, plane):
        if plane in (1, 2):
            return (height >> self.sub_h, width >> self.sub_w)
        return (height, width)


# name -> ClipFormat, the spellings BASELINE.json uses
FORMATS = {
    "Y8": ClipFormat(1, 0, 0, 8), "Y10": ClipFormat(1, 0, 0, 10), "Y12": ClipFormat(1, 0, 0, 12),
    "Y16": ClipFormat(1, 0, 0, 16), "Y32": ClipFormat(1, 0, 0, 32),
    "YV12": ClipFormat(3, 1, 1, 8), "YUV420P8": ClipFormat(3, 1, 1, 8), "YUV420P10": ClipFormat(3, 1, 1, 10),
    "YUV420P16": ClipFormat(3, 1, 1, 16), "YUV420PS": ClipFormat(3, 1, 1, 32),
    "YV16": ClipFormat(3, 1, 0, 8), "YUV422P8": ClipFormat(3, 1, 0, 8), "YUV422P10": ClipFormat(3, 1, 0, 10),
    "YUV422P16": ClipFormat(3, 1, 0, 16), "YUV422PS": ClipFormat(3, 1, 0, 32),
    "YV24": ClipFormat(3, 0, 0, 8), "YUV444P8": ClipFormat(3, 0, 0, 8), "YUV444P10": ClipFormat(3, 0, 0, 10),
    "YUV444P16": ClipFormat(3, 0, 0, 16), "YUV444PS": ClipFormat(3, 0, 0, 32),
    "YV411": ClipFormat(3, 2, 0, 8),
    "YUVA420P8": ClipFormat(4, 1, 1, 8), "YUVA444P16": ClipFormat(4, 0, 0, 16), "YUVA444PS": ClipFormat(4, 0, 0, 32),
    "RGB24": ClipFormat(3, 0, 0, 8, rgb=True, planar=False), "RGBP8": ClipFormat(3, 0, 0, 8, rgb=True, planar=True),
    "YUY2": ClipFormat(3, 1, 0, 8, rgb=False, planar=False),
}
