"""src/process/yuv422p8.ts: the 8-bit flavour of the planar 4:2:2 packer (see yuv422p10.py)."""
from __future__ import annotations

from . import yuv422p10 as _p

pixelsPerWorkItem = _p.pixelsPerWorkItem
getPitch = _p.getPitch


def getPitchBytes(width: int) -> int:   # yuv422p8.ts:223
    return _p.getPitchBytes(width, 8)


def fillBuf(buf, width: int, height: int) -> None:   # yuv422p8.ts:225-253
    _p.fillBuf(buf, width, height, 8)


class Reader(_p.Reader):   # yuv422p8.ts:297-327
    def __init__(self, width: int, height: int):
        super().__init__(width, height, 8)


class Writer(_p.Writer):   # yuv422p8.ts:329-356
    def __init__(self, width: int, height: int, interlaced: bool):
        super().__init__(width, height, interlaced, 8)
