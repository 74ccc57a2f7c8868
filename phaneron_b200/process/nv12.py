"""src/process/nv12.ts: 8-bit 4:2:0 with one interleaved chroma plane (see yuv420p.py)."""
from __future__ import annotations

from . import yuv420p as _p

pixelsPerWorkItem = _p.pixelsPerWorkItem
getPitch = _p.getPitch
getPitchBytes = _p.getPitchBytes


def fillBuf(buf, width: int, height: int) -> None:   # nv12.ts:246-281
    _p.fillBuf(buf, width, height, True)


class Reader(_p.Reader):   # nv12.ts:318-345
    def __init__(self, width: int, height: int):
        super().__init__(width, height, True)


class Writer(_p.Writer):   # nv12.ts:347-381
    def __init__(self, width: int, height: int, interlaced: bool):
        super().__init__(width, height, interlaced, True)
