"""src/process/v210.ts: v210 Reader / Writer PackImpls and the fillBuf fixture."""
from __future__ import annotations

from typing import Any, Dict

import numpy as np

from ..nodencl import KernelSpec
from .packer import Interlace, PackImpl


def getPitch(width: int) -> int:   # v210.ts:198-200
    return width + 47 - ((width - 1) % 48)


def getPitchBytes(width: int) -> int:   # v210.ts:202-204
    return (getPitch(width) * 8) // 3


def fillBuf(buf, width: int, height: int) -> None:
    """v210.ts:206-236: grey ramp, Y 64..940 stepping once per 6-pixel group, Cb=Cr=512."""
    host = buf.host if hasattr(buf, "host") else buf
    pitchBytes = getPitchBytes(width)
    host[:] = 0
    words = host[: pitchBytes * height].view("<u4").reshape(height, pitchBytes // 4)
    groups = (width - (width % 6)) // 6
    remain = width % 6
    per_line = groups + (1 if remain else 0)
    # Y advances once per full group and carries across lines; the tail group reuses the current Y
    idx = np.arange(height)[:, None] * groups + np.arange(per_line)[None, :]
    Y = (64 + idx % 877).astype(np.uint32)
    Cb = Cr = np.uint32(512)
    Yf = Y[:, :groups]
    words[:, 0:groups * 4:4] = (Cr << 20) | (Yf << 10) | Cb
    words[:, 1:groups * 4:4] = (Yf << 20) | (Cb << 10) | Yf
    words[:, 2:groups * 4:4] = (Cb << 20) | (Yf << 10) | Cr
    words[:, 3:groups * 4:4] = (Yf << 20) | (Cr << 10) | Yf
    if remain:
        Yt = Y[:, groups]
        o = groups * 4
        words[:, o] = (Cr << 20) | (Yt << 10) | Cb
        if remain == 2:
            words[:, o + 1] = Yt
        elif remain == 4:
            words[:, o + 1] = (Yt << 20) | (Cb << 10) | Yt
            words[:, o + 2] = (Yt << 10) | Cr


pixelsPerWorkItem = 48


class Reader(PackImpl):   # v210.ts:284-310
    def __init__(self, width: int, height: int):
        super().__init__("v210", width, height, KernelSpec("v210_read"), "read")
        self.numBits, self.lumaBlack, self.lumaWhite, self.chromaRange = 10, 64, 940, 896
        self.isRGB = False
        self.numBytes = [getPitchBytes(self.width) * self.height]
        self.workItemsPerGroup = getPitch(self.width) // pixelsPerWorkItem
        self.globalWorkItems = self.workItemsPerGroup * self.height

    def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        srcArray = params["sources"]
        if len(srcArray) != 1:
            raise RuntimeError(f"Reader for {self.name} requires sources parameter with 1 OpenCL buffer")
        return {"input": srcArray[0], "output": params["dest"], "width": self.width,
                "colMatrix": params.get("colMatrix"), "gammaLut": params.get("gammaLut"),
                "gamutMatrix": params.get("gamutMatrix")}


class Writer(PackImpl):   # v210.ts:312-339
    def __init__(self, width: int, height: int, interlaced: bool):
        super().__init__("v210", width, height, KernelSpec("v210_write"), "write")
        self.interlaced = interlaced
        self.numBits, self.lumaBlack, self.lumaWhite, self.chromaRange = 10, 64, 940, 896
        self.isRGB = False
        self.numBytes = [getPitchBytes(self.width) * self.height]
        self.workItemsPerGroup = getPitch(self.width) // pixelsPerWorkItem
        self.globalWorkItems = (self.workItemsPerGroup * self.height) // (2 if self.interlaced else 1)

    def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        dstArray = params["dests"]
        if len(dstArray) != 1:
            raise RuntimeError(f"Writer for {self.name} requires dests parameter with 1 OpenCL buffer")
        il = params.get("interlace")
        return {"input": params["source"], "output": dstArray[0], "width": self.width,
                "interlace": int(il if (self.interlaced and il is not None) else Interlace.Progressive),
                "colMatrix": params.get("colMatrix"), "gammaLut": params.get("gammaLut")}
