"""src/process/rgba8.ts and bgra8.ts: 8-bit RGBA/BGRA Reader / Writer PackImpls."""
from __future__ import annotations

from typing import Any, Dict

from ..nodencl import KernelSpec
from .packer import Interlace, PackImpl


def getPitchBytes(width: int) -> int:   # rgba8.ts:109-111
    return width * 4


def fillBuf(buf, width: int, height: int, bgra: bool = False) -> None:   # rgba8.ts:114-133
    host = buf.host if hasattr(buf, "host") else buf
    px = host[: width * height * 4].reshape(-1, 4)
    px[:] = (64, 32, 16, 255) if bgra else (16, 32, 64, 255)


class Reader(PackImpl):   # rgba8.ts:163-190
    def __init__(self, width: int, height: int, bgra: bool = False):
        super().__init__("bgra8" if bgra else "rgba8", width, height, KernelSpec("bgra8_read" if bgra else "rgba8_read"), "read")
        self.isRGB = True
        self.numBytes = [getPitchBytes(width) * height]
        self.workItemsPerGroup = width / 64   # Q13: no ceil in the reference
        self.globalWorkItems = self.workItemsPerGroup * height

    def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        srcArray = params["sources"]
        if len(srcArray) != 1:
            raise RuntimeError(f"Reader for {self.name} requires sources parameter with 1 OpenCL buffer")
        return {"input": srcArray[0], "output": params["dest"], "width": self.width,
                "gammaLut": params.get("gammaLut"), "gamutMatrix": params.get("gamutMatrix")}


class Writer(PackImpl):   # rgba8.ts:192-212
    def __init__(self, width: int, height: int, interlaced: bool, bgra: bool = False):
        super().__init__("bgra8" if bgra else "rgba8", width, height, KernelSpec("bgra8_write" if bgra else "rgba8_write"), "write")
        self.interlaced = interlaced
        self.isRGB = True
        self.numBytes = [getPitchBytes(width) * height]
        self.workItemsPerGroup = width / 64
        self.globalWorkItems = (self.workItemsPerGroup * height) / (2 if interlaced else 1)

    def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        dstArray = params["dests"]
        if len(dstArray) != 1:
            raise RuntimeError(f"Writer for {self.name} requires dests parameter with 1 OpenCL buffer")
        il = params.get("interlace")
        return {"input": params["source"], "output": dstArray[0], "width": self.width,
                "interlace": int(il if (self.interlaced and il is not None) else Interlace.Progressive),
                "gammaLut": params.get("gammaLut")}
