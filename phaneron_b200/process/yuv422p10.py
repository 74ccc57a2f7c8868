"""src/process/yuv422p10.ts and yuv422p8.ts: planar 4:2:2 Reader / Writer PackImpls (FFmpegProducer / FFmpegConsumer
formats).  The two reference files differ only in sample type and range constants; `bits` selects (yuv422p8.py
re-exports the 8-bit flavour under the reference's module name)."""
from __future__ import annotations

import math
from typing import Any, Dict

import numpy as np

from ..nodencl import KernelSpec
from .packer import Interlace, PackImpl

pixelsPerWorkItem = 64   # yuv422p10.ts:295


def getPitch(width: int) -> int:   # yuv422p10.ts:222
    return width + 7 - ((width - 1) % 8)


def getPitchBytes(width: int, bits: int = 10) -> int:   # yuv422p10.ts:223 / yuv422p8.ts:223
    return getPitch(width) * (2 if bits == 10 else 1)


def fillBuf(buf, width: int, height: int, bits: int = 10) -> None:
    """yuv422p10.ts:225-255 / yuv422p8.ts:225-253: luma ramp stepping by 2 per pixel pair, neutral chroma"""
    host = buf.host if hasattr(buf, "host") else buf
    dt = np.dtype("<u2") if bits == 10 else np.uint8
    black, grey, wrap = (64, 512, 938) if bits == 10 else (16, 128, 234)
    pitch = getPitch(width)
    luma = getPitchBytes(width, bits) * height
    Y = host[:luma].view(dt).reshape(height, pitch)
    U = host[luma: luma + luma // 2].view(dt).reshape(height, pitch // 2)
    V = host[luma + luma // 2: 2 * luma].view(dt).reshape(height, pitch // 2)
    Y[:] = black
    U[:] = grey
    V[:] = grey
    span = (wrap - black) // 2 + 1
    pairs = (np.arange(height * ((width + 1) // 2), dtype=np.int64) % span) * 2 + black
    pairs = pairs.reshape(height, -1)
    Y[:, 0:width:2] = pairs[:, : (width + 1) // 2]
    Y[:, 1:width:2] = pairs[:, : width // 2] + 1


class Reader(PackImpl):   # yuv422p10.ts:297-327
    def __init__(self, width: int, height: int, bits: int = 10):
        name = "yuv422p10le" if bits == 10 else "yuv422p8"
        super().__init__(name, width, height, KernelSpec(f"yuv422p{bits}_read"), "read")
        self.bits = bits
        self.numBits = bits
        self.lumaBlack, self.lumaWhite, self.chromaRange = (64, 940, 896) if bits == 10 else (16, 235, 224)
        self.isRGB = False
        lumaBytes = getPitchBytes(width, bits) * height
        self.numBytes = [lumaBytes, lumaBytes // 2, lumaBytes // 2]
        self.workItemsPerGroup = math.ceil(getPitch(width) / pixelsPerWorkItem)
        self.globalWorkItems = self.workItemsPerGroup * height

    def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        srcArray = params["sources"]
        if len(srcArray) != 3:
            raise RuntimeError(f"Reader for {self.name} requires sources parameter with 3 OpenCL buffers")
        return {"inputY": srcArray[0], "inputU": srcArray[1], "inputV": srcArray[2], "output": params["dest"], "width": self.width,
                "colMatrix": params.get("colMatrix"), "gammaLut": params.get("gammaLut"), "gamutMatrix": params.get("gamutMatrix")}


class Writer(PackImpl):   # yuv422p10.ts:329-356
    def __init__(self, width: int, height: int, interlaced: bool, bits: int = 10):
        name = "yuv422p10le" if bits == 10 else "yuv422p8"
        super().__init__(name, width, height, KernelSpec(f"yuv422p{bits}_write"), "write")
        self.bits = bits
        self.interlaced = interlaced
        self.numBits = bits
        self.lumaBlack, self.lumaWhite, self.chromaRange = (64, 940, 896) if bits == 10 else (16, 235, 224)
        self.isRGB = False
        lumaBytes = getPitchBytes(width, bits) * height
        self.numBytes = [lumaBytes, lumaBytes // 2, lumaBytes // 2]
        self.workItemsPerGroup = math.ceil(getPitch(width) / pixelsPerWorkItem)
        self.globalWorkItems = (self.workItemsPerGroup * height) / (2 if interlaced else 1)

    def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        dstArray = params["dests"]
        if len(dstArray) != 3:
            raise RuntimeError(f"Writer for {self.name} requires dests parameter with 3 OpenCL buffers")
        il = params.get("interlace")
        return {"input": params["source"], "outputY": dstArray[0], "outputU": dstArray[1], "outputV": dstArray[2], "width": self.width,
                "interlace": int(il if (self.interlaced and il is not None) else Interlace.Progressive),
                "colMatrix": params.get("colMatrix"), "gammaLut": params.get("gammaLut")}
