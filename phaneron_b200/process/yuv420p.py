"""src/process/yuv420p.ts and nv12.ts: 8-bit 4:2:0 Reader / Writer PackImpls (FFmpegProducer formats).  The two reference
files differ only in the chroma layout -- two planes (yuv420p) or one plane of interleaved (U, V) pairs (nv12); `nv12`
selects (nv12.py re-exports that flavour under the reference's module name)."""
from __future__ import annotations

import math
from typing import Any, Dict

import numpy as np

from ..nodencl import KernelSpec
from .packer import Interlace, PackImpl

pixelsPerWorkItem = 64   # yuv420p.ts:327 / nv12.ts:316 (one image line PAIR per work group)


def getPitch(width: int) -> int:   # yuv420p.ts:240
    return width + 7 - ((width - 1) % 8)


def getPitchBytes(width: int) -> int:   # yuv420p.ts:241
    return getPitch(width)


def fillBuf(buf, width: int, height: int, nv12: bool = False) -> None:
    """yuv420p.ts:243-280 / nv12.ts:246-281: per line pair, a luma ramp stepping up by 2 per pixel pair on the first line
    (Y0, Y0 + 1) and down on the second (Y1 + 1, Y1), both carried across pairs; neutral chroma"""
    host = buf.host if hasattr(buf, "host") else buf
    pitch = getPitchBytes(width)
    luma = pitch * height
    host[:luma] = 16
    host[luma:] = 128
    Y = host[:luma].reshape(height, pitch)
    n = (width + 1) // 2                       # pixel pairs per line
    k = np.arange((height // 2 + height % 2) * n, dtype=np.int64).reshape(-1, n)
    span = (234 - 16) // 2 + 1                 # 110 steps: 16, 18 ... 234, then wrap
    y0 = 16 + 2 * (k % span)
    y1 = 234 - 2 * (k % span)
    even, odd = Y[0::2], Y[1::2]
    even[:, 0:width:2] = y0[:, : (width + 1) // 2]
    even[:, 1:width:2] = y0[:, : width // 2] + 1
    odd[:, 0:width:2] = y1[: odd.shape[0], : (width + 1) // 2] + 1
    odd[:, 1:width:2] = y1[: odd.shape[0], : width // 2]


def _num_bytes(width: int, height: int, nv12: bool):
    lumaBytes = getPitchBytes(width) * height
    return [lumaBytes, lumaBytes // 2] if nv12 else [lumaBytes, lumaBytes // 4, lumaBytes // 4]


class Reader(PackImpl):   # yuv420p.ts:329-360 / nv12.ts:318-345
    def __init__(self, width: int, height: int, nv12: bool = False):
        name = "nv12" if nv12 else "yuv420p"
        super().__init__(name, width, height, KernelSpec(f"{name}_read"), "read")
        self.nv12 = nv12
        self.numBits = 8
        self.lumaBlack, self.lumaWhite, self.chromaRange = 16, 235, 224
        self.isRGB = False
        self.numBytes = _num_bytes(width, height, nv12)
        self.workItemsPerGroup = math.ceil(getPitch(width) / pixelsPerWorkItem)
        self.globalWorkItems = (self.workItemsPerGroup * height) / 2   # each item processes two lines

    def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        srcArray = params["sources"]
        want = 2 if self.nv12 else 3
        if len(srcArray) != want:
            raise RuntimeError(f"Reader for {self.name} requires 'sources' parameter with {want} OpenCL buffers")
        planes = ({"inputY": srcArray[0], "inputC": srcArray[1]} if self.nv12
                  else {"inputY": srcArray[0], "inputU": srcArray[1], "inputV": srcArray[2]})
        return {**planes, "output": params["dest"], "width": self.width, "colMatrix": params.get("colMatrix"),
                "gammaLut": params.get("gammaLut"), "gamutMatrix": params.get("gamutMatrix")}


class Writer(PackImpl):   # yuv420p.ts:362-395 / nv12.ts:347-381
    def __init__(self, width: int, height: int, interlaced: bool, nv12: bool = False):
        name = "nv12" if nv12 else "yuv420p"
        super().__init__(name, width, height, KernelSpec(f"{name}_write"), "write")
        self.nv12 = nv12
        self.interlaced = interlaced
        self.numBits = 8
        self.lumaBlack, self.lumaWhite, self.chromaRange = 16, 235, 224
        self.isRGB = False
        self.numBytes = _num_bytes(width, height, nv12)
        self.workItemsPerGroup = math.ceil(getPitch(width) / pixelsPerWorkItem)
        self.globalWorkItems = (self.workItemsPerGroup * height) / 2   # also for a field: one line of each pair

    def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        dstArray = params["dests"]
        want = 2 if self.nv12 else 3
        if len(dstArray) != want:
            raise RuntimeError(f"Writer for {self.name} requires 'dests' parameter with {want} OpenCL buffers")
        planes = ({"outputY": dstArray[0], "outputC": dstArray[1]} if self.nv12
                  else {"outputY": dstArray[0], "outputU": dstArray[1], "outputV": dstArray[2]})
        il = params.get("interlace")
        return {"input": params["source"], **planes, "width": self.width,
                "interlace": int(il if (self.interlaced and il is not None) else Interlace.Progressive),
                "colMatrix": params.get("colMatrix"), "gammaLut": params.get("gammaLut")}
