"""src/process/packer.ts: Interlace enum, PackImpl base, Packer base."""
from __future__ import annotations

import enum
from typing import Any, Dict, List, Optional

from ..cl_job_queue import ClJobs
from ..nodencl import KernelSpec, OpenCLProgram, clContext


class Interlace(enum.IntEnum):   # packer.ts:24-28
    Progressive = 0
    TopField = 1
    BottomField = 3


class PackImpl:   # packer.ts:30-91
    def __init__(self, name: str, width: int, height: int, kernel: KernelSpec, programName: str):
        self.name = name
        self.width = width
        self.height = height
        self.interlaced = False
        self.kernel = kernel
        self.programName = programName
        self.numBits = 10
        self.lumaBlack = 64
        self.lumaWhite = 940
        self.chromaRange = 896
        self.isRGB = True
        self.numBytes: List[int] = [0]
        self.globalWorkItems = 0
        self.workItemsPerGroup = 0

    def getName(self) -> str: return self.name
    def getWidth(self) -> int: return self.width
    def getHeight(self) -> int: return self.height
    def getNumBytes(self) -> List[int]: return self.numBytes
    def getNumBytesRGBA(self) -> int: return self.width * self.height * 4 * 4
    def getIsRGB(self) -> bool: return self.isRGB
    def getTotalBytes(self) -> int: return sum(self.numBytes)
    def getGlobalWorkItems(self) -> int: return self.globalWorkItems
    def getWorkItemsPerGroup(self) -> int: return self.workItemsPerGroup

    def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        raise NotImplementedError


class Packer:   # packer.ts:84-106
    def __init__(self, clContext_: clContext, packImpl: PackImpl, clJobs: ClJobs):
        self.clContext = clContext_
        self.packImpl = packImpl
        self.clJobs = clJobs
        self.program: Optional[OpenCLProgram] = None

    async def init(self) -> None:
        self.program = await self.clContext.createProgram(self.packImpl.kernel, {
            "name": self.packImpl.programName,
            "globalWorkItems": self.packImpl.getGlobalWorkItems(),
            "workItemsPerGroup": self.packImpl.getWorkItemsPerGroup(),
            "width": self.packImpl.getWidth(),
            "height": self.packImpl.getHeight(),
        })
