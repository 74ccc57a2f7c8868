"""src/process/yadif.ts: 3-frame window driver around YadifCl."""
from __future__ import annotations

from typing import Dict, List, Optional

from ..cl_job_queue import ClJobs
from ..nodencl import OpenCLBuffer, clContext
from .image_process import ImageProcess
from .yadif_cl import YadifCl

YadifModes = ("send_frame", "send_field", "send_frame_nospatial", "send_field_nospatial")


class Yadif:   # yadif.ts:30-150
    def __init__(self, clContext_: clContext, clJobs: ClJobs, width: int, height: int, config: Dict, interlaced: bool):
        if config["mode"] not in YadifModes:
            raise RuntimeError(f"Unknown yadif mode '{config['mode']}'")
        self.clContext = clContext_
        self.clJobs = clJobs
        self.width = width
        self.height = height
        self.config = config
        self.interlaced = interlaced
        self.sendField = interlaced and config["mode"] in ("send_field", "send_field_nospatial")
        self.skipSpatial = config["mode"] in ("send_frame_nospatial", "send_field_nospatial")
        self.yadifCl: Optional[ImageProcess] = None
        self.in_: List[OpenCLBuffer] = []
        self.out: Optional[OpenCLBuffer] = None

    async def init(self) -> None:
        self.yadifCl = ImageProcess(self.clContext, YadifCl(self.width, self.height), self.clJobs)
        await self.yadifCl.init()

    async def makeOutput(self, isSecond: bool, sourceID: str, timestamp: int) -> None:
        self.out = await self.clContext.createBuffer(self.width * self.height * 4 * 4, "readwrite", "coarse",
                                                     {"width": self.width, "height": self.height},
                                                     f"yadif {'2' if isSecond else '1'} {sourceID} {timestamp}")

    async def runYadif(self, isSecond: bool, sourceID: str) -> None:
        if not self.yadifCl:
            raise RuntimeError("Yadif needs to be initialised")
        srcs = self.in_[:]
        for s in srcs: s.addRef()
        out = self.out
        out.timestamp = srcs[1].timestamp + (1 if isSecond else 0)
        await self.yadifCl.run(
            {"prev": srcs[0], "cur": srcs[1], "next": srcs[2],
             "parity": (1 if self.config["tff"] else 0) ^ (1 if not isSecond else 0),
             "tff": self.config["tff"], "skipSpatial": self.skipSpatial, "output": out},
            {"source": sourceID, "timestamp": out.timestamp},
            lambda: [s.release() for s in srcs] and None)
        await self.clJobs.runQueue({"source": sourceID, "timestamp": out.timestamp})

    async def processFrame(self, input_: OpenCLBuffer, outputs: List[OpenCLBuffer], sourceID: str) -> None:
        if not self.interlaced:
            outputs.append(input_)
            return
        self.in_.append(input_)
        if len(self.in_) < 3:
            # complete any processing queued for input so the sources are released
            await self.clJobs.runQueue({"source": sourceID, "timestamp": input_.timestamp})
            return
        if len(self.in_) > 3:
            old = self.in_.pop(0)
            old.release()
        await self.makeOutput(False, sourceID, self.in_[1].timestamp)
        await self.runYadif(False, sourceID)
        outputs.append(self.out)
        if self.sendField:
            await self.makeOutput(True, sourceID, self.in_[1].timestamp + 1)
            await self.runYadif(True, sourceID)
            outputs.append(self.out)

    def release(self) -> None:
        for i in self.in_: i.release()
