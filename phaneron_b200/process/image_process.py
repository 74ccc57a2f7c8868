"""src/process/imageProcess.ts: ProcessImpl base and the ImageProcess runner."""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional

from ..cl_job_queue import ClJobs
from ..nodencl import KernelSpec, OpenCLProgram, clContext


class ProcessImpl:   # imageProcess.ts:24-56
    def __init__(self, name: str, width: int, height: int, kernel: KernelSpec, programName: str):
        self.name = name
        self.width = width
        self.height = height
        self.kernel = kernel
        self.programName = programName
        self.globalWorkItems = 0

    async def init(self) -> None:
        return None

    def getName(self) -> str: return self.name
    def getNumBytesRGBA(self) -> int: return self.width * self.height * 4 * 4
    def getGlobalWorkItems(self) -> List[int]: return [self.width, self.height]

    async def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        raise NotImplementedError

    def releaseRefs(self) -> None:
        return None


class ImageProcess:   # imageProcess.ts:58-88
    def __init__(self, clContext_: clContext, processImpl: ProcessImpl, clJobs: ClJobs):
        self.clContext = clContext_
        self.processImpl = processImpl
        self.clJobs = clJobs
        self.program: Optional[OpenCLProgram] = None

    async def init(self) -> None:
        self.program = await self.clContext.createProgram(self.processImpl.kernel, {
            "name": self.processImpl.programName,
            "globalWorkItems": self.processImpl.getGlobalWorkItems(),
            "width": self.processImpl.width,
            "height": self.processImpl.height,
        })
        return await self.processImpl.init()

    async def run(self, params: Dict[str, Any], id_, cb: Callable[[], None]) -> None:
        if self.program is None:
            raise RuntimeError("Loader.run failed with no program available")
        kernelParams = await self.processImpl.getKernelParams(params)

        def done() -> None:
            self.processImpl.releaseRefs()
            cb()
        self.clJobs.add(id_, self.processImpl.getName(), self.program, kernelParams, done)

    def finish(self) -> None:
        self.processImpl.releaseRefs()
