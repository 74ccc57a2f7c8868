"""src/process/io.ts: ToRGBA / FromRGBA facades (buffer factories, H2D, queue job, D2H)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from ..cl_job_queue import ClJobs
from ..nodencl import OpenCLBuffer, clContext
from .load_save import Loader, Saver
from .packer import Interlace, PackImpl


class ToRGBA:   # io.ts:26-114
    def __init__(self, clContext_: clContext, colSpecRead: str, colSpecWrite: str, readImpl: PackImpl, clJobs: ClJobs):
        self.clContext = clContext_
        self.loader = Loader(clContext_, colSpecRead, colSpecWrite, readImpl, clJobs)
        self.numBytes = readImpl.getNumBytes()
        self.numBytesRGBA = readImpl.getNumBytesRGBA()
        self.totalBytes = readImpl.getTotalBytes()

    async def init(self) -> None:
        await self.loader.init()

    def getNumBytes(self) -> List[int]: return self.numBytes
    def getNumBytesRGBA(self) -> int: return self.numBytesRGBA
    def getTotalBytes(self) -> int: return self.totalBytes

    async def createSources(self, srcID: str = "") -> List[OpenCLBuffer]:
        return [await self.clContext.createBuffer(b, "readonly", "coarse", None, f"ToRGBA src {srcID}") for b in self.numBytes]

    async def createDest(self, imageDims: Dict[str, int], srcID: str = "") -> OpenCLBuffer:
        return await self.clContext.createBuffer(self.numBytesRGBA, "readonly", "coarse", imageDims, f"ToRGBA {srcID}")

    async def loadFrame(self, input_: Union[np.ndarray, bytes, Sequence], sources: List[OpenCLBuffer],
                        clQueue: Optional[int] = None) -> None:
        inputs = list(input_) if isinstance(input_, (list, tuple)) else [input_]
        if len(sources) != len(inputs):
            raise RuntimeError(f"Expected buffer array of {len(sources)} sources, found {len(inputs)}")
        for i, inp in enumerate(inputs):
            arr = np.frombuffer(inp, np.uint8) if not isinstance(inp, np.ndarray) else inp.view(np.uint8).reshape(-1)
            await sources[i].hostAccess("writeonly", clQueue or 0, arr[: self.numBytes[i]])
            await sources[i].hostAccess("none", clQueue or 0)

    def processFrame(self, sourceID: str, sources: List[OpenCLBuffer], dest: OpenCLBuffer) -> None:
        self.loader.run({"sources": sources, "dest": dest},
                        {"source": sourceID, "timestamp": sources[0].timestamp},
                        lambda: [s.release() for s in sources] and None)

    def finish(self) -> None:
        self.loader.releaseRefs()


class FromRGBA:   # io.ts:116-179
    def __init__(self, clContext_: clContext, colSpecRead: str, writeImpl: PackImpl, clJobs: ClJobs):
        self.clContext = clContext_
        self.saver = Saver(clContext_, colSpecRead, writeImpl, clJobs)
        self.numBytes = writeImpl.getNumBytes()
        self.numBytesRGBA = writeImpl.getNumBytesRGBA()
        self.totalBytes = writeImpl.getTotalBytes()

    async def init(self) -> None:
        await self.saver.init()

    def getNumBytes(self) -> List[int]: return self.numBytes
    def getNumBytesRGBA(self) -> int: return self.numBytesRGBA
    def getTotalBytes(self) -> int: return self.totalBytes

    async def createDests(self, sourceID: str = "") -> List[OpenCLBuffer]:
        return [await self.clContext.createBuffer(b, "writeonly", "coarse", None, f"FromRGBA {sourceID}") for b in self.numBytes]

    def processFrame(self, sourceID: str, source: OpenCLBuffer, dests: List[OpenCLBuffer],
                     interlace: Optional[Interlace] = None) -> None:
        self.saver.run({"source": source, "dests": dests, "interlace": interlace},
                       {"source": sourceID, "timestamp": source.timestamp},
                       lambda: source.release())

    async def saveFrame(self, output: Union[OpenCLBuffer, List[OpenCLBuffer]], clQueue: Optional[int] = None) -> None:
        outputs = output if isinstance(output, list) else [output]
        for o in outputs:
            await o.hostAccess("readonly", clQueue or 0)

    def finish(self) -> None:
        self.saver.releaseRefs()
