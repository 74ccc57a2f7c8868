"""src/process/switch.ts: 2-input switcher -- Transform x2 -> Mix | Wipe -> Combine with overlays.

Dead code in the reference, and stale against its own operators: it constructs `new Combine(width, height,
numOverlays)` (switch.ts:111-115) where the constructor is `(numLayers, width, height)` (combine.ts:71), and runs the combiner
with `{bgIn, ovIn, output}` (switch.ts:182-189) where `Combine.getKernelParams` wants `{inputs, output}` (combine.ts:88-101).
This mirror keeps the class, its constructor, `init()` and `processFrame()` signatures and call structure, and makes the two
stale calls the ones the current Combine accepts: `Combine(1 + numOverlays, width, height)` and
`inputs = [background, ...overlays]`.  Without overlays the mixed picture is the output (combine needs two layers)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

from ..cl_job_queue import ClJobs
from ..nodencl import OpenCLBuffer, clContext
from .combine import Combine
from .image_process import ImageProcess
from .mix import Mix
from .transform import Transform
from .wipe import Wipe


class Switch:   # switch.ts:29-191
    def __init__(self, clContext_: clContext, chanID: str, clJobs: ClJobs, width: int, height: int, numInputs: int, numOverlays: int):
        self.clContext = clContext_
        self.chanID = f"{chanID} switch"
        self.clJobs = clJobs
        self.width, self.height = width, height
        self.numInputs, self.numOverlays = numInputs, numOverlays
        self.xform0: Optional[ImageProcess] = None
        self.xform1: Optional[ImageProcess] = None
        self.rgbaXf0: Optional[OpenCLBuffer] = None
        self.rgbaXf1: Optional[OpenCLBuffer] = None
        self.rgbaMx: Optional[OpenCLBuffer] = None
        self.mixer: Optional[ImageProcess] = None
        self.wiper: Optional[ImageProcess] = None
        self.combiner: Optional[ImageProcess] = None

    async def _image(self) -> OpenCLBuffer:
        return await self.clContext.createBuffer(self.width * self.height * 4 * 4, "readwrite", "coarse",
                                                 {"width": self.width, "height": self.height}, "switch")

    async def init(self) -> None:   # switch.ts:65-128
        self.xform0 = ImageProcess(self.clContext, Transform(self.clContext, self.width, self.height), self.clJobs)
        await self.xform0.init()
        self.rgbaXf0 = await self._image()
        if self.numInputs > 1:
            self.xform1 = ImageProcess(self.clContext, Transform(self.clContext, self.width, self.height), self.clJobs)
            await self.xform1.init()
            self.rgbaXf1 = await self._image()
            self.mixer = ImageProcess(self.clContext, Mix(self.width, self.height), self.clJobs)
            await self.mixer.init()
            self.wiper = ImageProcess(self.clContext, Wipe(self.width, self.height), self.clJobs)
            await self.wiper.init()
        self.combiner = ImageProcess(self.clContext, Combine(1 + self.numOverlays, self.width, self.height), self.clJobs)
        await self.combiner.init()
        self.rgbaMx = await self._image()

    async def processFrame(self, inParams: List[Dict[str, Any]], mixParams: Dict[str, Any], overlays: List[OpenCLBuffer],
                           output: OpenCLBuffer) -> None:   # switch.ts:130-190
        if not (self.xform0 and (self.numInputs == 1 or (self.xform1 and self.mixer and self.wiper)) and self.combiner):
            raise RuntimeError(f"Switch needs to be initialised {self.numInputs}")
        ident = lambda b: {"source": self.chanID, "timestamp": b.timestamp}
        inBuf0 = inParams[0]["input"]
        mixed = self.rgbaMx if overlays else output
        if self.numInputs > 1:
            inParams[0]["output"] = self.rgbaXf0
            await self.xform0.run(inParams[0], ident(inBuf0), lambda: inBuf0.release())
            inParams[1]["output"] = self.rgbaXf1
            inBuf1 = inParams[1]["input"]
            await self.xform1.run(inParams[1], ident(inBuf1), lambda: inBuf1.release())
            xf0, xf1 = self.rgbaXf0, self.rgbaXf1

            def released() -> None:
                xf0.release()
                xf1.release()
            if mixParams.get("wipe"):
                await self.wiper.run({"input0": xf0, "input1": xf1, "wipe": mixParams["frac"], "output": mixed}, ident(inBuf0), released)
            else:
                await self.mixer.run({"input0": xf0, "input1": xf1, "mix": mixParams["frac"], "output": mixed}, ident(inBuf0), released)
        else:   # switch.ts:178-180: the transformed input is the mixed picture
            inParams[0]["output"] = mixed
            await self.xform0.run(inParams[0], ident(inBuf0), lambda: inBuf0.release())
        if not overlays:
            return
        mx = self.rgbaMx

        def done() -> None:
            mx.release()
            for o in overlays:
                o.release()
        await self.combiner.run({"inputs": [mx, *overlays], "output": output}, ident(inBuf0), done)
