"""src/process/resize.ts (dead code in the reference, named by the north star)."""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np

from ..nodencl import KernelSpec, OpenCLBuffer, clContext
from .image_process import ProcessImpl


class Resize(ProcessImpl):   # resize.ts:61-138
    def __init__(self, clContext_: clContext, width: int, height: int):
        super().__init__("resize", width, height, KernelSpec("resize"), "resize")
        self.clContext = clContext_
        self.flipH = False
        self.flipV = False
        self.flipArr = np.array([0.0, 1.0, 0.0, 1.0], np.float32)
        self.flipVals: Optional[OpenCLBuffer] = None

    async def updateFlip(self, flipH: bool, flipV: bool, clQueue: int) -> None:
        if self.flipVals is None:
            raise RuntimeError("Resize.updateFlip failed with no program available")
        self.flipH, self.flipV = flipH, flipV
        self.flipArr = np.array([1.0 if flipH else 0.0, -1.0 if flipH else 1.0,
                                 1.0 if flipV else 0.0, -1.0 if flipV else 1.0], np.float32)
        await self.flipVals.hostAccess("writeonly", clQueue, self.flipArr)
        await self.flipVals.hostAccess("none", clQueue)

    async def init(self) -> None:
        self.flipVals = await self.clContext.createBuffer(self.flipArr.nbytes, "readonly", "coarse", None, "flipVals")
        await self.updateFlip(False, False, self.clContext.queue.load)

    async def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        flipH, flipV = bool(params.get("flipH")), bool(params.get("flipV"))
        scale, offsetX, offsetY = params.get("scale"), params.get("offsetX"), params.get("offsetY")
        if not (self.flipH == flipH and self.flipV == flipV):
            await self.updateFlip(flipH, flipV, self.clContext.queue.load)
        if scale and not (scale > 0.0):
            raise RuntimeError("resize scale factor must be greater than zero")
        if offsetX and not (-1.0 <= offsetX <= 1.0):
            raise RuntimeError("resize offsetX must be between -1.0 and +1.0")
        if offsetY and not (-1.0 <= offsetY <= 1.0):
            raise RuntimeError("resize offsetX must be between -1.0 and +1.0")
        if self.flipVals: self.flipVals.addRef()
        return {"input": params["input"], "scale": scale or 1.0, "offsetX": offsetX or 0.0, "offsetY": offsetY or 0.0,
                "flip": self.flipVals, "output": params["output"]}

    def releaseRefs(self) -> None:
        if self.flipVals: self.flipVals.release()
