"""src/process/transform.ts: the live resize/DVE (anchor, scale, rotate, translate)."""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np

from ..nodencl import KernelSpec, OpenCLBuffer, clContext
from .colour_maths import matrixFlatten, transformMatrix
from .image_process import ProcessImpl

_KEYS = ("flipH", "flipV", "anchorX", "anchorY", "scaleX", "scaleY", "offsetX", "offsetY", "rotate")


class Transform(ProcessImpl):   # transform.ts:62-189
    def __init__(self, clContext_: clContext, width: int, height: int):
        super().__init__("transform", width, height, KernelSpec("transform"), "transform")
        self.clContext = clContext_
        self.transformMatrix = np.eye(3, dtype=np.float32)
        self.transformArray = matrixFlatten(self.transformMatrix)
        self.matrixBuffer: Optional[OpenCLBuffer] = None
        self.curParams: Optional[Dict[str, Any]] = None

    async def updateMatrix(self, clQueue: int) -> None:
        if not self.matrixBuffer:
            raise RuntimeError("Transform needs to be initialised")
        self.transformArray = matrixFlatten(self.transformMatrix)
        await self.matrixBuffer.hostAccess("writeonly", clQueue, self.transformArray)
        await self.matrixBuffer.hostAccess("none", clQueue)

    async def init(self) -> None:
        self.matrixBuffer = await self.clContext.createBuffer(self.transformArray.nbytes, "readonly", "coarse", None,
                                                              "transformMatrix")
        await self.updateMatrix(self.clContext.queue.load)
        await self.clContext.waitFinish(self.clContext.queue.load)

    def checkParamsChange(self, params: Dict[str, Any]) -> bool:
        return self.curParams is not None and all(params.get(k) == self.curParams.get(k) for k in _KEYS)

    async def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        if not self.checkParamsChange(params):
            # transform.ts:119-171, built by the library so TS/Python/C++ hosts agree bit for bit
            self.transformMatrix = transformMatrix(
                self.width, self.height, params.get("flipH") or False, params.get("flipV") or False,
                params.get("anchorX") or 0.0, params.get("anchorY") or 0.0, params.get("scaleX") or 1.0,
                params.get("scaleY") or 1.0, params.get("offsetX") or 0.0, params.get("offsetY") or 0.0,
                params.get("rotate") or 0.0)
            await self.updateMatrix(self.clContext.queue.load)
            await self.clContext.waitFinish(self.clContext.queue.load)
        self.curParams = params
        if self.matrixBuffer: self.matrixBuffer.addRef()
        kp = {"input": params["input"], "transformMatrix": self.matrixBuffer, "output": params["output"]}
        # extension (not in the reference, BASELINE.json config 5): filter 'lanczosN' selects an N-lobe Lanczos filter
        # for axis-aligned transforms; default = the reference's bilinear image sampler
        flt = params.get("filter")
        if flt:
            if not (isinstance(flt, str) and flt.startswith("lanczos") and flt[7:].isdigit()):
                raise RuntimeError(f"Transform filter must be 'lanczosN', found '{flt}'")
            kp["lanczos"] = int(flt[7:])
        return kp

    def releaseRefs(self) -> None:
        if self.matrixBuffer: self.matrixBuffer.release()
