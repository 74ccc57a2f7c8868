"""src/process/loadSave.ts: Loader (read side constants + job) and Saver (write side)."""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional

import numpy as np

from ..cl_job_queue import ClJobs
from ..nodencl import OpenCLBuffer, clContext
from .colour_maths import (gamma2linearLUT, linear2gammaLUT, matrixFlatten, rgb2rgbMatrix, rgb2ycbcrMatrix,
                           ycbcr2rgbMatrix)
from .packer import Packer, PackImpl


async def _upload(ctx: clContext, arr: np.ndarray, svm: str, owner: str) -> OpenCLBuffer:
    # createBuffer + hostAccess('writeonly') + Buffer.copy (loadSave.ts:69-77)
    buf = await ctx.createBuffer(arr.nbytes, "readonly", svm, None, owner)
    await buf.hostAccess("writeonly")
    buf.host[:] = arr.view(np.uint8).reshape(-1)
    return buf


class Loader(Packer):   # loadSave.ts:33-128
    def __init__(self, clContext_: clContext, colSpec: str, outColSpec: str, packImpl: PackImpl, clJobs: ClJobs):
        super().__init__(clContext_, packImpl, clJobs)
        self.gammaArray = gamma2linearLUT(colSpec)
        self.colMatrixArray: Optional[np.ndarray] = None
        if not self.packImpl.getIsRGB():
            self.colMatrixArray = matrixFlatten(ycbcr2rgbMatrix(colSpec, packImpl.numBits, packImpl.lumaBlack,
                                                                packImpl.lumaWhite, packImpl.chromaRange))
        self.gamutMatrixArray = matrixFlatten(rgb2rgbMatrix(colSpec, outColSpec))
        self.gammaLut: Optional[OpenCLBuffer] = None
        self.colMatrix: Optional[OpenCLBuffer] = None
        self.gamutMatrix: Optional[OpenCLBuffer] = None

    async def init(self) -> None:
        await super().init()
        self.gammaLut = await _upload(self.clContext, self.gammaArray, "coarse", "loader gammaLut")
        if self.colMatrixArray is not None:
            self.colMatrix = await _upload(self.clContext, self.colMatrixArray, "none", "loader colMatrix")
        self.gamutMatrix = await _upload(self.clContext, self.gamutMatrixArray, "none", "loader gamutMatrix")

    def addRefs(self) -> None:
        for b in (self.gammaLut, self.colMatrix, self.gamutMatrix):
            if b: b.addRef()

    def releaseRefs(self) -> None:
        for b in (self.gammaLut, self.colMatrix, self.gamutMatrix):
            if b: b.release()

    def run(self, params: Dict[str, Any], id_, cb: Callable[[], None]) -> None:
        if self.program is None:
            raise RuntimeError("Loader.run failed with no program available")
        self.addRefs()
        kernelParams = self.packImpl.getKernelParams(params)
        kernelParams["gammaLut"] = self.gammaLut
        kernelParams["gamutMatrix"] = self.gamutMatrix
        if self.colMatrix: kernelParams["colMatrix"] = self.colMatrix

        def done() -> None:
            self.releaseRefs()
            cb()
        self.clJobs.add(id_, self.packImpl.getName(), self.program, kernelParams, done)


class Saver(Packer):   # loadSave.ts:130-201
    def __init__(self, clContext_: clContext, colSpec: str, packImpl: PackImpl, clJobs: ClJobs):
        super().__init__(clContext_, packImpl, clJobs)
        self.gammaArray = linear2gammaLUT(colSpec)
        self.colMatrixArray: Optional[np.ndarray] = None
        if not self.packImpl.getIsRGB():
            self.colMatrixArray = matrixFlatten(rgb2ycbcrMatrix(colSpec, packImpl.numBits, packImpl.lumaBlack,
                                                                packImpl.lumaWhite, packImpl.chromaRange))
        self.gammaLut: Optional[OpenCLBuffer] = None
        self.colMatrix: Optional[OpenCLBuffer] = None

    async def init(self) -> None:
        await super().init()
        self.gammaLut = await _upload(self.clContext, self.gammaArray, "coarse", "saver gammaLut")
        if self.colMatrixArray is not None:
            self.colMatrix = await _upload(self.clContext, self.colMatrixArray, "none", "saver colMatrix")

    def addRefs(self) -> None:
        for b in (self.gammaLut, self.colMatrix):
            if b: b.addRef()

    def releaseRefs(self) -> None:
        for b in (self.gammaLut, self.colMatrix):
            if b: b.release()

    def run(self, params: Dict[str, Any], id_, cb: Callable[[], None]) -> None:
        if self.program is None:
            raise RuntimeError("Saver.run failed with no program available")
        self.addRefs()
        kernelParams = self.packImpl.getKernelParams(params)
        kernelParams["gammaLut"] = self.gammaLut
        if self.colMatrix: kernelParams["colMatrix"] = self.colMatrix

        def done() -> None:
            self.releaseRefs()
            cb()
        self.clJobs.add(id_, self.packImpl.getName(), self.program, kernelParams, done)
