"""src/process/combine.ts: N-layer premultiplied 'over' (combine_N)."""
from __future__ import annotations

from typing import Any, Dict

from ..nodencl import KernelSpec
from .image_process import ProcessImpl


class Combine(ProcessImpl):   # combine.ts:70-104
    def __init__(self, numLayers: int, width: int, height: int):
        n = 2 if numLayers < 2 else numLayers   # combine will not actually be used if numLayers < 2
        super().__init__(f"combine-{numLayers}", width, height, KernelSpec("combine"), f"combine_{n}")

    async def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        kernelParams: Dict[str, Any] = {"output": params["output"]}
        inArray = params["inputs"]
        if len(inArray) < 2:
            raise RuntimeError("Combine requires an 'inputs' array parameter with at least 2 OpenCL buffers")
        for i, b in enumerate(inArray):
            kernelParams[f"l{i}In"] = b
        return kernelParams
