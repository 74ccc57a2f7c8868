"""src/process/transition.ts: dissolve / wipe transitions."""
from __future__ import annotations

from typing import Any, Dict

from ..nodencl import KernelSpec
from .image_process import ProcessImpl


class Transition(ProcessImpl):   # transition.ts:83-116
    def __init__(self, type_: str, width: int, height: int):
        if type_ not in ("dissolve", "wipe"):
            raise RuntimeError(f"Transition requires a 'type' parameter that is either 'dissolve' or 'wipe' - found '{type_}'")
        super().__init__(type_, width, height, KernelSpec("dissolve" if type_ == "dissolve" else "wipe_mask"),
                         f"transition_{type_}")

    async def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        kernelParams: Dict[str, Any] = {"output": params["output"]}
        inArray = params["inputs"]
        if len(inArray) != 2:
            raise RuntimeError("Transition requires an 'inputs' array parameter with 2 OpenCL buffers")
        for i, b in enumerate(inArray):
            kernelParams[f"input{i}"] = b
        if self.name == "dissolve":
            kernelParams["mix"] = params["mix"]
        elif params.get("mask") is not None:
            kernelParams["maskIn"] = params["mask"]
        else:
            raise RuntimeError(f"Transition '{self.name}' expected a 'mask' buffer which wasn't found")
        return kernelParams
