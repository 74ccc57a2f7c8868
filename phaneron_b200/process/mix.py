"""src/process/mix.ts (dead code in the reference, named by the north star)."""
from __future__ import annotations

from typing import Any, Dict

from ..nodencl import KernelSpec
from .image_process import ProcessImpl


class Mix(ProcessImpl):   # mix.ts:48-69
    def __init__(self, width: int, height: int):
        super().__init__("mixer", width, height, KernelSpec("mix"), "mixer")

    async def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        return {"input0": params["input0"], "input1": params["input1"], "mix": params["mix"], "output": params["output"]}
