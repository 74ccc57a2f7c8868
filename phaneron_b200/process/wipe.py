"""src/process/wipe.ts (dead code in the reference, named by the north star)."""
from __future__ import annotations

from typing import Any, Dict

from ..nodencl import KernelSpec
from .image_process import ProcessImpl


class Wipe(ProcessImpl):   # wipe.ts:49-70
    def __init__(self, width: int, height: int):
        super().__init__("wipe", width, height, KernelSpec("wipe"), "wipe")

    async def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        return {"input0": params["input0"], "input1": params["input1"], "wipe": params["wipe"], "output": params["output"]}
