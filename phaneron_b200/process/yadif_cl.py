"""src/process/yadifCl.ts: the YADIF de-interlace kernel wrapper."""
from __future__ import annotations

from typing import Any, Dict

from ..nodencl import KernelSpec
from .image_process import ProcessImpl


class YadifCl(ProcessImpl):   # yadifCl.ts:170-194
    def __init__(self, width: int, height: int):
        super().__init__("yadif", width, height, KernelSpec("yadif"), "yadif")

    async def getKernelParams(self, params: Dict[str, Any]) -> Dict[str, Any]:
        return {"prev": params["prev"], "cur": params["cur"], "next": params["next"], "parity": params["parity"],
                "tff": 1 if params.get("tff") else 0, "skipSpatial": 1 if params.get("skipSpatial") else 0,
                "output": params["output"]}
