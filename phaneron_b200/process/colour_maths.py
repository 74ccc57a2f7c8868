"""src/process/colourMaths.ts exports, computed by the library's host code
(phaneron_b200/csrc/pb_colour.cpp) so every host language feeds the kernels the same bits."""
from __future__ import annotations

import numpy as np

from .. import _lib


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


def gamma2linearLUT(colSpec: str) -> np.ndarray:
    out = np.empty(65536, np.float32)
    if not _lib.lib().pb_gamma2linear_lut(colSpec.encode(), _p(out)):
        print(f"Unrecognised colourspace {colSpec} - defaulting to BT.709")
    return out


def linear2gammaLUT(colSpec: str) -> np.ndarray:
    out = np.empty(65536, np.float32)
    if not _lib.lib().pb_linear2gamma_lut(colSpec.encode(), _p(out)):
        print(f"Unrecognised colourspace {colSpec} - defaulting to BT.709")
    return out


def ycbcr2rgbMatrix(colSpec: str, numBits: int, lumaBlack: int, lumaWhite: int, chrRange: int) -> np.ndarray:
    out = np.empty((3, 4), np.float32)
    _lib.lib().pb_ycbcr2rgb_matrix(colSpec.encode(), numBits, lumaBlack, lumaWhite, chrRange, _p(out))
    return out


def rgb2ycbcrMatrix(colSpec: str, numBits: int, lumaBlack: int, lumaWhite: int, chrRange: int) -> np.ndarray:
    out = np.empty((3, 4), np.float32)
    _lib.lib().pb_rgb2ycbcr_matrix(colSpec.encode(), numBits, lumaBlack, lumaWhite, chrRange, _p(out))
    return out


def rgb2rgbMatrix(srcColSpec: str, dstColSpec: str) -> np.ndarray:
    out = np.empty((3, 3), np.float32)
    _lib.lib().pb_rgb2rgb_matrix(srcColSpec.encode(), dstColSpec.encode(), _p(out))
    return out


def matrixFlatten(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, np.float32).reshape(-1)


def transformMatrix(width, height, flipH, flipV, anchorX, anchorY, scaleX, scaleY, offsetX, offsetY, rotate) -> np.ndarray:
    out = np.empty((3, 3), np.float32)
    _lib.check(_lib.lib().pb_transform_matrix(int(width), int(height), int(bool(flipH)), int(bool(flipV)), float(anchorX),
                                              float(anchorY), float(scaleX), float(scaleY), float(offsetX), float(offsetY),
                                              float(rotate), _p(out)))
    return out
