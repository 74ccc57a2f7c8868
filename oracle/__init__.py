"""ctypes wrapper around oracle/liboracle.so -- the CPU restatement of the
reference's pixel path (see oracle/oracle.h).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(phaneron_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def use_native() -> bool:
    """bench.py's CPU-baseline legs: switch this process to a -O3 -march=native build made on THIS machine
    (oracle/_native/, git-ignored).  Same source, same -ffp-contract=off: same results, the host's full ISA."""
    global _lib
    path = os.path.join(_HERE, "_native", "liboracle_native.so")
    try:
        subprocess.check_call(["make", "-C", _HERE, "-B", "_native/liboracle_native.so"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        l = C.CDLL(path)
        _declare(l)
        _lib = l
        return True
    except Exception:
        return False


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def _declare(l: C.CDLL) -> None:
    l.orc_set_threads.argtypes = [C.c_int]
    l.orc_get_threads.restype = C.c_int
    l.orc_gamma2linear_lut.argtypes = [C.c_char_p, _f32p]
    l.orc_linear2gamma_lut.argtypes = [C.c_char_p, _f32p]
    for f in (l.orc_ycbcr2rgb_matrix, l.orc_rgb2ycbcr_matrix):
        f.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
    l.orc_rgb2rgb_matrix.argtypes = [C.c_char_p, C.c_char_p, _f32p]
    l.orc_transform_matrix.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_double] * 7 + [_f32p]
    l.orc_v210_pitch.argtypes = [C.c_uint32]
    l.orc_v210_pitch.restype = C.c_uint32
    l.orc_v210_pitch_bytes.argtypes = [C.c_uint32]
    l.orc_v210_pitch_bytes.restype = C.c_uint32
    l.orc_v210_fill.argtypes = [_u8p, C.c_uint32, C.c_uint32]
    l.orc_v210_read.argtypes = [_u32p, _f32p, C.c_uint32, C.c_uint32, _f32p, _f32p, _f32p]
    l.orc_v210_write.argtypes = [_f32p, _u32p, C.c_uint32, C.c_uint32, C.c_uint32, _f32p, _f32p, C.c_int]
    l.orc_combine.argtypes = [C.POINTER(C.c_void_p), C.c_int, _f32p, C.c_int, C.c_int]
    l.orc_dissolve.argtypes = [_f32p, _f32p, C.c_float, _f32p, C.c_int, C.c_int]
    l.orc_mix.argtypes = [_f32p, _f32p, C.c_float, _f32p, C.c_int, C.c_int]
    l.orc_wipe.argtypes = [_f32p, _f32p, C.c_float, _f32p, C.c_int, C.c_int]
    l.orc_wipe_mask.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int]
    l.orc_transform.argtypes = [_f32p, C.c_int, C.c_int, _f32p, _f32p, C.c_int, C.c_int]
    l.orc_resize.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, _f32p, _f32p, C.c_int, C.c_int]
    l.orc_yadif.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int]
    l.orc_rgba8_read.argtypes = [_u8p, _f32p, C.c_uint32, C.c_uint32, _f32p, _f32p, C.c_int]
    l.orc_rgba8_write.argtypes = [_f32p, _u8p, C.c_uint32, C.c_uint32, C.c_uint32, _f32p, C.c_int]


def set_threads(n: int) -> None:
    lib().orc_set_threads(int(n))


# --- colourMaths.ts -------------------------------------------------------
def gamma2linear_lut(colspec: str) -> np.ndarray:
    out = np.empty(65536, np.float32)
    lib().orc_gamma2linear_lut(colspec.encode(), out)
    return out


def linear2gamma_lut(colspec: str) -> np.ndarray:
    out = np.empty(65536, np.float32)
    lib().orc_linear2gamma_lut(colspec.encode(), out)
    return out


def ycbcr2rgb_matrix(colspec: str, num_bits=10, luma_black=64, luma_white=940, chr_range=896) -> np.ndarray:
    out = np.empty(12, np.float32)
    lib().orc_ycbcr2rgb_matrix(colspec.encode(), num_bits, luma_black, luma_white, chr_range, out)
    return out.reshape(3, 4)


def rgb2ycbcr_matrix(colspec: str, num_bits=10, luma_black=64, luma_white=940, chr_range=896) -> np.ndarray:
    out = np.empty(12, np.float32)
    lib().orc_rgb2ycbcr_matrix(colspec.encode(), num_bits, luma_black, luma_white, chr_range, out)
    return out.reshape(3, 4)


def rgb2rgb_matrix(src: str, dst: str) -> np.ndarray:
    out = np.empty(9, np.float32)
    lib().orc_rgb2rgb_matrix(src.encode(), dst.encode(), out)
    return out.reshape(3, 3)


def transform_matrix(width, height, flip_h=False, flip_v=False, anchor_x=0.0, anchor_y=0.0, scale_x=1.0,
                     scale_y=1.0, offset_x=0.0, offset_y=0.0, rotate=0.0) -> np.ndarray:
    out = np.empty(9, np.float32)
    lib().orc_transform_matrix(width, height, int(flip_h), int(flip_v), anchor_x, anchor_y, scale_x, scale_y,
                               offset_x, offset_y, rotate, out)
    return out.reshape(3, 3)


# --- v210 -----------------------------------------------------------------
def v210_pitch_bytes(width: int) -> int:
    return int(lib().orc_v210_pitch_bytes(width))


def v210_fill(width: int, height: int) -> np.ndarray:
    buf = np.empty(v210_pitch_bytes(width) * height, np.uint8)
    lib().orc_v210_fill(buf, width, height)
    return buf


def v210_read(v210: np.ndarray, width: int, height: int, col_matrix, gamma_lut, gamut) -> np.ndarray:
    src = np.ascontiguousarray(v210).view(np.uint32)
    out = np.empty((height, width, 4), np.float32)
    lib().orc_v210_read(src, out, width, height, _f(col_matrix), _f(gamma_lut), _f(gamut))
    return out


def v210_write(rgba: np.ndarray, width: int, height: int, interlace: int, col_matrix, gamma_lut,
               q3_literal: bool = False, out: np.ndarray | None = None) -> np.ndarray:
    if out is None:
        out = np.zeros(v210_pitch_bytes(width) * height, np.uint8)
    lib().orc_v210_write(_f(rgba), out.view(np.uint32), width, height, interlace, _f(col_matrix),
                         _f(gamma_lut), int(q3_literal))
    return out


# --- image ops ------------------------------------------------------------
def _f(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def combine(layers) -> np.ndarray:
    layers = [_f(l) for l in layers]
    h, w, _ = layers[0].shape
    ptrs = (C.c_void_p * len(layers))(*[l.ctypes.data for l in layers])
    out = np.empty((h, w, 4), np.float32)
    lib().orc_combine(ptrs, len(layers), out, w, h)
    return out


def dissolve(in0, in1, mix: float) -> np.ndarray:
    in0, in1 = _f(in0), _f(in1)
    h, w, _ = in0.shape
    out = np.empty_like(in0)
    lib().orc_dissolve(in0, in1, mix, out, w, h)
    return out


def mix(in0, in1, m: float) -> np.ndarray:
    in0, in1 = _f(in0), _f(in1)
    h, w, _ = in0.shape
    out = np.empty_like(in0)
    lib().orc_mix(in0, in1, m, out, w, h)
    return out


def wipe(in0, in1, wipe_: float) -> np.ndarray:
    in0, in1 = _f(in0), _f(in1)
    h, w, _ = in0.shape
    out = np.empty_like(in0)
    lib().orc_wipe(in0, in1, wipe_, out, w, h)
    return out


def wipe_mask(in0, in1, mask) -> np.ndarray:
    in0, in1, mask = _f(in0), _f(in1), _f(mask)
    h, w, _ = in0.shape
    out = np.empty_like(in0)
    lib().orc_wipe_mask(in0, in1, mask, out, w, h)
    return out


def transform(img, mat, out_w: int, out_h: int) -> np.ndarray:
    img = _f(img)
    sh, sw, _ = img.shape
    out = np.empty((out_h, out_w, 4), np.float32)
    lib().orc_transform(img, sw, sh, _f(mat).reshape(-1), out, out_w, out_h)
    return out


def resize(img, scale, offset_x, offset_y, flip4, out_w: int, out_h: int) -> np.ndarray:
    img = _f(img)
    sh, sw, _ = img.shape
    out = np.empty((out_h, out_w, 4), np.float32)
    lib().orc_resize(img, sw, sh, scale, offset_x, offset_y, _f(flip4), out, out_w, out_h)
    return out


def yadif(prev, cur, nxt, parity: int, tff: bool, skip_spatial: bool) -> np.ndarray:
    prev, cur, nxt = _f(prev), _f(cur), _f(nxt)
    h, w, _ = cur.shape
    out = np.empty_like(cur)
    lib().orc_yadif(prev, cur, nxt, parity, int(tff), int(skip_spatial), out, w, h)
    return out


def rgba8_read(buf, width, height, gamma_lut, gamut, bgra=False) -> np.ndarray:
    out = np.empty((height, width, 4), np.float32)
    lib().orc_rgba8_read(np.ascontiguousarray(buf, np.uint8), out, width, height, _f(gamma_lut), _f(gamut), int(bgra))
    return out


def rgba8_write(rgba, width, height, interlace, gamma_lut, bgra=False, out=None) -> np.ndarray:
    if out is None:
        out = np.zeros(width * height * 4, np.uint8)
    lib().orc_rgba8_write(_f(rgba), out, width, height, interlace, _f(gamma_lut), int(bgra))
    return out


# ---- yuv422p10le / yuv422p8 (yuv422p10.ts, yuv422p8.ts) -------------------------------------------------------
def yuv422p_pitch(width: int) -> int:
    return int(lib().orc_yuv422p_pitch(width))


def yuv422p_plane_bytes(bits: int, width: int, height: int):
    luma = yuv422p_pitch(width) * (1 if bits == 8 else 2) * height
    return [luma, luma // 2, luma // 2]


def yuv422p_fill(bits: int, width: int, height: int) -> np.ndarray:
    buf = np.zeros(sum(yuv422p_plane_bytes(bits, width, height)), np.uint8)
    lib().orc_yuv422p_fill(bits, buf.ctypes.data_as(C.c_void_p), width, height)
    return buf


def yuv422p_read(bits: int, y, u, v, width: int, height: int, col_matrix, gamma_lut, gamut) -> np.ndarray:
    out = np.empty((height, width, 4), np.float32)
    y, u, v = (np.ascontiguousarray(a, np.uint8) for a in (y, u, v))
    lib().orc_yuv422p_read(bits, y.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p),
                           out.ctypes.data_as(C.c_void_p), width, height, _f(col_matrix).ctypes.data_as(C.c_void_p),
                           _f(gamma_lut).ctypes.data_as(C.c_void_p), _f(gamut).ctypes.data_as(C.c_void_p))
    return out


def yuv422p_write(bits: int, rgba, width: int, height: int, interlace: int, col_matrix, gamma_lut, outs=None):
    nb = yuv422p_plane_bytes(bits, width, height)
    if outs is None:
        outs = [np.zeros(n, np.uint8) for n in nb]
    rgba = _f(rgba)
    lib().orc_yuv422p_write(bits, rgba.ctypes.data_as(C.c_void_p), *(o.ctypes.data_as(C.c_void_p) for o in outs), width, height, interlace,
                            _f(col_matrix).ctypes.data_as(C.c_void_p), _f(gamma_lut).ctypes.data_as(C.c_void_p))
    return outs


# ---- yuv420p / nv12 (yuv420p.ts, nv12.ts) ------------------------------------------------------------------------
def yuv420_plane_bytes(nv12: bool, width: int, height: int):
    """Reader.numBytes: [luma, luma/4, luma/4] (yuv420p.ts:332-333) or [luma, luma/2] (nv12.ts:328-329)"""
    luma = yuv422p_pitch(width) * height
    return [luma, luma // 2] if nv12 else [luma, luma // 4, luma // 4]


def yuv420_fill(nv12: bool, width: int, height: int) -> np.ndarray:
    buf = np.zeros(sum(yuv420_plane_bytes(nv12, width, height)), np.uint8)
    lib().orc_yuv420_fill(int(nv12), buf.ctypes.data_as(C.c_void_p), width, height)
    return buf


def yuv420_read(nv12: bool, planes, width: int, height: int, col_matrix, gamma_lut, gamut) -> np.ndarray:
    out = np.empty((height, width, 4), np.float32)
    planes = [np.ascontiguousarray(a, np.uint8) for a in planes]
    ptrs = [a.ctypes.data_as(C.c_void_p) for a in planes] + ([C.c_void_p(0)] if nv12 else [])
    lib().orc_yuv420_read(int(nv12), *ptrs, out.ctypes.data_as(C.c_void_p), width, height, _f(col_matrix).ctypes.data_as(C.c_void_p),
                          _f(gamma_lut).ctypes.data_as(C.c_void_p), _f(gamut).ctypes.data_as(C.c_void_p))
    return out


def yuv420_write(nv12: bool, rgba, width: int, height: int, interlace: int, col_matrix, gamma_lut, outs=None):
    nb = yuv420_plane_bytes(nv12, width, height)
    if outs is None:
        outs = [np.zeros(n, np.uint8) for n in nb]
    rgba = _f(rgba)
    ptrs = [o.ctypes.data_as(C.c_void_p) for o in outs] + ([C.c_void_p(0)] if nv12 else [])
    lib().orc_yuv420_write(int(nv12), rgba.ctypes.data_as(C.c_void_p), *ptrs, width, height, interlace,
                           _f(col_matrix).ctypes.data_as(C.c_void_p), _f(gamma_lut).ctypes.data_as(C.c_void_p))
    return outs


# ---- Lanczos Transform filter (not in the reference; definition in oracle.c) -----------------------------------------
def transform_lanczos(img, mat, out_w: int, out_h: int, lobes: int = 3) -> np.ndarray:
    img = _f(img)
    sh, sw = img.shape[:2]
    out = np.empty((out_h, out_w, 4), np.float32)
    rc = lib().orc_transform_lanczos(img.ctypes.data_as(C.c_void_p), sw, sh, _f(mat).ctypes.data_as(C.c_void_p), int(lobes),
                                     out.ctypes.data_as(C.c_void_p), out_w, out_h)
    if rc != 0:
        raise ValueError("lanczos: axis-aligned transforms with at most 64 taps per axis only")
    return out
