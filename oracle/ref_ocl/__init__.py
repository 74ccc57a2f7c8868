"""TEST INFRASTRUCTURE: the reference's own OpenCL kernels, run on the NVIDIA OpenCL driver of the GPU box.

`oracle/_ref/kernels/*.cl` are the kernel strings of /root/reference/src/process/*.ts, extracted verbatim by
extract_kernels.py (git-ignored, shipped to the GPU box); `oracle/_ref/libocl_ref.so` is ocl_runner.c.
Each function below launches a kernel with the NDRange and argument list its reference wrapper uses
(file:line cited), so its output IS the reference's output on this device.  tests/test_gpu_reference_opencl.py
compares the CPU oracle and the CUDA path against it.  Nothing under phaneron_b200/ imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(os.path.dirname(HERE), "_ref")
LIB = os.path.join(REF_DIR, "libocl_ref.so")
KERNELS = os.path.join(REF_DIR, "kernels")

_lib = None
_kernels = {}


def build() -> str:
    """gcc ocl_runner.c -> oracle/_ref/libocl_ref.so, and (when /root/reference is present) refresh the kernel strings"""
    os.makedirs(REF_DIR, exist_ok=True)
    src = os.path.join(HERE, "ocl_runner.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.run([cc, "-O2", "-std=c99", "-fPIC", "-shared", "-o", LIB, src, "-ldl"], check=True)
    if os.path.isdir("/root/reference/src/process"):
        subprocess.run(["python", os.path.join(HERE, "extract_kernels.py"), "/root/reference"], check=True, stdout=subprocess.DEVNULL)
    return LIB


def available() -> bool:
    """True when the runner is built, the kernel strings are present and an NVIDIA OpenCL device answers"""
    global _lib
    if _lib is not None:
        return True
    if not (os.path.exists(LIB) and os.path.isdir(KERNELS)):
        return False
    l = C.CDLL(LIB)
    l.ocl_log.restype = C.c_char_p
    l.ocl_device_name.restype = C.c_char_p
    l.ocl_kernel.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    l.ocl_buffer.argtypes = [C.c_size_t, C.c_void_p]
    l.ocl_image.argtypes = [C.c_int, C.c_int, C.c_void_p]
    l.ocl_read_buffer.argtypes = [C.c_int, C.c_void_p, C.c_size_t]
    l.ocl_read_image.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int]
    l.ocl_arg_u32.argtypes = [C.c_int, C.c_int, C.c_uint32]
    l.ocl_arg_f32.argtypes = [C.c_int, C.c_int, C.c_float]
    l.ocl_run.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t]
    if l.ocl_init() != 0:
        return False
    _lib = l
    return True


def why_unavailable() -> str:
    if not os.path.exists(LIB):
        return f"{LIB} not built"
    if not os.path.isdir(KERNELS):
        return f"{KERNELS} missing (run oracle/ref_ocl/extract_kernels.py where /root/reference exists)"
    l = C.CDLL(LIB)
    l.ocl_log.restype = C.c_char_p
    l.ocl_init()
    return l.ocl_log().decode()


def device_name() -> str:
    return _lib.ocl_device_name().decode()


def _ck(rc):
    if rc < 0:
        raise RuntimeError(_lib.ocl_log().decode())
    return rc


def _kernel(file: str, entry: str, options: str = "") -> int:
    key = (file, entry, options)
    if key not in _kernels:
        with open(os.path.join(KERNELS, file)) as f:
            src = f.read()
        _kernels[key] = _ck(_lib.ocl_kernel(src.encode(), entry.encode(), options.encode()))
    return _kernels[key]


def program_text(file: str, entry: str, options: str = "") -> str:
    """what the driver compiled the kernel to (PTX text on NVIDIA): the ground truth for dot()/fma contraction"""
    k = _kernel(file, entry, options)
    _lib.ocl_program_binary.restype = C.c_long
    _lib.ocl_program_binary.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
    buf = C.create_string_buffer(1 << 20)
    n = _lib.ocl_program_binary(k, buf, len(buf))
    if n < 0:
        raise RuntimeError(_lib.ocl_log().decode())
    return buf.raw[:n].decode(errors="replace")


def _buf(arr: Optional[np.ndarray] = None, nbytes: int = 0) -> int:
    if arr is not None:
        a = np.ascontiguousarray(arr)
        return _ck(_lib.ocl_buffer(a.nbytes, a.ctypes.data))
    return _ck(_lib.ocl_buffer(nbytes, None))


def _img(w: int, h: int, arr: Optional[np.ndarray] = None) -> int:
    if arr is not None:
        a = np.ascontiguousarray(arr, np.float32)
        assert a.size == w * h * 4
        return _ck(_lib.ocl_image(w, h, a.ctypes.data))
    return _ck(_lib.ocl_image(w, h, None))


def _read_img(m: int, w: int, h: int) -> np.ndarray:
    out = np.empty((h, w, 4), np.float32)
    _ck(_lib.ocl_read_image(m, out.ctypes.data, w, h))
    return out


def _free(*mems):
    for m in mems:
        _lib.ocl_release(m)


def _pad(a, n):
    flat = np.asarray(a, np.float32).reshape(-1)
    out = np.zeros(n, np.float32)
    out[: flat.size] = flat
    return out


def _pitch(width: int) -> int:   # v210.ts:198-204, pixels
    return ((width + 47) // 48) * 48


def v210_read(src: np.ndarray, width: int, height: int, col_matrix, gamma_lut, gamut, options: str = "") -> np.ndarray:
    """v210.ts:25-111 with Reader's NDRange (v210.ts:293-294): one work-group per line"""
    k = _kernel("v210.cl", "read", options)
    wpg = _pitch(width) // 48
    i, o = _buf(src), _buf(nbytes=width * height * 16)
    cm, lut, gm = _buf(_pad(col_matrix, 12)), _buf(np.asarray(gamma_lut, np.float32)), _buf(_pad(gamut, 16))   # Q4: 3 float4 are read
    for n, m in enumerate((i, o)):
        _ck(_lib.ocl_arg_mem(k, n, m))
    _ck(_lib.ocl_arg_u32(k, 2, width))
    for n, m in zip((3, 4, 5), (cm, lut, gm)):
        _ck(_lib.ocl_arg_mem(k, n, m))
    _ck(_lib.ocl_run(k, 1, wpg * height, 1, wpg))
    out = np.empty((height, width, 4), np.float32)
    _ck(_lib.ocl_read_buffer(o, out.ctypes.data, out.nbytes))
    _free(i, o, cm, lut, gm)
    return out


def v210_write(rgba: np.ndarray, width: int, height: int, interlace: int, col_matrix, gamma_lut, out: Optional[np.ndarray] = None,
               options: str = "") -> np.ndarray:
    """v210.ts:113-195 with Writer's NDRange (v210.ts:322-323)"""
    k = _kernel("v210.cl", "write", options)
    wpg = _pitch(width) // 48
    nbytes = wpg * 128 * height
    if out is None:
        out = np.zeros(nbytes, np.uint8)
    i, o = _buf(np.asarray(rgba, np.float32)), _buf(out)
    cm, lut = _buf(_pad(col_matrix, 12)), _buf(np.asarray(gamma_lut, np.float32))
    _ck(_lib.ocl_arg_mem(k, 0, i))
    _ck(_lib.ocl_arg_mem(k, 1, o))
    _ck(_lib.ocl_arg_u32(k, 2, width))
    _ck(_lib.ocl_arg_u32(k, 3, interlace))
    _ck(_lib.ocl_arg_mem(k, 4, cm))
    _ck(_lib.ocl_arg_mem(k, 5, lut))
    _ck(_lib.ocl_run(k, 1, wpg * height // (2 if interlace else 1), 1, wpg))
    _ck(_lib.ocl_read_buffer(o, out.ctypes.data, nbytes))
    _free(i, o, cm, lut)
    return out


def _image_op(file: str, entry: str, images: List[np.ndarray], w: int, h: int, scalar: Optional[float] = None) -> np.ndarray:
    """ProcessImpl NDRange = [width, height] (imageProcess.ts:48-50); inputs, [float], output"""
    k = _kernel(file, entry)
    mems = [_img(im.shape[1], im.shape[0], im) for im in images]
    o = _img(w, h)
    n = 0
    for m in mems:
        _ck(_lib.ocl_arg_mem(k, n, m))
        n += 1
    if scalar is not None:
        _ck(_lib.ocl_arg_f32(k, n, scalar))
        n += 1
    _ck(_lib.ocl_arg_mem(k, n, o))
    _ck(_lib.ocl_run(k, 2, w, h, 0))
    out = _read_img(o, w, h)
    _free(o, *mems)
    return out


def combine(layers: List[np.ndarray]) -> np.ndarray:
    """combine.ts:24-68"""
    h, w, _ = layers[0].shape
    return _image_op(f"combine_{len(layers)}.cl", f"combine_{len(layers)}", layers, w, h)


def dissolve(in0: np.ndarray, in1: np.ndarray, mix: float) -> np.ndarray:
    """transition.ts:60-65"""
    h, w, _ = in0.shape
    return _image_op("transition_dissolve.cl", "transition_dissolve", [in0, in1], w, h, scalar=mix)


def wipe_mask(in0: np.ndarray, in1: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """transition.ts:66-73"""
    h, w, _ = in0.shape
    return _image_op("transition_wipe.cl", "transition_wipe", [in0, in1, mask], w, h)


def mix(in0: np.ndarray, in1: np.ndarray, m: float) -> np.ndarray:
    """mix.ts:24-47 (entry 'mixer')"""
    h, w, _ = in0.shape
    return _image_op("mix.cl", "mixer", [in0, in1], w, h, scalar=m)


def wipe(in0: np.ndarray, in1: np.ndarray, wipe_: float) -> np.ndarray:
    """wipe.ts:24-48"""
    h, w, _ = in0.shape
    return _image_op("wipe.cl", "wipe", [in0, in1], w, h, scalar=wipe_)


def resize(src: np.ndarray, scale: float, offset_x: float, offset_y: float, flip4, w: int, h: int) -> np.ndarray:
    """resize.ts:24-58: (input image, scale, offsetX, offsetY, flip buffer, output image)"""
    k = _kernel("resize.cl", "resize")
    i, o, f = _img(src.shape[1], src.shape[0], src), _img(w, h), _buf(np.asarray(flip4, np.float32))
    _ck(_lib.ocl_arg_mem(k, 0, i))
    _ck(_lib.ocl_arg_f32(k, 1, scale))
    _ck(_lib.ocl_arg_f32(k, 2, offset_x))
    _ck(_lib.ocl_arg_f32(k, 3, offset_y))
    _ck(_lib.ocl_arg_mem(k, 4, f))
    _ck(_lib.ocl_arg_mem(k, 5, o))
    _ck(_lib.ocl_run(k, 2, w, h, 0))
    out = _read_img(o, w, h)
    _free(i, o, f)
    return out


def transform(src: np.ndarray, mat9, w: int, h: int) -> np.ndarray:
    """transform.ts:36-59: (input image, transformMatrix buffer, output image)"""
    k = _kernel("transform.cl", "transform")
    i, o, m = _img(src.shape[1], src.shape[0], src), _img(w, h), _buf(_pad(mat9, 12))
    _ck(_lib.ocl_arg_mem(k, 0, i))
    _ck(_lib.ocl_arg_mem(k, 1, m))
    _ck(_lib.ocl_arg_mem(k, 2, o))
    _ck(_lib.ocl_run(k, 2, w, h, 0))
    out = _read_img(o, w, h)
    _free(i, o, m)
    return out


def yadif(prev: np.ndarray, cur: np.ndarray, nxt: np.ndarray, parity: int, tff: bool, skip_spatial: bool) -> np.ndarray:
    """yadifCl.ts:105-167 with YadifCL's parameter list (yadifCl.ts:178-188)"""
    h, w, _ = cur.shape
    k = _kernel("yadif.cl", "yadif")
    mems = [_img(w, h, im) for im in (prev, cur, nxt)]
    o = _img(w, h)
    for n, m in enumerate(mems):
        _ck(_lib.ocl_arg_mem(k, n, m))
    _ck(_lib.ocl_arg_u32(k, 3, int(parity)))
    _ck(_lib.ocl_arg_u32(k, 4, 1 if tff else 0))
    _ck(_lib.ocl_arg_u32(k, 5, 1 if skip_spatial else 0))
    _ck(_lib.ocl_arg_mem(k, 6, o))
    _ck(_lib.ocl_run(k, 2, w, h, 0))
    out = _read_img(o, w, h)
    _free(o, *mems)
    return out


def rgba8_read(src: np.ndarray, width: int, height: int, gamma_lut, gamut) -> np.ndarray:
    """rgba8.ts:25-67 with Reader's NDRange (rgba8.ts:171-173): 64 pixels per work-item, one group per line"""
    assert width % 64 == 0, "Q13: the reference's work-group size is width/64 without a ceil"
    k = _kernel("rgba8.cl", "read")
    wpg = width // 64
    i, o = _buf(np.asarray(src, np.uint8)), _buf(nbytes=width * height * 16)
    lut, gm = _buf(np.asarray(gamma_lut, np.float32)), _buf(_pad(gamut, 16))
    _ck(_lib.ocl_arg_mem(k, 0, i)); _ck(_lib.ocl_arg_mem(k, 1, o)); _ck(_lib.ocl_arg_u32(k, 2, width))
    _ck(_lib.ocl_arg_mem(k, 3, lut)); _ck(_lib.ocl_arg_mem(k, 4, gm))
    _ck(_lib.ocl_run(k, 1, wpg * height, 1, wpg))
    out = np.empty((height, width, 4), np.float32)
    _ck(_lib.ocl_read_buffer(o, out.ctypes.data, out.nbytes))
    _free(i, o, lut, gm)
    return out


def rgba8_write(rgba: np.ndarray, width: int, height: int, interlace: int, gamma_lut) -> np.ndarray:
    """rgba8.ts:69-103 with Writer's NDRange (rgba8.ts:195-197)"""
    assert width % 64 == 0
    k = _kernel("rgba8.cl", "write")
    wpg = width // 64
    out = np.zeros(width * height * 4, np.uint8)
    i, o, lut = _buf(np.asarray(rgba, np.float32)), _buf(out), _buf(np.asarray(gamma_lut, np.float32))
    _ck(_lib.ocl_arg_mem(k, 0, i)); _ck(_lib.ocl_arg_mem(k, 1, o)); _ck(_lib.ocl_arg_u32(k, 2, width)); _ck(_lib.ocl_arg_u32(k, 3, interlace))
    _ck(_lib.ocl_arg_mem(k, 4, lut))
    _ck(_lib.ocl_run(k, 1, wpg * height // (2 if interlace else 1), 1, wpg))
    _ck(_lib.ocl_read_buffer(o, out.ctypes.data, out.nbytes))
    _free(i, o, lut)
    return out


def _yuv422p_geom(bits: int, width: int, height: int):
    pitch = width + 7 - ((width - 1) % 8)                      # yuv422p10.ts:222
    luma = pitch * (2 if bits == 10 else 1) * height
    return -(-pitch // 64), [luma, luma // 2, luma // 2]       # Math.ceil(getPitch / 64) work-items per line (:308)


def yuv422p_read(bits: int, y, u, v, width: int, height: int, col_matrix, gamma_lut, gamut) -> np.ndarray:
    """yuv422p10.ts:25-124 / yuv422p8.ts:25-124 with Reader's NDRange"""
    k = _kernel(f"yuv422p{bits}.cl", "read")
    wpg, _ = _yuv422p_geom(bits, width, height)
    mems = [_buf(np.asarray(a, np.uint8)) for a in (y, u, v)]
    o = _buf(nbytes=width * height * 16)
    cm, lut, gm = _buf(_pad(col_matrix, 12)), _buf(np.asarray(gamma_lut, np.float32)), _buf(_pad(gamut, 16))
    for n, m in enumerate(mems + [o]):
        _ck(_lib.ocl_arg_mem(k, n, m))
    _ck(_lib.ocl_arg_u32(k, 4, width))
    for n, m in zip((5, 6, 7), (cm, lut, gm)):
        _ck(_lib.ocl_arg_mem(k, n, m))
    _ck(_lib.ocl_run(k, 1, wpg * height, 1, wpg))
    out = np.empty((height, width, 4), np.float32)
    _ck(_lib.ocl_read_buffer(o, out.ctypes.data, out.nbytes))
    _free(o, cm, lut, gm, *mems)
    return out


def yuv422p_write(bits: int, rgba, width: int, height: int, interlace: int, col_matrix, gamma_lut, outs=None):
    """yuv422p10.ts:126-219 / yuv422p8.ts:126-219 with Writer's NDRange"""
    k = _kernel(f"yuv422p{bits}.cl", "write")
    wpg, nb = _yuv422p_geom(bits, width, height)
    if outs is None:
        outs = [np.zeros(n, np.uint8) for n in nb]
    i = _buf(np.asarray(rgba, np.float32))
    mems = [_buf(o) for o in outs]
    cm, lut = _buf(_pad(col_matrix, 12)), _buf(np.asarray(gamma_lut, np.float32))
    _ck(_lib.ocl_arg_mem(k, 0, i))
    for n, m in enumerate(mems):
        _ck(_lib.ocl_arg_mem(k, 1 + n, m))
    _ck(_lib.ocl_arg_u32(k, 4, width))
    _ck(_lib.ocl_arg_u32(k, 5, interlace))
    _ck(_lib.ocl_arg_mem(k, 6, cm))
    _ck(_lib.ocl_arg_mem(k, 7, lut))
    _ck(_lib.ocl_run(k, 1, wpg * height // (2 if interlace else 1), 1, wpg))
    for o, m in zip(outs, mems):
        _ck(_lib.ocl_read_buffer(m, o.ctypes.data, o.nbytes))
    _free(i, cm, lut, *mems)
    return outs


def _yuv420_geom(nv12: bool, width: int, height: int):
    pitch = width + 7 - ((width - 1) % 8)                      # yuv420p.ts:240
    luma = pitch * height
    return -(-pitch // 64), ([luma, luma // 2] if nv12 else [luma, luma // 4, luma // 4])


def yuv420_read(nv12: bool, planes, width: int, height: int, col_matrix, gamma_lut, gamut) -> np.ndarray:
    """yuv420p.ts:25-140 / nv12.ts:24-132 with Reader's NDRange (one work-group per line pair)"""
    k = _kernel("nv12.cl" if nv12 else "yuv420p.cl", "read")
    wpg, _ = _yuv420_geom(nv12, width, height)
    mems = [_buf(np.asarray(a, np.uint8)) for a in planes]
    o = _buf(nbytes=width * height * 16)
    cm, lut, gm = _buf(_pad(col_matrix, 12)), _buf(np.asarray(gamma_lut, np.float32)), _buf(_pad(gamut, 16))
    a = 0
    for m in mems + [o]:
        _ck(_lib.ocl_arg_mem(k, a, m)); a += 1
    _ck(_lib.ocl_arg_u32(k, a, width)); a += 1
    for m in (cm, lut, gm):
        _ck(_lib.ocl_arg_mem(k, a, m)); a += 1
    _ck(_lib.ocl_run(k, 1, wpg * height // 2, 1, wpg))
    out = np.empty((height, width, 4), np.float32)
    _ck(_lib.ocl_read_buffer(o, out.ctypes.data, out.nbytes))
    _free(o, cm, lut, gm, *mems)
    return out


def yuv420_write(nv12: bool, rgba, width: int, height: int, interlace: int, col_matrix, gamma_lut, outs=None):
    """yuv420p.ts:142-238 / nv12.ts:134-240 with Writer's NDRange (height / 2 work-groups, fields included)"""
    k = _kernel("nv12.cl" if nv12 else "yuv420p.cl", "write")
    wpg, nb = _yuv420_geom(nv12, width, height)
    if outs is None:
        outs = [np.zeros(n, np.uint8) for n in nb]
    i = _buf(np.asarray(rgba, np.float32))
    mems = [_buf(o) for o in outs]
    cm, lut = _buf(_pad(col_matrix, 12)), _buf(np.asarray(gamma_lut, np.float32))
    _ck(_lib.ocl_arg_mem(k, 0, i))
    a = 1
    for m in mems:
        _ck(_lib.ocl_arg_mem(k, a, m)); a += 1
    _ck(_lib.ocl_arg_u32(k, a, width)); a += 1
    _ck(_lib.ocl_arg_u32(k, a, interlace)); a += 1
    _ck(_lib.ocl_arg_mem(k, a, cm)); a += 1
    _ck(_lib.ocl_arg_mem(k, a, lut))
    _ck(_lib.ocl_run(k, 1, wpg * height // 2, 1, wpg))
    for o, m in zip(outs, mems):
        _ck(_lib.ocl_read_buffer(m, o.ctypes.data, o.nbytes))
    _free(i, cm, lut, *mems)
    return outs


class ReferenceChain:
    """The reference's UNFUSED launch sequence for a harness scene, with persistent device buffers, for timing on the
    same GPU: per source `read` (+ `transform`), per transition layer `transition_*`, `combine_N`, `write`
    (SURVEY.md 3.2-3.4).  Kernels are the reference's own; buffers stay resident, as phaneron's do between stages."""

    def __init__(self, scene, consts, xf_matrix):
        cm_r, lut_r, gamut, cm_w, lut_w = consts
        self.W, self.H = scene["width"], scene["height"]
        W, H = self.W, self.H
        self.launches = []   # (kernel id, dims, g0, g1, local0)
        self.cm_r, self.lut_r, self.gm = _buf(_pad(cm_r, 12)), _buf(np.asarray(lut_r, np.float32)), _buf(_pad(gamut, 16))
        self.cm_w, self.lut_w = _buf(_pad(cm_w, 12)), _buf(np.asarray(lut_w, np.float32))
        n_kernel = [0]

        def fresh(file, entry):   # one kernel object per launch site: arguments stay bound
            n_kernel[0] += 1
            return _kernel(file, entry, "-DPB_SITE=%d" % n_kernel[0])

        def source(src, sw, sh, xf):
            k = fresh("v210.cl", "read")
            wpg = _pitch(sw) // 48
            i, o = _buf(src), _buf(nbytes=sw * sh * 16)
            for n, m in enumerate((i, o)):
                _ck(_lib.ocl_arg_mem(k, n, m))
            _ck(_lib.ocl_arg_u32(k, 2, sw))
            for n, m in zip((3, 4, 5), (self.cm_r, self.lut_r, self.gm)):
                _ck(_lib.ocl_arg_mem(k, n, m))
            self.launches.append((k, 1, wpg * sh, 1, wpg))
            # nodencl's image view of the RGBA buffer: here an image2d_t filled by a buffer->image copy is not available
            # through the runner, so the read kernel's output buffer is re-uploaded once as an image at build time
            # and the read launch is still timed (its output buffer is what a packer-only chain would consume).
            rgba = np.empty((sh, sw, 4), np.float32)
            _ck(_lib.ocl_run(k, 1, wpg * sh, 1, wpg))
            _ck(_lib.ocl_read_buffer(o, rgba.ctypes.data, rgba.nbytes))
            img = _img(sw, sh, rgba)
            if xf is None:
                return img
            kt = fresh("transform.cl", "transform")
            out = _img(W, H)
            mb = _buf(_pad(xf_matrix(W, H, xf), 12))
            _ck(_lib.ocl_arg_mem(kt, 0, img))
            _ck(_lib.ocl_arg_mem(kt, 1, mb))
            _ck(_lib.ocl_arg_mem(kt, 2, out))
            self.launches.append((kt, 2, W, H, 0))
            _ck(_lib.ocl_run(kt, 2, W, H, 0))
            return out

        layers = []
        for L in scene["layers"]:
            a = source(L["src"], L["sw"], L["sh"], L.get("xf"))
            t = L.get("transition")
            if t:
                b = source(t["src"], t["sw"], t["sh"], t.get("xf"))
                out = _img(W, H)
                if t["type"] == "dissolve":
                    k = fresh("transition_dissolve.cl", "transition_dissolve")
                    _ck(_lib.ocl_arg_mem(k, 0, a)); _ck(_lib.ocl_arg_mem(k, 1, b)); _ck(_lib.ocl_arg_f32(k, 2, t["mix"])); _ck(_lib.ocl_arg_mem(k, 3, out))
                else:
                    m = source(t["mask"], t["mask_sw"], t["mask_sh"], t.get("mask_xf"))
                    k = fresh("transition_wipe.cl", "transition_wipe")
                    _ck(_lib.ocl_arg_mem(k, 0, a)); _ck(_lib.ocl_arg_mem(k, 1, b)); _ck(_lib.ocl_arg_mem(k, 2, m)); _ck(_lib.ocl_arg_mem(k, 3, out))
                self.launches.append((k, 2, W, H, 0))
                _ck(_lib.ocl_run(k, 2, W, H, 0))
                a = out
            layers.append(a)
        comp = layers[0]
        if len(layers) > 1:
            k = fresh(f"combine_{len(layers)}.cl", f"combine_{len(layers)}")
            comp = _img(W, H)
            for n, m in enumerate(layers):
                _ck(_lib.ocl_arg_mem(k, n, m))
            _ck(_lib.ocl_arg_mem(k, len(layers), comp))
            self.launches.append((k, 2, W, H, 0))
            _ck(_lib.ocl_run(k, 2, W, H, 0))
        # the writer reads a buffer: bring the composite back as one (done once; the write launch is what is timed)
        rgba = _read_img(comp, W, H)
        kw = fresh("v210.cl", "write")
        wpg = _pitch(W) // 48
        self.out_bytes = wpg * 128 * H
        wi, self.wo = _buf(rgba), _buf(nbytes=self.out_bytes)
        _ck(_lib.ocl_arg_mem(kw, 0, wi)); _ck(_lib.ocl_arg_mem(kw, 1, self.wo)); _ck(_lib.ocl_arg_u32(kw, 2, W)); _ck(_lib.ocl_arg_u32(kw, 3, 0))
        _ck(_lib.ocl_arg_mem(kw, 4, self.cm_w)); _ck(_lib.ocl_arg_mem(kw, 5, self.lut_w))
        self.launches.append((kw, 1, wpg * H, 1, wpg))
        _ck(_lib.ocl_run(kw, 1, wpg * H, 1, wpg))

    def result(self) -> np.ndarray:
        out = np.empty(self.out_bytes, np.uint8)
        _ck(_lib.ocl_read_buffer(self.wo, out.ctypes.data, self.out_bytes))
        return out

    def run_frames(self, n: int) -> float:
        """enqueue the whole launch sequence n times back to back, one clFinish at the end; seconds"""
        import time
        _lib.ocl_enqueue.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t]
        _ck(_lib.ocl_finish())
        t0 = time.perf_counter()
        for _ in range(n):
            for (k, dims, g0, g1, l0) in self.launches:
                _ck(_lib.ocl_enqueue(k, dims, g0, g1, l0))
        _ck(_lib.ocl_finish())
        return time.perf_counter() - t0
