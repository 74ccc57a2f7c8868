"""ctypes binding of libphaneron_b200.so (include/phaneron_b200.h).

There is no fallback: if the shared library is missing this raises, and
pb_ctx_create fails on a machine without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PB_LIB") or os.path.join(HERE, "libphaneron_b200.so")   # PB_LIB: kernel-variant experiments

PB_OK = 0
QUEUE_LOAD, QUEUE_PROCESS, QUEUE_UNLOAD = 0, 1, 2
DIR = {"readonly": 0, "writeonly": 1, "readwrite": 2}
SVM = {"none": 0, "coarse": 1, "fine": 2}
ACCESS = {"none": 0, "readonly": 1, "writeonly": 2}
CTX_DEFER = 1
CTX_NO_MARCH = 2
CTX_RAW_LUT = 4
CTX_NO_CULL = 8
CTX_FOOTPRINT = 16
CTX_NO_DIRECT = 32

OPS = {
    "v210_read": 1, "v210_write": 2, "rgba8_read": 3, "rgba8_write": 4, "bgra8_read": 5, "bgra8_write": 6,
    "combine": 10, "dissolve": 11, "wipe_mask": 12, "transform": 13, "yadif": 14, "mix": 15, "wipe": 16,
    "resize": 17, "yuv422p10_read": 20, "yuv422p10_write": 21, "yuv422p8_read": 22, "yuv422p8_write": 23,
    "yuv420p_read": 24, "yuv420p_write": 25, "nv12_read": 26, "nv12_write": 27,
}

# every symbol include/phaneron_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "pb_last_error", "pb_version", "pb_ctx_create", "pb_ctx_destroy", "pb_ctx_info", "pb_ctx_stats",
    "pb_ctx_set_flags", "pb_buf_create", "pb_buf_wrap", "pb_buf_addref", "pb_buf_release", "pb_buf_refs",
    "pb_buf_bytes", "pb_buf_trim", "pb_buf_host_ptr", "pb_buf_dev_ptr", "pb_buf_is_deferred", "pb_buf_host_access",
    "pb_buf_upload_async", "pb_buf_download_async", "pb_host_alloc", "pb_host_free", "pb_prog_create",
    "pb_prog_destroy", "pb_run_program", "pb_wait_finish", "pb_queue_wait_queue", "pb_ctx_stream",
    "pb_event_create", "pb_event_record", "pb_event_sync", "pb_event_elapsed_ms", "pb_event_destroy",
    "pb_chain_begin", "pb_chain_end", "pb_chain_info", "pb_chain_replay", "pb_chain_destroy",
    "pb_gamma2linear_lut", "pb_linear2gamma_lut", "pb_ycbcr2rgb_matrix", "pb_rgb2ycbcr_matrix",
    "pb_rgb2rgb_matrix", "pb_transform_matrix",
    "pb_comm_unique_id", "pb_comm_init", "pb_comm_info", "pb_comm_destroy", "pb_route_begin", "pb_route_send",
    "pb_route_recv", "pb_route_end", "pb_route_attach", "pb_route_transport", "pb_route_wait", "pb_route_wait_age", "pb_route_sync", "pb_route_copy_peer",
]


class Param(C.Structure):
    _fields_ = [("name", C.c_char_p), ("kind", C.c_int), ("buf", C.c_void_p), ("num", C.c_double)]


class Timings(C.Structure):
    _fields_ = [("dataToKernel", C.c_uint32), ("kernelExec", C.c_uint32), ("totalTime", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "kernel_launches", "fused_launches", "march_launches", "deferred_nodes", "materialised", "h2d_bytes", "d2h_bytes",
        "dev_bytes_live", "dev_bytes_pooled", "lut_tables", "lut_tables_d8", "march_src_bytes", "lut_tables_poly", "run_program_ns", "run_program_calls")]


class PhaneronError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PhaneronError(
            f"{LIB_PATH} is missing: build it with `python -m phaneron_b200.build` "
            "(phaneron_b200 has no CPU or PyTorch fallback)")
    l = C.CDLL(LIB_PATH)
    vp, i, sz, cp = C.c_void_p, C.c_int, C.c_size_t, C.c_char_p
    f32p = C.POINTER(C.c_float)
    sig = {
        "pb_last_error": (cp, []),
        "pb_version": (cp, []),
        "pb_ctx_create": (i, [i, C.c_uint, C.POINTER(vp)]),
        "pb_ctx_destroy": (i, [vp]),
        "pb_ctx_info": (i, [vp, C.c_char_p, sz]),
        "pb_ctx_stats": (i, [vp, C.POINTER(Stats)]),
        "pb_ctx_set_flags": (i, [vp, C.c_uint]),
        "pb_buf_create": (i, [vp, sz, i, i, i, i, cp, C.POINTER(vp)]),
        "pb_buf_wrap": (i, [vp, vp, sz, i, i, C.POINTER(vp)]),
        "pb_buf_addref": (i, [vp]),
        "pb_buf_release": (i, [vp]),
        "pb_buf_refs": (i, [vp]),
        "pb_buf_bytes": (sz, [vp]),
        "pb_buf_trim": (i, [vp]),
        "pb_buf_host_ptr": (vp, [vp]),
        "pb_buf_dev_ptr": (vp, [vp]),
        "pb_buf_is_deferred": (i, [vp]),
        "pb_buf_host_access": (i, [vp, i, i, vp, sz]),
        "pb_buf_upload_async": (i, [vp, i, vp, sz]),
        "pb_buf_download_async": (i, [vp, i, vp, sz]),
        "pb_host_alloc": (vp, [sz]),
        "pb_host_free": (None, [vp]),
        "pb_prog_create": (i, [vp, i, i, i, C.POINTER(vp)]),
        "pb_prog_destroy": (i, [vp]),
        "pb_run_program": (i, [vp, vp, C.POINTER(Param), i, i, C.POINTER(Timings)]),
        "pb_wait_finish": (i, [vp, i]),
        "pb_queue_wait_queue": (i, [vp, i, i]),
        "pb_ctx_stream": (vp, [vp, i]),
        "pb_event_create": (i, [vp, C.POINTER(vp)]),
        "pb_event_record": (i, [vp, i]),
        "pb_event_sync": (i, [vp]),
        "pb_event_elapsed_ms": (i, [vp, vp, f32p]),
        "pb_event_destroy": (i, [vp]),
        "pb_chain_begin": (i, [vp]),
        "pb_chain_end": (i, [vp, C.POINTER(vp)]),
        "pb_chain_info": (i, [vp, C.POINTER(i), C.POINTER(i)]),
        "pb_chain_replay": (i, [vp, i]),
        "pb_chain_destroy": (i, [vp]),
        "pb_gamma2linear_lut": (i, [cp, vp]),
        "pb_linear2gamma_lut": (i, [cp, vp]),
        "pb_ycbcr2rgb_matrix": (i, [cp, i, i, i, i, vp]),
        "pb_rgb2ycbcr_matrix": (i, [cp, i, i, i, i, vp]),
        "pb_rgb2rgb_matrix": (i, [cp, cp, vp]),
        "pb_transform_matrix": (i, [i, i, i, i] + [C.c_double] * 7 + [vp]),
        "pb_comm_unique_id": (i, [vp]),
        "pb_comm_init": (i, [vp, i, i, vp, C.POINTER(vp)]),
        "pb_comm_info": (i, [vp, C.POINTER(i), C.POINTER(i), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "pb_comm_destroy": (i, [vp]),
        "pb_route_begin": (i, [vp]),
        "pb_route_send": (i, [vp, vp, i]),
        "pb_route_recv": (i, [vp, vp, i]),
        "pb_route_end": (i, [vp]),
        "pb_route_attach": (i, [vp, C.POINTER(vp), i, i, i]),
        "pb_route_transport": (i, [vp]),
        "pb_route_wait": (i, [vp, i]),
        "pb_route_wait_age": (i, [vp, i, i]),
        "pb_route_sync": (i, [vp]),
        "pb_route_copy_peer": (i, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def check(rc: int) -> None:
    if rc != PB_OK:
        raise PhaneronError(lib().pb_last_error().decode() or f"phaneron_b200 error {rc}")


def call_in_worker(fn, *args):
    """for run_in_executor: pb_last_error() is thread-local, so the message of a failed call has to be fetched on the worker
    thread that made it (the event-loop thread would read its own, stale or empty, message).  -> (rc, message)"""
    rc = fn(*args)
    return rc, (lib().pb_last_error().decode() if rc != PB_OK else "")


def check_worker(res) -> None:
    rc, msg = res
    if rc != PB_OK:
        raise PhaneronError(msg or f"phaneron_b200 error {rc}")
