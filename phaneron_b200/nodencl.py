"""Host-side mirror of the `nodencl` addon surface that phaneron's src/process/*.ts and
src/clJobQueue.ts use, implemented over the C ABI of
libphaneron_b200.so.  Names, argument meaning and error behaviour follow the call sites
in the reference:

  new clContext({platformIndex, deviceIndex, overlapping}) / initialise()   index.ts:94-102
  createBuffer(numBytes, dir, svm, imageDims?, owner?)                      io.ts:61-77
  OpenCLBuffer.hostAccess / addRef / release / timestamp                    io.ts:89-94
  createProgram(source, {name, globalWorkItems, workItemsPerGroup})         packer.ts:97-103
  runProgram(program, params, queue) -> RunTimings                          clJobQueue.ts:122-128
  waitFinish(queue)                                                          clJobQueue.ts:131

Promises become coroutines.  The OpenCL source string argument of createProgram is a
KernelSpec naming the CUDA op that replaces that source (the reference reuses the entry
names 'read'/'write' for every packer, so identity has to come from the source).
"""
from __future__ import annotations

import asyncio
import ctypes as C
import json
import time
from dataclasses import dataclass
from typing import Any, Dict, Optional, Union

import numpy as np

from . import _lib
from ._lib import PhaneronError, check


@dataclass(frozen=True)
class KernelSpec:
    """Stands where the reference passes OpenCL C source text."""
    op: str

    def __post_init__(self):
        if self.op not in _lib.OPS:
            raise PhaneronError(f"unknown kernel op '{self.op}'")


@dataclass
class RunTimings:
    dataToKernel: int = 0
    kernelExec: int = 0
    totalTime: int = 0


class _Queues:
    load = _lib.QUEUE_LOAD
    process = _lib.QUEUE_PROCESS
    unload = _lib.QUEUE_UNLOAD


class OpenCLProgram:
    def __init__(self, ctx: "clContext", handle: int, spec: KernelSpec, name: str, width: int, height: int):
        self._ctx = ctx
        self._h = handle
        self.spec = spec
        self.name = name
        self.width = width
        self.height = height

    def __del__(self):
        try:
            if self._h and self._ctx._h:
                _lib.lib().pb_prog_destroy(self._h)
        except Exception:
            pass


_ASYNC_COPY_BYTES = 1 << 20


class OpenCLBuffer:
    """nodencl's OpenCLBuffer extends node's Buffer: host-addressable bytes plus
    addRef/release/hostAccess and a mutable timestamp.  `.host` is a numpy uint8 view of
    the pinned host face."""

    def __init__(self, ctx: "clContext", handle: int, num_bytes: int, owner: str, image_dims):
        self._ctx = ctx
        self._h = handle
        self.numBytes = num_bytes
        self.length = num_bytes
        self.owner = owner
        self.imageDims = image_dims
        self.timestamp = 0
        self.loadstamp = 0
        self.creationTime = time.perf_counter_ns()
        self._host: Optional[np.ndarray] = None

    # -- Buffer face ------------------------------------------------------------------
    @property
    def host(self) -> np.ndarray:
        if self._host is None:
            self._alive()
            p = _lib.lib().pb_buf_host_ptr(self._h)
            if not p:
                raise PhaneronError(_lib.lib().pb_last_error().decode())
            self._host = np.ctypeslib.as_array((C.c_uint8 * self.numBytes).from_address(p))
        return self._host

    def view(self, dtype) -> np.ndarray:
        return self.host.view(dtype)

    def fill(self, value: int) -> None:
        self.host[:] = value

    def compare(self, other: Union["OpenCLBuffer", np.ndarray, bytes]) -> int:
        """Buffer.compare(): 0 when equal (the reference's round-trip pass criterion)."""
        a = self.host
        b = other.host if isinstance(other, OpenCLBuffer) else np.frombuffer(other, np.uint8)
        if a.size == b.size and np.array_equal(a, b):
            return 0
        n = min(a.size, b.size)
        ne = np.nonzero(a[:n] != b[:n])[0]
        if ne.size == 0:
            return -1 if a.size < b.size else 1
        return -1 if a[ne[0]] < b[ne[0]] else 1

    # -- refcounting -------------------------------------------------------------------
    def _alive(self) -> None:
        if not self._h:
            raise PhaneronError(f"buffer '{self.owner}' has been released")

    def addRef(self) -> None:
        self._alive()
        check(_lib.lib().pb_buf_addref(self._h))

    def release(self) -> None:
        self._alive()
        l = _lib.lib()
        last = l.pb_buf_refs(self._h) == 1
        check(l.pb_buf_release(self._h))
        if last:
            self._h = 0
            self._host = None

    @property
    def refs(self) -> int:
        return _lib.lib().pb_buf_refs(self._h) if self._h else 0

    @property
    def deferred(self) -> bool:
        return bool(self._h and _lib.lib().pb_buf_is_deferred(self._h))

    def devicePointer(self) -> int:
        self._alive()
        p = _lib.lib().pb_buf_dev_ptr(self._h)
        if not p:
            raise PhaneronError(_lib.lib().pb_last_error().decode())
        return p

    # -- hostAccess ----------------------------------------------------------------------
    async def hostAccess(self, mode: str = "none", queue: int = 0, src=None) -> None:
        self._alive()
        if mode not in _lib.ACCESS:
            raise PhaneronError(f"hostAccess mode must be one of none|readonly|writeonly, found '{mode}'")
        ptr, n = None, 0
        if src is not None:
            arr = src.host if isinstance(src, OpenCLBuffer) else np.frombuffer(src, np.uint8) if not isinstance(src, np.ndarray) else src
            arr = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
            ptr, n = arr.ctypes.data, arr.size
        h, m, q = self._h, _lib.ACCESS[mode], int(queue or 0)
        if (n if src is not None else self.numBytes) >= _ASYNC_COPY_BYTES and mode != "none":
            # nodencl runs every call as async work on the libuv pool: frame-sized copies go to a
            # worker thread (ctypes drops the GIL), so `await Promise.all([...])`-style callers overlap H2D, D2H
            # and kernel launches exactly as they do under Node
            _lib.check_worker(await asyncio.get_running_loop().run_in_executor(None, _lib.call_in_worker, _lib.lib().pb_buf_host_access, h, m, q, ptr, n))
        else:
            check(_lib.lib().pb_buf_host_access(h, m, q, ptr, n))


class clContext:
    def __init__(self, options: Optional[Dict[str, Any]] = None):
        options = options or {}
        self.platformIndex = int(options.get("platformIndex", 0))
        self.deviceIndex = int(options.get("deviceIndex", 0))
        self.overlapping = bool(options.get("overlapping", True))
        # product extension: defer RGBA intermediates and fuse the chain at packed sinks
        self.deferred = bool(options.get("deferred", True))
        self.marchKernel = bool(options.get("marchKernel", True))   # False: always the generic fused kernel
        self.rawLut = bool(options.get("rawLut", False))            # True: march kernel gathers from the raw gamma tables
        self.occlusionCulling = bool(options.get("occlusionCulling", True))   # False: evaluate layers hidden under opaque ones too
        self.footprint = bool(options.get("footprint", False))      # True: stats()['march_src_bytes'] per march launch
        self.directKernel = bool(options.get("directKernel", True))  # False: 1:1 v210 -> v210 frames take the general march kernel (A/B)
        self.queue = _Queues()
        self._h = 0

    async def initialise(self) -> None:
        if self.platformIndex != 0:
            raise PhaneronError("phaneron_b200 exposes one platform (CUDA); platformIndex must be 0")
        h = C.c_void_p()
        check(_lib.lib().pb_ctx_create(self.deviceIndex, self._flags(), C.byref(h)))
        self._h = h.value

    def _flags(self) -> int:
        return ((_lib.CTX_DEFER if self.deferred else 0) | (0 if self.marchKernel else _lib.CTX_NO_MARCH)
                | (_lib.CTX_RAW_LUT if self.rawLut else 0) | (0 if self.occlusionCulling else _lib.CTX_NO_CULL)
                | (_lib.CTX_FOOTPRINT if self.footprint else 0) | (0 if self.directKernel else _lib.CTX_NO_DIRECT))

    def close(self) -> None:
        if self._h:
            _lib.lib().pb_ctx_destroy(self._h)
            self._h = 0

    def _need(self) -> int:
        if not self._h:
            raise PhaneronError("clContext.initialise() has not completed")
        return self._h

    def setDeferred(self, on: bool) -> None:
        self.deferred = bool(on)
        check(_lib.lib().pb_ctx_set_flags(self._need(), self._flags()))

    def setMarchKernel(self, on: bool, rawLut: Optional[bool] = None) -> None:
        self.marchKernel = bool(on)
        if rawLut is not None:
            self.rawLut = bool(rawLut)
        check(_lib.lib().pb_ctx_set_flags(self._need(), self._flags()))

    def setOcclusionCulling(self, on: bool) -> None:
        self.occlusionCulling = bool(on)
        check(_lib.lib().pb_ctx_set_flags(self._need(), self._flags()))

    def getPlatformInfo(self) -> Dict[str, Any]:
        buf = C.create_string_buffer(1024)
        check(_lib.lib().pb_ctx_info(self._need(), buf, len(buf)))
        return json.loads(buf.value.decode())

    def stats(self) -> Dict[str, int]:
        s = _lib.Stats()
        check(_lib.lib().pb_ctx_stats(self._need(), C.byref(s)))
        return {n: int(getattr(s, n)) for n, _ in s._fields_}

    async def createBuffer(self, numBytes: int, bufDir: str, bufType: str, imageDims: Optional[Dict[str, int]] = None,
                           owner: Optional[str] = None) -> OpenCLBuffer:
        if bufDir not in _lib.DIR:
            raise PhaneronError(f"buffer direction must be readonly|writeonly|readwrite, found '{bufDir}'")
        if bufType not in _lib.SVM:
            raise PhaneronError(f"buffer type must be none|coarse|fine, found '{bufType}'")
        w = int(imageDims["width"]) if imageDims else 0
        h = int(imageDims["height"]) if imageDims else 0
        out = C.c_void_p()
        check(_lib.lib().pb_buf_create(self._need(), int(numBytes), _lib.DIR[bufDir], _lib.SVM[bufType], w, h,
                                       (owner or "").encode(), C.byref(out)))
        return OpenCLBuffer(self, out.value, int(numBytes), owner or "", imageDims)

    def wrapDeviceMemory(self, devicePointer: int, numBytes: int, imageDims: Optional[Dict[str, int]] = None,
                         owner: str = "wrapped") -> OpenCLBuffer:
        """an OpenCLBuffer over device memory the caller owns (pb_buf_wrap): how a ROUTE frame received over NCCL
        enters a channel as a layer source (route.py); the memory must outlive the buffer"""
        w = int(imageDims["width"]) if imageDims else 0
        h = int(imageDims["height"]) if imageDims else 0
        out = C.c_void_p()
        check(_lib.lib().pb_buf_wrap(self._need(), C.c_void_p(devicePointer), int(numBytes), w, h, C.byref(out)))
        return OpenCLBuffer(self, out.value, int(numBytes), owner, imageDims)

    async def createProgram(self, kernel: KernelSpec, options: Dict[str, Any]) -> OpenCLProgram:
        if not isinstance(kernel, KernelSpec):
            raise PhaneronError("createProgram expects a KernelSpec where the reference passes OpenCL source")
        width, height = int(options["width"]), int(options["height"])
        out = C.c_void_p()
        check(_lib.lib().pb_prog_create(self._need(), _lib.OPS[kernel.op], width, height, C.byref(out)))
        return OpenCLProgram(self, out.value, kernel, str(options.get("name", kernel.op)), width, height)

    async def runProgram(self, program: OpenCLProgram, params: Dict[str, Any], queue: int = _lib.QUEUE_PROCESS,
                         timed: bool = False) -> RunTimings:
        arr = (_lib.Param * len(params))()
        for i, (name, val) in enumerate(params.items()):
            arr[i].name = name.encode()
            if isinstance(val, OpenCLBuffer):
                val._alive()
                arr[i].kind, arr[i].buf = 0, val._h
            elif isinstance(val, (bool, int, float, np.integer, np.floating)):
                arr[i].kind, arr[i].num = 1, float(val)
            elif val is None:
                raise PhaneronError(f"kernel parameter '{name}' is null")
            else:
                raise PhaneronError(f"kernel parameter '{name}' has unsupported type {type(val).__name__}")
        t = _lib.Timings()
        check(_lib.lib().pb_run_program(self._need(), program._h, arr, len(params), int(queue),
                                        C.byref(t) if timed else None))
        return RunTimings(t.dataToKernel, t.kernelExec, t.totalTime)

    async def waitFinish(self, queue: int = _lib.QUEUE_PROCESS) -> None:
        q = int(queue or 0)
        if q == _lib.QUEUE_PROCESS:
            check(_lib.lib().pb_wait_finish(self._need(), q))
            return
        # the copy queues are waited on from a worker thread, like nodencl's waitFinish on the libuv pool: a producer
        # waiting for its upload (macadamProducer.ts:186) must not stall the loop that is composing the previous frame
        _lib.check_worker(await asyncio.get_running_loop().run_in_executor(None, _lib.call_in_worker, _lib.lib().pb_wait_finish, self._need(), q))


class Chain:
    """Recorded fused launches of one frame (pb_chain_*); replay re-issues them."""

    def __init__(self, ctx: clContext, handle: int):
        self._ctx = ctx
        self._h = handle
        n, ok = C.c_int(), C.c_int()
        check(_lib.lib().pb_chain_info(handle, C.byref(n), C.byref(ok)))
        self.launches = n.value
        self.complete = bool(ok.value)

    def replay(self, queue: int = _lib.QUEUE_PROCESS) -> None:
        check(_lib.lib().pb_chain_replay(self._h, queue))

    def destroy(self) -> None:
        if self._h:
            _lib.lib().pb_chain_destroy(self._h)
            self._h = 0


class Event:
    def __init__(self, ctx: clContext):
        h = C.c_void_p()
        check(_lib.lib().pb_event_create(ctx._need(), C.byref(h)))
        self._h = h.value

    def record(self, queue: int = _lib.QUEUE_PROCESS) -> None:
        check(_lib.lib().pb_event_record(self._h, queue))

    def synchronize(self) -> None:
        check(_lib.lib().pb_event_sync(self._h))

    def elapsed_ms(self, stop: "Event") -> float:
        ms = C.c_float()
        check(_lib.lib().pb_event_elapsed_ms(self._h, stop._h, C.byref(ms)))
        return float(ms.value)


def _begin_chain(self: clContext) -> None:
    check(_lib.lib().pb_chain_begin(self._need()))


def _end_chain(self: clContext) -> Chain:
    h = C.c_void_p()
    check(_lib.lib().pb_chain_end(self._need(), C.byref(h)))
    return Chain(self, h.value)


clContext.beginChain = _begin_chain
clContext.endChain = _end_chain
clContext.createEvent = lambda self: Event(self)
