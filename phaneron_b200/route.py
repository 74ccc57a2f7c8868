"""ROUTE across GPUs.

In the reference a ROUTE producer forks another channel's pipes inside ONE process and ONE device: the routed
"frame" is a reference to that channel's combined RGBA-f32 OpenCLBuffer (routeProducer.ts:63-70,
channel.ts:290-300).  Here channels shard one per GPU, one process per GPU, so a ROUTE whose source channel
lives on another GPU becomes the path's single exchange step: the source rank sends its channel frame
(RGBA-f32, exactly the bytes the reference would have shared; a frame that is still a deferred expression is
materialised by the send) point-to-point, the destination rank receives it into an image buffer that enters
its layer stack like any other source.  No reduction, no collective in steady state.

  * RouteComm / GpuRouteExchange: the GPU path.  NCCL point-to-point through the library's own C ABI
    (pb_comm_*, pb_route_*: csrc/pb_route.cu) on a side stream of the context, ordered against the process queue
    by CUDA events: what a Node.js host would call, no torch on the data path.
  * RouteExchange: the same plan over torch.distributed on CPU tensors (gloo): the host-logic tests.
  * RouteProducer: mirror of src/producer/routeProducer.ts (video side).

The payload is the RGBA frame and not the packed output, because packing quantises to 10-bit YCbCr: a routed
layer must see the same floats the reference's shared buffer holds.
"""
from __future__ import annotations

import ctypes as C
import re
from typing import Any, Callable, Dict, List, Optional, Tuple

from . import _lib
from ._lib import check


def channel_rank(channel: int, world: int) -> int:
    """channel i lives on GPU i mod N (index.ts:156-160 builds the channels; we shard them)"""
    return channel % world


class RouteTable:
    """which (channel, layer) slots are fed by which channel, as AMCP `PLAY 2-10 route://1` declares them"""

    def __init__(self, routes: List[Tuple[int, int]]):
        """routes: (source channel, destination channel) pairs"""
        self.routes = list(routes)

    def plan(self, rank: int, world: int) -> Tuple[List[Tuple[int, int]], List[Tuple[int, int]]]:
        """-> (sends, recvs) for this rank: sends = [(route index, peer rank)], recvs likewise; local routes
        (both channels on this rank) appear in neither list: they stay buffer references as in the reference"""
        sends, recvs = [], []
        for i, (src, dst) in enumerate(self.routes):
            rs, rd = channel_rank(src, world), channel_rank(dst, world)
            if rs == rd:
                continue
            if rs == rank:
                sends.append((i, rd))
            if rd == rank:
                recvs.append((i, rs))
        return sends, recvs


# ---- GPU path: the C ABI ------------------------------------------------------------------------------------------
class RouteComm:
    """pb_comm_* / pb_route_* (include/phaneron_b200.h): one NCCL communicator per context"""

    ID_BYTES = 128

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(RouteComm.ID_BYTES)
        check(_lib.lib().pb_comm_unique_id(buf))
        return buf.raw

    def __init__(self, ctx, rank: int, world: int, unique_id: bytes):
        if len(unique_id) != self.ID_BYTES:
            raise ValueError(f"NCCL unique id must be {self.ID_BYTES} bytes")
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        h = C.c_void_p()
        check(_lib.lib().pb_comm_init(ctx._need(), self.rank, self.world, C.create_string_buffer(unique_id, self.ID_BYTES), C.byref(h)))
        self._h = h.value

    def attach(self, landing, peer_in: int, peer_out: int) -> bool:
        """collective, once: map the neighbours' landing buffers (CUDA IPC) so that frames cross with the copy engines instead of
        NCCL's copy kernel.  `landing`: this rank's buffers in the order they will be passed to recv().  -> True if attached"""
        arr = (C.c_void_p * len(landing))(*[b._h for b in landing])
        check(_lib.lib().pb_route_attach(self._h, arr, len(landing), int(peer_in), int(peer_out)))
        return bool(_lib.lib().pb_route_transport(self._h))

    def begin(self) -> None:
        check(_lib.lib().pb_route_begin(self._h))

    def send(self, buf, peer: int) -> None:
        check(_lib.lib().pb_route_send(self._h, buf._h, int(peer)))

    def recv(self, buf, peer: int) -> None:
        check(_lib.lib().pb_route_recv(self._h, buf._h, int(peer)))

    def end(self) -> None:
        check(_lib.lib().pb_route_end(self._h))

    def wait(self, queue: int = _lib.QUEUE_PROCESS, age: int = 0) -> None:
        """device-side: `queue` waits for the exchange last ended (age 0) or the one `age` exchanges before it; the host
        does not block"""
        check(_lib.lib().pb_route_wait_age(self._h, int(queue), int(age)))

    def sync(self) -> None:
        check(_lib.lib().pb_route_sync(self._h))

    def info(self) -> Dict[str, int]:
        r, w, s, g = C.c_int(), C.c_int(), C.c_uint64(), C.c_uint64()
        check(_lib.lib().pb_comm_info(self._h, C.byref(r), C.byref(w), C.byref(s), C.byref(g)))
        return {"rank": r.value, "world": w.value, "bytes_sent": s.value, "bytes_received": g.value}

    def close(self) -> None:
        if self._h:
            _lib.lib().pb_comm_destroy(self._h)
            self._h = 0


class GpuRouteExchange:
    """One frame period's ROUTE traffic of this rank over RouteComm, one frame ahead: start() posts the sends of the
    frame just composed and the receives of the peers' (nothing blocks, the copies run on the side stream while the next
    frame is composed); finish() hands out the frames that arrived, making the process queue wait for them on the device."""

    def __init__(self, ctx, comm: RouteComm, table: RouteTable, width: int, height: int):
        self.ctx, self.comm, self.table, self.width, self.height = ctx, comm, table, width, height
        self.sends, self.recvs = table.plan(comm.rank, comm.world)
        self.landing: Dict[int, list] = {}
        self.phase = 0
        self._started = False

    async def init(self) -> None:
        n = self.width * self.height * 16
        dims = {"width": self.width, "height": self.height}
        for i, _ in self.recvs:   # two landing frames per incoming route: frame n is consumed while frame n+1 arrives
            self.landing[i] = [await self.ctx.createBuffer(n, "readwrite", "coarse", dims, f"route {i} landing {k}") for k in range(2)]

    def start(self, outgoing: Dict[int, Any]) -> None:
        self.comm.begin()
        for i, peer in self.sends:
            self.comm.send(outgoing[i], peer)
        for i, peer in self.recvs:
            self.comm.recv(self.landing[i][self.phase], peer)
        self.comm.end()
        self._started = True

    def finish(self) -> Dict[int, Any]:
        """-> {route index: OpenCLBuffer holding the received frame}; valid until the start() after next"""
        if self._started:
            self.comm.wait(_lib.QUEUE_PROCESS)
        got = {i: self.landing[i][self.phase] for i, _ in self.recvs} if self._started else {}
        self.phase ^= 1
        self._started = False
        return got

    def close(self) -> None:
        self.comm.sync()
        for bufs in self.landing.values():
            for b in bufs:
                b.release()
        self.landing = {}


# ---- host logic over torch.distributed (CPU / gloo): what the world-size-2 tests run ---------------------------------
class RouteExchange:
    """one frame period's worth of ROUTE traffic for this rank over torch.distributed P2P on CPU tensors (gloo).
    The GPU path is GpuRouteExchange (the library's own NCCL calls, ordered against its queues)."""

    def __init__(self, table: RouteTable, frame_bytes: int, device=None, group=None):
        import torch
        import torch.distributed as dist
        self._torch, self._dist = torch, dist
        device = device or torch.device("cpu")
        if torch.device(device).type != "cpu":
            raise ValueError("RouteExchange is the CPU (gloo) planning path; on GPUs use RouteComm / GpuRouteExchange")
        self.table, self.frame_bytes, self.device, self.group = table, int(frame_bytes), device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.sends, self.recvs = table.plan(self.rank, self.world)
        # two landing buffers per incoming route: frame n is consumed while frame n+1 arrives
        self.landing = {i: [torch.empty(self.frame_bytes, dtype=torch.uint8, device=device) for _ in range(2)] for i, _ in self.recvs}
        self.phase = 0
        self._pending: list = []
        self._keep: list = []

    def start(self, outgoing) -> None:
        """post the sends of this frame's routed outputs and the receives of the peers' (non-blocking)"""
        torch, dist = self._torch, self._dist
        ops = []
        self._keep = []
        for i, peer in self.sends:
            t = outgoing[i]
            assert t.numel() * t.element_size() == self.frame_bytes and t.is_contiguous()
            self._keep.append(t)   # the payload stays referenced until finish()
            ops.append(dist.P2POp(dist.isend, t.view(torch.uint8).reshape(-1), peer, self.group))
        for i, peer in self.recvs:
            ops.append(dist.P2POp(dist.irecv, self.landing[i][self.phase], peer, self.group))
        self._pending = dist.batch_isend_irecv(ops) if ops else []

    def finish(self):
        """wait for this period's transfers; -> {route index: received frame, valid for one more period}"""
        for w in self._pending:
            w.wait()
        self._pending = []
        self._keep = []
        got = {i: self.landing[i][self.phase] for i, _ in self.recvs}
        self.phase ^= 1
        return got


# ---- src/producer/routeProducer.ts (video side) -------------------------------------------------------------------
class InvalidProducerError(Exception):
    pass


def chanLayerFromString(chanLayStr: str) -> Dict[str, Any]:
    """chanLayer.ts:51-66"""
    m = re.search(r"(?P<channel>\d+)-?(?P<layer>\d*)", chanLayStr or "")
    if not m:
        return {"valid": False, "channel": 0, "layer": 0}
    return {"valid": True, "channel": int(m.group("channel")), "layer": int(m.group("layer")) if m.group("layer") != "" else 0}


class RouteProducer:
    """routeProducer.ts:33-186, video side.  `channels` is the directory index.ts exports (list of objects with
    `async getRoutePipes(layer)` -> {video: async callable returning the next OpenCLBuffer | None, format, release}).  A
    channel that lives on another GPU is represented in that directory by a RemoteChannel, whose pipes hand out the frames
    GpuRouteExchange received; RouteProducer itself does not care where the frame came from -- as in the reference it only
    forks the pipes and adds one reference per extra fork (routeProducer.ts:106-113)."""

    def __init__(self, id_: int, params: Dict[str, Any], channels: List[Any]):
        self.sourceID = f"P{id_} {params['url']} L{params['layer']}"
        self.params = params
        self.channels = channels
        self.srcPipes: Optional[Dict[str, Any]] = None
        self.srcFormat = None
        self.numForks = 0
        self.paused = True
        self.running = True
        if params["url"][:5].upper() != "ROUTE":
            raise InvalidProducerError("Route producer supports route command")

    async def initialise(self) -> None:
        url = self.params["url"]
        routeIndex = url.find("://")
        if routeIndex < 0:
            raise RuntimeError("Route producer failed to find route source in parameters")
        chanLayer = chanLayerFromString(url[routeIndex + 3:])
        if not chanLayer["valid"]:
            raise RuntimeError(f"Route producer failed to parse channel and layer from params {url[routeIndex + 3:]}")
        idx = chanLayer["channel"] - 1
        channel = self.channels[idx] if 0 <= idx < len(self.channels) else None
        if not channel:
            raise RuntimeError(f"Route producer failed to find source of channel {chanLayer['channel']}")
        self.srcPipes = await channel.getRoutePipes(chanLayer["layer"])
        self.srcFormat = self.srcPipes["format"]

    async def _vid(self):
        """srcPipes.video.valve(vidForkRef).pause(...) (routeProducer.ts:106-126): one reference per extra fork; a paused
        producer keeps re-presenting the frame with one more reference each time"""
        if not self.running:
            return None
        frame = await self.srcPipes["video"]()
        if frame is None:
            return None
        for _ in range(1, self.numForks):
            frame.addRef()
        if self.paused:
            frame.addRef()
        return frame

    def getSourcePipes(self) -> Dict[str, Any]:
        if not (self.srcPipes and self.srcFormat is not None):
            raise RuntimeError("Route producer failed to find source pipes for route")
        self.numForks += 1
        released = {"done": False}

        def release() -> None:
            if not released["done"]:
                released["done"] = True
                self.numForks -= 1

        return {"video": self._vid, "format": self.srcFormat, "release": release}

    def srcID(self) -> str:
        return self.sourceID

    def setPaused(self, pause: bool) -> None:
        self.paused = pause

    def release(self) -> None:
        self.running = False
        if self.srcPipes:
            self.srcPipes["release"]()


class RemoteChannel:
    """stand-in, in the `channels` directory of a rank, for a channel that lives on another GPU: its route pipes deliver the
    frames that GpuRouteExchange (or any callable) received for route `route_index`"""

    def __init__(self, fetch: Callable[[], Any], fmt: Any):
        self._fetch, self._fmt = fetch, fmt

    async def getRoutePipes(self, layerNum: int) -> Dict[str, Any]:
        if layerNum != 0:
            raise RuntimeError(f"Failed to find source pipes for layer {layerNum}")   # only channel outputs cross GPUs

        async def video():
            f = self._fetch()
            if f is not None:
                f.addRef()   # the consumer releases it like any source frame; the landing buffer itself stays with the exchange
            return f

        return {"video": video, "format": self._fmt, "release": lambda: None}


class _CudaArray:
    """__cuda_array_interface__ view of raw device memory (an OpenCLBuffer's device face) for torch.as_tensor"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def buffer_as_tensor(buf, device):
    """zero-copy uint8 tensor over an OpenCLBuffer's device memory (materialises a deferred frame first and waits for the
    process queue: torch's streams are not ordered against the library's queues)"""
    import torch
    ptr = buf.devicePointer()
    check(_lib.lib().pb_wait_finish(buf._ctx._need(), _lib.QUEUE_PROCESS))
    return torch.as_tensor(_CudaArray(ptr, buf.numBytes), device=device)


def tensor_as_buffer(ctx, t, width: int, height: int, owner: str = "route"):
    """an RGBA-f32 OpenCLBuffer over a device tensor; `t` must stay alive while the buffer is in use, and whatever produced
    `t` on torch's stream must have completed (torch.cuda.current_stream().synchronize()) before the buffer is used"""
    assert t.is_cuda and t.is_contiguous() and t.numel() * t.element_size() >= width * height * 16
    return ctx.wrapDeviceMemory(t.data_ptr(), width * height * 16, {"width": width, "height": height}, owner)
