"""ROUTE across GPUs.

In the reference a ROUTE producer forks another channel's pipes inside ONE process and ONE device: the routed
"frame" is a reference to that channel's combined RGBA-f32 OpenCLBuffer (routeProducer.ts:63-70,
channel.ts:290-300).  Here channels shard one per GPU, one process per GPU, so a ROUTE whose source channel
lives on another GPU becomes the path's single exchange step: the source rank materialises its channel
frame (RGBA-f32, exactly the bytes the reference would have shared) and sends it point-to-point; the
destination rank receives into device memory and wraps it as an OpenCLBuffer (pb_buf_wrap) that enters its
layer stack like any other source.  No reduction, no collective in steady state: torch.distributed P2P over
NCCL (NVLink 5 / NVSwitch) on GPUs, gloo on CPU for the host-logic tests.

The payload is the RGBA frame and not the packed output, because packing quantises to 10-bit YCbCr: a routed
layer must see the same floats the reference's shared buffer holds.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


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


class RouteExchange:
    """one frame period's worth of ROUTE traffic for this rank, double-buffered one frame ahead"""

    def __init__(self, table: RouteTable, frame_bytes: int, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        self.table, self.frame_bytes, self.device, self.group = table, int(frame_bytes), device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.sends, self.recvs = table.plan(self.rank, self.world)
        # two landing buffers per incoming route: frame n is consumed while frame n+1 arrives
        self.landing: Dict[int, List[torch.Tensor]] = {i: [torch.empty(self.frame_bytes, dtype=torch.uint8, device=device) for _ in range(2)]
                                                       for i, _ in self.recvs}
        self.phase = 0
        self._pending: list = []

    def start(self, outgoing: Dict[int, torch.Tensor]) -> None:
        """post the sends of this frame's routed outputs and the receives of the peers' (non-blocking)"""
        ops = []
        for i, peer in self.sends:
            t = outgoing[i]
            assert t.numel() * t.element_size() == self.frame_bytes and t.is_contiguous()
            ops.append(dist.P2POp(dist.isend, t.view(torch.uint8).reshape(-1), peer, self.group))
        for i, peer in self.recvs:
            ops.append(dist.P2POp(dist.irecv, self.landing[i][self.phase], peer, self.group))
        self._pending = dist.batch_isend_irecv(ops) if ops else []

    def finish(self) -> Dict[int, torch.Tensor]:
        """wait for this period's transfers; -> {route index: received frame (device tensor, valid for one more period)}"""
        for w in self._pending:
            w.wait()
        self._pending = []
        got = {i: self.landing[i][self.phase] for i, _ in self.recvs}
        self.phase ^= 1
        return got


class _CudaArray:
    """__cuda_array_interface__ view of raw device memory (an OpenCLBuffer's device face) for torch.as_tensor"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def buffer_as_tensor(buf, device: torch.device) -> torch.Tensor:
    """zero-copy uint8 tensor over an OpenCLBuffer's device memory (materialises a deferred frame first)"""
    return torch.as_tensor(_CudaArray(buf.devicePointer(), buf.numBytes), device=device)


def tensor_as_buffer(ctx, t: torch.Tensor, width: int, height: int, owner: str = "route"):
    """an RGBA-f32 OpenCLBuffer over a received frame; `t` must stay alive while the buffer is in use"""
    assert t.is_cuda and t.is_contiguous() and t.numel() * t.element_size() >= width * height * 16
    return ctx.wrapDeviceMemory(t.data_ptr(), width * height * 16, {"width": width, "height": height}, owner)
