"""A minimal stand-in for phaneron's graph layer (channel.ts / layer.ts / mixer.ts /
transitioner.ts / combiner.ts / macadamConsumer.ts) used by tests and bench.py: it calls
the src/process operators and the job queue in exactly the order the reference's callers
do, one runQueue per stage, so whatever works here works under the
real, unchanged TypeScript callers.

Scene description (plain dicts, shared with the oracle-side evaluator in tests/):
  scene = {width, height, colRead, colWork, interlace (0|None),
           layers: [ {src: np.uint8 v210 frame, sw, sh,
                      xf: None | {anchorX, anchorY, scaleX, scaleY, rotate, offsetX, offsetY, flipH, flipV},
                      transition: None | {type:'dissolve', mix, src, sw, sh, xf}
                                       | {type:'wipe', src, sw, sh, xf, mask, mask_sw, mask_sh, mask_xf}} ]}
"""
from __future__ import annotations

import asyncio
from typing import Any, Dict, List, Optional

import numpy as np

from .cl_job_queue import ClJobs, ClProcessJobs
from .nodencl import OpenCLBuffer, clContext
from .process import v210
from .process.combine import Combine
from .process.image_process import ImageProcess
from .process.io import FromRGBA, ToRGBA
from .process.packer import Interlace
from .process.transform import Transform
from .process.transition import Transition

_XF_DEFAULT = dict(flipH=False, flipV=False, anchorX=0.0, anchorY=0.0, scaleX=1.0, scaleY=1.0, rotate=0.0,
                   offsetX=0.0, offsetY=0.0)


def make_reader(fmt: str, w: int, h: int):
    """the Reader PackImpl of a source format (the formats MacadamProducer / FFmpegProducer hand to ToRGBA)"""
    from .process import nv12, rgba8, yuv420p, yuv422p8, yuv422p10
    if fmt == "v210":
        return v210.Reader(w, h)
    if fmt in ("rgba8", "bgra8"):
        return rgba8.Reader(w, h, fmt == "bgra8")
    mods = {"yuv422p10": yuv422p10, "yuv422p8": yuv422p8, "yuv420p": yuv420p, "nv12": nv12}
    if fmt not in mods:
        raise ValueError(f"unknown source format '{fmt}'")
    return mods[fmt].Reader(w, h)


def make_writer(fmt: str, w: int, h: int, interlaced: bool):
    """the Writer PackImpl of a consumer format (MacadamConsumer v210, FFmpegConsumer yuv422p8, ScreenConsumer rgba8 ...)"""
    from .process import nv12, rgba8, yuv420p, yuv422p8, yuv422p10
    if fmt == "v210":
        return v210.Writer(w, h, interlaced)
    if fmt in ("rgba8", "bgra8"):
        return rgba8.Writer(w, h, interlaced, fmt == "bgra8")
    mods = {"yuv422p10": yuv422p10, "yuv422p8": yuv422p8, "yuv420p": yuv420p, "nv12": nv12}
    if fmt not in mods:
        raise ValueError(f"unknown consumer format '{fmt}'")
    return mods[fmt].Writer(w, h, interlaced)


class _Source:
    """producer side of one input: ToRGBA (+ the Mixer's Transform)"""

    def __init__(self, h: "ChannelHarness", sid: str, sw: int, sh: int, xf: Optional[Dict[str, Any]], fmt: str = "v210",
                 colRead: Optional[str] = None):
        self.h, self.sid, self.sw, self.sh, self.xf = h, sid, sw, sh, xf
        # fmt 'rgbaf32': the source already is an RGBA-f32 frame in the working colour space (what a Yadif stage or any other
        # image process hands on): uploaded as an image buffer, no Reader
        self.rgbaf32 = fmt == "rgbaf32"
        self.toRGBA = None if self.rgbaf32 else ToRGBA(h.ctx, colRead or h.colRead, h.colWork, make_reader(fmt, sw, sh), h.clJobs)
        self.transform: Optional[ImageProcess] = None
        if xf is not None:
            self.transform = ImageProcess(h.ctx, Transform(h.ctx, h.width, h.height), h.clJobs)
        self.resident: Optional[List[OpenCLBuffer]] = None

    async def init(self) -> None:
        if self.toRGBA:
            await self.toRGBA.init()
        if self.transform:
            await self.transform.init()

    async def upload(self, frame: np.ndarray, timestamp: int) -> List[OpenCLBuffer]:
        if self.rgbaf32:
            img = await self.h.ctx.createBuffer(self.sw * self.sh * 16, "readwrite", "coarse", {"width": self.sw, "height": self.sh}, self.sid)
            img.timestamp = timestamp
            await img.hostAccess("writeonly", self.h.ctx.queue.load, np.ascontiguousarray(frame, np.float32).view(np.uint8).reshape(-1))
            return [img]
        srcs = await self.toRGBA.createSources(self.sid)        # macadamProducer.ts:171-191
        for s in srcs:
            s.timestamp = timestamp
        await self.toRGBA.loadFrame(frame, srcs, self.h.ctx.queue.load)
        await self.h.ctx.waitFinish(self.h.ctx.queue.load)
        return srcs

    async def frame(self, srcs: List[OpenCLBuffer], timestamp: int) -> OpenCLBuffer:
        h = self.h
        if self.rgbaf32:
            dest = srcs[0]
            if not self.transform:
                return dest
        else:
            dest = await self.toRGBA.createDest({"width": self.sw, "height": self.sh}, self.sid)   # macadamProducer.ts:193-210
            dest.timestamp = timestamp
            self.toRGBA.processFrame(self.sid, srcs, dest)
            if not self.transform:
                await h.clJobs.runQueue({"source": self.sid, "timestamp": timestamp})
                return dest
        xfDest = await h.ctx.createBuffer(h.width * h.height * 16, "readwrite", "coarse",          # mixer.ts:196-207
                                          {"width": h.width, "height": h.height}, f"mixer {self.sid} {timestamp}")
        xfDest.timestamp = timestamp
        p = dict(_XF_DEFAULT)
        p.update(self.xf or {})
        await self.transform.run(dict(input=dest, output=xfDest, **p), {"source": self.sid, "timestamp": timestamp},
                                 lambda: dest.release())                                          # mixer.ts:209-226
        await h.clJobs.runQueue({"source": self.sid, "timestamp": timestamp})
        return xfDest


class ChannelHarness:
    def __init__(self, ctx: clContext, scene: Dict[str, Any], processJobs: Optional[ClProcessJobs] = None, chanID: str = "ch1"):
        self.ctx = ctx
        self.scene = scene
        self.width, self.height = scene["width"], scene["height"]
        self.colRead, self.colWork = scene.get("colRead", "709"), scene.get("colWork", "709")
        self.interlaced = bool(scene.get("interlaced", False))
        self.chanID = chanID
        self.processJobs = processJobs or ClProcessJobs(ctx)
        self.clJobs: ClJobs = self.processJobs.getJobs()
        self.layers: List[Dict[str, Any]] = []
        self.combiner: Optional[ImageProcess] = None
        self.fromRGBA: Optional[FromRGBA] = None
        self.timestamp = 0

    async def init(self) -> None:
        for li, L in enumerate(self.scene["layers"]):
            ent: Dict[str, Any] = {"a": _Source(self, f"{self.chanID}-L{li}a", L["sw"], L["sh"], L.get("xf"), L.get("fmt", "v210"),
                                                L.get("colRead"))}
            t = L.get("transition")
            if t:
                ent["type"] = t["type"]
                ent["b"] = _Source(self, f"{self.chanID}-L{li}b", t["sw"], t["sh"], t.get("xf"))
                if t["type"] == "wipe":
                    ent["m"] = _Source(self, f"{self.chanID}-L{li}m", t["mask_sw"], t["mask_sh"], t.get("mask_xf"))
                ent["transition"] = ImageProcess(self.ctx, Transition(t["type"], self.width, self.height), self.clJobs)
                await ent["transition"].init()
            for k in ("a", "b", "m"):
                if k in ent:
                    await ent[k].init()
            self.layers.append(ent)
        if len(self.layers) >= 2:
            self.combiner = ImageProcess(self.ctx, Combine(len(self.layers), self.width, self.height), self.clJobs)
            await self.combiner.init()
        self.fromRGBA = FromRGBA(self.ctx, self.scene.get("colWrite", self.colWork),
                                 make_writer(self.scene.get("outFmt", "v210"), self.width, self.height, self.interlaced), self.clJobs)
        await self.fromRGBA.init()

    def _frames(self, li: int) -> Dict[str, np.ndarray]:
        L = self.scene["layers"][li]
        out = {"a": L["src"]}
        t = L.get("transition")
        if t:
            out["b"] = t["src"]
            if t["type"] == "wipe":
                out["m"] = t["mask"]
        return out

    async def upload_all(self, timestamp: int) -> List[Dict[str, List[OpenCLBuffer]]]:
        # every layer's producer runs its own pipe (producer/*.ts): the uploads of one frame time are concurrent
        jobs, where = [], []
        for li, ent in enumerate(self.layers):
            fr = self._frames(li)
            for k in fr:
                jobs.append(ent[k].upload(fr[k], timestamp))
                where.append((li, k))
        done = await asyncio.gather(*jobs)
        ups: List[Dict[str, List[OpenCLBuffer]]] = [{} for _ in self.layers]
        for (li, k), bufs in zip(where, done):
            ups[li][k] = bufs
        return ups

    async def compose(self, ups: List[Dict[str, List[OpenCLBuffer]]], timestamp: int) -> OpenCLBuffer:
        """layer pipes + combiner for one frame -> the channel's RGBA frame (possibly deferred)"""
        layerFrames: List[OpenCLBuffer] = []
        for li, ent in enumerate(self.layers):
            a = await ent["a"].frame(ups[li]["a"], timestamp)
            if "transition" not in ent:
                layerFrames.append(a)
                continue
            b = await ent["b"].frame(ups[li]["b"], timestamp)
            t = self.scene["layers"][li]["transition"]
            layerID = f"{self.chanID}-L{li}"
            dest = await self.ctx.createBuffer(self.width * self.height * 16, "readwrite", "coarse",        # transitioner.ts:152-163
                                               {"width": self.width, "height": self.height}, f"{layerID} {timestamp}")
            dest.timestamp = timestamp
            params: Dict[str, Any] = {"inputs": [a, b], "output": dest}
            extra = [a, b]
            if ent["type"] == "dissolve":
                params["mix"] = t["mix"]
            else:
                m = await ent["m"].frame(ups[li]["m"], timestamp)
                params["mask"] = m
                extra.append(m)
            await ent["transition"].run(params, {"source": layerID, "timestamp": timestamp}, lambda: None)   # transitioner.ts:176-182
            await self.clJobs.runQueue({"source": layerID, "timestamp": timestamp})
            for f in extra:   # transitioner.ts:199 frames.forEach(release)
                f.release()
            layerFrames.append(dest)
        if len(layerFrames) == 1:   # combiner.ts:219-228 passthrough
            return layerFrames[0]
        combineDest = await self.ctx.createBuffer(self.width * self.height * 16, "readwrite", "coarse",      # combiner.ts:230-241
                                                  {"width": self.width, "height": self.height}, f"{self.chanID} {timestamp}")
        combineDest.timestamp = timestamp
        await self.combiner.run({"inputs": layerFrames, "output": combineDest},
                                {"source": self.chanID, "timestamp": timestamp}, lambda: None)
        await self.clJobs.runQueue({"source": self.chanID, "timestamp": timestamp})
        for f in layerFrames:   # combiner.ts:257
            f.release()
        return combineDest

    async def consume(self, frame: OpenCLBuffer, dests: Optional[List[OpenCLBuffer]] = None,
                      interlace: Optional[Interlace] = None, download: bool = True) -> List[OpenCLBuffer]:
        """macadamConsumer.ts:220-260: FromRGBA.processFrame + runQueue (+ saveFrame)"""
        if dests is None:
            dests = await self.fromRGBA.createDests(self.chanID)
        cid = f"{self.chanID}-out"
        frame_ts = frame.timestamp
        self.fromRGBA.processFrame(cid, frame, dests, interlace)
        await self.clJobs.runQueue({"source": cid, "timestamp": frame_ts})
        if download:
            await self.fromRGBA.saveFrame(dests, self.ctx.queue.unload)
            await self.ctx.waitFinish(self.ctx.queue.unload)
        return dests

    async def run_frame(self, download: bool = True) -> np.ndarray:
        """one whole frame through the public API, host buffers in, host buffer out"""
        ts = self.timestamp
        self.timestamp += 1
        ups = await self.upload_all(ts)
        frame = await self.compose(ups, ts)
        dests = await self.consume(frame, download=download)
        out = (dests[0].host.copy() if len(dests) == 1 else np.concatenate([d.host for d in dests])) if download else None
        for d in dests:
            d.release()
        return out

    async def record_chain(self):
        """Upload once, run one frame while recording fused launches; returns (chain, dests, ups).
        The caller keeps `ups`/`dests` alive implicitly through the chain."""
        ts = self.timestamp
        self.timestamp += 1
        ups = await self.upload_all(ts)
        self.ctx.beginChain()
        try:
            frame = await self.compose(ups, ts)
            dests = await self.consume(frame, download=False)
        finally:
            chain = self.ctx.endChain()
        return chain, dests

    def algorithmic_bytes(self) -> int:
        """(distinct packed inputs read + 1 packed output) x packed frame bytes, in each one's own format"""
        total = sum(make_writer(self.scene.get("outFmt", "v210"), self.width, self.height, self.interlaced).numBytes)
        for L in self.scene["layers"]:
            total += sum(make_reader(L.get("fmt", "v210"), L["sw"], L["sh"]).numBytes)
            t = L.get("transition")
            if t:
                total += v210.getPitchBytes(t["sw"]) * t["sh"]
                if t["type"] == "wipe":
                    total += v210.getPitchBytes(t["mask_sw"]) * t["mask_sh"]
        return total
