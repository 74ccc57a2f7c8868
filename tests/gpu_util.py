"""Helpers for the -m gpu tests: everything goes through the Python mirror of the
reference's operator API, i.e. through the C ABI of libphaneron_b200.so."""
from __future__ import annotations

import asyncio

import numpy as np

from phaneron_b200 import ClProcessJobs, clContext
from phaneron_b200.process.image_process import ImageProcess


def run(coro):
    return asyncio.run(coro)


class Env:
    """one context + job queue, like index.ts:139-146"""

    def __init__(self, deferred=True):
        self.deferred = deferred

    async def __aenter__(self):
        self.ctx = clContext({"platformIndex": 0, "deviceIndex": 0, "overlapping": True, "deferred": self.deferred})
        await self.ctx.initialise()
        self.pj = ClProcessJobs(self.ctx)
        self.jobs = self.pj.getJobs()
        return self

    async def __aexit__(self, *exc):
        self.ctx.close()

    async def image(self, arr: np.ndarray, owner="img"):
        """an RGBA-f32 frame uploaded the way blackSilence.ts / tests do: createBuffer + hostAccess"""
        h, w, _ = arr.shape
        b = await self.ctx.createBuffer(w * h * 16, "readwrite", "coarse", {"width": w, "height": h}, owner)
        await b.hostAccess("writeonly", 0, np.ascontiguousarray(arr, np.float32))
        return b

    async def out_image(self, w, h, owner="out"):
        return await self.ctx.createBuffer(w * h * 16, "readwrite", "coarse", {"width": w, "height": h}, owner)

    async def fetch(self, buf, w, h) -> np.ndarray:
        await buf.hostAccess("readonly")
        return buf.host.view(np.float32).reshape(h, w, 4).copy()

    async def run_op(self, impl, params, w, h, sid="op"):
        ip = ImageProcess(self.ctx, impl, self.jobs)
        await ip.init()
        out = await self.out_image(w, h)
        await ip.run(dict(params, output=out), {"source": sid, "timestamp": 0}, lambda: None)
        await self.jobs.runQueue({"source": sid, "timestamp": 0})
        res = await self.fetch(out, w, h)
        out.release()
        return res


def rand_rgba(h, w, seed, lo=0.0, hi=1.0):
    rng = np.random.default_rng(seed)
    return (lo + (hi - lo) * rng.random((h, w, 4), dtype=np.float32)).astype(np.float32)


def assert_bits_equal(a: np.ndarray, b: np.ndarray, what=""):
    """bit-exact float comparison, treating +0/-0 as equal"""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape, (a.shape, b.shape)
    bad = a != b
    if bad.any():
        idx = np.argwhere(bad)[:5]
        raise AssertionError(f"{what}: {int(bad.sum())} of {a.size} floats differ, first at {idx.tolist()}: "
                             f"{a[tuple(idx[0])]!r} vs {b[tuple(idx[0])]!r}")
