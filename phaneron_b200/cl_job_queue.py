"""Mirror of phaneron's src/clJobQueue.ts (ClJobs / ClProcessJobs) on asyncio.

Same contract: jobs are collected per (source, timestamp) key by add(); runQueue(id)
turns the list into one request, processed FIFO by ClProcessJobs.processQueue(), which
runs every job's program on queue.process, waits for the queue, fires the job callbacks
(these release input buffers, clJobQueue.ts:132) and resolves the runQueue() awaitable.

What changed underneath: runProgram on a deferred context only records RGBA-producing
jobs, so a request without a packed writer launches nothing and its waitFinish is free;
the request that carries the writer launches the whole chain as one fused kernel.
"""
from __future__ import annotations

import asyncio
import time
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional

from .nodencl import OpenCLProgram, RunTimings, clContext

JobCB = Callable[[], None]


@dataclass
class JobID:
    source: str
    timestamp: int


@dataclass
class ClJob:
    name: str
    program: OpenCLProgram
    params: Dict[str, Any]
    cb: JobCB


@dataclass
class JobsRequest:
    id: str
    jobs: List[ClJob]
    start: float
    done: Callable[[], None]
    error: Optional[Callable[[BaseException], None]] = None


def _as_id(id_) -> JobID:
    return id_ if isinstance(id_, JobID) else JobID(id_["source"], id_["timestamp"])


class ClJobs:
    """clJobQueue.ts:40-95"""

    def __init__(self, processJobs: "ClProcessJobs"):
        self.processJobs = processJobs
        self.jobs: Dict[str, List[ClJob]] = {}

    def makeKey(self, id_) -> str:
        id_ = _as_id(id_)
        return f"{id_.source} ts {id_.timestamp}"

    def add(self, id_, name: str, program: OpenCLProgram, params: Dict[str, Any], cb: JobCB) -> None:
        self.jobs.setdefault(self.makeKey(id_), []).append(ClJob(name, program, params, cb))

    def get(self, id_) -> Optional[List[ClJob]]:
        return self.jobs.get(self.makeKey(id_))

    def delete(self, id_) -> None:
        self.jobs.pop(self.makeKey(id_), None)

    def clear(self) -> None:
        self.jobs.clear()

    async def runQueue(self, id_) -> None:
        key = self.makeKey(id_)
        tsJobs = self.jobs.get(key)
        if not tsJobs:
            raise RuntimeError(f"Failed to run queue for id {key}")
        fut = asyncio.get_running_loop().create_future()

        def done() -> None:
            if not fut.done():
                fut.set_result(None)

        def error(e: BaseException) -> None:
            if not fut.done():
                fut.set_exception(e)

        self.processJobs.requestRun(key, JobsRequest(key, tsJobs, time.perf_counter(), done, error))
        self.delete(id_)
        await fut

    def clearQueue(self, src: str) -> None:
        for key, jobs in self.jobs.items():
            if key.startswith(src):
                # run the callbacks so sources are released
                for j in jobs:
                    j.cb()


class ClProcessJobs:
    """clJobQueue.ts:97-216"""

    def __init__(self, clContext_: clContext):
        self.clContext = clContext_
        self.requests: Dict[str, JobsRequest] = {}
        self.clJobs = ClJobs(self)
        self.showTimings = 0
        self._running: Optional[asyncio.Task] = None
        self.lastTimings: Dict[str, RunTimings] = {}

    async def processQueue(self) -> None:
        while self.requests:
            chan = next(iter(self.requests))
            req = self.requests[chan]
            timings: Dict[str, RunTimings] = {}
            jobQueued = time.perf_counter() - req.start
            try:
                for job in req.jobs:
                    timings[job.name] = await self.clContext.runProgram(
                        job.program, job.params, self.clContext.queue.process, timed=self.showTimings > 0)
                submit = time.perf_counter() - req.start
                await self.clContext.waitFinish(self.clContext.queue.process)
            except BaseException as e:  # reject the runQueue() promise instead of wedging the loop
                del self.requests[chan]
                for j in req.jobs:      # the release callbacks still run (as clearQueue does, clJobQueue.ts:87-94): no buffer leaks
                    try:
                        j.cb()
                    except Exception:
                        pass
                if req.error:
                    req.error(e)
                continue
            for j in req.jobs:
                j.cb()
            end = time.perf_counter() - req.start
            self.lastTimings = timings
            self.logTimings(req.id, jobQueued, submit, end, timings)
            req.done()
            del self.requests[chan]

    def getJobs(self) -> ClJobs:
        return self.clJobs

    def requestRun(self, id_: str, request: JobsRequest) -> None:
        self.requests[id_] = request
        if self._running is None or self._running.done():
            self._running = asyncio.get_running_loop().create_task(self.processQueue())

    def logRequests(self) -> None:
        for i, r in enumerate(self.requests.values()):
            print(f"{i}: {r.id} {[j.name for j in r.jobs]}")

    def logTimings(self, id_: str, jobQueued: float, submit: float, end: float, timings: Dict[str, RunTimings]) -> None:
        if self.showTimings <= 0:
            return
        if self.showTimings > 1:
            print(f"\n{id_[-20:]}: | toGPU | process | total (microseconds)")
            for name, t in timings.items():
                print(f"{name:<26}| {t.dataToKernel:>7} | {t.kernelExec:>7} | {t.totalTime:>7}")
        print(f"{id_[-20:]}: {end * 1e3:.2f}ms elapsed ({jobQueued * 1e3:.2f}ms job queued, "
              f"{(submit - jobQueued) * 1e3:.2f}ms submit, {(end - submit) * 1e3:.2f}ms execute)")
