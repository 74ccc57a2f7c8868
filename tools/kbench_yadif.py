"""De-interlacing producer timing (VERDICT r1 missing #3):

  interlaced v210 frames -> ToRGBA -> Yadif send_field (2 output frames per input frame) -> Mixer Transform (identity) ->
  Combine with a 0.5x PiP -> FromRGBA v210

  python tools/kbench_yadif.py [--size 1920x1080] [--frames 100]

fused : the product's deferred mode.  Per input frame: ONE direct-kernel launch makes the new ToRGBA output real; per output frame
        a pre-pass computes the interpolated lines of the field (half a frame, each pixel once) and ONE march launch composites,
        reading the field's other lines from the current frame in place.  Timed by replaying the recorded launches of one input
        frame (CUDA events).
eager : the reference's launch structure (one kernel per job, RGBA-f32 frames between all stages): sum of the per-job kernel
        times the library reports (RunTimings.kernelExec, CUDA events around each launch), host time excluded.
"""
import argparse, asyncio, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phaneron_b200 import ClProcessJobs, clContext
from phaneron_b200.process import v210
from phaneron_b200.process.combine import Combine
from phaneron_b200.process.image_process import ImageProcess
from phaneron_b200.process.io import FromRGBA, ToRGBA
from phaneron_b200.process.packer import Interlace
from phaneron_b200.process.transform import Transform
from phaneron_b200.process.yadif import Yadif
from phaneron_b200.scenes import IDENTITY_XF, make_frame, pip


class Rig:
    def __init__(self, ctx, w, h):
        self.ctx, self.w, self.h = ctx, w, h
        self.pj = ClProcessJobs(ctx)
        self.jobs = self.pj.getJobs()
        self.to_a = ToRGBA(ctx, "709", "2020", v210.Reader(w, h), self.jobs)
        self.to_b = ToRGBA(ctx, "709", "2020", v210.Reader(w, h), self.jobs)
        self.frm = FromRGBA(ctx, "2020", v210.Writer(w, h, False), self.jobs)
        self.xa = ImageProcess(ctx, Transform(ctx, w, h), self.jobs)
        self.xb = ImageProcess(ctx, Transform(ctx, w, h), self.jobs)
        self.comb = ImageProcess(ctx, Combine(2, w, h), self.jobs)
        self.yad = Yadif(ctx, self.jobs, w, h, {"mode": "send_field", "tff": True}, True)
        self.t = 0

    async def init(self):
        for o in (self.to_a, self.to_b, self.frm, self.xa, self.xb, self.comb, self.yad):
            await o.init()
        self.pip_srcs = await self.to_b.createSources("pip")
        await self.to_b.loadFrame(make_frame("noise", self.w, self.h, 9), self.pip_srcs, self.ctx.queue.load)
        self.src_frames = []
        for i in range(4):
            s = await self.to_a.createSources("src")
            await self.to_a.loadFrame(make_frame("noise", self.w, self.h, 20 + i), s, self.ctx.queue.load)
            self.src_frames.append(s)
        await self.ctx.waitFinish(self.ctx.queue.load)

    async def input_frame(self):
        """one interlaced frame arrives: producer side + the channel's work for every de-interlaced frame it yields"""
        ctx, jobs, w, h = self.ctx, self.jobs, self.w, self.h
        dims = {"width": w, "height": h}
        t = self.t
        self.t += 1
        srcs = self.src_frames[t % len(self.src_frames)]
        for s in srcs:
            s.addRef()
            s.timestamp = t * 2
        rgba = await self.to_a.createDest(dims, "src")
        rgba.timestamp = t * 2
        self.to_a.processFrame("src", srcs, rgba)
        if len(self.yad.in_) >= 2:
            await jobs.runQueue({"source": "src", "timestamp": t * 2})
        outs = []
        await self.yad.processFrame(rgba, outs, "src")
        n = 0
        for deint in outs:
            ts = deint.timestamp
            xfa = await ctx.createBuffer(w * h * 16, "readwrite", "coarse", dims, "mixer a")
            await self.xa.run(dict(input=deint, output=xfa, **IDENTITY_XF), {"source": "L0", "timestamp": ts}, lambda d=deint: d.release())
            await jobs.runQueue({"source": "L0", "timestamp": ts})
            for s in self.pip_srcs:
                s.addRef()
                s.timestamp = ts
            rgb = await self.to_b.createDest(dims, "pip")
            self.to_b.processFrame("pip", self.pip_srcs, rgb)
            xfb = await ctx.createBuffer(w * h * 16, "readwrite", "coarse", dims, "mixer b")
            await self.xb.run(dict(input=rgb, output=xfb, **pip(0.5, 0.3, 0.2)), {"source": "pip", "timestamp": ts}, lambda r_=rgb: r_.release())
            await jobs.runQueue({"source": "pip", "timestamp": ts})
            cdest = await ctx.createBuffer(w * h * 16, "readwrite", "coarse", dims, "comb")
            cdest.timestamp = ts
            await self.comb.run({"inputs": [xfa, xfb], "output": cdest}, {"source": "ch", "timestamp": ts}, lambda: None)
            await jobs.runQueue({"source": "ch", "timestamp": ts})
            xfa.release()
            xfb.release()
            if not hasattr(self, "dests"):
                self.dests = await self.frm.createDests("out")
            self.frm.processFrame("out", cdest, self.dests, Interlace.Progressive)
            await jobs.runQueue({"source": "out", "timestamp": ts})
            n += 1
        return n


async def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("--frames", type=int, default=100)
    a = ap.parse_args()
    w, h = (int(v) for v in a.size.split("x"))
    packed = v210.getPitchBytes(w) * h
    # ---- fused (deferred) ----
    ctx = clContext({"deviceIndex": 0})
    await ctx.initialise()
    rig = Rig(ctx, w, h)
    await rig.init()
    for _ in range(4):
        await rig.input_frame()
    await ctx.waitFinish(ctx.queue.process)
    st0 = ctx.stats()
    ctx.beginChain()
    fields = await rig.input_frame()
    chain = ctx.endChain()
    st1 = ctx.stats()
    assert chain.complete and fields == 2, (chain.complete, fields)
    for _ in range(5):
        chain.replay()
    await ctx.waitFinish(ctx.queue.process)
    e0, e1 = ctx.createEvent(), ctx.createEvent()
    e0.record()
    for _ in range(a.frames):
        chain.replay()
    e1.record(); e1.synchronize()
    us = e0.elapsed_ms(e1) * 1e3 / a.frames
    print(f"yadif fused  {w}x{h}: {us:8.1f} us per input frame = {us / 2:7.1f} us per output field ({2e6 / us:7.0f} fields/s); launches per input frame "
          f"{chain.launches} (materialised {st1['materialised'] - st0['materialised']}, march {st1['march_launches'] - st0['march_launches']}); "
          f"packed bytes per field in+out {2 * packed}", flush=True)
    ctx.close()
    # ---- eager: the reference's launch structure ----
    ctx = clContext({"deviceIndex": 0, "deferred": False})
    await ctx.initialise()
    total = {"us": 0, "n": 0}
    per = {}
    orig = ctx.runProgram

    async def timed_run(program, params, queue, timed=False):
        t = await orig(program, params, queue, timed=True)
        total["us"] += t.kernelExec
        total["n"] += 1
        name = getattr(program, "name", None) or getattr(program, "op_name", None) or str(program)
        per[name] = per.get(name, 0) + t.kernelExec
        return t
    ctx.runProgram = timed_run
    rig = Rig(ctx, w, h)
    await rig.init()
    for _ in range(4):
        await rig.input_frame()
    total.update(us=0, n=0)
    per.clear()
    m = max(10, a.frames // 5)
    for _ in range(m):
        await rig.input_frame()
    print(f"yadif eager  {w}x{h}: {total['us'] / m:8.1f} us per input frame = {total['us'] / m / 2:7.1f} us per output field; kernels per input frame "
          f"{total['n'] / m:.0f} (sum of per-kernel CUDA-event times, host time excluded)", flush=True)
    print("             per input frame: " + ", ".join(f"{k} {v / m:.1f}" for k, v in sorted(per.items(), key=lambda kv: -kv[1])), flush=True)
    ctx.close()

asyncio.run(main())
