"""where does a pipelined end-to-end frame spend its time? (host-side breakdown of bench.py's e2e loop)"""
import asyncio, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from phaneron_b200 import _lib, clContext
from phaneron_b200.harness import ChannelHarness
from phaneron_b200.scenes import layered_scene

async def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    lib = _lib.lib()
    ctx = clContext({"deviceIndex": 0}); await ctx.initialise()
    scene = bench.pin_scene(lib, layered_scene(3840, 2160, 4, "noise", "mix", "709", "2020"))
    he = ChannelHarness(ctx, scene, chanID="e2e"); await he.init()
    for _ in range(3):
        await he.run_frame()
    n_all = 48
    pending = [asyncio.ensure_future(he.upload_all(1000 + j)) for j in range(depth)]
    T = {"wait_up": 0.0, "compose": 0.0, "run": 0.0, "save": 0.0, "rel": 0.0}
    t_start = None
    for i in range(n_all):
        if i == 8:
            t_start = time.perf_counter(); T = {k: 0.0 for k in T}
        t0 = time.perf_counter()
        ups = await pending.pop(0)
        if i + depth < n_all:
            pending.append(asyncio.ensure_future(he.upload_all(1000 + i + depth)))
        t1 = time.perf_counter()
        frame = await he.compose(ups, 1000 + i)
        t2 = time.perf_counter()
        dests = await he.fromRGBA.createDests(he.chanID)
        cid = f"{he.chanID}-out"; ts = frame.timestamp
        he.fromRGBA.processFrame(cid, frame, dests, None)
        await he.clJobs.runQueue({"source": cid, "timestamp": ts})
        t3 = time.perf_counter()
        await he.fromRGBA.saveFrame(dests, ctx.queue.unload)
        await ctx.waitFinish(ctx.queue.unload)
        t4 = time.perf_counter()
        for d in dests: d.release()
        t5 = time.perf_counter()
        T["wait_up"] += t1 - t0; T["compose"] += t2 - t1; T["run"] += t3 - t2; T["save"] += t4 - t3; T["rel"] += t5 - t4
    dt = time.perf_counter() - t_start
    n = n_all - 8
    print(f"depth {depth}: {n / dt:.0f} fps, per frame ms:", {k: round(v / n * 1e3, 3) for k, v in T.items()}, "total", round(dt / n * 1e3, 3))
asyncio.run(main())
