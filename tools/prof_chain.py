"""profiling helper: record the bench scene once and replay it N times (for ncu)"""
import asyncio, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phaneron_b200 import clContext
from phaneron_b200.harness import ChannelHarness
from phaneron_b200.scenes import layered_scene, overlay_scene

async def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    inputs = sys.argv[2] if len(sys.argv) > 2 else "noise"
    w, h = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (3840, 2160)
    ctx = clContext({"deviceIndex": 0}); await ctx.initialise()
    kind = sys.argv[5] if len(sys.argv) > 5 else "north"
    scene = overlay_scene(w, h, inputs, "709", "2020") if kind == "overlay" else layered_scene(w, h, 4, inputs, "mix", "709", "2020")
    hs = ChannelHarness(ctx, scene); await hs.init()
    chain, dests = await hs.record_chain()
    for _ in range(n):
        chain.replay()
    await ctx.waitFinish(ctx.queue.process)
    print("done", ctx.stats())
asyncio.run(main())
