"""profiling helper: one 2160p v210 layer, read 1:1 ('direct') or through the Mixer's identity Transform ('single'), replayed for ncu"""
import asyncio, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phaneron_b200 import clContext
from phaneron_b200.harness import ChannelHarness
from phaneron_b200.scenes import single_layer_scene
async def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "single"
    ctx = clContext({"deviceIndex": 0}); await ctx.initialise()
    hs = ChannelHarness(ctx, single_layer_scene(3840, 2160, "noise", kind == "single", "709", "2020")); await hs.init()
    chain, dests = await hs.record_chain()
    for _ in range(4): chain.replay()
    await ctx.waitFinish(ctx.queue.process)
asyncio.run(main())
