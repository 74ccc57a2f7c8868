import asyncio, sys, os
sys.path.insert(0, "/root/repo")
from phaneron_b200 import clContext
from phaneron_b200.harness import ChannelHarness
from phaneron_b200.scenes import single_layer_scene
async def main():
    ctx = clContext({"deviceIndex": 0}); await ctx.initialise()
    hs = ChannelHarness(ctx, single_layer_scene(3840, 2160, "noise", True, "709", "2020")); await hs.init()
    chain, dests = await hs.record_chain()
    for _ in range(4): chain.replay()
    await ctx.waitFinish(ctx.queue.process)
asyncio.run(main())
