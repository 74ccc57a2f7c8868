"""debug helper: run a scene with the strip kernel on/off and show where outputs differ"""
import asyncio, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phaneron_b200 import clContext, ClProcessJobs
from phaneron_b200.harness import ChannelHarness
from phaneron_b200.scenes import layered_scene, single_layer_scene

def unpack(buf, w, h):
    g = buf.view(np.uint32).reshape(h, -1, 4)[:, : w // 6]
    Y = np.stack([(g[..., 0] >> 10) & 1023, g[..., 1] & 1023, (g[..., 1] >> 20) & 1023, (g[..., 2] >> 10) & 1023, g[..., 3] & 1023, (g[..., 3] >> 20) & 1023], -1).reshape(h, w)
    Cb = np.stack([g[..., 0] & 1023, (g[..., 1] >> 10) & 1023, (g[..., 2] >> 20) & 1023], -1).reshape(h, w // 2)
    Cr = np.stack([(g[..., 0] >> 20) & 1023, g[..., 2] & 1023, (g[..., 3] >> 10) & 1023], -1).reshape(h, w // 2)
    return Y.astype(int), Cb.astype(int), Cr.astype(int)

async def one(scene, strip):
    ctx = clContext({"deviceIndex": 0}); await ctx.initialise(); ctx.setStripKernel(strip)
    h = ChannelHarness(ctx, scene, ClProcessJobs(ctx)); await h.init()
    out = await h.run_frame(); st = ctx.stats(); ctx.close(); return out, st

async def main():
    w, hh = 480, 270
    for name, scene in (("1layer-direct", single_layer_scene(w, hh, "noise", False)),
                        ("1layer-identity", single_layer_scene(w, hh, "noise", True)),
                        ("2layer", layered_scene(w, hh, 2, "noise", "plain", "709", "709")),
                        ("4layer", layered_scene(w, hh, 4, "noise", "plain", "709", "2020"))):
        a, sa = await one(scene, True); b, sb = await one(scene, False)
        Ya, Cba, Cra = unpack(a, w, hh); Yb, Cbb, Crb = unpack(b, w, hh)
        dy, dcb, dcr = Ya != Yb, Cba != Cbb, Cra != Crb
        print(name, "strip launches", sa["strip_launches"], "Y diffs", dy.sum(), "Cb diffs", dcb.sum(), "Cr diffs", dcr.sum())
        if dy.any():
            ys, xs = np.nonzero(dy)
            print("  first Y diffs (y,x,strip,generic):", [(int(y), int(x), int(Ya[y, x]), int(Yb[y, x])) for y, x in list(zip(ys, xs))[:12]])
            print("  rows with diffs:", np.unique(ys)[:40], " cols mod 192:", np.unique(xs % 192)[:40])
            print("  max |dY|", np.abs(Ya - Yb).max())
asyncio.run(main())
