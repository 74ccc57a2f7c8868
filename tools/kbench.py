"""kernel experiments: time the fused launch of a scene for several kernel variants in one process.

  python tools/kbench.py [--scene north|single|single_xf] [--size 3840x2160] [--inputs noise,ramp]
                         [--kernels march,march_raw,generic] [--frames 60]
Prints one line per (inputs, kernel): us/frame, frames/s, achieved GB/s (algorithmic bytes), fraction
of the measured HBM peak.  Inputs rotate over enough sets to exceed L2.
"""
import argparse, asyncio, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phaneron_b200 import clContext
from phaneron_b200.harness import ChannelHarness
from phaneron_b200.scenes import layered_scene, overlay_scene, planar_layered_scene, single_layer_scene

L2 = 126 << 20


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def make_scene(kind, w, h, inputs, fs):
    if kind == "north":
        return layered_scene(w, h, 4, inputs, "mix", "709", "2020", frame_set=fs)
    if kind == "plain4":
        return layered_scene(w, h, 4, inputs, "plain", "709", "2020", frame_set=fs)
    if kind == "two":
        return layered_scene(w, h, 2, inputs, "plain", "709", "2020", frame_set=fs)
    if kind == "single":
        return single_layer_scene(w, h, inputs, False, "709", "709", frame_set=fs)
    if kind == "single_xf":
        return single_layer_scene(w, h, inputs, True, "709", "709", frame_set=fs)
    if kind == "rot":   # the bench scene with its second layer rotated (the Mixer's DVE rotation)
        sc = layered_scene(w, h, 4, inputs, "mix", "709", "2020", frame_set=fs)
        sc["layers"][1]["xf"] = dict(sc["layers"][1]["xf"], rotate=0.04)
        return sc
    if kind == "overlay":   # 3 video layers + a full-frame rgba8 graphic with alpha
        return overlay_scene(w, h, inputs, "709", "2020", frame_set=fs)
    if kind.startswith("planar1:"):   # e.g. planar1:yuv422p10 -- one FFmpegProducer-format clip through the Mixer's identity Transform
        return planar_layered_scene(w, h, kind.split(":")[1], 1, "709", "2020", frame_set=fs)
    if kind.startswith("planar4:"):   # e.g. planar4:yuv422p10 -- 4 layers of an FFmpegProducer format
        return planar_layered_scene(w, h, kind.split(":")[1], 4, "709", "2020", frame_set=fs)
    raise SystemExit(f"unknown scene {kind}")


async def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="north")
    ap.add_argument("--size", default="3840x2160")
    ap.add_argument("--inputs", default="noise,ramp")
    ap.add_argument("--kernels", default="march,march_raw,generic")
    ap.add_argument("--frames", type=int, default=60)
    a = ap.parse_args()
    w, h = (int(v) for v in a.size.split("x"))
    pk = peak()
    for inputs in a.inputs.split(","):
        for kern in a.kernels.split(","):
            ctx = clContext({"deviceIndex": 0, "marchKernel": kern != "generic", "rawLut": kern == "march_raw",
                             "directKernel": os.environ.get("PB_NO_DIRECT") is None,      # A/B: the dedicated single-layer kernels
                             "occlusionCulling": os.environ.get("PB_NO_CULL") is None})
            await ctx.initialise()
            hs, chains, keep = [], [], []
            h0 = ChannelHarness(ctx, make_scene(a.scene, w, h, inputs, 0))
            set_bytes = h0.algorithmic_bytes()
            n_sets = max(2, -(-2 * L2 // set_bytes) + 1)
            for s in range(n_sets):
                hh = h0 if s == 0 else ChannelHarness(ctx, make_scene(a.scene, w, h, inputs, s), chanID=f"s{s}")
                await hh.init()
                chain, dests = await hh.record_chain()
                assert chain.complete
                hs.append(hh); chains.append(chain); keep.append(dests)
            await ctx.waitFinish(ctx.queue.process)
            for i in range(6):
                chains[i % n_sets].replay()
            await ctx.waitFinish(ctx.queue.process)
            e0, e1 = ctx.createEvent(), ctx.createEvent()
            e0.record()
            for i in range(a.frames):
                chains[i % n_sets].replay()
            e1.record(); e1.synchronize()
            ms = e0.elapsed_ms(e1)
            us = ms * 1e3 / a.frames
            st = ctx.stats()
            gbs = set_bytes / (us * 1e-6) / 1e9
            print(f"{a.scene:9s} {w}x{h} {inputs:5s} {kern:9s} {us:9.1f} us/frame {1e6 / us:9.0f} fps {gbs:8.1f} GB/s "
                  f"frac {gbs / pk:6.3f}  launches/frame {chains[0].launches} march {st['march_launches']} "
                  f"luts {st['lut_tables']}/{st['lut_tables_d8']}/poly {st['lut_tables_poly']} sets {n_sets}", flush=True)
            for c in chains:
                c.destroy() if hasattr(c, "destroy") else None
            ctx.close()

asyncio.run(main())
