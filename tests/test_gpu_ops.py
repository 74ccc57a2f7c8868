"""Parity of every stand-alone CUDA kernel with the oracle, through the reference-shaped
operator API (ToRGBA/FromRGBA/ImageProcess + ClJobs) and the C ABI.  Bit-exact: the packed
outputs as bytes, the RGBA-f32 intermediates as floats (tolerance stated: 0 ulp)."""
import numpy as np
import pytest

import oracle
from phaneron_b200 import PhaneronError
from phaneron_b200.process import rgba8, v210
from phaneron_b200.process.combine import Combine
from phaneron_b200.process.io import FromRGBA, ToRGBA
from phaneron_b200.process.mix import Mix
from phaneron_b200.process.packer import Interlace
from phaneron_b200.process.resize import Resize
from phaneron_b200.process.transform import Transform
from phaneron_b200.process.transition import Transition
from phaneron_b200.process.wipe import Wipe
from phaneron_b200.process.yadif import Yadif
from phaneron_b200.scenes import noise_frame, ramp_frame

from gpu_util import Env, assert_bits_equal, rand_rgba, run

pytestmark = pytest.mark.gpu


async def _read_write(env, w, h, src_bytes, spec_r, spec_w, interlaced=False):
    toRGBA = ToRGBA(env.ctx, spec_r, spec_w, v210.Reader(w, h), env.jobs)
    await toRGBA.init()
    fromRGBA = FromRGBA(env.ctx, spec_w, v210.Writer(w, h, interlaced), env.jobs)
    await fromRGBA.init()
    srcs = await toRGBA.createSources("")
    rgbaDst = await toRGBA.createDest({"width": w, "height": h}, "")
    dsts = await fromRGBA.createDests("")
    await toRGBA.loadFrame(src_bytes, srcs)
    rgbaDst.addRef()   # keep it for the read-back below
    toRGBA.processFrame("yuvRead", srcs, rgbaDst)
    await env.jobs.runQueue({"source": "yuvRead", "timestamp": 0})
    rgba = await env.fetch(rgbaDst, w, h)
    if not interlaced:
        fromRGBA.processFrame("yuvWrite", rgbaDst, dsts, Interlace.Progressive)
        await env.jobs.runQueue({"source": "yuvWrite", "timestamp": 0})
    else:
        dsts[0].fill(0)
        await dsts[0].hostAccess("writeonly")
        rgbaDst.addRef()
        fromRGBA.processFrame("yuvWrite", rgbaDst, dsts, Interlace.TopField)
        fromRGBA.processFrame("yuvWrite", rgbaDst, dsts, Interlace.BottomField)
        await env.jobs.runQueue({"source": "yuvWrite", "timestamp": 0})
    await fromRGBA.saveFrame(dsts)
    out = dsts[0].host.copy()
    rgbaDst.release()
    dsts[0].release()
    return rgba, out


@pytest.mark.parametrize("deferred", [False, True])
@pytest.mark.parametrize("w,h", [(1920, 1080), (1280, 720), (3840, 2160)])
def test_v210_ramp_roundtrip_like_the_reference_test_script(w, h, deferred):
    """src/process/test/yuv422p10Test.ts:40-112 applied to v210: fillBuf -> ToRGBA -> FromRGBA -> compare == 0"""
    async def go():
        async with Env(deferred) as env:
            src = np.empty(v210.getPitchBytes(w) * h, np.uint8)
            v210.fillBuf(src, w, h)
            rgba, out = await _read_write(env, w, h, src, "709", "709")
            ora = oracle.v210_read(src, w, h, oracle.ycbcr2rgb_matrix("709"), oracle.gamma2linear_lut("709"),
                                   oracle.rgb2rgb_matrix("709", "709"))
            assert_bits_equal(rgba, ora, "v210 read")
            oout = oracle.v210_write(ora, w, h, 0, oracle.rgb2ycbcr_matrix("709"), oracle.linear2gamma_lut("709"))
            assert np.array_equal(out, oout)
            if w % 6 == 0:
                assert np.array_equal(out, src)   # Compare returned 0
    run(go())


@pytest.mark.parametrize("deferred", [False, True])
def test_v210_interlaced_fields_fill_one_destination(deferred):
    """BASELINE config 1 on the GPU: TopField + BottomField writes into one dest reproduce the source"""
    async def go():
        async with Env(deferred) as env:
            w, h = 1920, 1080
            src = ramp_frame(w, h)
            _, out = await _read_write(env, w, h, src, "709", "709", interlaced=True)
            assert np.array_equal(out, src)
    run(go())


@pytest.mark.parametrize("spec_r,spec_w", [("709", "709"), ("709", "2020"), ("2020", "709"), ("601-625", "sRGB")])
def test_v210_noise_read_write_match_oracle(spec_r, spec_w):
    async def go():
        async with Env(False) as env:
            w, h = 1920, 270
            src = noise_frame(w, h, 11)
            rgba, out = await _read_write(env, w, h, src, spec_r, spec_w)
            ora = oracle.v210_read(src, w, h, oracle.ycbcr2rgb_matrix(spec_r), oracle.gamma2linear_lut(spec_r),
                                   oracle.rgb2rgb_matrix(spec_r, spec_w))
            assert_bits_equal(rgba, ora, "v210 read")
            oout = oracle.v210_write(ora, w, h, 0, oracle.rgb2ycbcr_matrix(spec_w), oracle.linear2gamma_lut(spec_w))
            assert np.array_equal(out, oout)
    run(go())


def test_v210_write_saturates_and_handles_nan():
    """convert_ushort_sat_rte edge cases: negatives, > 1, NaN, inf"""
    async def go():
        async with Env(False) as env:
            w, h = 96, 2
            img = rand_rgba(h, w, 5, -0.5, 1.5)
            img[0, 0] = [np.nan, np.inf, -np.inf, 0]
            img[0, 1] = [1e30, -1e30, 0.5, 0]
            fromRGBA = FromRGBA(env.ctx, "709", v210.Writer(w, h, False), env.jobs)
            await fromRGBA.init()
            src = await env.image(img)
            dsts = await fromRGBA.createDests("")
            fromRGBA.processFrame("w", src, dsts, Interlace.Progressive)
            await env.jobs.runQueue({"source": "w", "timestamp": 0})
            await fromRGBA.saveFrame(dsts)
            ref = oracle.v210_write(img, w, h, 0, oracle.rgb2ycbcr_matrix("709"), oracle.linear2gamma_lut("709"))
            assert np.array_equal(dsts[0].host, ref)
    run(go())


@pytest.mark.parametrize("bgra", [False, True])
def test_rgba8_read_write(bgra):
    async def go():
        async with Env(False) as env:
            w, h = 128, 6
            rng = np.random.default_rng(3)
            src = rng.integers(0, 256, w * h * 4, dtype=np.uint8)
            toRGBA = ToRGBA(env.ctx, "sRGB", "709", rgba8.Reader(w, h, bgra), env.jobs)
            await toRGBA.init()
            fromRGBA = FromRGBA(env.ctx, "sRGB", rgba8.Writer(w, h, False, bgra), env.jobs)
            await fromRGBA.init()
            srcs = await toRGBA.createSources("")
            dest = await toRGBA.createDest({"width": w, "height": h}, "")
            await toRGBA.loadFrame(src, srcs)
            dest.addRef()
            toRGBA.processFrame("r", srcs, dest)
            await env.jobs.runQueue({"source": "r", "timestamp": 0})
            rgba = await env.fetch(dest, w, h)
            ora = oracle.rgba8_read(src, w, h, oracle.gamma2linear_lut("sRGB"), oracle.rgb2rgb_matrix("sRGB", "709"), bgra)
            assert_bits_equal(rgba, ora, "rgba8 read")
            dsts = await fromRGBA.createDests("")
            fromRGBA.processFrame("w", dest, dsts, Interlace.Progressive)
            await env.jobs.runQueue({"source": "w", "timestamp": 0})
            await fromRGBA.saveFrame(dsts)
            assert np.array_equal(dsts[0].host, oracle.rgba8_write(ora, w, h, 0, oracle.linear2gamma_lut("sRGB"), bgra))
    run(go())


@pytest.mark.parametrize("deferred", [False, True])
@pytest.mark.parametrize("n", [2, 3, 4, 8])
def test_combine(n, deferred):
    async def go():
        async with Env(deferred) as env:
            w, h = 200, 37
            imgs = [rand_rgba(h, w, 10 + i) for i in range(n)]
            bufs = [await env.image(i) for i in imgs]
            got = await env.run_op(Combine(n, w, h), {"inputs": bufs}, w, h)
            assert_bits_equal(got, oracle.combine(imgs), f"combine_{n}")
    run(go())


def test_combine_needs_two_inputs():
    async def go():
        async with Env() as env:
            b = await env.image(rand_rgba(4, 4, 0))
            with pytest.raises(RuntimeError, match="at least 2"):
                await env.run_op(Combine(1, 4, 4), {"inputs": [b]}, 4, 4)
    run(go())


@pytest.mark.parametrize("deferred", [False, True])
def test_transition_dissolve_wipe_mix(deferred):
    async def go():
        async with Env(deferred) as env:
            w, h = 321, 45
            a, b, m = rand_rgba(h, w, 1), rand_rgba(h, w, 2), rand_rgba(h, w, 3)
            A, B, M = await env.image(a), await env.image(b), await env.image(m)
            for mix in (0.0, 0.25, 1.0 - 7 / 25, 1.0):
                got = await env.run_op(Transition("dissolve", w, h), {"inputs": [A, B], "mix": mix}, w, h)
                assert_bits_equal(got, oracle.dissolve(a, b, np.float32(mix)), f"dissolve {mix}")
            got = await env.run_op(Transition("wipe", w, h), {"inputs": [A, B], "mask": M}, w, h)
            assert_bits_equal(got, oracle.wipe_mask(a, b, m), "wipe")
            got = await env.run_op(Mix(w, h), {"input0": A, "input1": B, "mix": 0.3}, w, h)
            assert_bits_equal(got, oracle.mix(a, b, np.float32(0.3)), "mixer")
            got = await env.run_op(Wipe(w, h), {"input0": A, "input1": B, "wipe": 0.37}, w, h)
            assert_bits_equal(got, oracle.wipe(a, b, np.float32(0.37)), "wipe (dead op)")
            with pytest.raises(RuntimeError, match="expected a 'mask'"):
                await env.run_op(Transition("wipe", w, h), {"inputs": [A, B]}, w, h)
    run(go())


XFS = [
    {},   # identity: the half-pixel 2x2 box blur of Q6
    dict(anchorX=-0.5, anchorY=-0.5, scaleX=0.5, scaleY=0.5, offsetX=-0.05, offsetY=-0.05),
    dict(scaleX=1.7, scaleY=0.6, offsetX=0.13, offsetY=-0.21),
    dict(rotate=0.07, scaleX=0.8, scaleY=0.8, anchorX=0.1, anchorY=-0.2),
    dict(flipH=True, flipV=True, rotate=-0.31),
    dict(offsetX=3.0),   # entirely outside: all border
]


@pytest.mark.parametrize("deferred", [False, True])
@pytest.mark.parametrize("xf", XFS)
def test_transform(xf, deferred):
    async def go():
        async with Env(deferred) as env:
            sw, sh, w, h = 160, 90, 192, 108
            img = rand_rgba(sh, sw, 21)
            src = await env.image(img)
            got = await env.run_op(Transform(env.ctx, w, h), dict(input=src, **xf), w, h)
            from scene_oracle import xf_matrix
            assert_bits_equal(got, oracle.transform(img, xf_matrix(w, h, xf), w, h), f"transform {xf}")
    run(go())


# ---- Lanczos filter for Transform: NOT in the reference (BASELINE.json config 5); definition = oracle/oracle.c ----
LANCZOS_XFS = [
    dict(anchorX=-0.5, anchorY=-0.5),                                                    # scale 1: a = 1/2 everywhere (Q6)
    dict(anchorX=-0.5, anchorY=-0.5, scaleX=0.5, scaleY=0.5, offsetX=-0.2, offsetY=-0.1),   # 2x down-scale: widened support
    dict(anchorX=-0.5, anchorY=-0.5, scaleX=0.37, scaleY=0.61, offsetX=-0.3),
    dict(anchorX=-0.5, anchorY=-0.5, scaleX=1.7, scaleY=2.3, offsetX=0.2, offsetY=0.1),     # up-scale
    dict(anchorX=-0.5, anchorY=-0.5, scaleX=0.8, scaleY=0.8, flipH=True, flipV=True),
    dict(anchorX=-0.5, anchorY=-0.5, scaleX=0.5, scaleY=0.5, offsetX=-0.9, offsetY=0.7),    # mostly outside
]


@pytest.mark.parametrize("deferred", [False, True])
@pytest.mark.parametrize("lobes", [2, 3])
@pytest.mark.parametrize("xf", LANCZOS_XFS)
def test_transform_lanczos(xf, lobes, deferred):
    async def go():
        async with Env(deferred) as env:
            sw, sh, w, h = 160, 90, 192, 108
            img = rand_rgba(sh, sw, 23)
            src = await env.image(img)
            got = await env.run_op(Transform(env.ctx, w, h), dict(input=src, filter=f"lanczos{lobes}", **xf), w, h)
            from scene_oracle import xf_matrix
            assert_bits_equal(got, oracle.transform_lanczos(img, xf_matrix(w, h, xf), w, h, lobes), f"lanczos{lobes} {xf}")
    run(go())


def test_transform_lanczos_properties_and_errors():
    async def go():
        async with Env(True) as env:
            sw, sh, w, h = 96, 64, 96, 64
            flat = np.full((sh, sw, 4), 0.25, np.float32)
            src = await env.image(flat)
            got = await env.run_op(Transform(env.ctx, w, h), dict(input=src, filter="lanczos3", anchorX=-0.5, anchorY=-0.5, scaleX=0.75, scaleY=0.75), w, h)
            inside = got[8:36, 8:56]                        # well inside the shrunken picture (72 x 48, top left): weights sum to 1
            assert np.abs(inside - 0.25).max() < 1e-6
            assert got[-1, -1, 3] == 0.0                    # outside: border colour
            with pytest.raises(Exception, match="axis-aligned"):
                await env.run_op(Transform(env.ctx, w, h), dict(input=src, filter="lanczos3", rotate=0.1), w, h)
            with pytest.raises(Exception, match="taps"):
                await env.run_op(Transform(env.ctx, w, h), dict(input=src, filter="lanczos3", scaleX=0.05, scaleY=0.05), w, h)
            with pytest.raises(RuntimeError, match="lanczosN"):
                await env.run_op(Transform(env.ctx, w, h), dict(input=src, filter="bicubic"), w, h)
    run(go())


def test_resize():
    async def go():
        async with Env(False) as env:
            sw, sh, w, h = 160, 90, 128, 72
            img = rand_rgba(sh, sw, 22)
            src = await env.image(img)
            for scale, ox, oy, fh, fv in ((1.0, 0.0, 0.0, False, False), (0.5, 0.1, -0.2, True, False), (2.0, -0.3, 0.3, False, True)):
                got = await env.run_op(Resize(env.ctx, w, h), dict(input=src, scale=scale, offsetX=ox, offsetY=oy, flipH=fh, flipV=fv), w, h)
                flip = np.array([1.0 if fh else 0.0, -1.0 if fh else 1.0, 1.0 if fv else 0.0, -1.0 if fv else 1.0], np.float32)
                assert_bits_equal(got, oracle.resize(img, scale, ox, oy, flip, w, h), f"resize {scale}")
            with pytest.raises(RuntimeError, match="greater than zero"):
                await env.run_op(Resize(env.ctx, w, h), dict(input=src, scale=-1.0), w, h)
    run(go())


@pytest.mark.parametrize("deferred", [False, True])
@pytest.mark.parametrize("mode,tff", [("send_frame", True), ("send_field", True), ("send_field", False), ("send_field_nospatial", True)])
@pytest.mark.parametrize("size", [(96, 40), (150, 37), (7, 5)])   # (deferred: the tiled pre-pass k_yadif_rows -- 64 x 8 tiles -- at ragged widths, odd heights, frames smaller than its halo)
def test_yadif_window(mode, tff, deferred, size):
    """yadif.ts:115-145: 3-frame window, 1 or 2 outputs per input, parity rule of yadif.ts:104"""
    async def go():
        async with Env(deferred) as env:
            w, h = size
            frames = [rand_rgba(h, w, 40 + i) for i in range(5)]
            yad = Yadif(env.ctx, env.jobs, w, h, {"mode": mode, "tff": tff}, True)
            await yad.init()
            got = []
            for t, f in enumerate(frames):
                b = await env.image(f, f"in{t}")
                b.timestamp = t * 2
                outs = []
                if len(yad.in_) < 2:
                    # yadif.ts:128 flushes the producer's pending ToRGBA job for the first two
                    # frames; stand in for it with a throw-away job under the same id
                    from phaneron_b200.process.image_process import ImageProcess
                    from phaneron_b200.process.mix import Mix as _M
                    ip = ImageProcess(env.ctx, _M(w, h), env.jobs)
                    await ip.init()
                    tmp = await env.out_image(w, h)
                    await ip.run({"input0": b, "input1": b, "mix": 1.0, "output": tmp}, {"source": "y", "timestamp": b.timestamp}, lambda: None)
                await yad.processFrame(b, outs, "y")
                for o in outs:
                    got.append(await env.fetch(o, w, h))
                    o.release()
            skip = mode.endswith("nospatial")
            exp = []
            for i in range(1, 4):
                p, c, n = frames[i - 1], frames[i], frames[i + 1]
                exp.append(oracle.yadif(p, c, n, (1 if tff else 0) ^ 1, tff, skip))
                if mode.startswith("send_field"):
                    exp.append(oracle.yadif(p, c, n, (1 if tff else 0) ^ 0, tff, skip))
            assert len(got) == len(exp)
            for g, e in zip(got, exp):
                assert_bits_equal(g, e, f"yadif {mode}")
    run(go())


def test_progressive_yadif_is_passthrough():
    async def go():
        async with Env() as env:
            yad = Yadif(env.ctx, env.jobs, 8, 8, {"mode": "send_frame", "tff": True}, False)
            await yad.init()
            b = await env.image(rand_rgba(8, 8, 1))
            outs = []
            await yad.processFrame(b, outs, "y")
            assert outs == [b]
    run(go())


def test_errors_are_loud():
    async def go():
        async with Env() as env:
            b = await env.image(rand_rgba(4, 4, 0))
            with pytest.raises(PhaneronError, match="mode must be one of"):
                await b.hostAccess("sideways")
            with pytest.raises(PhaneronError, match="direction"):
                await env.ctx.createBuffer(16, "inout", "coarse")
            b.release()
            with pytest.raises(PhaneronError, match="released"):
                b.addRef()
            small = await env.ctx.createBuffer(64, "readwrite", "coarse", None, "small")
            prog = await env.ctx.createProgram(v210.Reader(1920, 1080).kernel, {"name": "read", "width": 1920, "height": 1080})
            with pytest.raises(PhaneronError, match="missing buffer parameter"):
                await env.ctx.runProgram(prog, {"input": small})
    run(go())


@pytest.mark.parametrize("deferred", [False, True])
@pytest.mark.parametrize("wipe,overlays", [(False, 1), (True, 2), (False, 0)])
def test_switch(wipe, overlays, deferred):
    """src/process/switch.ts (dead in the reference): Transform x2 -> Mix | Wipe -> Combine with overlays"""
    from phaneron_b200.process.switch import Switch
    from scene_oracle import xf_matrix

    async def go():
        async with Env(deferred) as env:
            w, h = 192, 108
            a, b = rand_rgba(h, w, 31), rand_rgba(h, w, 32)
            ovs = [rand_rgba(h, w, 40 + i) for i in range(overlays)]
            xfa = dict(anchorX=-0.5, anchorY=-0.5, scaleX=0.9, scaleY=0.9, offsetX=-0.05)
            xfb = dict(anchorX=-0.5, anchorY=-0.5, scaleX=1.2, scaleY=1.1, rotate=0.03)
            sw = Switch(env.ctx, "ch1", env.jobs, w, h, 2, overlays)
            await sw.init()
            ia, ib = await env.image(a), await env.image(b)
            ia.timestamp = ib.timestamp = 7
            ovb = [await env.image(o) for o in ovs]
            for bf in (sw.rgbaXf0, sw.rgbaXf1, sw.rgbaMx, *ovb):
                bf.addRef()   # the switcher's callbacks release what they consumed (switch.ts:147,158-175,186-188)
            out = await env.out_image(w, h)
            await sw.processFrame([dict(input=ia, **xfa), dict(input=ib, **xfb)], {"wipe": wipe, "frac": 0.4}, ovb, out)
            await env.jobs.runQueue({"source": "ch1 switch", "timestamp": 7})
            got = await env.fetch(out, w, h)
            ta, tb = oracle.transform(a, xf_matrix(w, h, xfa), w, h), oracle.transform(b, xf_matrix(w, h, xfb), w, h)
            mixed = oracle.wipe(ta, tb, 0.4) if wipe else oracle.mix(ta, tb, 0.4)
            ref = oracle.combine([mixed, *ovs]) if ovs else mixed
            assert_bits_equal(got, ref, f"switch wipe={wipe} overlays={overlays}")
    run(go())
