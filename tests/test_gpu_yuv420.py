"""8-bit 4:2:0 packers (yuv420p.ts, nv12.ts) through the C ABI: the reference's own test scripts
(src/process/test/yuv420pTest.ts, nv12Test.ts: fillBuf -> ToRGBA -> FromRGBA -> compare() == 0), bit-exactness against
the oracle on random frames (progressive and two-field writes, 718-wide tails), and -- where the OpenCL driver is
present -- the oracle against the reference's own kernels."""
import numpy as np
import pytest

import oracle
from oracle import ref_ocl
from phaneron_b200.process import nv12, yuv420p
from phaneron_b200.process.io import FromRGBA, ToRGBA
from phaneron_b200.process.packer import Interlace

from gpu_util import Env, run

pytestmark = pytest.mark.gpu

RANGE = (8, 16, 235, 224)


def _impl(is_nv12):
    return nv12 if is_nv12 else yuv420p


def _planes(buf, nb):
    out, o = [], 0
    for n in nb:
        out.append(buf[o: o + n])
        o += n
    return out


def _random_planes(is_nv12, w, h, seed):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 256, n, dtype=np.uint8) for n in oracle.yuv420_plane_bytes(is_nv12, w, h)]


async def _round_trip(env, is_nv12, w, h, planes, colRead="709", colWrite="709", fields=False):
    m = _impl(is_nv12)
    toRGBA = ToRGBA(env.ctx, colRead, colWrite, m.Reader(w, h), env.jobs)
    fromRGBA = FromRGBA(env.ctx, colWrite, m.Writer(w, h, fields), env.jobs)
    await toRGBA.init()
    await fromRGBA.init()
    srcs = await toRGBA.createSources("t")
    rgba = await toRGBA.createDest({"width": w, "height": h}, "t")
    dsts = await fromRGBA.createDests("t")
    assert len(srcs) == len(planes) == len(dsts)
    await toRGBA.loadFrame(planes, srcs)
    toRGBA.processFrame("yuvRead", srcs, rgba)
    await env.jobs.runQueue({"source": "yuvRead", "timestamp": 0})
    await rgba.hostAccess("readonly")
    rgba_host = rgba.host.view(np.float32).reshape(h, w, 4).copy()
    if fields:
        for d in dsts:
            d.fill(0)
            await d.hostAccess("writeonly")
        for il in (Interlace.TopField, Interlace.BottomField):
            rgba.addRef()
            fromRGBA.processFrame("yuvWrite", rgba, dsts, il)
            await env.jobs.runQueue({"source": "yuvWrite", "timestamp": 0})
    else:
        rgba.addRef()
        fromRGBA.processFrame("yuvWrite", rgba, dsts, Interlace.Progressive)
        await env.jobs.runQueue({"source": "yuvWrite", "timestamp": 0})
    await fromRGBA.saveFrame(dsts)
    return rgba_host, [d.host.copy() for d in dsts]


@pytest.mark.parametrize("is_nv12,w,h", [(False, 1920, 1080), (True, 1920, 1080), (False, 718, 270), (True, 718, 270)])
def test_reference_test_script_round_trip(is_nv12, w, h):
    """yuv420pTest.ts:40-112 / nv12Test.ts:40-104: `console.log('Compare returned', yuvSrc.compare(yuvDst))` must print 0"""
    src = oracle.yuv420_fill(is_nv12, w, h)
    mirror = np.zeros_like(src)
    _impl(is_nv12).fillBuf(mirror, w, h)
    assert np.array_equal(mirror, src)
    nb = oracle.yuv420_plane_bytes(is_nv12, w, h)

    async def go():
        async with Env(deferred=True) as env:
            return await _round_trip(env, is_nv12, w, h, _planes(src, nb))
    _, outs = run(go())
    assert np.array_equal(np.concatenate(outs), src)


@pytest.mark.parametrize("is_nv12,w,h,fields", [(False, 1280, 64, False), (True, 1280, 64, False), (False, 718, 48, False), (True, 718, 48, True),
                                                (False, 1920, 32, True), (True, 714, 20, False), (False, 722, 12, True)])
def test_read_and_write_bit_exact_vs_oracle(is_nv12, w, h, fields):
    planes = _random_planes(is_nv12, w, h, 41 + int(is_nv12))
    cm_r, cm_w = oracle.ycbcr2rgb_matrix("709", *RANGE), oracle.rgb2ycbcr_matrix("2020", *RANGE)
    lut_r, lut_w, gamut = oracle.gamma2linear_lut("709"), oracle.linear2gamma_lut("2020"), oracle.rgb2rgb_matrix("709", "2020")

    async def go():
        async with Env(deferred=True) as env:
            return await _round_trip(env, is_nv12, w, h, planes, "709", "2020", fields)
    rgba, outs = run(go())
    ref_rgba = oracle.yuv420_read(is_nv12, planes, w, h, cm_r, lut_r, gamut)
    assert np.array_equal(rgba.view(np.uint32), ref_rgba.view(np.uint32))
    if fields:
        ref = [np.zeros(n, np.uint8) for n in oracle.yuv420_plane_bytes(is_nv12, w, h)]
        oracle.yuv420_write(is_nv12, ref_rgba, w, h, 1, cm_w, lut_w, ref)
        oracle.yuv420_write(is_nv12, ref_rgba, w, h, 3, cm_w, lut_w, ref)
    else:
        ref = oracle.yuv420_write(is_nv12, ref_rgba, w, h, 0, cm_w, lut_w)
    for o, r in zip(outs, ref):
        assert np.array_equal(o, r)


def test_odd_height_is_rejected():
    """the reference launches height / 2 work-groups; a fractional NDRange is an error there, an explicit one here"""
    async def go():
        async with Env(deferred=True) as env:
            with pytest.raises(Exception, match="even height"):
                await _round_trip(env, False, 64, 7, [np.zeros(64 * 7, np.uint8), np.zeros(64 * 7 // 4 + 64, np.uint8), np.zeros(64 * 7 // 4 + 64, np.uint8)])
    run(go())


@pytest.mark.skipif(not ref_ocl.available(), reason="reference OpenCL kernels not runnable here")
@pytest.mark.parametrize("is_nv12,w,h", [(False, 1920, 64), (True, 1920, 64), (False, 718, 32), (True, 718, 32), (False, 714, 8), (True, 722, 8)])
def test_oracle_bit_exact_vs_reference_kernels(is_nv12, w, h):
    planes = _random_planes(is_nv12, w, h, 51 + int(is_nv12))
    cm_r, cm_w = oracle.ycbcr2rgb_matrix("709", *RANGE), oracle.rgb2ycbcr_matrix("709", *RANGE)
    lut_r, lut_w, gamut = oracle.gamma2linear_lut("709"), oracle.linear2gamma_lut("709"), oracle.rgb2rgb_matrix("709", "709")
    ref_rgba = ref_ocl.yuv420_read(is_nv12, planes, w, h, cm_r, lut_r, gamut)
    assert np.array_equal(ref_rgba.view(np.uint32), oracle.yuv420_read(is_nv12, planes, w, h, cm_r, lut_r, gamut).view(np.uint32))
    for il in (0, 1, 3):
        a = ref_ocl.yuv420_write(is_nv12, ref_rgba, w, h, il, cm_w, lut_w)
        b = oracle.yuv420_write(is_nv12, ref_rgba, w, h, il, cm_w, lut_w)
        for x, y in zip(a, b):
            assert np.array_equal(x, y), (is_nv12, w, il)
