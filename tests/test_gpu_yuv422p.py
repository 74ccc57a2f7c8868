"""Planar 4:2:2 packers (yuv422p10.ts, yuv422p8.ts) through the C ABI: the reference's own test
scripts (src/process/test/yuv422p10Test.ts, yuv422p8Test.ts: fillBuf -> ToRGBA -> FromRGBA -> compare() == 0, the latter
at width 718 to exercise the line tails), bit-exactness against the oracle on random frames, and -- where the OpenCL driver
is present -- the oracle against the reference's own kernels."""
import numpy as np
import pytest

import oracle
from oracle import ref_ocl
from phaneron_b200.process import yuv422p8, yuv422p10
from phaneron_b200.process.io import FromRGBA, ToRGBA
from phaneron_b200.process.packer import Interlace

from gpu_util import Env, run

pytestmark = pytest.mark.gpu

RANGE = {10: (10, 64, 940, 896), 8: (8, 16, 235, 224)}


def _impl(bits):
    return yuv422p10 if bits == 10 else yuv422p8


def _planes(buf, nb):
    return [buf[: nb[0]], buf[nb[0]: nb[0] + nb[1]], buf[nb[0] + nb[1]: nb[0] + nb[1] + nb[2]]]


def _random_planes(bits, w, h, seed):
    rng = np.random.default_rng(seed)
    nb = oracle.yuv422p_plane_bytes(bits, w, h)
    hi = 1024 if bits == 10 else 256
    out = []
    for n in nb:
        samples = rng.integers(0, hi, n // (2 if bits == 10 else 1), dtype=np.uint16)
        out.append(samples.astype("<u2").view(np.uint8) if bits == 10 else samples.astype(np.uint8))
    return out


async def _round_trip(env, bits, w, h, planes, colRead="709", colWrite="709", fields=False):
    m = _impl(bits)
    toRGBA = ToRGBA(env.ctx, colRead, colWrite, m.Reader(w, h), env.jobs)
    fromRGBA = FromRGBA(env.ctx, colWrite, m.Writer(w, h, fields), env.jobs)
    await toRGBA.init()
    await fromRGBA.init()
    srcs = await toRGBA.createSources("t")
    rgba = await toRGBA.createDest({"width": w, "height": h}, "t")
    dsts = await fromRGBA.createDests("t")
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


@pytest.mark.parametrize("bits,w,h", [(10, 1920, 1080), (8, 718, 1080), (8, 1920, 1080), (10, 718, 270)])
def test_reference_test_script_round_trip(bits, w, h):
    """yuv422p10Test.ts:40-112 / yuv422p8Test.ts:40-112: `console.log('Compare returned', yuvSrc.compare(yuvDst))` must print 0"""
    src = oracle.yuv422p_fill(bits, w, h)
    mirror = np.zeros_like(src)
    _impl(bits).fillBuf(mirror, w, h)
    assert np.array_equal(mirror, src)
    nb = oracle.yuv422p_plane_bytes(bits, w, h)

    async def go():
        async with Env(deferred=True) as env:
            return await _round_trip(env, bits, w, h, _planes(src, nb))
    _, outs = run(go())
    assert np.array_equal(np.concatenate(outs), src)


@pytest.mark.parametrize("bits,w,h,fields", [(10, 1280, 64, False), (8, 718, 48, False), (10, 718, 48, True), (8, 1920, 32, True), (10, 714, 20, False)])
def test_read_and_write_bit_exact_vs_oracle(bits, w, h, fields):
    planes = _random_planes(bits, w, h, 21 + bits)
    cm_r, cm_w = oracle.ycbcr2rgb_matrix("709", *RANGE[bits]), oracle.rgb2ycbcr_matrix("2020", *RANGE[bits])
    lut_r, lut_w, gamut = oracle.gamma2linear_lut("709"), oracle.linear2gamma_lut("2020"), oracle.rgb2rgb_matrix("709", "2020")

    async def go():
        async with Env(deferred=True) as env:
            return await _round_trip(env, bits, w, h, planes, "709", "2020", fields)
    rgba, outs = run(go())
    ref_rgba = oracle.yuv422p_read(bits, *planes, w, h, cm_r, lut_r, gamut)
    assert np.array_equal(rgba.view(np.uint32), ref_rgba.view(np.uint32))
    if fields:
        ref = [np.zeros(n, np.uint8) for n in oracle.yuv422p_plane_bytes(bits, w, h)]
        oracle.yuv422p_write(bits, ref_rgba, w, h, 1, cm_w, lut_w, ref)
        oracle.yuv422p_write(bits, ref_rgba, w, h, 3, cm_w, lut_w, ref)
    else:
        ref = oracle.yuv422p_write(bits, ref_rgba, w, h, 0, cm_w, lut_w)
    for o, r in zip(outs, ref):
        assert np.array_equal(o, r)


@pytest.mark.skipif(not ref_ocl.available(), reason="reference OpenCL kernels not runnable here")
@pytest.mark.parametrize("bits,w,h", [(10, 1920, 64), (8, 718, 64), (10, 718, 32), (8, 1920, 16)])
def test_oracle_bit_exact_vs_reference_kernels(bits, w, h):
    planes = _random_planes(bits, w, h, 31 + bits)
    cm_r, cm_w = oracle.ycbcr2rgb_matrix("709", *RANGE[bits]), oracle.rgb2ycbcr_matrix("709", *RANGE[bits])
    lut_r, lut_w, gamut = oracle.gamma2linear_lut("709"), oracle.linear2gamma_lut("709"), oracle.rgb2rgb_matrix("709", "709")
    ref_rgba = ref_ocl.yuv422p_read(bits, *planes, w, h, cm_r, lut_r, gamut)
    assert np.array_equal(ref_rgba.view(np.uint32), oracle.yuv422p_read(bits, *planes, w, h, cm_r, lut_r, gamut).view(np.uint32))
    for il in (0, 1, 3):
        a = ref_ocl.yuv422p_write(bits, ref_rgba, w, h, il, cm_w, lut_w)
        b = oracle.yuv422p_write(bits, ref_rgba, w, h, il, cm_w, lut_w)
        for x, y in zip(a, b):
            assert np.array_equal(x, y), (bits, w, il)
