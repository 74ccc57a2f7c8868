"""The reference's OWN OpenCL kernels (src/process/*.ts of Streampunk/phaneron, extracted verbatim into the
git-ignored oracle/_ref/kernels/) run on the NVIDIA OpenCL driver of the GPU box, against (a) the CPU oracle and
(b) the fused CUDA path.  This is what pins the oracle: everything on the path except the image sampler of
`transform` is BIT-EXACT with the reference on the same B200; the sampler is hardware (9-bit filter weights,
tex.2d) on NVIDIA's OpenCL and follows the OpenCL 1.2 spec formula in the oracle / CUDA path.
Skipped when the OpenCL driver or the extracted kernels are absent."""
import numpy as np
import pytest

import oracle
from oracle import ref_ocl
from phaneron_b200.harness import ChannelHarness
from phaneron_b200.scenes import IDENTITY_XF, make_frame, pip, single_layer_scene

from gpu_util import Env, run
from scene_oracle import xf_matrix

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_ocl.available(), reason="reference OpenCL kernels not runnable here")]

W, H = 1920, 1080


def _consts(read="709", work="2020"):
    return (oracle.ycbcr2rgb_matrix(read), oracle.gamma2linear_lut(read), oracle.rgb2rgb_matrix(read, work),
            oracle.rgb2ycbcr_matrix(work), oracle.linear2gamma_lut(work))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("inputs", ["ramp", "noise"])
def test_oracle_v210_read_write_bit_exact_vs_reference_kernels(inputs):
    cm_r, lut_r, gamut, cm_w, lut_w = _consts()
    src = make_frame(inputs, W, H, 0)
    ref_rgba = ref_ocl.v210_read(src, W, H, cm_r, lut_r, gamut)
    assert np.array_equal(_bits(ref_rgba), _bits(oracle.v210_read(src, W, H, cm_r, lut_r, gamut)))
    ref_out = ref_ocl.v210_write(ref_rgba, W, H, 0, cm_w, lut_w)
    assert np.array_equal(ref_out, oracle.v210_write(ref_rgba, W, H, 0, cm_w, lut_w))


def test_reference_round_trip_invariant_on_its_own_fixture():
    """the reference's pass criterion `src.compare(dst) === 0` (src/process/test/*.ts), on v210.fillBuf, progressive and as two fields"""
    cm_r, lut_r, gamut, cm_w, lut_w = _consts("709", "709")
    src = make_frame("ramp", W, H, 0)
    rgba = ref_ocl.v210_read(src, W, H, cm_r, lut_r, gamut)
    assert np.array_equal(ref_ocl.v210_write(rgba, W, H, 0, cm_w, lut_w), src)
    dst = np.zeros_like(src)
    ref_ocl.v210_write(rgba, W, H, 1, cm_w, lut_w, out=dst)
    ref_ocl.v210_write(rgba, W, H, 3, cm_w, lut_w, out=dst)
    assert np.array_equal(dst, src)


def test_oracle_image_ops_bit_exact_vs_reference_kernels():
    rng = np.random.default_rng(11)
    ims = [rng.random((270, 480, 4), dtype=np.float32) for _ in range(5)]
    for n in (2, 3, 5):
        assert np.array_equal(_bits(ref_ocl.combine(ims[:n])), _bits(oracle.combine(ims[:n])))
    assert np.array_equal(_bits(ref_ocl.dissolve(ims[0], ims[1], 0.37)), _bits(oracle.dissolve(ims[0], ims[1], 0.37)))
    assert np.array_equal(_bits(ref_ocl.wipe_mask(ims[0], ims[1], ims[2])), _bits(oracle.wipe_mask(ims[0], ims[1], ims[2])))


def test_transform_differs_from_reference_only_by_the_hardware_sampler():
    """NVIDIA's OpenCL samples with tex.2d: filter weights quantised to 1/256.  The spec-formula result must sit within
    one weight step (times the local contrast <= 1) of it, and identity / half-pixel cases within float noise."""
    rng = np.random.default_rng(12)
    im = rng.random((270, 480, 4), dtype=np.float32)
    for xf, tol in ((dict(IDENTITY_XF), 5e-5), (pip(0.5, 0.25, 0.45), 5e-5), (dict(IDENTITY_XF, scaleX=0.731, scaleY=0.577, offsetX=0.21), 1.0 / 128)):
        m = xf_matrix(480, 270, xf)
        assert float(np.abs(ref_ocl.transform(im, m, 480, 270) - oracle.transform(im, m, 480, 270)).max()) < tol


@pytest.mark.parametrize("inputs", ["ramp", "noise"])
def test_fused_cuda_chain_bit_exact_vs_reference_kernel_chain(inputs):
    """ToRGBA -> FromRGBA at 1080p through our C ABI (one fused launch) == reference read kernel -> reference write kernel"""
    cm_r, lut_r, gamut, cm_w, lut_w = _consts("709", "2020")
    scene = single_layer_scene(W, H, inputs, False, "709", "2020")

    async def go():
        async with Env() as env:
            h = ChannelHarness(env.ctx, scene, env.pj)
            await h.init()
            out = await h.run_frame()
            assert env.ctx.stats()["march_launches"] == 1
            return out
    ours = run(go())
    src = scene["layers"][0]["src"]
    ref = ref_ocl.v210_write(ref_ocl.v210_read(src, W, H, cm_r, lut_r, gamut), W, H, 0, cm_w, lut_w)
    assert np.array_equal(ours, ref)


def test_fused_cuda_dissolve_and_combine_bit_exact_vs_reference_kernel_chain():
    """two direct sources dissolved, over a third: reference = 3x read, transition_dissolve, combine_2, write"""
    cm_r, lut_r, gamut, cm_w, lut_w = _consts("709", "2020")
    w, h = 960, 540
    a, b, c = (make_frame("noise", w, h, i) for i in range(3))
    scene = dict(width=w, height=h, colRead="709", colWork="2020", interlaced=False,
                 layers=[dict(src=c, sw=w, sh=h, xf=None, transition=None),
                         dict(src=a, sw=w, sh=h, xf=None, transition=dict(type="dissolve", mix=0.3, src=b, sw=w, sh=h, xf=None))])

    async def go():
        async with Env() as env:
            hh = ChannelHarness(env.ctx, scene, env.pj)
            await hh.init()
            return await hh.run_frame()
    ours = run(go())
    rd = lambda s: ref_ocl.v210_read(s, w, h, cm_r, lut_r, gamut)
    comp = ref_ocl.combine([rd(c), ref_ocl.dissolve(rd(a), rd(b), 0.3)])
    assert np.array_equal(ours, ref_ocl.v210_write(comp, w, h, 0, cm_w, lut_w))


@pytest.mark.parametrize("parity,tff,skip", [(0, True, False), (1, True, False), (1, False, True), (0, False, False)])
def test_oracle_yadif_bit_exact_vs_reference_kernel(parity, tff, skip):
    """yadifCl.ts:105-167 on the driver vs the restatement: compares, selects and adds only, so every bit must agree"""
    rng = np.random.default_rng(13)
    prev, cur, nxt = (rng.random((72, 128, 4), dtype=np.float32) for _ in range(3))
    ref = ref_ocl.yadif(prev, cur, nxt, parity, tff, skip)
    assert np.array_equal(_bits(ref), _bits(oracle.yadif(prev, cur, nxt, parity, tff, skip)))


def test_oracle_rgba8_bit_exact_vs_reference_kernels():
    """rgba8.ts:25-103 (ScreenConsumer / FFmpegProducer path): sRGB tables, alpha through the LUT on read, forced 255 on write"""
    w, h = 640, 90
    rng = np.random.default_rng(14)
    src = rng.integers(0, 256, w * h * 4, dtype=np.uint8)
    lut_r, gamut, lut_w = oracle.gamma2linear_lut("sRGB"), oracle.rgb2rgb_matrix("sRGB", "709"), oracle.linear2gamma_lut("sRGB")
    ref = ref_ocl.rgba8_read(src, w, h, lut_r, gamut)
    assert np.array_equal(_bits(ref), _bits(oracle.rgba8_read(src, w, h, lut_r, gamut)))
    assert np.array_equal(ref_ocl.rgba8_write(ref, w, h, 0, lut_w), oracle.rgba8_write(ref, w, h, 0, lut_w))


def test_oracle_mix_and_wipe_bit_exact_vs_reference_kernels():
    """mix.ts:24-47 ('mixer': fma(in0, mix, in1 * (1 - mix))) and wipe.ts:24-48 (x > w * wipe): point-sampled, so every bit must agree"""
    rng = np.random.default_rng(15)
    a, b = (rng.random((135, 240, 4), dtype=np.float32) for _ in range(2))
    for m in (0.0, 0.25, 0.37, 1.0):
        assert np.array_equal(_bits(ref_ocl.mix(a, b, m)), _bits(oracle.mix(a, b, m)))
    for wp in (0.0, 0.3, 0.5, 0.999, 1.0):
        assert np.array_equal(_bits(ref_ocl.wipe(a, b, wp)), _bits(oracle.wipe(a, b, wp)))


def test_resize_differs_from_reference_only_by_the_hardware_sampler():
    """resize.ts:24-58 samples with normalised coordinates + CLK_FILTER_LINEAR: as for transform, the driver's tex.2d quantises the
    filter weights to 1/256; the restatement follows the OpenCL 1.2 formula and must sit within one weight step of it"""
    rng = np.random.default_rng(16)
    im = rng.random((135, 240, 4), dtype=np.float32)
    for scale, ox, oy, flip in ((1.0, 0.0, 0.0, (0.0, 1.0, 0.0, 1.0)), (0.5, 0.1, -0.2, (0.0, 1.0, 0.0, 1.0)), (1.7, 0.0, 0.25, (1.0, -1.0, 0.0, 1.0)),
                                (0.8, -0.3, 0.0, (1.0, -1.0, 1.0, -1.0))):
        ref = ref_ocl.resize(im, scale, ox, oy, flip, 240, 135)
        got = oracle.resize(im, scale, ox, oy, flip, 240, 135)
        assert float(np.abs(ref - got).max()) < 1.0 / 128
