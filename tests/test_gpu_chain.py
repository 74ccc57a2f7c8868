"""The fused chain against the oracle's unfused stage sequence, driven through the
harness that calls the operators like phaneron's mixer/transitioner/combiner/consumer do.
Packed output: bit-exact bytes.  RGBA composite (materialised): bit-exact floats (0 ulp)."""
import numpy as np
import pytest

import oracle

from phaneron_b200.harness import ChannelHarness
from phaneron_b200.process.packer import Interlace
from phaneron_b200.scenes import IDENTITY_XF, layered_scene, make_frame, pip, single_layer_scene

from gpu_util import Env, assert_bits_equal, run
from scene_oracle import SceneOracle

pytestmark = pytest.mark.gpu


async def _run_scene(scene, deferred=True):
    async with Env(deferred) as env:
        h = ChannelHarness(env.ctx, scene, env.pj)
        await h.init()
        before = env.ctx.stats()
        out = await h.run_frame()
        after = env.ctx.stats()
        return out, {k: after[k] - before[k] for k in after}


@pytest.mark.parametrize("inputs", ["ramp", "noise"])
@pytest.mark.parametrize("variant", ["plain", "mix", "wipe"])
def test_four_layer_scene_fused_vs_oracle(inputs, variant):
    """BASELINE config 3 shape (4-layer Mix/Wipe composite, 709 -> 2020) at a size the oracle finishes in seconds"""
    scene = layered_scene(480, 270, 4, inputs, variant, "709", "2020")
    out, st = run(_run_scene(scene))
    assert np.array_equal(out, SceneOracle(scene).packed())
    # the whole layer graph collapsed into ONE launch and no RGBA frame touched HBM
    assert st["kernel_launches"] == 1 and st["fused_launches"] == 1 and st["materialised"] == 0


@pytest.mark.parametrize("n_layers", [1, 2, 3, 8])
def test_layer_counts(n_layers):
    scene = layered_scene(384, 216, n_layers, "noise", "plain", "709", "709")
    out, st = run(_run_scene(scene))
    assert np.array_equal(out, SceneOracle(scene).packed())
    assert st["kernel_launches"] == 1


def test_eager_mode_matches_deferred_and_oracle():
    scene = layered_scene(384, 216, 3, "noise", "mix", "709", "2020")
    eager, st_e = run(_run_scene(scene, deferred=False))
    fused, st_f = run(_run_scene(scene, deferred=True))
    ref = SceneOracle(scene).packed()
    assert np.array_equal(eager, ref) and np.array_equal(fused, ref)
    # eager = the reference's launch structure: 4 reads + 4 transforms + 1 dissolve + 1 combine + 1 write
    assert st_e["kernel_launches"] == 11 and st_f["kernel_launches"] == 1


def test_single_layer_config2_passthrough_combine():
    """BASELINE config 2: ToRGBA -> Combine (0/1 layers: passthrough, combiner.ts:219-228) -> FromRGBA"""
    for with_mixer in (False, True):
        scene = single_layer_scene(1920, 1080, "ramp", with_mixer)
        out, st = run(_run_scene(scene))
        assert np.array_equal(out, SceneOracle(scene).packed())
        assert st["kernel_launches"] == 1
        if not with_mixer:
            assert np.array_equal(out, scene["layers"][0]["src"])   # round trip of the fixture


def test_rotated_and_scaled_layers():
    w, h = 480, 270
    scene = layered_scene(w, h, 3, "noise", "plain", "709", "709")
    scene["layers"][1]["xf"] = dict(pip(0.6, 0.2, 0.1), rotate=0.04)
    scene["layers"][2]["xf"] = dict(IDENTITY_XF, scaleX=1.3, scaleY=0.7, offsetX=0.11, flipH=True)
    out, st = run(_run_scene(scene))
    assert np.array_equal(out, SceneOracle(scene).packed())
    # the rotated layer's source is made real once (RGBA-f32, the direct kernel) and sampled as a frame by the one composite launch
    assert st["kernel_launches"] == 2 and st["materialised"] == 1 and st["march_launches"] == 2, st


def test_sources_of_different_sizes():
    """a 720p and a ragged 1280-wide source (Q1 tail) transformed into a 480x270 channel"""
    w, h = 480, 270
    scene = layered_scene(w, h, 2, "ramp", "plain", "709", "709")
    scene["layers"][0].update(src=make_frame("ramp", 1280, 72, 0), sw=1280, sh=72)
    scene["layers"][1].update(src=make_frame("noise", 960, 540, 3), sw=960, sh=540)
    out, _ = run(_run_scene(scene))
    assert np.array_equal(out, SceneOracle(scene).packed())


def test_ragged_output_width_1280():
    scene = layered_scene(1280, 72, 2, "ramp", "plain", "709", "709")
    out, _ = run(_run_scene(scene))
    assert np.array_equal(out, SceneOracle(scene).packed())


def test_interlaced_consumer_two_fields_one_buffer():
    """macadamConsumer.ts:224-244: two consecutive channel frames fill top then bottom lines"""
    async def go():
        scene = layered_scene(480, 270, 2, "noise", "plain", "709", "709")
        scene["interlaced"] = True
        async with Env() as env:
            h = ChannelHarness(env.ctx, scene, env.pj)
            await h.init()
            dests = await h.fromRGBA.createDests("il")
            dests[0].fill(0)
            await dests[0].hostAccess("writeonly")
            for il in (Interlace.TopField, Interlace.BottomField):
                ups = await h.upload_all(int(il))
                frame = await h.compose(ups, int(il))
                await h.consume(frame, dests, il, download=(il == Interlace.BottomField))
            so = SceneOracle(scene)
            ref = np.zeros_like(dests[0].host)
            so.packed(1, ref)
            so.packed(3, ref)
            assert np.array_equal(dests[0].host, ref)
    run(go())


def test_deferred_frame_materialises_on_host_read():
    """ScreenConsumer / ROUTE style access: hostAccess('readonly') on a frame that only
    exists as an expression must produce the RGBA floats the unfused path would"""
    async def go():
        scene = layered_scene(384, 216, 3, "noise", "wipe", "709", "2020")
        async with Env() as env:
            h = ChannelHarness(env.ctx, scene, env.pj)
            await h.init()
            ups = await h.upload_all(0)
            frame = await h.compose(ups, 0)
            assert frame.deferred
            got = await env.fetch(frame, 384, 216)
            assert not frame.deferred
            assert_bits_equal(got, SceneOracle(scene).composite(), "materialised composite")
            # and it can still feed the writer afterwards
            dests = await h.consume(frame)
            assert np.array_equal(dests[0].host, SceneOracle(scene).packed())
    run(go())


def test_chain_replay_reproduces_the_frame():
    async def go():
        scene = layered_scene(480, 270, 4, "noise", "mix", "709", "2020")
        async with Env() as env:
            h = ChannelHarness(env.ctx, scene, env.pj)
            await h.init()
            chain, dests = await h.record_chain()
            assert chain.complete and chain.launches == 1
            dests[0].fill(0)
            await dests[0].hostAccess("writeonly")
            await dests[0].hostAccess("none")
            chain.replay()
            await env.ctx.waitFinish(env.ctx.queue.process)
            await dests[0].hostAccess("readonly")
            assert np.array_equal(dests[0].host, SceneOracle(scene).packed())
            chain.destroy()
    run(go())


def test_buffers_are_recycled_not_leaked():
    async def go():
        scene = layered_scene(384, 216, 3, "ramp", "mix", "709", "709")
        async with Env() as env:
            h = ChannelHarness(env.ctx, scene, env.pj)
            await h.init()
            await h.run_frame()
            live0 = env.ctx.stats()["dev_bytes_live"]
            for _ in range(5):
                await h.run_frame()
            st = env.ctx.stats()
            assert st["dev_bytes_live"] == live0
    run(go())


def test_full_size_2160p_properties():
    """BASELINE config 3 at full size.  The oracle would take minutes here, so use
    size-independent properties: (1) an opaque full-frame identity top layer over anything
    equals that layer alone; (2) a dissolve at mix=1 equals input0, at mix=0 input1;
    (3) fused == eager byte for byte."""
    async def go():
        w, h = 3840, 2160
        async with Env() as env:
            base = layered_scene(w, h, 4, "ramp", "plain", "709", "2020")
            top_only = dict(base, layers=[base["layers"][0]])
            covered = dict(base, layers=[base["layers"][1], base["layers"][2], base["layers"][0]])
            async def frame_of(scene):
                hs = ChannelHarness(env.ctx, scene, env.pj)
                await hs.init()
                return await hs.run_frame()
            a = await frame_of(top_only)
            b = await frame_of(covered)
            # identity Transform blends row 0 / column 0 with the transparent border (Q6), so
            # the top layer is not opaque there; compare everything else
            words = lambda x: x.view(np.uint32).reshape(h, -1)
            assert np.array_equal(words(a)[1:, 4:], words(b)[1:, 4:])
            mix1 = dict(base, layers=[dict(base["layers"][0], transition=dict(type="dissolve", mix=1.0, src=base["layers"][1]["src"], sw=w, sh=h, xf=dict(IDENTITY_XF)))])
            assert np.array_equal(await frame_of(mix1), a)
            fused = await frame_of(base)
            env.ctx.setDeferred(False)
            eager = await frame_of(base)
            assert np.array_equal(fused, eager)
    run(go())


# ---- the march kernel (pb_march.cu) vs the generic fused kernel vs the oracle --------------------------
async def _run_scene_variant(scene, mode):
    """mode: 'march' (gamma tables in the one-byte shared-memory form), 'march_raw' (raw tables from
    global memory), 'generic' (pb_fused.cu)"""
    async with Env(True) as env:
        env.ctx.setMarchKernel(mode != "generic", rawLut=(mode == "march_raw"))
        env.ctx.footprint = True
        env.ctx.setOcclusionCulling(mode != "march_nocull")
        h = ChannelHarness(env.ctx, scene, env.pj)
        await h.init()
        before = env.ctx.stats()
        out = await h.run_frame()
        after = env.ctx.stats()
        st = {k: after[k] - before[k] for k in after}
        st["lut_tables"], st["lut_tables_d8"] = after["lut_tables"], after["lut_tables_d8"]
        st["march_src_bytes"] = after["march_src_bytes"]
        return out, st


def _xf(**kw):
    return dict(IDENTITY_XF, **kw)


def _with_xf(scene, xfs):
    for L, xf in zip(scene["layers"], xfs):
        L["xf"] = xf
    return scene


MARCH_SCENES = {
    "north_star_mix": lambda: layered_scene(480, 270, 4, "noise", "mix", "709", "2020"),
    "north_star_wipe": lambda: layered_scene(480, 270, 4, "noise", "wipe", "709", "2020"),
    "north_star_ramp": lambda: layered_scene(480, 270, 4, "ramp", "mix", "709", "2020"),
    "full_strip_width": lambda: layered_scene(768, 54, 3, "noise", "plain", "709", "709"),
    "two_strips_direct": lambda: single_layer_scene(384, 100, "noise", False),
    "odd_height": lambda: layered_scene(480, 135, 3, "noise", "mix", "709", "2020"),   # W + H odd: the 16-byte strip table must stay aligned
    "srgb_working_space": lambda: layered_scene(480, 64, 2, "noise", "plain", "709", "sRGB"),
    "flips_and_upscale": lambda: _with_xf(layered_scene(480, 270, 3, "noise", "plain", "709", "709"),
                                          [_xf(flipH=True), _xf(flipV=True, scaleX=1.5, scaleY=2.25, offsetX=0.1),
                                           _xf(flipH=True, flipV=True, scaleX=0.75, scaleY=0.6, offsetY=-0.2)]),
    "odd_fractions": lambda: _with_xf(layered_scene(480, 270, 2, "noise", "plain", "709", "709"),
                                      [_xf(scaleX=1.0001, scaleY=0.9999, offsetX=0.00013), _xf(scaleX=0.731, scaleY=0.577, offsetX=0.21, offsetY=0.13)]),
    # deep down-scales (multiviewer tiles): up to 64 source groups per 90-px strip, the big row buffers
    "multiview_thirds": lambda: _with_xf(layered_scene(960, 270, 4, "noise", "plain", "709", "2020"),
                                         [_xf(), pip(1 / 3, 0.0, 0.0), pip(1 / 3, 1 / 3, 0.2), pip(0.3, 0.6, 0.5)]),
    "quarter_tiles_mix": lambda: _with_xf(layered_scene(960, 270, 3, "noise", "mix", "709", "2020"), [_xf(), pip(0.25, 0.1, 0.1), pip(0.26, 0.5, 0.4)]),
    "stray_top_bits": lambda: _ragged_scene(480, 48, [(480, 48, _xf()), (480, 48, pip(0.5, 0.3, 0.3))]),
    "mostly_outside": lambda: _with_xf(layered_scene(480, 270, 2, "noise", "plain", "709", "709"),
                                       [_xf(offsetX=0.97, offsetY=-0.96), _xf(scaleX=0.5, scaleY=0.5, offsetX=-0.7)]),
}


@pytest.mark.parametrize("name", sorted(MARCH_SCENES))
def test_march_kernel_matches_generic_and_oracle(name):
    scene = MARCH_SCENES[name]()
    ref = SceneOracle(scene).packed()
    slow, st_slow = run(_run_scene_variant(scene, "generic"))
    assert st_slow["march_launches"] == 0 and st_slow["kernel_launches"] == 1
    assert np.array_equal(slow, ref)
    for mode in ("march", "march_raw"):
        fast, st = run(_run_scene_variant(scene, mode))
        # (the big-row and general-load variants exist for the shared-memory tables only: raw-table mode falls back there)
        want = 0 if (mode == "march_raw" and name in ("multiview_thirds", "quarter_tiles_mix")) else 1
        assert st["march_launches"] == want and st["kernel_launches"] == 1, (mode, st)
        assert np.array_equal(fast, ref), f"{mode}: {int((fast != ref).sum())} bytes differ"


# ---- exact occlusion culling (pb_march_prep.cu leaf_opacity) -------------------------------
def _stack(w, h, xfs, variant="plain", inputs="noise"):
    return _with_xf(layered_scene(w, h, len(xfs), inputs, variant, "709", "2020"), xfs)


CULL_SCENES = {
    # name: (scene factory, culling must reduce the bytes read)
    "north_star_mix": (lambda: layered_scene(960, 540, 4, "noise", "mix", "709", "2020"), True),
    "north_star_wipe": (lambda: layered_scene(960, 540, 4, "noise", "wipe", "709", "2020"), True),   # L2, L3 still hide L1
    "full_frame_on_top": (lambda: _stack(960, 270, [_xf(), _xf(scaleX=0.5, scaleY=0.5, offsetX=-0.2), _xf()]), True),
    "direct_on_top": (lambda: _with_xf(layered_scene(960, 270, 3, "noise", "plain", "709", "709"), [_xf(), _xf(scaleX=0.5, scaleY=0.5), None]), True),
    "pip_over_pip_over_pip": (lambda: _stack(960, 540, [_xf(), _xf(scaleX=0.75, scaleY=0.75, offsetX=-0.1, offsetY=-0.1),
                                                        _xf(scaleX=0.625, scaleY=0.625, offsetX=-0.2, offsetY=-0.15),
                                                        _xf(scaleX=0.5, scaleY=0.5, offsetX=-0.3, offsetY=-0.3)]), True),
    "flipped_pip": (lambda: _stack(960, 540, [_xf(), _xf(flipH=True, flipV=True, scaleX=0.5, scaleY=0.5, offsetX=-0.3, offsetY=-0.2)]), True),
    "upscaled_top": (lambda: _stack(960, 270, [_xf(), _xf(scaleX=2.0, scaleY=2.0)]), True),
    "odd_scale_pip": (lambda: _stack(960, 540, [_xf(), _xf(scaleX=0.731, scaleY=0.577, offsetX=-0.21, offsetY=-0.13)]), False),
    "ramp_inputs": (lambda: layered_scene(960, 540, 4, "ramp", "mix", "709", "2020"), True),
}


@pytest.mark.parametrize("name", sorted(CULL_SCENES))
def test_occlusion_culling_is_exact(name):
    """hidden layers are skipped where an upper layer's alpha is exactly 1.0f: same bytes out, fewer bytes in"""
    make, must_cull = CULL_SCENES[name]
    scene = make()
    ref = SceneOracle(scene).packed()
    plain, st0 = run(_run_scene_variant(scene, "march_nocull"))
    culled, st1 = run(_run_scene_variant(scene, "march"))
    assert st0["march_launches"] == 1 and st1["march_launches"] == 1
    assert np.array_equal(plain, ref)
    assert np.array_equal(culled, ref), f"{int((culled != ref).sum())} bytes differ with culling on"
    assert st1["march_src_bytes"] <= st0["march_src_bytes"]
    if must_cull:
        assert st1["march_src_bytes"] < st0["march_src_bytes"], (st0["march_src_bytes"], st1["march_src_bytes"])


def test_dissolve_layer_is_opaque_only_when_its_alpha_rounds_to_one():
    """top layer in a dissolve: alpha = RN(mix + RN(1 - mix)); culled under it only when that is exactly 1"""
    for mix in (0.5, 0.3, 1.0 / 3.0, 0.9999999):
        scene = layered_scene(960, 270, 3, "noise", "mix", "709", "2020")
        scene["layers"][-1]["transition"]["mix"] = mix
        ref = SceneOracle(scene).packed()
        culled, st = run(_run_scene_variant(scene, "march"))
        assert st["march_launches"] == 1
        assert np.array_equal(culled, ref), mix


def test_gamma_tables_are_deduplicated_and_compressed():
    """five Loaders + one Saver upload six tables; the context keeps two (709 gamma->linear, 2020
    linear->gamma), both in the lossless one-byte form (pb_lut.cuh)"""
    scene = layered_scene(480, 270, 4, "noise", "mix", "709", "2020")
    _, st = run(_run_scene_variant(scene, "march"))
    assert st["lut_tables"] == 2 and st["lut_tables_d8"] == 2, st


def test_march_kernel_declines_what_it_cannot_do():
    """a deep downscale (footprint wider than a row buffer) falls back to the generic kernel; a rotated packed source is made real
    as RGBA-f32 first (one launch of the direct kernel) and then sampled by the march kernel at its affine positions"""
    rot = _with_xf(layered_scene(480, 270, 2, "noise", "plain"), [_xf(), _xf(rotate=0.01)])
    deep = _with_xf(layered_scene(480, 270, 2, "noise", "plain"), [_xf(), _xf(scaleX=0.2, scaleY=0.2)])
    out, st = run(_run_scene_variant(deep, "march"))
    assert st["march_launches"] == 0 and st["fused_launches"] == 1, st
    assert np.array_equal(out, SceneOracle(deep).packed())
    out, st = run(_run_scene_variant(rot, "march"))
    assert st["march_launches"] == 2 and st["fused_launches"] == 2 and st["materialised"] == 1 and st["kernel_launches"] == 2, st
    assert np.array_equal(out, SceneOracle(rot).packed())
    out, st = run(_run_scene_variant(rot, "generic"))
    assert st["march_launches"] == 0 and np.array_equal(out, SceneOracle(rot).packed())


@pytest.mark.parametrize("angle,scale,ox,oy", [(0.04, 0.6, 0.2, 0.1), (-0.31, 1.0, 0.0, 0.0), (0.25, 0.35, -0.3, 0.3), (0.5, 1.4, 0.4, -0.45), (0.013, 0.5, 0.9, 0.9)])
def test_rotated_layers_on_the_march_kernel(angle, scale, ox, oy):
    """DVE rotation (transform.ts:132-171 builds rotate into the matrix): layers 2 and 3 rotated -- one of them in a dissolve --
    over a full-frame layer: bounding boxes of the rotated quads (incl. mostly / entirely outside the frame), bit-exact against
    the oracle, and march == generic"""
    w, h = 480, 270
    scene = layered_scene(w, h, 3, "noise", "mix", "709", "2020")
    scene["layers"][1]["xf"] = dict(pip(scale, ox, oy), rotate=angle)
    xf2 = dict(pip(0.5, 0.3, 0.25), rotate=-angle * 0.5)
    scene["layers"][2]["xf"] = xf2
    scene["layers"][2]["transition"]["xf"] = xf2
    ref = SceneOracle(scene).packed()
    out, st = run(_run_scene_variant(scene, "march"))
    assert np.array_equal(out, ref), f"{int((out != ref).sum())} bytes differ"
    assert st["materialised"] == 3 and st["march_launches"] == 4, st   # three rotated sources made real, one composite
    slow, st2 = run(_run_scene_variant(scene, "generic"))
    assert st2["march_launches"] == 0 and np.array_equal(slow, ref)


# ---- widths that are not whole v210 groups / 48-pixel blocks: 1280-wide 720p (213 groups + 2 pixels per line) ----
def _random_v210(w, h, seed):
    """any 32-bit words: 10-bit fields of all values, and stray bits 30-31 (the reference masks the fields: legal input)"""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 2 ** 32, v210_pitch(w) * h // 4, dtype=np.uint64).astype(np.uint32).view(np.uint8)


def v210_pitch(w):
    return (w + 47) // 48 * 128


def _ragged_scene(w, h, specs, variant="plain"):
    """specs: [(source width, source height, xf)]"""
    layers = [dict(src=_random_v210(sw, sh, 70 + i), sw=sw, sh=sh, xf=xf, transition=None) for i, (sw, sh, xf) in enumerate(specs)]
    if variant == "mix":
        sw, sh, xf = specs[-1]
        layers[-1]["transition"] = dict(type="dissolve", mix=0.5, src=_random_v210(sw, sh, 99), sw=sw, sh=sh, xf=xf)
    return dict(width=w, height=h, colRead="709", colWork="2020", interlaced=False, layers=layers)


RAGGED_SCENES = {
    "720p_direct": lambda: _ragged_scene(1280, 40, [(1280, 40, None)]),
    "720p_identity_transform": lambda: _ragged_scene(1280, 40, [(1280, 40, _xf())]),
    "720p_four_layers_mix": lambda: _ragged_scene(1280, 72, [(1280, 72, _xf()), (1280, 72, pip(0.5, 0.05, 0.05)), (1280, 72, pip(0.5, 0.45, 0.1)),
                                                             (1280, 72, pip(0.75, 0.3, 0.3))], "mix"),
    "720p_source_into_1080p": lambda: _ragged_scene(1920, 54, [(1920, 54, _xf()), (1280, 36, _xf(scaleX=0.6, scaleY=0.6, offsetX=-0.3))]),
    "1080p_source_into_720p": lambda: _ragged_scene(1280, 36, [(1280, 36, _xf()), (1920, 54, pip(0.8, 0.1, 0.1))]),
    "2k_dci": lambda: _ragged_scene(2048, 24, [(2048, 24, _xf()), (2048, 24, pip(0.5, 0.5, 0.2))]),
    "narrow_50": lambda: _ragged_scene(50, 20, [(50, 20, _xf())]),
}


@pytest.mark.parametrize("name", sorted(RAGGED_SCENES))
def test_ragged_widths_take_the_march_kernel_plus_a_tail_launch(name):
    """the march kernel writes the whole v210 groups, the generic kernel the tail columns of every line (partial group with the
    reference's Q2 rounding, padding groups); sources read their partial last group with Q1 (v210.ts:90-110)"""
    scene = RAGGED_SCENES[name]()
    ref = SceneOracle(scene).packed()
    slow, st0 = run(_run_scene_variant(scene, "generic"))
    assert st0["march_launches"] == 0 and np.array_equal(slow, ref)
    out, st = run(_run_scene_variant(scene, "march"))
    ragged_out = scene["width"] % 48 != 0
    assert st["march_launches"] == 1 and st["kernel_launches"] == (2 if ragged_out else 1) and st["materialised"] == 0, st
    assert np.array_equal(out, ref), f"{int((out != ref).sum())} bytes differ"


def test_ragged_width_interlaced_and_replay():
    async def go():
        scene = _ragged_scene(1280, 72, [(1280, 72, _xf()), (1280, 72, pip(0.5, 0.2, 0.2))])
        scene["interlaced"] = True
        async with Env() as env:
            h = ChannelHarness(env.ctx, scene, env.pj)
            await h.init()
            dests = await h.fromRGBA.createDests("il")
            dests[0].fill(0)
            await dests[0].hostAccess("writeonly")
            for il in (Interlace.TopField, Interlace.BottomField):
                ups = await h.upload_all(int(il))
                frame = await h.compose(ups, int(il))
                await h.consume(frame, dests, il, download=(il == Interlace.BottomField))
            so = SceneOracle(scene)
            ref = np.zeros_like(dests[0].host)
            so.packed(1, ref)
            so.packed(3, ref)
            assert np.array_equal(dests[0].host, ref)
            # a recorded chain re-issues both launches
            scene["interlaced"] = False
            h2 = ChannelHarness(env.ctx, scene, env.pj)
            await h2.init()
            chain, d2 = await h2.record_chain()
            assert chain.complete and chain.launches == 1
            d2[0].fill(0)
            await d2[0].hostAccess("writeonly")
            await d2[0].hostAccess("none")
            chain.replay()
            await env.ctx.waitFinish(env.ctx.queue.process)
            await d2[0].hostAccess("readonly")
            assert np.array_equal(d2[0].host, SceneOracle(scene).packed())
    run(go())


def test_march_kernel_interlaced_fields():
    async def go():
        scene = layered_scene(480, 270, 3, "noise", "mix", "709", "2020")
        scene["interlaced"] = True
        async with Env() as env:
            h = ChannelHarness(env.ctx, scene, env.pj)
            await h.init()
            dests = await h.fromRGBA.createDests("il")
            dests[0].fill(0)
            await dests[0].hostAccess("writeonly")
            for il in (Interlace.TopField, Interlace.BottomField):
                ups = await h.upload_all(int(il))
                frame = await h.compose(ups, int(il))
                await h.consume(frame, dests, il, download=(il == Interlace.BottomField))
            assert env.ctx.stats()["march_launches"] == 2
            so = SceneOracle(scene)
            ref = np.zeros_like(dests[0].host)
            so.packed(1, ref)
            so.packed(3, ref)
            assert np.array_equal(dests[0].host, ref)
    run(go())


def test_march_kernel_full_size_equals_generic_2160p():
    """BASELINE config 3 at full size: the march kernel must reproduce the generic kernel byte for byte"""
    scene = layered_scene(3840, 2160, 4, "noise", "mix", "709", "2020")
    slow, _ = run(_run_scene_variant(scene, "generic"))
    for mode in ("march", "march_raw"):
        fast, st = run(_run_scene_variant(scene, mode))
        assert st["march_launches"] == 1
        assert np.array_equal(fast, slow), mode


# ---- other packed source formats read in place by the fused kernel ("behind the same fused front end") ----
def _rand_source(fmt, w, h, seed):
    rng = np.random.default_rng(seed)
    if fmt in ("rgba8", "bgra8"):
        return rng.integers(0, 256, w * h * 4, dtype=np.uint8)           # alpha included: Combine becomes non-trivial
    if fmt in ("yuv422p10", "yuv422p8"):
        bits = 10 if fmt == "yuv422p10" else 8
        nb = oracle.yuv422p_plane_bytes(bits, w, h)
        if bits == 10:
            return [rng.integers(0, 1024, n // 2, dtype=np.uint16).astype("<u2").view(np.uint8) for n in nb]
        return [rng.integers(0, 256, n, dtype=np.uint8) for n in nb]
    return [rng.integers(0, 256, n, dtype=np.uint8) for n in oracle.yuv420_plane_bytes(fmt == "nv12", w, h)]


def _mixed_format_scene(w, h, specs):
    """specs: [(fmt, colRead, xf)]; sources have the channel's dimensions"""
    layers = []
    for i, (fmt, col, xf) in enumerate(specs):
        src = make_frame("noise", w, h, 40 + i) if fmt == "v210" else _rand_source(fmt, w, h, 60 + i)
        layers.append(dict(src=src, sw=w, sh=h, xf=xf, transition=None, fmt=fmt, colRead=col))
    return dict(width=w, height=h, colRead="709", colWork="2020", interlaced=False, layers=layers)


MIXED_SCENES = {
    "ffmpeg_formats_stack": lambda: _mixed_format_scene(960, 540, [
        ("yuv420p", "709", _xf()), ("rgba8", "sRGB", _xf(scaleX=0.5, scaleY=0.5, offsetX=-0.05, offsetY=-0.05)),
        ("yuv422p10", "709", _xf(scaleX=0.5, scaleY=0.5, offsetX=-0.45, offsetY=-0.1)),
        ("nv12", "601_525", _xf(scaleX=0.5, scaleY=0.5, offsetX=-0.25, offsetY=-0.45))]),
    "bgra8_over_v210": lambda: _mixed_format_scene(960, 270, [("v210", None, _xf()), ("bgra8", "sRGB", _xf(scaleX=0.75, scaleY=0.75, rotate=0.02))]),
    "yuv422p8_direct": lambda: _mixed_format_scene(768, 64, [("yuv422p8", "709", None)]),
    # graphics with alpha over video, the way a CG overlay reaches the combiner (rgba8.ts:61: alpha through the LUT)
    "rgba8_overlay_on_video": lambda: _mixed_format_scene(960, 540, [("v210", None, _xf()), ("v210", None, pip(0.5, 0.4, 0.1)),
                                                                     ("rgba8", "sRGB", _xf()), ("bgra8", "sRGB", pip(0.6, 0.1, 0.3))]),
    "rgba8_only_upscaled_flipped": lambda: _mixed_format_scene(960, 270, [("rgba8", "sRGB", _xf(scaleX=1.4, scaleY=1.7, flipH=True, offsetX=0.1))]),
    "rgba8_alpha_stack_direct": lambda: _mixed_format_scene(448, 36, [("yuv422p8", "709", None), ("rgba8", "sRGB", None), ("bgra8", "sRGB", None)]),
}


@pytest.mark.parametrize("name", sorted(MIXED_SCENES))
def test_packed_source_formats_fuse_into_one_launch(name):
    """rgba8 / bgra8 / yuv422p10 / yuv422p8 / yuv420p / nv12 sources are leaves of the fused kernel: the frame is one
    launch (no RGBA-f32 intermediate), bit-exact against the unfused oracle chain"""
    scene = MIXED_SCENES[name]()
    ref = SceneOracle(scene).packed()
    out, st = run(_run_scene_variant(scene, "march"))
    tail = 1 if (st["march_launches"] and scene["width"] % 48) else 0   # ragged v210 widths: march kernel + line-tail launch
    rot = 1 if name == "bgra8_over_v210" else 0   # a rotated packed source is made real (RGBA-f32) before it is sampled
    assert st["kernel_launches"] == 1 + tail + rot and st["fused_launches"] == 1 + rot and st["materialised"] == rot, st
    assert st["march_launches"] >= 1, st
    assert np.array_equal(out, ref), f"{int((out != ref).sum())} bytes differ"
    slow, st2 = run(_run_scene_variant(scene, "generic"))
    assert st2["march_launches"] == 0 and np.array_equal(slow, ref)


def test_packed_sources_eager_equals_deferred():
    scene = MIXED_SCENES["ffmpeg_formats_stack"]()
    async def go(deferred):
        async with Env(deferred) as env:
            h = ChannelHarness(env.ctx, scene, env.pj)
            await h.init()
            return await h.run_frame()
    assert np.array_equal(run(go(True)), run(go(False)))


# ---- the other consumer formats as fused sinks ("behind the same fused back end") ----
@pytest.mark.parametrize("out_fmt,w,h", [("yuv422p8", 960, 270), ("yuv422p8", 726, 64), ("yuv422p10", 960, 128), ("yuv422p10", 714, 32),
                                         ("rgba8", 960, 135), ("bgra8", 384, 36), ("yuv420p", 960, 270), ("yuv420p", 708, 20),
                                         ("nv12", 960, 128), ("nv12", 726, 32)])
def test_consumer_formats_are_fused_sinks(out_fmt, w, h):
    """FFmpegConsumer (yuv422p8) / ScreenConsumer (rgba8) / the other Writer PackImpls: the layer graph is evaluated inside
    the writer -- one launch, no RGBA-f32 composite in HBM -- bit-exact against the unfused oracle chain (tails included)"""
    scene = layered_scene(w, h, 3, "noise", "mix", "709", "2020")
    scene["outFmt"] = out_fmt
    if out_fmt in ("rgba8", "bgra8"):
        scene["colWrite"] = "sRGB"
    ref = SceneOracle(scene).packed()
    out, st = run(_run_scene_variant(scene, "march"))
    assert st["kernel_launches"] == 1 and st["fused_launches"] == 1 and st["materialised"] == 0, st
    # at march-kernel widths every consumer format is written by the march kernel itself
    assert st["march_launches"] == (1 if w % 48 == 0 else 0), st
    assert out.shape == ref.shape
    assert np.array_equal(out, ref), f"{int((out != ref).sum())} bytes differ"
    slow, st2 = run(_run_scene_variant(scene, "generic"))
    assert st2["march_launches"] == 0 and np.array_equal(slow, ref)


def test_mixed_sources_into_a_planar_sink_one_launch():
    scene = MIXED_SCENES["ffmpeg_formats_stack"]()
    scene["outFmt"] = "yuv422p8"
    ref = SceneOracle(scene).packed()
    out, st = run(_run_scene_variant(scene, "march"))
    assert st["kernel_launches"] == 1 and st["materialised"] == 0 and st["march_launches"] == 1, st
    assert np.array_equal(out, ref)


def test_lanczos_layers_fuse_into_one_launch():
    """BASELINE.json config 5's shape (2 layers, the upper one a resized PiP) with the Lanczos filter on both layers' Transforms,
    v210 and yuv420p sources: one launch, bit-exact against the unfused oracle chain (oracle.transform_lanczos)"""
    scene = _mixed_format_scene(960, 540, [("v210", None, _xf(filter="lanczos3")),
                                           ("yuv420p", "709", _xf(scaleX=0.5, scaleY=0.5, offsetX=-0.25, offsetY=-0.2, filter="lanczos3"))])
    ref = SceneOracle(scene).packed()
    out, st = run(_run_scene_variant(scene, "march"))
    # Lanczos leaves ride the march kernel: ONE fused launch, now the fast one, plus the horizontal first pass of each Lanczos
    # leaf (k_lanczos_hpass: every source row converted once); no RGBA-f32 frame of the reference's kind is written
    assert st["fused_launches"] == 1 and st["materialised"] == 0 and st["march_launches"] == 1 and 1 <= st["kernel_launches"] <= 3, st
    assert np.array_equal(out, ref), f"{int((out != ref).sum())} bytes differ"
    slow, st0 = run(_run_scene_variant(scene, "generic"))
    assert st0["march_launches"] == 0 and np.array_equal(slow, ref)


LANCZOS_MARCH_SCENES = {
    # BASELINE.json config 5's layer structure: full-frame bilinear background, Lanczos-3 half-size PiP
    "config5_shape": lambda: _lanczos_scene(960, 540, [("v210", _xf()), ("v210", dict(pip(0.5, 0.25, 0.25), filter="lanczos3"))]),
    "two_lobes_upscale_flip": lambda: _lanczos_scene(768, 432, [("v210", _xf(filter="lanczos2")),
                                                                ("v210", _xf(scaleX=1.4, scaleY=1.2, offsetX=0.1, flipH=True, filter="lanczos2"))]),
    "mostly_outside_and_dissolve": lambda: _lanczos_scene(768, 432, [("v210", _xf()), ("v210", dict(pip(0.6, 0.7, 0.65), filter="lanczos3"))], dissolve=True),
    "deep_downscale_big_rows": lambda: _lanczos_scene(960, 540, [("v210", _xf()), ("v210", dict(pip(0.4, 0.1, 0.3), filter="lanczos3"))]),
}


def _lanczos_scene(w, h, specs, dissolve=False):
    layers = [dict(src=make_frame("noise", w, h, i), sw=w, sh=h, xf=xf, transition=None) for i, (_, xf) in enumerate(specs)]
    if dissolve:
        layers[-1]["transition"] = dict(type="dissolve", mix=0.25, src=make_frame("noise", w, h, 9), sw=w, sh=h, xf=layers[-1]["xf"])
    return dict(width=w, height=h, colRead="709", colWork="2020", interlaced=False, layers=layers)


@pytest.mark.parametrize("name", sorted(LANCZOS_MARCH_SCENES))
def test_lanczos_leaves_in_the_march_kernel(name):
    scene = LANCZOS_MARCH_SCENES[name]()
    ref = SceneOracle(scene).packed()
    passes = {"config5_shape": 1, "two_lobes_upscale_flip": 2, "mostly_outside_and_dissolve": 2, "deep_downscale_big_rows": 1}[name]
    for mode in ("march", "march_nocull"):
        out, st = run(_run_scene_variant(scene, mode))
        # one fused launch + at most one horizontal first pass per Lanczos leaf (where its strip footprints fit a row buffer; a leaf
        # that is not separable this way is filtered inside the fused launch)
        assert st["march_launches"] == 1 and st["fused_launches"] == 1 and 1 <= st["kernel_launches"] <= 1 + passes and st["materialised"] == 0, (mode, st)
        if name == "config5_shape":
            assert st["kernel_launches"] == 2, st
        assert np.array_equal(out, ref), f"{mode}: {int((out != ref).sum())} bytes differ"
    import os
    os.environ["PB_LANCZOS_ONE_PASS"] = "1"   # the filter evaluated inside the one launch (pb_march.cu eval_leaf_lanczos): same bytes
    try:
        out, st = run(_run_scene_variant(scene, "march"))
    finally:
        del os.environ["PB_LANCZOS_ONE_PASS"]
    assert st["march_launches"] == 1 and st["kernel_launches"] == 1, st
    assert np.array_equal(out, ref)


# ---- planar YCbCr leaves in the march kernel (gathered into the v210 group layout, pb_march.cu load_group) ----
PLANAR_MARCH_SCENES = {
    "yuv422p10_stack": lambda: _mixed_format_scene(960, 540, [("yuv422p10", "709", _xf()), ("yuv422p10", "709", pip(0.5, 0.05, 0.05)),
                                                              ("yuv422p10", "709", pip(0.5, 0.45, 0.1))]),
    "all_ycbcr_formats": lambda: _mixed_format_scene(960, 540, [("yuv420p", "709", _xf()), ("nv12", "601_525", pip(0.5, 0.05, 0.05)),
                                                                ("yuv422p8", "709", pip(0.5, 0.45, 0.1)), ("v210", None, pip(0.5, 0.25, 0.45)),
                                                                ("yuv422p10", "2020", pip(0.75, 0.2, 0.2))]),
    "yuv420p_direct": lambda: _mixed_format_scene(768, 64, [("yuv420p", "709", None)]),
    "nv12_upscaled_flipped": lambda: _mixed_format_scene(960, 270, [("v210", None, _xf()), ("nv12", "709", _xf(scaleX=1.5, scaleY=2.0, flipH=True, offsetX=0.1))]),
}


@pytest.mark.parametrize("name", sorted(PLANAR_MARCH_SCENES))
def test_planar_sources_take_the_march_kernel(name):
    scene = PLANAR_MARCH_SCENES[name]()
    ref = SceneOracle(scene).packed()
    slow, st0 = run(_run_scene_variant(scene, "generic"))
    assert st0["march_launches"] == 0 and np.array_equal(slow, ref)
    for mode in ("march", "march_nocull"):
        out, st = run(_run_scene_variant(scene, mode))
        assert st["march_launches"] == 1 and st["kernel_launches"] == 1 and st["materialised"] == 0, (mode, st)
        assert np.array_equal(out, ref), f"{mode}: {int((out != ref).sum())} bytes differ"


def test_planar_10bit_samples_above_1023_match_the_reader():
    """yuv422p10 planes are 16-bit: out-of-range samples must convert exactly as the stand-alone reader converts them"""
    scene = _mixed_format_scene(480, 64, [("yuv422p10", "709", _xf())])
    rng = np.random.default_rng(9)
    scene["layers"][0]["src"] = [rng.integers(0, 65536, p.size // 2, dtype=np.uint16).astype("<u2").view(np.uint8) for p in scene["layers"][0]["src"]]
    ref = SceneOracle(scene).packed()
    out, st = run(_run_scene_variant(scene, "march"))
    assert np.array_equal(out, ref), (st, int((out != ref).sum()))


@pytest.mark.parametrize("out_fmt", ["yuv422p8", "yuv422p10", "yuv420p", "nv12"])
def test_march_kernel_planar_sink_two_fields(out_fmt):
    """interlaced consumer: top then bottom field into one set of planes; 4:2:0 keeps the bottom field's chroma (yuv420p.ts:153-200)"""
    async def go():
        scene = layered_scene(480, 270 if out_fmt.startswith("yuv422") else 268, 3, "noise", "mix", "709", "2020")
        scene["interlaced"] = True
        scene["outFmt"] = out_fmt
        async with Env() as env:
            h = ChannelHarness(env.ctx, scene, env.pj)
            await h.init()
            dests = await h.fromRGBA.createDests("il")
            for d in dests:
                d.fill(0)
                await d.hostAccess("writeonly")
            for il in (Interlace.TopField, Interlace.BottomField):
                ups = await h.upload_all(int(il))
                frame = await h.compose(ups, int(il))
                await h.consume(frame, dests, il, download=(il == Interlace.BottomField))
            assert env.ctx.stats()["march_launches"] == 2
            so = SceneOracle(scene)
            comp = so.composite()
            cw, W, H = scene["colWork"], scene["width"], scene["height"]
            lut = oracle.linear2gamma_lut(cw)
            if out_fmt.startswith("yuv422"):
                bits = 10 if out_fmt == "yuv422p10" else 8
                cm = oracle.rgb2ycbcr_matrix(cw, *((10, 64, 940, 896) if bits == 10 else (8, 16, 235, 224)))
                ref = [np.zeros(n, np.uint8) for n in oracle.yuv422p_plane_bytes(bits, W, H)]
                for il in (1, 3):
                    oracle.yuv422p_write(bits, comp, W, H, il, cm, lut, ref)
            else:
                cm = oracle.rgb2ycbcr_matrix(cw, 8, 16, 235, 224)
                ref = [np.zeros(n, np.uint8) for n in oracle.yuv420_plane_bytes(out_fmt == "nv12", W, H)]
                for il in (1, 3):
                    oracle.yuv420_write(out_fmt == "nv12", comp, W, H, il, cm, lut, ref)
            for d, r in zip(dests, ref):
                assert np.array_equal(d.host, r)
    run(go())


def test_planar_sources_with_padded_pitch_and_other_sizes():
    """planar sources whose line pitch (width rounded up to 8 samples) differs from their width, smaller than the channel:
    714 x 270 (pitch 720) and 1278 x 540 (pitch 1280) into a 960 x 540 channel, through the march kernel"""
    layers = [dict(src=make_frame("noise", 960, 540, 3), sw=960, sh=540, xf=_xf(), transition=None)]
    for i, (fmt, sw, sh, xf) in enumerate([("yuv422p10", 714, 270, _xf(scaleX=0.7, scaleY=0.7, offsetX=-0.1)),
                                           ("yuv420p", 1278, 540, pip(0.6, 0.3, 0.1)), ("nv12", 714, 270, _xf(scaleX=0.9, scaleY=0.8, offsetX=0.2, offsetY=0.1)),
                                           ("yuv422p8", 1278, 540, pip(0.55, 0.05, 0.4))]):
        layers.append(dict(src=_rand_source(fmt, sw, sh, 80 + i), sw=sw, sh=sh, xf=xf, transition=None, fmt=fmt, colRead="709"))
    scene = dict(width=960, height=540, colRead="709", colWork="2020", interlaced=False, layers=layers)
    ref = SceneOracle(scene).packed()
    out, st = run(_run_scene_variant(scene, "march"))
    assert st["march_launches"] == 1 and st["kernel_launches"] == 1 and st["materialised"] == 0, st
    assert np.array_equal(out, ref), f"{int((out != ref).sum())} bytes differ"


# ---- k_march_direct: one v210 source 1:1 into a v210 output (BASELINE.json config 2) ----
@pytest.mark.parametrize("w,h,inputs,cols", [(1920, 1080, "noise", ("709", "709")), (1920, 270, "ramp", ("709", "2020")), (384, 100, "noise", ("709", "709")),
                                             (48, 37, "noise", ("2020", "709")), (3840, 64, "noise", ("709", "2020")), (240, 11, "noise", ("601_525", "709"))])
def test_direct_kernel_matches_the_general_kernel_and_oracle(w, h, inputs, cols):
    scene = single_layer_scene(w, h, inputs, False, cols[0], cols[1])
    ref = SceneOracle(scene).packed()

    async def go(direct):
        async with Env(True) as env:
            env.ctx.directKernel = direct
            env.ctx.setOcclusionCulling(True)   # pushes the flags
            hh = ChannelHarness(env.ctx, scene, env.pj)
            await hh.init()
            out = await hh.run_frame()
            return out, env.ctx.stats()
    fast, st = run(go(True))
    slow, st2 = run(go(False))
    assert st["march_launches"] == 1 and st2["march_launches"] == 1 and st["kernel_launches"] == 1
    assert np.array_equal(slow, ref)
    assert np.array_equal(fast, ref), f"{int((fast != ref).sum())} bytes differ"


def test_direct_kernel_two_fields():
    async def go():
        scene = single_layer_scene(480, 54, "noise", False, "709", "2020")
        scene["interlaced"] = True
        async with Env() as env:
            h = ChannelHarness(env.ctx, scene, env.pj)
            await h.init()
            dests = await h.fromRGBA.createDests("il")
            dests[0].fill(0)
            await dests[0].hostAccess("writeonly")
            for il in (Interlace.TopField, Interlace.BottomField):
                ups = await h.upload_all(int(il))
                frame = await h.compose(ups, int(il))
                await h.consume(frame, dests, il, download=(il == Interlace.BottomField))
            so = SceneOracle(scene)
            ref = np.zeros_like(dests[0].host)
            so.packed(1, ref)
            so.packed(3, ref)
            assert np.array_equal(dests[0].host, ref)
    run(go())


# ---- k_march_single: one v210 layer through an axis-aligned Transform (a channel playing one clip through its Mixer) ----
SINGLE_XFS = {
    "identity_1080p": (1920, 1080, _xf()),
    "identity_small": (480, 135, _xf()),
    "shifted": (960, 270, _xf(offsetX=0.1, offsetY=-0.07)),
    "upscaled_flipped": (960, 270, _xf(scaleX=1.6, scaleY=2.2, flipH=True, flipV=True, offsetX=0.05)),
    "slightly_upscaled": (960, 540, _xf(scaleX=1.0003, scaleY=1.0007)),
    "mostly_outside": (960, 270, _xf(offsetX=0.9, offsetY=0.95)),
    "one_strip": (48, 37, _xf()),
    "uhd_slice": (3840, 48, _xf(scaleX=1.25, scaleY=1.25)),
}


@pytest.mark.parametrize("name", sorted(SINGLE_XFS))
def test_single_layer_kernel_matches_the_general_kernel_and_oracle(name):
    w, h, xf = SINGLE_XFS[name]
    scene = single_layer_scene(w, h, "noise", True, "709", "2020")
    scene["layers"][0]["xf"] = xf
    ref = SceneOracle(scene).packed()

    async def go(dedicated):
        async with Env(True) as env:
            env.ctx.directKernel = dedicated
            env.ctx.setOcclusionCulling(True)   # pushes the flags
            hh = ChannelHarness(env.ctx, scene, env.pj)
            await hh.init()
            out = await hh.run_frame()
            return out, env.ctx.stats()
    fast, st = run(go(True))
    slow, st2 = run(go(False))
    assert st["march_launches"] == 1 and st2["march_launches"] == 1 and st["kernel_launches"] == 1
    assert np.array_equal(slow, ref)
    assert np.array_equal(fast, ref), f"{int((fast != ref).sum())} bytes differ"


# ---- the background pass: the bottom layer's background-only strip-pair lines as a second phase of the general kernel ----
@pytest.mark.parametrize("name", ["north_star_mix", "north_star_wipe", "flips_and_upscale", "odd_fractions", "mostly_outside", "odd_height"])
def test_background_pass_is_exact(name, monkeypatch):
    """march_single_items<true>: gated to large frames in production (PB_BG_MIN forces it here); same bytes as without it"""
    monkeypatch.setenv("PB_BG_MIN", "0")
    scene = MARCH_SCENES[name]()
    ref = SceneOracle(scene).packed()
    out, st = run(_run_scene_variant(scene, "march"))
    assert st["march_launches"] == 1 and st["kernel_launches"] == 1
    assert np.array_equal(out, ref), f"{int((out != ref).sum())} bytes differ"
    out2, _ = run(_run_scene_variant(scene, "march_nocull"))
    assert np.array_equal(out2, ref)


def test_background_pass_full_size_4320p():
    """BASELINE config 5's shape at full size, where the background pass is on by default: march == generic byte for byte"""
    scene = layered_scene(7680, 4320, 2, "noise", "plain", "709", "2020")
    slow, _ = run(_run_scene_variant(scene, "generic"))
    fast, st = run(_run_scene_variant(scene, "march"))
    assert st["march_launches"] == 1 and st["kernel_launches"] == 1
    assert np.array_equal(fast, slow)


@pytest.mark.parametrize("fmt,w,h,xf", [("yuv422p10", 960, 270, _xf()), ("yuv420p", 960, 270, _xf(scaleX=1.3, scaleY=1.6, offsetX=0.05)),
                                        ("nv12", 480, 136, _xf(flipH=True)), ("yuv422p8", 1920, 64, _xf(offsetY=0.1))])
def test_single_layer_kernel_planar_clip(fmt, w, h, xf):
    """a channel playing one FFmpegProducer clip through its Mixer: k_march_single<planar> == general kernel == oracle"""
    scene = _mixed_format_scene(w, h, [(fmt, "709", xf)])
    if fmt == "yuv422p10":   # words above 1023 take the sample-exact path inside the single-layer kernel too
        rng = np.random.default_rng(11)
        planes = scene["layers"][0]["src"]
        y16 = planes[0].view("<u2").copy()
        y16[rng.integers(0, y16.size, 200)] = rng.integers(1024, 65536, 200, dtype=np.uint16)
        planes[0] = y16.view(np.uint8)
    ref = SceneOracle(scene).packed()

    async def go(dedicated):
        async with Env(True) as env:
            env.ctx.directKernel = dedicated
            env.ctx.setOcclusionCulling(True)   # pushes the flags
            hh = ChannelHarness(env.ctx, scene, env.pj)
            await hh.init()
            out = await hh.run_frame()
            return out, env.ctx.stats()
    fast, st = run(go(True))
    slow, st2 = run(go(False))
    assert st["march_launches"] == 1 and st2["march_launches"] == 1 and st["kernel_launches"] == 1
    assert np.array_equal(slow, ref)
    assert np.array_equal(fast, ref), f"{int((fast != ref).sum())} bytes differ"


@pytest.mark.parametrize("src_fmt,out_fmt,w,h", [("yuv422p10", "yuv422p8", 960, 270), ("v210", "yuv420p", 960, 270), ("nv12", "nv12", 480, 136),
                                                 ("yuv420p", "yuv422p10", 1920, 64), ("v210", "yuv422p8", 480, 135)])
def test_single_layer_kernel_transcodes(src_fmt, out_fmt, w, h):
    """one clip through the Mixer into an FFmpegConsumer-style planar output: k_march_single<true> == general kernel == oracle"""
    scene = _mixed_format_scene(w, h, [(src_fmt, None if src_fmt == "v210" else "709", _xf(scaleX=1.1, scaleY=1.1))])
    scene["outFmt"] = out_fmt
    ref = SceneOracle(scene).packed()

    async def go(dedicated):
        async with Env(True) as env:
            env.ctx.directKernel = dedicated
            env.ctx.setOcclusionCulling(True)   # pushes the flags
            hh = ChannelHarness(env.ctx, scene, env.pj)
            await hh.init()
            out = await hh.run_frame()
            return out, env.ctx.stats()
    fast, st = run(go(True))
    slow, st2 = run(go(False))
    assert st["march_launches"] == 1 and st2["march_launches"] == 1 and st["kernel_launches"] == 1
    assert np.array_equal(slow, ref)
    assert np.array_equal(fast, ref), f"{int((fast != ref).sum())} bytes differ"


@pytest.mark.parametrize("kind", ["direct", "single", "single_planar", "background"])
def test_chain_replay_of_the_dedicated_kernels(kind, monkeypatch):
    """a recorded chain re-issues k_march_direct / k_march_single / the two-phase general kernel with their descriptors"""
    if kind == "background":
        monkeypatch.setenv("PB_BG_MIN", "0")
    scene = {"direct": lambda: single_layer_scene(480, 135, "noise", False, "709", "2020"),
             "single": lambda: single_layer_scene(480, 135, "noise", True, "709", "2020"),
             "single_planar": lambda: _mixed_format_scene(480, 136, [("yuv420p", "709", _xf(scaleX=1.2, scaleY=1.2))]),
             "background": lambda: layered_scene(480, 270, 3, "noise", "plain", "709", "2020")}[kind]()

    async def go():
        async with Env() as env:
            h = ChannelHarness(env.ctx, scene, env.pj)
            await h.init()
            chain, dests = await h.record_chain()
            assert chain.complete and chain.launches == 1
            dests[0].fill(0)
            await dests[0].hostAccess("writeonly")
            await dests[0].hostAccess("none")
            for _ in range(2):
                chain.replay()
            await env.ctx.waitFinish(env.ctx.queue.process)
            await dests[0].hostAccess("readonly")
            assert np.array_equal(dests[0].host, SceneOracle(scene).packed())
            chain.destroy()
    run(go())


# ---- RGBA-f32 frames as leaves of the march kernel (Yadif outputs, materialised sub-expressions, image-process results) ----
def _f32_scene(w, h, specs, nan=False):
    """specs: [(fmt, xf)] with fmt 'v210' or 'rgbaf32'"""
    layers = []
    for i, (fmt, xf) in enumerate(specs):
        if fmt == "rgbaf32":
            rng = np.random.default_rng(120 + i)
            img = rng.random((h, w, 4), dtype=np.float32)
            img[..., :3] *= img[..., 3:4]          # premultiplied, like the frames the chain hands on
            if nan:
                img[h // 2, w // 3] = np.float32("nan")
                img[h // 3, w // 2, 1] = np.float32("inf")
            src = img
        else:
            src = make_frame("noise", w, h, 130 + i)
        layers.append(dict(src=src, sw=w, sh=h, xf=xf, transition=None, fmt=fmt))
    return dict(width=w, height=h, colRead="709", colWork="2020", interlaced=False, layers=layers)


F32_SCENES = {
    "f32_over_video": lambda: _f32_scene(960, 270, [("v210", _xf()), ("rgbaf32", _xf()), ("v210", pip(0.5, 0.3, 0.3))]),
    "f32_background_upscaled": lambda: _f32_scene(480, 135, [("rgbaf32", _xf(scaleX=1.5, scaleY=1.2, flipH=True)), ("v210", pip(0.5, 0.1, 0.1))]),
    "f32_pip_direct_under": lambda: _f32_scene(480, 136, [("rgbaf32", None), ("rgbaf32", pip(0.6, 0.2, 0.2))]),
    "f32_with_nan_under_opaque_video": lambda: _f32_scene(480, 135, [("rgbaf32", _xf()), ("v210", pip(0.5, 0.2, 0.2))], nan=True),
}


@pytest.mark.parametrize("name", sorted(F32_SCENES))
def test_rgba_f32_leaves_take_the_march_kernel(name):
    """bit-exact against the oracle chain and the generic kernel; NaN / inf in a frame propagate exactly as the reference's
    fma(prev, 1 - alpha, layer) propagates them (no culling under opaque layers when an RGBA-f32 frame is in the graph)"""
    scene = F32_SCENES[name]()
    ref = SceneOracle(scene).packed()
    slow, st0 = run(_run_scene_variant(scene, "generic"))
    assert st0["march_launches"] == 0 and np.array_equal(slow, ref)
    out, st = run(_run_scene_variant(scene, "march"))
    assert st["march_launches"] == 1 and st["kernel_launches"] == 1 and st["materialised"] == 0, st
    assert np.array_equal(out, ref), f"{int((out != ref).sum())} bytes differ"


@pytest.mark.parametrize("use_march", [True, False])
def test_interlaced_source_through_yadif_composites_on_the_march_kernel(use_march):
    """the north star's whole chain with its de-interlace stage: v210 (1080i-style) -> ToRGBA -> Yadif (3-frame window) ->
    Mixer Transform -> Combine with a PiP layer -> FromRGBA v210.  The Yadif output is an RGBA-f32 frame; the composite that
    follows it is one march-kernel launch, bit-exact against the oracle's stage-by-stage chain"""
    from phaneron_b200.process import v210 as v210m
    from phaneron_b200.process.combine import Combine
    from phaneron_b200.process.image_process import ImageProcess
    from phaneron_b200.process.io import FromRGBA, ToRGBA
    from phaneron_b200.process.transform import Transform
    from phaneron_b200.process.yadif import Yadif
    from scene_oracle import xf_matrix
    w, h = 480, 136
    fields = [make_frame("noise", w, h, 140 + i) for i in range(3)]
    pip_src = make_frame("noise", w, h, 150)
    pip_xf = pip(0.5, 0.3, 0.2)

    async def go():
        async with Env(True) as env:
            env.ctx.setMarchKernel(use_march)
            ctx, jobs = env.ctx, env.jobs
            to_a = ToRGBA(ctx, "709", "2020", v210m.Reader(w, h), jobs)
            to_b = ToRGBA(ctx, "709", "2020", v210m.Reader(w, h), jobs)
            frm = FromRGBA(ctx, "2020", v210m.Writer(w, h, False), jobs)
            xa = ImageProcess(ctx, Transform(ctx, w, h), jobs)
            xb = ImageProcess(ctx, Transform(ctx, w, h), jobs)
            comb = ImageProcess(ctx, Combine(2, w, h), jobs)
            yad = Yadif(ctx, jobs, w, h, {"mode": "send_frame", "tff": True}, True)
            for o in (to_a, to_b, frm, xa, xb, comb, yad):
                await o.init()
            outs = []
            for t, f in enumerate(fields):   # producer side: load, convert, de-interlace (yadif.ts:115-145)
                srcs = await to_a.createSources("src")
                for s_ in srcs:
                    s_.timestamp = t * 2
                await to_a.loadFrame(f, srcs, ctx.queue.load)
                rgba = await to_a.createDest({"width": w, "height": h}, "src")
                rgba.timestamp = t * 2
                to_a.processFrame("src", srcs, rgba)
                await yad.processFrame(rgba, outs, "src")
            assert len(outs) == 1   # the window fills at the third field: one de-interlaced frame
            deint = outs[0]
            deint.addRef()
            got_deint = await env.fetch(deint, w, h)
            before = ctx.stats()
            # mixer + second layer + combiner + consumer
            xfa = await ctx.createBuffer(w * h * 16, "readwrite", "coarse", {"width": w, "height": h}, "mixer a")
            await xa.run(dict(input=deint, output=xfa, **_xf()), {"source": "L0", "timestamp": 2}, lambda: None)
            await jobs.runQueue({"source": "L0", "timestamp": 2})
            srcs = await to_b.createSources("pip")
            await to_b.loadFrame(pip_src, srcs, ctx.queue.load)
            rgb = await to_b.createDest({"width": w, "height": h}, "pip")
            to_b.processFrame("pip", srcs, rgb)
            xfb = await ctx.createBuffer(w * h * 16, "readwrite", "coarse", {"width": w, "height": h}, "mixer b")
            await xb.run(dict(input=rgb, output=xfb, **pip_xf), {"source": "pip", "timestamp": 0}, lambda: None)
            await jobs.runQueue({"source": "pip", "timestamp": 0})
            cdest = await ctx.createBuffer(w * h * 16, "readwrite", "coarse", {"width": w, "height": h}, "comb")
            cdest.timestamp = 2
            await comb.run({"inputs": [xfa, xfb], "output": cdest}, {"source": "ch", "timestamp": 2}, lambda: None)
            await jobs.runQueue({"source": "ch", "timestamp": 2})
            dests = await frm.createDests("out")
            frm.processFrame("out", cdest, dests, Interlace.Progressive)
            await jobs.runQueue({"source": "out", "timestamp": 2})
            await frm.saveFrame(dests)
            after = ctx.stats()
            return dests[0].host.copy(), {k: after[k] - before[k] for k in after}, got_deint
    out, st, got_deint = run(go())
    cm_r, lut_r, gam = oracle.ycbcr2rgb_matrix("709"), oracle.gamma2linear_lut("709"), oracle.rgb2rgb_matrix("709", "2020")
    # (the Yadif stage itself is pinned in tests/test_gpu_ops.py; here its actual output is the oracle's input for the stages
    # that follow it, which is what this test is about)
    c_ = oracle.v210_read(fields[1], w, h, cm_r, lut_r, gam)
    kept = (got_deint.view(np.uint32) == c_.view(np.uint32)).all(axis=(1, 2))
    assert kept.sum() == h // 2, "one field of the de-interlaced frame is the current frame's"
    deint = got_deint
    la = oracle.transform(deint, xf_matrix(w, h, _xf()), w, h)
    lb = oracle.transform(oracle.v210_read(pip_src, w, h, cm_r, lut_r, gam), xf_matrix(w, h, pip_xf), w, h)
    ref = oracle.v210_write(oracle.combine([la, lb]), w, h, 0, oracle.rgb2ycbcr_matrix("2020"), oracle.linear2gamma_lut("2020"))
    assert np.array_equal(out, ref), f"{int((out != ref).sum())} bytes differ"
    assert st["march_launches"] == (1 if use_march else 0) and st["fused_launches"] == 1 and st["kernel_launches"] == 1, st


@pytest.mark.parametrize("mode,size", [("send_field", (480, 136)), ("send_frame_nospatial", (480, 136)), ("send_field", (1920, 1080))])
def test_yadif_fields_are_fused_into_the_composite_launch(mode, size):
    """Interlaced v210 frames -> ToRGBA -> Yadif (yadif.ts:88-145) -> Mixer Transform -> Combine with a PiP -> FromRGBA.
    Each ToRGBA output is made real ONCE (one direct-kernel launch per input frame: the window re-reads it for six fields);
    the de-interlaced field is a leaf of the ONE march launch of its output frame: only its interpolated lines (half a frame)
    are computed, once, by that launch's pre-pass; the lines of its own parity are read from the current frame in place.
    Bit-exact against the oracle's stage-by-stage chain (v210 read x3 -> yadif -> transform -> combine -> v210 write)."""
    from phaneron_b200.process import v210 as v210m
    from phaneron_b200.process.combine import Combine
    from phaneron_b200.process.image_process import ImageProcess
    from phaneron_b200.process.io import FromRGBA, ToRGBA
    from phaneron_b200.process.transform import Transform
    from phaneron_b200.process.yadif import Yadif
    from scene_oracle import xf_matrix
    w, h = size   # (1920 x 1080: the reference's own operating point, 1080i50 channels, at full size)
    frames = [make_frame("noise", w, h, 170 + i) for i in range(4)]
    pip_src = make_frame("noise", w, h, 180)
    pip_xf = pip(0.5, 0.3, 0.2)

    async def go():
        async with Env(True) as env:
            ctx, jobs = env.ctx, env.jobs
            to_a = ToRGBA(ctx, "709", "2020", v210m.Reader(w, h), jobs)
            to_b = ToRGBA(ctx, "709", "2020", v210m.Reader(w, h), jobs)
            frm = FromRGBA(ctx, "2020", v210m.Writer(w, h, False), jobs)
            xa = ImageProcess(ctx, Transform(ctx, w, h), jobs)
            xb = ImageProcess(ctx, Transform(ctx, w, h), jobs)
            comb = ImageProcess(ctx, Combine(2, w, h), jobs)
            yad = Yadif(ctx, jobs, w, h, {"mode": mode, "tff": True}, True)
            for o in (to_a, to_b, frm, xa, xb, comb, yad):
                await o.init()
            results, stats = [], []
            for t, f in enumerate(frames):
                s0 = ctx.stats()
                srcs = await to_a.createSources("src")
                for s_ in srcs:
                    s_.timestamp = t * 2
                await to_a.loadFrame(f, srcs, ctx.queue.load)
                rgba = await to_a.createDest({"width": w, "height": h}, "src")
                rgba.timestamp = t * 2
                to_a.processFrame("src", srcs, rgba)
                # (In the reference the newest frame's read job is still queued under its own timestamp when the window first
                # uses it as `next` -- yadif.ts:101-112 runs the queue of the CURRENT frame's timestamp -- so `next` is read
                # before it is written; test_gpu_ops.py covers that order.  Here the read is flushed first: the window holds
                # three converted frames, which is what the fused path is about.)
                if len(yad.in_) >= 2:   # (while the window fills, Yadif.processFrame flushes the read itself, yadif.ts:126-130)
                    await jobs.runQueue({"source": "src", "timestamp": t * 2})
                outs = []
                await yad.processFrame(rgba, outs, "src")
                s1 = ctx.stats()
                stats.append(("deinterlace", {k: s1[k] - s0[k] for k in s1}))
                for deint in outs:   # every de-interlaced frame goes down the channel: mixer, second layer, combiner, consumer
                    assert deint.deferred
                    b0 = ctx.stats()
                    ts = deint.timestamp
                    xfa = await ctx.createBuffer(w * h * 16, "readwrite", "coarse", {"width": w, "height": h}, "mixer a")
                    await xa.run(dict(input=deint, output=xfa, **_xf()), {"source": "L0", "timestamp": ts}, lambda d=deint: d.release())
                    await jobs.runQueue({"source": "L0", "timestamp": ts})
                    psrcs = await to_b.createSources("pip")
                    for s_ in psrcs:
                        s_.timestamp = ts
                    await to_b.loadFrame(pip_src, psrcs, ctx.queue.load)
                    rgb = await to_b.createDest({"width": w, "height": h}, "pip")
                    to_b.processFrame("pip", psrcs, rgb)
                    xfb = await ctx.createBuffer(w * h * 16, "readwrite", "coarse", {"width": w, "height": h}, "mixer b")
                    await xb.run(dict(input=rgb, output=xfb, **pip_xf), {"source": "pip", "timestamp": ts}, lambda r_=rgb: r_.release())
                    await jobs.runQueue({"source": "pip", "timestamp": ts})
                    cdest = await ctx.createBuffer(w * h * 16, "readwrite", "coarse", {"width": w, "height": h}, "comb")
                    cdest.timestamp = ts
                    await comb.run({"inputs": [xfa, xfb], "output": cdest}, {"source": "ch", "timestamp": ts}, lambda: None)
                    await jobs.runQueue({"source": "ch", "timestamp": ts})
                    xfa.release()
                    xfb.release()
                    dests = await frm.createDests("out")
                    frm.processFrame("out", cdest, dests, Interlace.Progressive)
                    await jobs.runQueue({"source": "out", "timestamp": ts})
                    await frm.saveFrame(dests)
                    b1 = ctx.stats()
                    results.append(dests[0].host.copy())
                    stats.append(("compose", {k: b1[k] - b0[k] for k in b1}))
                    for d_ in dests:
                        d_.release()
            yad.release()
            return results, stats
    results, stats = run(go())
    send_field = mode.startswith("send_field")
    skip = mode.endswith("nospatial")
    assert len(results) == (4 if send_field else 2)   # the window fills at the third frame; 1 or 2 outputs per frame from then on
    # launch structure
    for kind, st in stats:
        if kind == "compose":
            assert st["kernel_launches"] == 2 and st["march_launches"] == 1 and st["fused_launches"] == 1 and st["materialised"] == 0, st
        else:   # producer side: frames are made real when the window first needs them (one direct-kernel launch each), nothing else
            assert st["kernel_launches"] == st["materialised"] == st["march_launches"], st
    assert sum(st["materialised"] for kind, st in stats if kind == "deinterlace") == 4   # each input frame exactly once
    # oracle chain
    cm_r, lut_r, gam = oracle.ycbcr2rgb_matrix("709"), oracle.gamma2linear_lut("709"), oracle.rgb2rgb_matrix("709", "2020")
    rgba = [oracle.v210_read(f, w, h, cm_r, lut_r, gam) for f in frames]
    lb = oracle.transform(oracle.v210_read(pip_src, w, h, cm_r, lut_r, gam), xf_matrix(w, h, pip_xf), w, h)
    k = 0
    for t in (2, 3):   # window (t-2, t-1, t): the de-interlaced frame is frame t-1's
        for second in ((False, True) if send_field else (False,)):
            parity = 1 ^ (0 if second else 1)   # tff: (tff ? 1 : 0) ^ (!isSecond ? 1 : 0), yadif.ts:104
            deint = oracle.yadif(rgba[t - 2], rgba[t - 1], rgba[t], parity, True, skip)
            la = oracle.transform(deint, xf_matrix(w, h, _xf()), w, h)
            ref = oracle.v210_write(oracle.combine([la, lb]), w, h, 0, oracle.rgb2ycbcr_matrix("2020"), oracle.linear2gamma_lut("2020"))
            assert np.array_equal(results[k], ref), f"output {k}: {int((results[k] != ref).sum())} bytes differ"
            k += 1
