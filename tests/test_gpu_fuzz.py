"""Seeded random layer graphs through the public operator surface, fused (march kernel and generic kernel) against the oracle's
unfused chain, byte for byte: layer counts 1-6, every Transform flavour the Mixer can produce (none, identity, scales 0.3-1.7,
offsets that push layers partly or wholly out of frame, flips, small and large rotations), dissolve and wipe transitions on any
layer, three frame sizes (ragged strip counts, odd heights), random colour-spec pairs."""
import os

import numpy as np
import pytest

from phaneron_b200.scenes import IDENTITY_XF, layered_scene, make_frame, ramp_frame

from gpu_util import run
from scene_oracle import SceneOracle
from test_gpu_chain import _run_scene_variant

pytestmark = pytest.mark.gpu

SIZES = [(480, 270), (528, 97), (960, 136)]
SPECS = ["601_525", "709", "2020"]


def _random_xf(rng):
    kind = rng.integers(0, 6)
    if kind == 0:
        return None                      # ToRGBA output straight into the combiner
    if kind == 1:
        return dict(IDENTITY_XF)         # the Mixer's identity Transform (half-texel blur: transform.ts samples at x/w)
    xf = dict(IDENTITY_XF)
    s = float(rng.choice([0.3, 0.5, 0.62, 0.75, 1.0, 1.25, 1.7]))
    xf["scaleX"], xf["scaleY"] = s, float(s * rng.choice([1.0, 1.0, 0.8, 1.3]))
    xf["offsetX"], xf["offsetY"] = float(rng.uniform(-0.7, 0.7)), float(rng.uniform(-0.7, 0.7))
    if kind == 3:
        xf["flipH"], xf["flipV"] = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    if kind >= 4:
        xf["rotate"] = float(rng.choice([0.01, -0.04, 0.25, -0.5, 0.125]))
    if rng.integers(0, 4) == 0:
        xf["anchorX"], xf["anchorY"] = float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-0.3, 0.3))
    return xf


def _random_scene(seed):
    rng = np.random.default_rng(1000 + seed)
    w, h = SIZES[seed % len(SIZES)]
    n = int(rng.integers(1, 7))
    spec_r, spec_w = str(rng.choice(SPECS)), str(rng.choice(SPECS))
    scene = layered_scene(w, h, n, "noise", "plain", spec_r, spec_w, frame_set=seed)
    for i, L in enumerate(scene["layers"]):
        L["xf"] = _random_xf(rng) if (i > 0 or rng.integers(0, 3)) else dict(IDENTITY_XF)
        t = rng.integers(0, 5)
        if t == 0:
            L["transition"] = dict(type="dissolve", mix=float(rng.choice([0.0, 0.25, 0.5, 0.9, 1.0])), src=make_frame("noise", w, h, 500 + seed * 8 + i),
                                   sw=w, sh=h, xf=L["xf"] if rng.integers(0, 2) else _random_xf(rng))
        elif t == 1:
            L["transition"] = dict(type="wipe", src=make_frame("noise", w, h, 600 + seed * 8 + i), sw=w, sh=h, xf=L["xf"],
                                   mask=ramp_frame(w, h, 13 + i), mask_sw=w, mask_sh=h, mask_xf=dict(IDENTITY_XF) if rng.integers(0, 2) else None)
    return scene


@pytest.mark.parametrize("seed", range(int(os.environ.get("PB_FUZZ_FIRST", "0")), int(os.environ.get("PB_FUZZ_FIRST", "0")) + int(os.environ.get("PB_FUZZ_SEEDS", "24"))))
def test_random_layer_graphs_match_the_oracle(seed):
    scene = _random_scene(seed)
    ref = SceneOracle(scene).packed()
    out, st = run(_run_scene_variant(scene, "march"))
    assert np.array_equal(out, ref), f"march path, seed {seed}: {int((out != ref).sum())} bytes differ ({st})"
    slow, _ = run(_run_scene_variant(scene, "generic"))
    assert np.array_equal(slow, ref), f"generic path, seed {seed}: {int((slow != ref).sum())} bytes differ"


FORMATS = [("v210", None), ("yuv422p10", "709"), ("yuv422p8", "709"), ("yuv420p", "709"), ("nv12", "601_525"), ("rgba8", "sRGB"), ("bgra8", "sRGB")]
SINKS = [None, "yuv422p10", "yuv422p8", "yuv420p", "nv12", "rgba8", "bgra8"]


def _random_format_scene(seed):
    from test_gpu_chain import _mixed_format_scene
    rng = np.random.default_rng(5000 + seed)
    w, h = [(480, 270), (960, 136), (528, 98)][seed % 3]   # (4:2:0 formats need an even height)
    n = int(rng.integers(1, 5))
    specs = []
    for i in range(n):
        fmt, col = FORMATS[int(rng.integers(0, len(FORMATS)))]
        xf = _random_xf(rng)
        if xf is not None and rng.integers(0, 5) == 0 and "rotate" not in xf and fmt not in ("rgba8", "bgra8"):
            xf["filter"] = f"lanczos{int(rng.choice([2, 3]))}"   # the extension of DESIGN.md 4.6 (axis-aligned, YCbCr sources)
        specs.append((fmt, col, xf))
    scene = _mixed_format_scene(w, h, specs)
    sink = SINKS[int(rng.integers(0, len(SINKS)))]
    if sink:
        scene["outFmt"] = sink
        if sink in ("rgba8", "bgra8"):
            scene["colWrite"] = "sRGB"
    return scene


@pytest.mark.parametrize("seed", range(int(os.environ.get("PB_FUZZ_FIRST", "0")), int(os.environ.get("PB_FUZZ_FIRST", "0")) + int(os.environ.get("PB_FUZZ_SEEDS", "24"))))
def test_random_source_and_consumer_formats_match_the_oracle(seed):
    """every Reader format as a layer (planar 4:2:2 / 4:2:0, rgba8 with its alpha), every Writer format as the sink, Lanczos on some"""
    scene = _random_format_scene(seed)
    ref = SceneOracle(scene).packed()
    out, st = run(_run_scene_variant(scene, "march"))
    assert out.shape == ref.shape and np.array_equal(out, ref), f"march path, seed {seed}: {int((out != ref).sum())} bytes differ ({st})"
    slow, _ = run(_run_scene_variant(scene, "generic"))
    assert np.array_equal(slow, ref), f"generic path, seed {seed}: {int((slow != ref).sum())} bytes differ"


def test_interleaved_replays_of_different_chains_stay_exact():
    """Eight different frames (random layer graphs and random formats: fast, general, single-layer, direct, rotated and Lanczos
    launches, pre-passes included) recorded as chains in ONE context and replayed 160 times in random order with no
    synchronisation in between: consecutive launches of different kernels overlap their prologues (programmatic dependent launch),
    scratch blocks cycle through the pool.  Every chain's destination must still hold its oracle frame."""
    from phaneron_b200.harness import ChannelHarness
    from gpu_util import Env

    scenes = [_random_scene(s) for s in (3, 7, 11, 18)] + [_random_format_scene(s) for s in (2, 5, 9, 14)]
    refs = [SceneOracle(sc).packed() for sc in scenes]

    async def go():
        async with Env(True) as env:
            hs, chains, dests = [], [], []
            for i, sc in enumerate(scenes):
                h = ChannelHarness(env.ctx, sc, env.pj, chanID=f"fz{i}")
                await h.init()
                chain, d = await h.record_chain()
                assert chain.complete, f"scene {i} is not replayable"
                hs.append(h); chains.append(chain); dests.append(d)
            order = np.random.default_rng(77).integers(0, len(scenes), 160)
            for k in order:
                chains[int(k)].replay()
            await env.ctx.waitFinish(env.ctx.queue.process)
            outs = []
            for h, d in zip(hs, dests):
                await h.fromRGBA.saveFrame(d, env.ctx.queue.unload)
                await env.ctx.waitFinish(env.ctx.queue.unload)
                outs.append(d[0].host.copy() if len(d) == 1 else np.concatenate([b.host for b in d]))   # (as ChannelHarness.run_frame)
            return outs
    outs = run(go())
    for i, (o, r) in enumerate(zip(outs, refs)):
        assert o.shape == r.shape and np.array_equal(o, r), f"chain {i}: {int((o != r).sum()) if o.shape == r.shape else 'shape'} bytes differ"
