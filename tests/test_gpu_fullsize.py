"""BASELINE.json's configurations at FULL size against the oracle, byte for byte.

The frames bench.py times (and the other configs' frames) are compared with the oracle's unfused
stage sequence (v210.ts:25-195, transform.ts:36-59, transition.ts:60-73, combine.ts:24-68 as restated
in oracle/oracle.c), not with another kernel of this repo.  The oracle needs a few seconds per frame
at these sizes (OpenMP over lines)."""
import numpy as np
import pytest

from phaneron_b200.harness import ChannelHarness
from phaneron_b200.scenes import layered_scene, single_layer_scene

from gpu_util import Env, run
from scene_oracle import SceneOracle

pytestmark = pytest.mark.gpu


async def _frame_and_stats(scene):
    async with Env() as env:
        h = ChannelHarness(env.ctx, scene, env.pj)
        await h.init()
        before = env.ctx.stats()
        out = await h.run_frame()
        after = env.ctx.stats()
        return out, {k: after[k] - before[k] for k in after}


def _assert_same(out, ref, what):
    if not np.array_equal(out, ref):
        bad = np.flatnonzero(out != ref)
        raise AssertionError(f"{what}: {bad.size} of {ref.size} bytes differ from the oracle, first at byte {int(bad[0])}")


@pytest.mark.parametrize("inputs", ["noise", "ramp"])
def test_bench_scene_2160p_matches_oracle(inputs):
    """config 3 -- exactly the scene bench.py replays: 3840x2160, 4 layers, top layer dissolved with a 5th source, 709 -> 2020"""
    scene = layered_scene(3840, 2160, 4, inputs, "mix", "709", "2020")
    out, st = run(_frame_and_stats(scene))
    _assert_same(out, SceneOracle(scene).packed(), f"2160p 4-layer mix {inputs}")
    assert st["kernel_launches"] == 1 and st["march_launches"] == 1 and st["materialised"] == 0


@pytest.mark.parametrize("variant", ["plain", "wipe"])
def test_2160p_other_variants_match_oracle(variant):
    scene = layered_scene(3840, 2160, 4, "noise", variant, "709", "2020")
    out, st = run(_frame_and_stats(scene))
    _assert_same(out, SceneOracle(scene).packed(), f"2160p 4-layer {variant}")
    assert st["kernel_launches"] == 1 and st["march_launches"] == 1


def test_4320p_two_layer_matches_oracle():
    """config 5's frame geometry with the reference's bilinear sampler: 7680x4320, L1 identity + L2 0.5x PiP"""
    scene = layered_scene(7680, 4320, 2, "noise", "plain", "709", "2020")
    out, st = run(_frame_and_stats(scene))
    _assert_same(out, SceneOracle(scene).packed(), "4320p 2-layer")
    assert st["kernel_launches"] == 1 and st["march_launches"] == 1


@pytest.mark.parametrize("with_mixer", [False, True])
@pytest.mark.parametrize("inputs", ["noise", "ramp"])
def test_1080p_single_layer_matches_oracle(with_mixer, inputs):
    """config 2: 1920x1080 ToRGBA -> Combine passthrough -> FromRGBA, with and without the Mixer's identity Transform"""
    scene = single_layer_scene(1920, 1080, inputs, with_mixer, "709", "709")
    out, st = run(_frame_and_stats(scene))
    _assert_same(out, SceneOracle(scene).packed(), f"1080p single mixer={with_mixer} {inputs}")
    assert st["kernel_launches"] == 1


def test_1080p_two_layer_channel_matches_oracle():
    """a config-4 channel without its ROUTE: 1920x1080, own source + a 0.5x PiP"""
    scene = layered_scene(1920, 1080, 2, "noise", "plain", "709", "709")
    out, st = run(_frame_and_stats(scene))
    _assert_same(out, SceneOracle(scene).packed(), "1080p 2-layer")
    assert st["kernel_launches"] == 1
