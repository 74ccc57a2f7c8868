"""The oracle against everything the reference pins for this path:
the fillBuf fixture, the round-trip invariant of the src/process/test scripts, and the
known-answers of colourMaths.ts.  CPU only."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
KA = json.load(open(os.path.join(HERE, "golden", "known_answers.json")))


def test_known_answer_matrices():
    np.testing.assert_array_equal(oracle.ycbcr2rgb_matrix("709"), np.array(KA["ycbcr2rgb_709"], np.float32))
    np.testing.assert_array_equal(oracle.rgb2ycbcr_matrix("709"), np.array(KA["rgb2ycbcr_709"], np.float32))
    np.testing.assert_array_equal(oracle.rgb2rgb_matrix("709", "709"), np.array(KA["rgb2rgb_709_709"], np.float32))
    np.testing.assert_array_equal(oracle.rgb2rgb_matrix("709", "2020"), np.array(KA["rgb2rgb_709_2020"], np.float32))


def test_known_answer_luts():
    g = oracle.gamma2linear_lut("709")
    l = oracle.linear2gamma_lut("709")
    for k, v in KA["gamma2linear_709"].items():
        assert g[int(k)] == np.float32(v)
    for k, v in KA["linear2gamma_709"].items():
        assert l[int(k)] == np.float32(v)
    assert g[0] == 0 and l[0] == 0
    assert np.all(np.diff(l) >= 0)
    # the EOTF as coded (knee at beta*delta) steps DOWN once, between entries 5308 and 5309
    assert list(np.nonzero(np.diff(g) < 0)[0]) == [5308]


def test_known_answers_are_reproducible_from_the_naive_script():
    out = subprocess.check_output([sys.executable, os.path.join(HERE, "golden", "make_known_answers.py")])
    fresh = json.loads(out)
    for k, v in fresh.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                assert np.float32(vv) == np.float32(KA[k][kk]), (k, kk)
        else:
            np.testing.assert_array_equal(np.array(v, np.float32), np.array(KA[k], np.float32))


def test_unknown_colourspace_falls_back_to_709():
    np.testing.assert_array_equal(oracle.gamma2linear_lut("nonsense"), oracle.gamma2linear_lut("709"))
    np.testing.assert_array_equal(oracle.rgb2rgb_matrix("nonsense", "2020"), oracle.rgb2rgb_matrix("709", "2020"))


@pytest.mark.parametrize("wh,bytes_", [((1920, 1080), 5529600), ((3840, 2160), 22118400), ((1280, 720), 2488320)])
def test_v210_geometry(wh, bytes_):
    w, h = wh
    assert oracle.v210_pitch_bytes(w) * h == bytes_ == KA["v210_frame_bytes"][f"{w}x{h}"]


def test_fillbuf_fixture():
    buf = oracle.v210_fill(1920, 1080).view(np.uint32)
    assert int(buf[0]) == int(KA["v210_fill_first_word_hex"], 16)
    # Y steps once per 6-pixel group and wraps 940 -> 64, carrying across lines
    groups = buf.reshape(-1, 4)
    y = (groups[:, 0] >> 10) & 0x3FF
    assert y[0] == 64 and y[1] == 65 and y[876] == 940 and y[877] == 64
    assert np.all((groups[:, 0] & 0x3FF) == 512) and np.all((groups[:, 0] >> 20) == 512)


def _rt(w, h, spec_r="709", spec_w="709", interlaced=False):
    src = oracle.v210_fill(w, h)
    rgba = oracle.v210_read(src, w, h, oracle.ycbcr2rgb_matrix(spec_r), oracle.gamma2linear_lut(spec_r),
                            oracle.rgb2rgb_matrix(spec_r, spec_w))
    if not interlaced:
        dst = oracle.v210_write(rgba, w, h, 0, oracle.rgb2ycbcr_matrix(spec_w), oracle.linear2gamma_lut(spec_w))
    else:
        dst = np.zeros_like(src)
        for il in (1, 3):
            oracle.v210_write(rgba, w, h, il, oracle.rgb2ycbcr_matrix(spec_w), oracle.linear2gamma_lut(spec_w), out=dst)
    return src, rgba, dst


def test_roundtrip_invariant_1080p():
    """the reference's pass criterion: src.compare(dst) === 0 (test/yuv422p10Test.ts:109)"""
    src, rgba, dst = _rt(1920, 1080)
    assert np.array_equal(src, dst)
    assert rgba[..., 3].min() == 1.0 and rgba[..., 3].max() == 1.0
    assert rgba[0, 0, 0] == 0.0 and abs(rgba[0, 876 * 6 % 1920, 0]) <= 1.0


def test_roundtrip_interlaced_1080i():
    """BASELINE config 1: read, then write TopField + BottomField into one destination"""
    src, _, dst = _rt(1920, 1080, interlaced=True)
    assert np.array_equal(src, dst)


def test_roundtrip_2160p_2020():
    src, _, dst = _rt(3840, 2160, "2020", "2020")
    assert np.array_equal(src, dst)


def test_tail_width_1280_q1_q2():
    """1280 % 48 = 32, remain 2: Q1 (alpha 0 in the read tail) and Q2 (rtz/round in the write tail)"""
    w, h = 1280, 16
    src = oracle.v210_fill(w, h)
    cm, lut, gm = oracle.ycbcr2rgb_matrix("709"), oracle.gamma2linear_lut("709"), oracle.rgb2rgb_matrix("709", "709")
    rgba = oracle.v210_read(src, w, h, cm, lut, gm)
    # the last two pixels of each line drop the matrix offset column: grey Y=~500 no longer maps to ~0.5
    body, tail = rgba[:, : w - 2, :3], rgba[:, w - 2:, :3]
    assert np.all(np.abs(body[:, :, 0] - body[:, :, 1]) < 1e-3)
    assert np.all(tail[..., 0] > 0.999) and np.all(tail[..., 2] > 0.999)   # offset dropped: R' and B' saturate the LUT
    dst = oracle.v210_write(rgba, w, h, 0, oracle.rgb2ycbcr_matrix("709"), oracle.linear2gamma_lut("709"))
    s, d = src.view(np.uint32).reshape(h, -1), dst.view(np.uint32).reshape(h, -1)
    assert np.array_equal(s[:, : (w // 6) * 4], d[:, : (w // 6) * 4])   # full groups round-trip
    assert np.all(d[:, (w // 6) * 4 + 2:] == 0)                         # padding is cleared


def test_q3_literal_offset_differs_only_for_ragged_widths():
    rng = np.random.default_rng(0)
    for w in (1920, 1280):
        rgba = rng.random((8, w, 4), dtype=np.float32)
        cm, lut = oracle.rgb2ycbcr_matrix("709"), oracle.linear2gamma_lut("709")
        a = oracle.v210_write(rgba, w, 8, 0, cm, lut, q3_literal=False)
        b = oracle.v210_write(rgba, w, 8, 0, cm, lut, q3_literal=True)
        assert np.array_equal(a, b) == (w % 48 == 0)


def test_combine_is_premultiplied_over_with_top_alpha():
    rng = np.random.default_rng(1)
    l0, l1, l2 = (rng.random((4, 5, 4), dtype=np.float32) for _ in range(3))
    out = oracle.combine([l0, l1, l2])
    k1, k2 = np.float32(1) - l1[..., 3:], np.float32(1) - l2[..., 3:]
    ref = (l0[..., :3].astype(np.float64) * k1 + l1[..., :3]) * k2 + l2[..., :3]
    np.testing.assert_allclose(out[..., :3], ref, rtol=1e-6)
    np.testing.assert_array_equal(out[..., 3], l2[..., 3])
    # an opaque top layer hides everything below, exactly
    l2[..., 3] = 1.0
    np.testing.assert_array_equal(oracle.combine([l0, l1, l2])[..., :3], l2[..., :3])


def test_dissolve_endpoints_and_wipe():
    rng = np.random.default_rng(2)
    a, b = rng.random((3, 7, 4), dtype=np.float32), rng.random((3, 7, 4), dtype=np.float32)
    np.testing.assert_array_equal(oracle.dissolve(a, b, 1.0), a)
    np.testing.assert_array_equal(oracle.dissolve(a, b, 0.0), b)
    np.testing.assert_array_equal(oracle.mix(a, b, 0.25), oracle.dissolve(a, b, 0.25))
    mask = np.zeros_like(a)
    mask[:, 4:, 0] = 1.0
    out = oracle.wipe_mask(a, b, mask)
    np.testing.assert_array_equal(out[:, :4], a[:, :4])
    np.testing.assert_array_equal(out[:, 4:], b[:, 4:])
    w = oracle.wipe(a, b, 0.5)   # x > 3.5
    np.testing.assert_array_equal(w[:, :4], a[:, :4])
    np.testing.assert_array_equal(w[:, 4:], b[:, 4:])


def test_identity_transform_is_half_pixel_box_blur_q6():
    rng = np.random.default_rng(3)
    img = rng.random((6, 8, 4), dtype=np.float32)
    out = oracle.transform(img, np.eye(3, dtype=np.float32), 8, 6)
    x, y = 3, 2
    ref = 0.25 * (img[y - 1, x - 1].astype(np.float64) + img[y - 1, x] + img[y, x - 1] + img[y, x])
    np.testing.assert_allclose(out[y, x], ref, rtol=1e-6)
    # row 0 / column 0 blend with the transparent-black CLAMP border
    np.testing.assert_allclose(out[0, 0], 0.25 * img[0, 0], rtol=1e-6)


def test_transform_far_outside_is_border():
    img = np.ones((4, 4, 4), np.float32)
    m = oracle.transform_matrix(4, 4, offset_x=5.0)
    assert np.all(oracle.transform(img, m, 4, 4) == 0)


def test_yadif_keeps_primary_field_and_interpolates_static_scene():
    rng = np.random.default_rng(4)
    f = rng.random((12, 16, 4), dtype=np.float32)
    for parity in (0, 1):
        out = oracle.yadif(f, f, f, parity, True, False)
        np.testing.assert_array_equal(out[parity::2], f[parity::2])
        out = oracle.yadif(f, f, f, parity, True, True)
        np.testing.assert_array_equal(out[parity::2], f[parity::2])
        # static scene without the spatial check: the temporal clamp (diff = 0) pins the
        # interpolated lines to the co-sited lines of the neighbouring frames
        inner = [y for y in range(2, 10) if y % 2 != parity]
        np.testing.assert_array_equal(out[inner][..., :3], f[inner][..., :3])
        np.testing.assert_array_equal(out[..., 3], f[..., 3])


def test_rgba8_roundtrip_fixture():
    """rgba8.fillBuf fixture (16, 32, 64, 255) -> read -> write round-trips (rgba8.ts:114-133)"""
    w, h = 64, 4
    src = np.tile(np.array([16, 32, 64, 255], np.uint8), w * h)
    lut_r, lut_w = oracle.gamma2linear_lut("sRGB"), oracle.linear2gamma_lut("sRGB")
    rgba = oracle.rgba8_read(src, w, h, lut_r, oracle.rgb2rgb_matrix("sRGB", "sRGB"))
    dst = oracle.rgba8_write(rgba, w, h, 0, lut_w)
    assert np.array_equal(src, dst)
    assert rgba[0, 0, 3] == 1.0


@pytest.mark.parametrize("bits,w,h", [(10, 1920, 1080), (8, 718, 1080), (8, 1920, 64), (10, 718, 64)])
def test_yuv422p_fixture_round_trips(bits, w, h):
    """the pass criterion of src/process/test/yuv422p10Test.ts / yuv422p8Test.ts (`compare() === 0`; 718 wide for the tails)"""
    rng = (10, 64, 940, 896) if bits == 10 else (8, 16, 235, 224)
    src = oracle.yuv422p_fill(bits, w, h)
    nb = oracle.yuv422p_plane_bytes(bits, w, h)
    y, u, v = src[: nb[0]], src[nb[0]: nb[0] + nb[1]], src[nb[0] + nb[1]:]
    rgba = oracle.yuv422p_read(bits, y, u, v, w, h, oracle.ycbcr2rgb_matrix("709", *rng), oracle.gamma2linear_lut("709"),
                               oracle.rgb2rgb_matrix("709", "709"))
    outs = oracle.yuv422p_write(bits, rgba, w, h, 0, oracle.rgb2ycbcr_matrix("709", *rng), oracle.linear2gamma_lut("709"))
    assert np.array_equal(np.concatenate(outs), src)


@pytest.mark.parametrize("nv12,w,h", [(False, 1920, 1080), (True, 1920, 1080), (False, 718, 64), (True, 718, 64), (False, 64, 4)])
def test_yuv420_fixture_round_trips(nv12, w, h):
    """the pass criterion of src/process/test/yuv420pTest.ts:109 / nv12Test.ts:101 (`compare() === 0`), plus 718 wide for the tails"""
    rng = (8, 16, 235, 224)
    src = oracle.yuv420_fill(nv12, w, h)
    planes, o = [], 0
    for n in oracle.yuv420_plane_bytes(nv12, w, h):
        planes.append(src[o: o + n])
        o += n
    assert o == src.size
    rgba = oracle.yuv420_read(nv12, planes, w, h, oracle.ycbcr2rgb_matrix("709", *rng), oracle.gamma2linear_lut("709"),
                              oracle.rgb2rgb_matrix("709", "709"))
    assert rgba[..., 3].min() == 1.0
    outs = oracle.yuv420_write(nv12, rgba, w, h, 0, oracle.rgb2ycbcr_matrix("709", *rng), oracle.linear2gamma_lut("709"))
    assert np.array_equal(np.concatenate(outs), src)


def test_yuv420_fixture_layout():
    """yuv420p.ts:243-280: first line of a pair ramps up (Y0, Y0 + 1), the second down (Y1 + 1, Y1); chroma 128"""
    w, h = 16, 4
    src = oracle.yuv420_fill(False, w, h)
    Y = src[: 16 * 4].reshape(4, 16)
    assert list(Y[0, :6]) == [16, 17, 18, 19, 20, 21]
    assert list(Y[1, :6]) == [235, 234, 233, 232, 231, 230]
    assert list(Y[2, :4]) == [32, 33, 34, 35]          # Y0 carried across the pair: 16 + 2 * 8
    assert list(Y[3, :4]) == [219, 218, 217, 216]
    assert set(src[64:]) == {128}
    nv = oracle.yuv420_fill(True, w, h)
    assert np.array_equal(nv[:64], src[:64]) and set(nv[64:]) == {128} and nv.size == 64 + 32


def test_yuv420_field_writes_share_the_chroma_plane():
    """a field launch writes one luma line per pair and the pair's chroma from that line (yuv420p.ts:153-200): after
    top then bottom field the luma equals the progressive result, the chroma is the bottom field's"""
    w, h = 64, 8
    rng = np.random.default_rng(5)
    rgba = rng.random((h, w, 4), dtype=np.float32)
    cm, lut = oracle.rgb2ycbcr_matrix("709", 8, 16, 235, 224), oracle.linear2gamma_lut("709")
    prog = oracle.yuv420_write(False, rgba, w, h, 0, cm, lut)
    outs = [np.zeros_like(p) for p in prog]
    oracle.yuv420_write(False, rgba, w, h, 1, cm, lut, outs)
    top_chroma = outs[1].copy()
    assert np.array_equal(top_chroma, prog[1])          # progressive chroma comes from the first (top) line too
    oracle.yuv420_write(False, rgba, w, h, 3, cm, lut, outs)
    assert np.array_equal(outs[0], prog[0])
    swapped = rgba.reshape(h // 2, 2, w, 4)[:, ::-1].reshape(h, w, 4).copy()   # bottom lines moved to the top slots
    assert np.array_equal(outs[1], oracle.yuv420_write(False, swapped, w, h, 0, cm, lut)[1])


def test_lanczos_definition_properties():
    """the Lanczos Transform filter is this repo's own definition (oracle.c; not in the reference): partition of unity,
    interpolation at integer positions when the source grid is hit exactly, rejection of rotated transforms"""
    from scene_oracle import xf_matrix
    sw, sh = 64, 48
    flat = np.full((sh, sw, 4), 0.5, np.float32)
    m = xf_matrix(sw, sh, dict(anchorX=-0.5, anchorY=-0.5, scaleX=0.5, scaleY=0.5))
    out = oracle.transform_lanczos(flat, m, sw, sh, 3)
    assert np.abs(out[10:14, 10:20] - 0.5).max() < 1e-6
    assert out[-1, -1, 3] == 0.0
    rng = np.random.default_rng(3)
    img = rng.random((sh, sw, 4), dtype=np.float32)
    # shift by exactly half a texel so that the sampling position lands on texel centres (Q6: um = x - 1/2 at identity)
    m = xf_matrix(sw, sh, dict(anchorX=-0.5, anchorY=-0.5, offsetX=0.5 / sw, offsetY=0.5 / sh))
    out = oracle.transform_lanczos(img, m, sw, sh, 3)
    assert np.abs(out[8:40, 8:56] - img[8:40, 8:56]).max() < 1e-5
    with pytest.raises(ValueError):
        oracle.transform_lanczos(img, xf_matrix(sw, sh, dict(rotate=0.05)), sw, sh, 3)
