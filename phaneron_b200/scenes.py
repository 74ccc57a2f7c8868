"""Synthetic frames and layer scenes shared by tests and bench.py.

  ramp  = the reference's own fixture v210.fillBuf (v210.ts:206-236), rotated per layer
  noise = seeded uniform R'G'B' in [0,1) converted to legal, in-gamut 10-bit 4:2:2
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np

from .process import v210


def pack_v210(Y: np.ndarray, Cb: np.ndarray, Cr: np.ndarray, width: int, height: int) -> np.ndarray:
    """Y: (h, w) codes; Cb/Cr: (h, w/2) codes -> v210 bytes.  width must be a multiple of 6."""
    assert width % 6 == 0
    pitch = v210.getPitchBytes(width)
    out = np.zeros((height, pitch // 4), np.uint32)
    g = width // 6
    Y = Y.astype(np.uint32).reshape(height, g, 6)
    Cb = Cb.astype(np.uint32).reshape(height, g, 3)
    Cr = Cr.astype(np.uint32).reshape(height, g, 3)
    w = out[:, : g * 4].reshape(height, g, 4)
    w[:, :, 0] = (Cr[:, :, 0] << 20) | (Y[:, :, 0] << 10) | Cb[:, :, 0]
    w[:, :, 1] = (Y[:, :, 2] << 20) | (Cb[:, :, 1] << 10) | Y[:, :, 1]
    w[:, :, 2] = (Cb[:, :, 2] << 20) | (Y[:, :, 3] << 10) | Cr[:, :, 1]
    w[:, :, 3] = (Y[:, :, 5] << 20) | (Cr[:, :, 2] << 10) | Y[:, :, 4]
    return out.view(np.uint8).reshape(-1)


def ramp_frame(width: int, height: int, rotate_groups: int = 0) -> np.ndarray:
    buf = np.zeros(v210.getPitchBytes(width) * height, np.uint8)
    v210.fillBuf(buf, width, height)
    if rotate_groups and width % 48 == 0:
        g = buf.view(np.uint32).reshape(-1, 4)
        buf = np.roll(g, rotate_groups, axis=0).reshape(-1).view(np.uint8).copy()
    return buf


def noise_frame(width: int, height: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    rgb = rng.random((height, width, 3), dtype=np.float32)
    kr, kb = 0.2126, 0.0722
    kg = 1.0 - kr - kb
    yp = kr * rgb[..., 0] + kg * rgb[..., 1] + kb * rgb[..., 2]
    cb = (rgb[..., 2] - yp) / (2 * (1 - kb))
    cr = (rgb[..., 0] - yp) / (2 * (1 - kr))
    Y = np.clip(np.rint(64 + 876 * yp), 64, 940)
    Cb = np.clip(np.rint(512 + 896 * cb[:, 0::2]), 64, 960)
    Cr = np.clip(np.rint(512 + 896 * cr[:, 0::2]), 64, 960)
    return pack_v210(Y, Cb, Cr, width, height)


def make_frame(kind: str, width: int, height: int, index: int) -> np.ndarray:
    if kind == "ramp":
        return ramp_frame(width, height, index * 97)
    if kind == "noise":
        return noise_frame(width, height, 1000 + index)
    raise ValueError(f"unknown input kind '{kind}'")


def pip(scale: float, x: float, y: float) -> Dict[str, Any]:
    """AMCP `MIXER FILL x y scale scale` as the Mixer hands it to Transform (mixer.ts:209-220):
    anchor = (0,0)-0.5, offset = -fill offset, scale = fill scale."""
    return dict(anchorX=-0.5, anchorY=-0.5, scaleX=scale, scaleY=scale, offsetX=-x, offsetY=-y, rotate=0.0,
                flipH=False, flipV=False)


IDENTITY_XF = dict(anchorX=-0.5, anchorY=-0.5, scaleX=1.0, scaleY=1.0, offsetX=0.0, offsetY=0.0, rotate=0.0,
                   flipH=False, flipV=False)


def layered_scene(width: int, height: int, n_layers: int = 4, inputs: str = "ramp", variant: str = "plain",
                  colRead: str = "709", colWork: str = "2020", frame_set: int = 0) -> Dict[str, Any]:
    """The 'honest N-layer' scene of SURVEY 8(d): L1 identity full frame, upper layers
    MIXER FILL at scale 0.5 with overlapping offsets so every source texel of every layer is
    sampled.  variant: 'plain' | 'mix' (top layer in a dissolve with an extra source at
    mix=0.5) | 'wipe' (top layer wiped against an extra source through a mask frame)."""
    offsets = [(0.05, 0.05), (0.45, 0.10), (0.25, 0.45), (0.10, 0.40), (0.40, 0.40), (0.30, 0.20), (0.20, 0.30)]
    base = frame_set * 16
    layers: List[Dict[str, Any]] = []
    for i in range(n_layers):
        xf = dict(IDENTITY_XF) if i == 0 else pip(0.5, *offsets[(i - 1) % len(offsets)])
        layers.append(dict(src=make_frame(inputs, width, height, base + i), sw=width, sh=height, xf=xf, transition=None))
    if variant == "mix":
        layers[-1]["transition"] = dict(type="dissolve", mix=0.5, src=make_frame(inputs, width, height, base + 8),
                                        sw=width, sh=height, xf=layers[-1]["xf"])
    elif variant == "wipe":
        layers[-1]["transition"] = dict(type="wipe", src=make_frame(inputs, width, height, base + 8), sw=width, sh=height,
                                        xf=layers[-1]["xf"], mask=ramp_frame(width, height, 13), mask_sw=width,
                                        mask_sh=height, mask_xf=dict(IDENTITY_XF))
    elif variant != "plain":
        raise ValueError(variant)
    return dict(width=width, height=height, colRead=colRead, colWork=colWork, interlaced=False, layers=layers)


def single_layer_scene(width: int, height: int, inputs: str = "ramp", with_mixer: bool = False,
                       colRead: str = "709", colWork: str = "709", frame_set: int = 0) -> Dict[str, Any]:
    """BASELINE.json config 2: one layer ToRGBA -> (Combine passthrough) -> FromRGBA."""
    return dict(width=width, height=height, colRead=colRead, colWork=colWork, interlaced=False,
                layers=[dict(src=make_frame(inputs, width, height, frame_set * 16), sw=width, sh=height,
                             xf=dict(IDENTITY_XF) if with_mixer else None, transition=None)])


def planar_frame(fmt: str, width: int, height: int, seed: int) -> List[np.ndarray]:
    """seeded legal-range samples in the plane layout of an FFmpegProducer format (yuv422p10 / yuv422p8 / yuv420p / nv12)"""
    rng = np.random.default_rng(seed)
    pitch = (width + 7) // 8 * 8
    if fmt == "yuv422p10":
        mk = lambda n, lo, hi: rng.integers(lo, hi, n, dtype=np.uint16).astype("<u2").view(np.uint8)
        return [mk(pitch * height, 64, 941), mk(pitch // 2 * height, 64, 961), mk(pitch // 2 * height, 64, 961)]
    mk8 = lambda n, lo, hi: rng.integers(lo, hi, n, dtype=np.uint8)
    if fmt == "yuv422p8":
        return [mk8(pitch * height, 16, 236), mk8(pitch // 2 * height, 16, 241), mk8(pitch // 2 * height, 16, 241)]
    if fmt == "yuv420p":
        return [mk8(pitch * height, 16, 236), mk8(pitch * height // 4, 16, 241), mk8(pitch * height // 4, 16, 241)]
    if fmt == "nv12":
        return [mk8(pitch * height, 16, 236), mk8(pitch * height // 2, 16, 241)]
    raise ValueError(fmt)


def planar_layered_scene(width: int, height: int, fmt: str = "yuv422p10", n_layers: int = 4, colRead: str = "709", colWork: str = "2020",
                         frame_set: int = 0) -> Dict[str, Any]:
    """the layered scene of SURVEY 8(d) with FFmpegProducer-format sources instead of v210"""
    offsets = [(0.05, 0.05), (0.45, 0.10), (0.25, 0.45), (0.10, 0.40)]
    layers = []
    for i in range(n_layers):
        xf = dict(IDENTITY_XF) if i == 0 else pip(0.5, *offsets[(i - 1) % len(offsets)])
        layers.append(dict(src=planar_frame(fmt, width, height, 2000 + frame_set * 16 + i), sw=width, sh=height, xf=xf, transition=None, fmt=fmt))
    return dict(width=width, height=height, colRead=colRead, colWork=colWork, interlaced=False, layers=layers)


def overlay_scene(width: int, height: int, inputs: str = "noise", colRead: str = "709", colWork: str = "2020", frame_set: int = 0) -> Dict[str, Any]:
    """video + graphics: L1 full-frame v210, L2-L3 0.5x v210 PiPs, L4 a full-frame rgba8 graphic whose alpha is a soft-edged
    lower third (opaque band, 32-line ramps, transparent elsewhere) -- the CG-over-video picture of a broadcast channel"""
    sc = layered_scene(width, height, 3, inputs, "plain", colRead, colWork, frame_set)
    rng = np.random.default_rng(3000 + frame_set)
    g = rng.integers(0, 256, (height, width, 4), dtype=np.uint8)
    alpha = np.zeros(height, np.float32)
    top, bot = int(height * 0.72), int(height * 0.9)
    alpha[top:bot] = 1.0
    ramp = min(32, top, height - bot)
    if ramp > 0:
        alpha[top - ramp:top] = np.linspace(0, 1, ramp, endpoint=False)
        alpha[bot:bot + ramp] = np.linspace(1, 0, ramp, endpoint=False)
    g[..., 3] = np.rint(alpha * 255).astype(np.uint8)[:, None]
    sc["layers"].append(dict(src=g.reshape(-1), sw=width, sh=height, xf=dict(IDENTITY_XF), transition=None, fmt="rgba8", colRead="sRGB"))
    return sc
