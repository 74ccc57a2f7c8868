"""Oracle-side evaluation of a harness scene: the reference's UNFUSED stage sequence
(v210 read -> transform -> transition -> combine -> v210 write) with RGBA-f32
intermediates, every stage the CPU restatement in oracle/."""
from __future__ import annotations

import numpy as np

import oracle

_XF_KEYS = ("flipH", "flipV", "anchorX", "anchorY", "scaleX", "scaleY", "offsetX", "offsetY", "rotate")


def _or(v, d):
    return v if v else d   # JS `x || d`


def xf_matrix(W, H, xf):
    return oracle.transform_matrix(W, H, bool(xf.get("flipH")), bool(xf.get("flipV")), _or(xf.get("anchorX"), 0.0),
                                   _or(xf.get("anchorY"), 0.0), _or(xf.get("scaleX"), 1.0), _or(xf.get("scaleY"), 1.0),
                                   _or(xf.get("offsetX"), 0.0), _or(xf.get("offsetY"), 0.0), _or(xf.get("rotate"), 0.0))


class SceneOracle:
    def __init__(self, scene):
        self.s = scene
        self.W, self.H = scene["width"], scene["height"]
        cr, cw = scene.get("colRead", "709"), scene.get("colWork", "709")
        self.cm_r = oracle.ycbcr2rgb_matrix(cr)
        self.lut_r = oracle.gamma2linear_lut(cr)
        self.gamut = oracle.rgb2rgb_matrix(cr, cw)
        self.cm_w = oracle.rgb2ycbcr_matrix(cw)
        self.lut_w = oracle.linear2gamma_lut(cw)

    def read(self, src, sw, sh, fmt="v210", colRead=None):
        """the Reader kernel of the layer's source format with the constants its Loader would upload (loadSave.ts:41-64)"""
        if fmt == "rgbaf32":   # already an RGBA-f32 frame in the working space
            return np.ascontiguousarray(src, np.float32).reshape(sh, sw, 4)
        if fmt == "v210" and colRead is None:
            return oracle.v210_read(src, sw, sh, self.cm_r, self.lut_r, self.gamut)
        cr = colRead or self.s.get("colRead", "709")
        lut, gamut = oracle.gamma2linear_lut(cr), oracle.rgb2rgb_matrix(cr, self.s.get("colWork", "709"))
        if fmt == "v210":
            return oracle.v210_read(src, sw, sh, oracle.ycbcr2rgb_matrix(cr), lut, gamut)
        if fmt in ("rgba8", "bgra8"):
            return oracle.rgba8_read(src, sw, sh, lut, gamut, bgra=(fmt == "bgra8"))
        if fmt in ("yuv422p10", "yuv422p8"):
            bits = 10 if fmt == "yuv422p10" else 8
            cm = oracle.ycbcr2rgb_matrix(cr, *((10, 64, 940, 896) if bits == 10 else (8, 16, 235, 224)))
            return oracle.yuv422p_read(bits, *src, sw, sh, cm, lut, gamut)
        return oracle.yuv420_read(fmt == "nv12", src, sw, sh, oracle.ycbcr2rgb_matrix(cr, 8, 16, 235, 224), lut, gamut)

    def source(self, src, sw, sh, xf, fmt="v210", colRead=None):
        rgba = self.read(src, sw, sh, fmt, colRead)
        if xf is None:
            return rgba
        flt = xf.get("filter")
        if flt:   # extension: 'lanczosN' (definition in oracle/oracle.c)
            return oracle.transform_lanczos(rgba, xf_matrix(self.W, self.H, xf), self.W, self.H, int(flt[7:]))
        return oracle.transform(rgba, xf_matrix(self.W, self.H, xf), self.W, self.H)

    def layer(self, L):
        a = self.source(L["src"], L["sw"], L["sh"], L.get("xf"), L.get("fmt", "v210"), L.get("colRead"))
        t = L.get("transition")
        if not t:
            return a
        b = self.source(t["src"], t["sw"], t["sh"], t.get("xf"))
        if t["type"] == "dissolve":
            return oracle.dissolve(a, b, t["mix"])
        m = self.source(t["mask"], t["mask_sw"], t["mask_sh"], t.get("mask_xf"))
        return oracle.wipe_mask(a, b, m)

    def composite(self):
        layers = [self.layer(L) for L in self.s["layers"]]
        return layers[0] if len(layers) == 1 else oracle.combine(layers)

    def packed(self, interlace=0, out=None):
        fmt = self.s.get("outFmt", "v210")
        if fmt == "v210":
            return oracle.v210_write(self.composite(), self.W, self.H, interlace, self.cm_w, self.lut_w, out=out)
        # the other Writer PackImpls, with the constants their Saver uploads (loadSave.ts:130-150); planes concatenated
        assert out is None
        cw = self.s.get("colWrite", self.s.get("colWork", "709"))
        lut = oracle.linear2gamma_lut(cw)
        rgba = self.composite()
        if fmt in ("rgba8", "bgra8"):
            return oracle.rgba8_write(rgba, self.W, self.H, interlace, lut, bgra=(fmt == "bgra8"))
        if fmt in ("yuv422p10", "yuv422p8"):
            bits = 10 if fmt == "yuv422p10" else 8
            cm = oracle.rgb2ycbcr_matrix(cw, *((10, 64, 940, 896) if bits == 10 else (8, 16, 235, 224)))
            return np.concatenate(oracle.yuv422p_write(bits, rgba, self.W, self.H, interlace, cm, lut))
        return np.concatenate(oracle.yuv420_write(fmt == "nv12", rgba, self.W, self.H, interlace, oracle.rgb2ycbcr_matrix(cw, 8, 16, 235, 224), lut))
