"""How tests/golden/known_answers.json was produced.

The reference (TypeScript + nodencl + an OpenCL device) cannot execute in this image, so
these are NOT outputs of the reference binary.  They are the values its formulas
(src/process/colourMaths.ts, src/process/v210.ts:198-236) yield when restated in numpy
with the same Float32Array rounding points and double accumulation, computed by the
survey independently of oracle/oracle.c and recorded in SURVEY.md section 8c.  This
script re-derives them with a third, deliberately naive implementation so that the JSON
is reproducible from this file alone:  python tests/golden/make_known_answers.py
"""
import json

import numpy as np

f32 = np.float32
P709 = dict(kR=0.2126, kB=0.0722, rx=0.64, ry=0.33, gx=0.3, gy=0.6, bx=0.15, by=0.06, wx=0.3127, wy=0.329)
P2020 = dict(kR=0.2627, kB=0.0593, rx=0.708, ry=0.292, gx=0.17, gy=0.797, bx=0.131, by=0.046, wx=0.3127, wy=0.329)


def mm(a, b):
    a, b = np.asarray(a, f32), np.asarray(b, f32)
    out = np.zeros((a.shape[0], b.shape[1]), f32)
    for i in range(a.shape[0]):
        for j in range(b.shape[1]):
            s = 0.0
            for k in range(a.shape[1]):
                s = s + float(a[i, k]) * float(b[k, j])
            out[i, j] = f32(s)
    return out


def inv3(a):
    a = np.asarray(a, f32)
    minors = np.zeros((3, 3), f32)
    for i in range(3):
        for j in range(3):
            ys = [0, 2] if i == 1 else [(i + 1) % 3, (i + 2) % 3]
            xs = [0, 2] if j == 1 else [(j + 1) % 3, (j + 2) % 3]
            minors[i, j] = f32(float(a[ys[0], xs[0]]) * float(a[ys[1], xs[1]]) - float(a[ys[0], xs[1]]) * float(a[ys[1], xs[0]]))
    cof = np.array([[f32(float(minors[i, j]) * (-1.0) ** (i + j)) for j in range(3)] for i in range(3)], f32)
    det = float(a[0, 0]) * float(minors[0, 0]) - float(a[0, 1]) * float(minors[0, 1]) + float(a[0, 2]) * float(minors[0, 2])
    return np.array([[f32(float(cof.T[i, j]) * (1.0 / det)) for j in range(3)] for i in range(3)], f32)


def rgb2xyz(p):
    w = np.array([[p["wx"]], [p["wy"]], [1.0 - p["wx"] - p["wy"]]], f32)
    W = np.array([[f32(float(v[0]) * (1.0 / float(w[1, 0])))] for v in w], f32)
    xyz = np.array([[p["rx"], p["gx"], p["bx"]], [p["ry"], p["gy"], p["by"]],
                    [1.0 - p["rx"] - p["ry"], 1.0 - p["gx"] - p["gy"], 1.0 - p["bx"] - p["by"]]], f32)
    f = mm(inv3(xyz), W)
    return mm(xyz, np.diag(f[:, 0]).astype(f32))


def ycbcr2rgb(p):
    kR, kB = p["kR"], p["kB"]
    kG = 1.0 - kR - kB
    col = np.array([[1.0, 0.0, 1.0 - kR], [1.0, (-(1.0 - kB) * kB) / kG, (-(1.0 - kR) * kR) / kG], [1.0, 1.0 - kB, 0.0]], f32)
    sc = np.array([[1.0 / 876, 0, 0, -64 / 876], [0, (1.0 / 896) * 2, 0, -(512 / 896) * 2], [0, 0, (1.0 / 896) * 2, -(512 / 896) * 2]], f32)
    return mm(col, sc)


def rgb2ycbcr(p):
    kR, kB = p["kR"], p["kB"]
    kG = 1.0 - kR - kB
    sc = np.diag([876.0, 448.0, 448.0]).astype(f32)
    col = np.array([[kR, kG, kB, 64 / 876], [-kR / (1.0 - kB), -kG / (1.0 - kB), 1.0, (512 / 896) * 2.0],
                    [1.0, -kG / (1.0 - kR), -kB / (1.0 - kR), (512 / 896) * 2.0]], f32)
    return mm(sc, col)


def g2l(i):
    fi = i / 65535
    return float(f32(fi / 4.5)) if fi < 0.018 * 4.5 else float(f32(((fi + 0.099) / 1.099) ** (1 / 0.45)))


def l2g(i):
    fi = i / 65535
    return float(f32(fi * 4.5)) if fi < 0.018 else float(f32(1.099 * fi ** 0.45 - 0.099))


if __name__ == "__main__":
    out = {
        "ycbcr2rgb_709": ycbcr2rgb(P709).astype(float).tolist(),
        "rgb2ycbcr_709": rgb2ycbcr(P709).astype(float).tolist(),
        "rgb2rgb_709_709": mm(inv3(rgb2xyz(P709)), rgb2xyz(P709)).astype(float).tolist(),
        "rgb2rgb_709_2020": mm(inv3(rgb2xyz(P2020)), rgb2xyz(P709)).astype(float).tolist(),
        "gamma2linear_709": {str(i): g2l(i) for i in (1, 4718, 4719, 32768, 65535)},
        "linear2gamma_709": {str(i): l2g(i) for i in (1, 1179, 1180, 32768, 65535)},
    }
    print(json.dumps(out, indent=1))
