"""Compare the CPU oracle (and, optionally, the CUDA path) with the reference's own OpenCL kernels run on the
NVIDIA OpenCL driver (oracle/ref_ocl).  Prints one line per stage: elements, mismatching elements, max ulp distance."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from oracle import ref_ocl
from phaneron_b200.scenes import make_frame, pip, IDENTITY_XF
from scene_oracle import xf_matrix


def ulp(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64); b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a); b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


def report(name, ref, got):
    if ref.dtype == np.float32:
        d = ulp(ref, got); n = int((d != 0).sum())
        print(f"{name:46s} n={ref.size:9d} differ={n:8d} ({100.0*n/ref.size:7.4f}%) max_ulp={int(d.max())}", flush=True)
    else:
        n = int((ref != got).sum())
        print(f"{name:46s} n={ref.size:9d} differ={n:8d} bytes", flush=True)
    return n


def main():
    ref_ocl.build() if os.path.isdir("/root/reference") else None
    if not ref_ocl.available():
        print("reference OpenCL unavailable:", ref_ocl.why_unavailable()); return 1
    print("reference kernels running on:", ref_ocl.device_name())
    opts = sys.argv[1] if len(sys.argv) > 1 else ""
    w, h = 1920, 1080
    cm_r, lut_r, gam = oracle.ycbcr2rgb_matrix("709"), oracle.gamma2linear_lut("709"), oracle.rgb2rgb_matrix("709", "2020")
    cm_w, lut_w = oracle.rgb2ycbcr_matrix("2020"), oracle.linear2gamma_lut("2020")
    for kind in ("ramp", "noise"):
        src = make_frame(kind, w, h, 0)
        o = oracle.v210_read(src, w, h, cm_r, lut_r, gam)
        r = ref_ocl.v210_read(src, w, h, cm_r, lut_r, gam, opts)
        report(f"v210 read 709->2020 {kind} (oracle vs reference)", r, o)
        ow = oracle.v210_write(o, w, h, 0, cm_w, lut_w)
        rw = ref_ocl.v210_write(o, w, h, 0, cm_w, lut_w, options=opts)
        report(f"v210 write 2020 {kind}, same RGBA in", rw, ow)
        rw2 = ref_ocl.v210_write(r, w, h, 0, cm_w, lut_w, options=opts)
        report(f"v210 read->write chain {kind} (each its own)", rw2, ow)
    # 709 -> 709 round trip of the reference's own fixture through the reference's own kernels
    src = make_frame("ramp", w, h, 0)
    g709 = oracle.rgb2rgb_matrix("709", "709")
    r = ref_ocl.v210_read(src, w, h, cm_r, lut_r, g709, opts)
    back = ref_ocl.v210_write(r, w, h, 0, oracle.rgb2ycbcr_matrix("709"), oracle.linear2gamma_lut("709"), options=opts)
    report("reference round trip of fillBuf (compare()==0)", src, back)
    # interlaced write: two fields into one buffer
    dst = np.zeros_like(src)
    ref_ocl.v210_write(r, w, h, 1, oracle.rgb2ycbcr_matrix("709"), oracle.linear2gamma_lut("709"), out=dst, options=opts)
    ref_ocl.v210_write(r, w, h, 3, oracle.rgb2ycbcr_matrix("709"), oracle.linear2gamma_lut("709"), out=dst, options=opts)
    report("reference two-field write of fillBuf", src, dst)
    # image ops on small frames
    W, H = 480, 270
    rng = np.random.default_rng(7)
    ims = [rng.random((H, W, 4), dtype=np.float32) for _ in range(4)]
    report("combine_4", ref_ocl.combine(ims), oracle.combine(ims))
    report("combine_2", ref_ocl.combine(ims[:2]), oracle.combine(ims[:2]))
    report("transition dissolve mix=0.37", ref_ocl.dissolve(ims[0], ims[1], 0.37), oracle.dissolve(ims[0], ims[1], 0.37))
    report("transition wipe", ref_ocl.wipe_mask(ims[0], ims[1], ims[2]), oracle.wipe_mask(ims[0], ims[1], ims[2]))
    for name, xf in (("identity", dict(IDENTITY_XF)), ("pip 0.5", pip(0.5, 0.25, 0.45)), ("scale 0.731/0.577", dict(IDENTITY_XF, scaleX=0.731, scaleY=0.577, offsetX=0.21)),
                     ("rotate 0.04", dict(pip(0.6, 0.2, 0.1), rotate=0.04))):
        m = xf_matrix(W, H, xf)
        a, b = ref_ocl.transform(ims[0], m, W, H), oracle.transform(ims[0], m, W, H)
        report(f"transform {name}", a, b)
        print(f"{'':46s} max abs diff {float(np.abs(a - b).max()):.3e}")
    return 0

sys.exit(main())
