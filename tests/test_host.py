"""Host-side logic and the C-ABI surface, no GPU: the library loads and exports every
symbol include/phaneron_b200.h declares; the product's colour maths (pb_colour.cpp)
equals the oracle bit for bit; fixtures; job-queue semantics; error behaviour."""
import asyncio
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
from phaneron_b200 import PhaneronError, _lib, clContext
from phaneron_b200.cl_job_queue import ClProcessJobs
from phaneron_b200.process import colour_maths as cm
from phaneron_b200.process import v210
from phaneron_b200.process.packer import Interlace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPECS = ["601-625", "601_525", "709", "2020", "sRGB"]


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "phaneron_b200.h")).read()
    declared = set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), s


def test_no_cpu_fallback_without_a_device():
    async def go():
        ctx = clContext({"platformIndex": 0, "deviceIndex": 0})
        await ctx.initialise()
        return ctx
    try:
        ctx = asyncio.run(go())
    except PhaneronError as e:
        assert "no CPU fallback" in str(e)
    else:   # on a GPU box the context simply works
        assert ctx.getPlatformInfo()["devices"][0]["type"] == "GPU"
        ctx.close()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "phaneron_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, re.M), f
                assert "oracle.h" not in txt and "liboracle" not in txt, f


@pytest.mark.parametrize("spec", SPECS)
def test_colour_maths_matches_oracle(spec):
    np.testing.assert_array_equal(cm.gamma2linearLUT(spec), oracle.gamma2linear_lut(spec))
    np.testing.assert_array_equal(cm.linear2gammaLUT(spec), oracle.linear2gamma_lut(spec))
    for bits, lb, lw, cr in ((10, 64, 940, 896), (8, 16, 235, 224)):
        np.testing.assert_array_equal(cm.ycbcr2rgbMatrix(spec, bits, lb, lw, cr), oracle.ycbcr2rgb_matrix(spec, bits, lb, lw, cr))
        np.testing.assert_array_equal(cm.rgb2ycbcrMatrix(spec, bits, lb, lw, cr), oracle.rgb2ycbcr_matrix(spec, bits, lb, lw, cr))
    for dst in SPECS:
        np.testing.assert_array_equal(cm.rgb2rgbMatrix(spec, dst), oracle.rgb2rgb_matrix(spec, dst))


def test_transform_matrix_matches_oracle():
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = rng.uniform(-1, 1, 7)
        args = (1920, 1080, bool(rng.integers(2)), bool(rng.integers(2)), a[0], a[1], 0.1 + abs(a[2]) * 2, 0.1 + abs(a[3]) * 2,
                a[4], a[5], a[6])
        np.testing.assert_array_equal(cm.transformMatrix(*args), oracle.transform_matrix(*args))
    # `(params.scaleX as number) || 1.0`
    np.testing.assert_array_equal(cm.transformMatrix(16, 9, False, False, 0, 0, 0.0, 0.0, 0, 0, 0), np.eye(3, dtype=np.float32))


@pytest.mark.parametrize("w,h", [(1920, 1080), (1280, 720), (720, 576), (52, 3), (50, 3)])
def test_fillbuf_mirror_matches_oracle(w, h):
    assert v210.getPitchBytes(w) == oracle.v210_pitch_bytes(w)
    buf = np.empty(v210.getPitchBytes(w) * h, np.uint8)
    v210.fillBuf(buf, w, h)
    np.testing.assert_array_equal(buf, oracle.v210_fill(w, h))


def test_scene_noise_frames_are_legal_v210():
    from phaneron_b200.scenes import noise_frame
    f = noise_frame(96, 4, 7).view(np.uint32)
    assert np.all(f >> 30 == 0)
    for sh in (0, 10, 20):
        c = (f >> sh) & 0x3FF
        assert c.min() >= 64 and c.max() <= 960


def test_reader_writer_geometry():
    r = v210.Reader(1920, 1080)
    assert r.getNumBytes() == [5529600] and r.getNumBytesRGBA() == 33177600
    assert r.getWorkItemsPerGroup() == 40 and r.getGlobalWorkItems() == 43200
    w = v210.Writer(1920, 1080, True)
    assert w.getGlobalWorkItems() == 21600
    assert w.getKernelParams({"source": 1, "dests": [2], "interlace": Interlace.BottomField})["interlace"] == 3
    # a non-interlaced Writer forces Progressive whatever the consumer passes (Q14, v210.ts:334)
    assert v210.Writer(1920, 1080, False).getKernelParams({"source": 1, "dests": [2], "interlace": 3})["interlace"] == 0
    with pytest.raises(RuntimeError):
        r.getKernelParams({"sources": [1, 2], "dest": 3})


class _FakeCtx:
    """records runProgram / waitFinish calls; enough of clContext for the queue logic"""

    class queue:
        load, process, unload = 0, 1, 2

    def __init__(self):
        self.log = []

    async def runProgram(self, program, params, queue, timed=False):
        self.log.append(("run", program, queue))
        if program == "boom":
            raise RuntimeError("kernel failed")
        from phaneron_b200.nodencl import RunTimings
        return RunTimings()

    async def waitFinish(self, queue):
        self.log.append(("wait", queue))


def test_job_queue_semantics():
    async def go():
        ctx = _FakeCtx()
        pj = ClProcessJobs(ctx)
        jobs = pj.getJobs()
        fired = []
        jobs.add({"source": "a", "timestamp": 1}, "read", "p1", {}, lambda: fired.append("cb1"))
        jobs.add({"source": "a", "timestamp": 1}, "xf", "p2", {}, lambda: fired.append("cb2"))
        jobs.add({"source": "a", "timestamp": 2}, "read", "p3", {}, lambda: fired.append("cb3"))
        assert jobs.makeKey({"source": "a", "timestamp": 1}) == "a ts 1"
        assert len(jobs.get({"source": "a", "timestamp": 1})) == 2
        await jobs.runQueue({"source": "a", "timestamp": 1})
        # both jobs launched in order on queue.process, ONE wait per request, callbacks after the wait
        assert ctx.log == [("run", "p1", 1), ("run", "p2", 1), ("wait", 1)]
        assert fired == ["cb1", "cb2"]
        assert jobs.get({"source": "a", "timestamp": 1}) is None
        with pytest.raises(RuntimeError, match="Failed to run queue for id a ts 1"):
            await jobs.runQueue({"source": "a", "timestamp": 1})
        # clearQueue fires pending callbacks without running anything (clJobQueue.ts:87-94)
        jobs.clearQueue("a")
        assert fired == ["cb1", "cb2", "cb3"]
        assert len(ctx.log) == 3
        # a failing kernel rejects the runQueue awaitable and the loop keeps serving later requests
        jobs.add({"source": "b", "timestamp": 0}, "bad", "boom", {}, lambda: fired.append("released-b"))
        with pytest.raises(RuntimeError, match="kernel failed"):
            await jobs.runQueue({"source": "b", "timestamp": 0})
        jobs.add({"source": "c", "timestamp": 0}, "ok", "p4", {}, lambda: fired.append("cb4"))
        await jobs.runQueue({"source": "c", "timestamp": 0})
        # the failed request's release callbacks ran too (its buffers must not leak; the reference would wedge here)
        assert fired[-2:] == ["released-b", "cb4"]

    asyncio.run(go())


def test_requests_are_fifo_across_sources():
    async def go():
        ctx = _FakeCtx()
        jobs = ClProcessJobs(ctx).getJobs()
        for s in ("x", "y", "z"):
            jobs.add({"source": s, "timestamp": 0}, "k", f"prog-{s}", {}, lambda: None)
        await asyncio.gather(*(jobs.runQueue({"source": s, "timestamp": 0}) for s in ("x", "y", "z")))
        assert [e[1] for e in ctx.log if e[0] == "run"] == ["prog-x", "prog-y", "prog-z"]

    asyncio.run(go())


# ---- the N-API addon and the TypeScript-face edits ----------------------------------------------------
def test_napi_shim_compiles_against_the_stub_and_binds_only_declared_entry_points():
    """napi/phaneron_napi.cc: node-addon-api is absent from the image, so the shim is type-checked against napi/stub/napi.h;
    every pb_* function it calls is declared in include/phaneron_b200.h and exported by the shared library"""
    import re
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gxx = shutil.which("g++")
    assert gxx, "g++ is part of the image"
    res = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Inapi/stub", "-Iinclude", "napi/phaneron_napi.cc"],
                         cwd=root, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    src = open(os.path.join(root, "napi", "phaneron_napi.cc")).read()
    hdr = open(os.path.join(root, "include", "phaneron_b200.h")).read()
    called = set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", src))
    declared = set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", hdr))
    assert called and called <= declared, sorted(called - declared)
    from phaneron_b200 import _lib
    l = _lib.lib()
    assert all(hasattr(l, s) for s in called)
    # the nodencl surface phaneron's sources use is all there
    for method in ("initialise", "getPlatformInfo", "createBuffer", "createProgram", "runProgram", "waitFinish", "hostAccess", "addRef", "release"):
        assert f'"{method}"' in src, method
    assert "..." not in re.sub(r"//.*", "", src), "no elisions in the shim"


def test_typescript_patches_apply_to_the_reference():
    """ts/*.patch are real unified diffs against the reference's files (checked with patch --dry-run where the reference is present)"""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    patches = sorted(f for f in os.listdir(os.path.join(root, "ts")) if f.endswith(".patch"))
    assert {"packer.ts.patch", "index.ts.patch", "package.json.patch"} <= set(patches)
    if not os.path.isdir("/root/reference/src") or not shutil.which("patch"):
        pytest.skip("reference tree or patch(1) not present")
    for p in patches:
        res = subprocess.run(["patch", "--dry-run", "-p1", "-d", "/root/reference", "-i", os.path.join(root, "ts", p), "-o", "/dev/null"],
                             capture_output=True, text=True)
        assert res.returncode == 0, (p, res.stdout + res.stderr)
