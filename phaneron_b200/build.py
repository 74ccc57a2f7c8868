"""In-tree build of libphaneron_b200.so (CUDA kernels + C ABI) for sm_100a.

nvcc cross-compiles without a GPU; the .so lands next to this file so that it
travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("PB_LIB_OUT") or os.path.join(HERE, "libphaneron_b200.so")   # PB_LIB_OUT: kernel-variant experiments (load with PB_LIB)
SOURCES = ["pb_kernels.cu", "pb_fused.cu", "pb_march.cu", "pb_march_general.cu", "pb_march_bigrows.cu", "pb_recorder.cu", "pb_lut_cache.cu", "pb_march_prep.cu", "pb_abi.cu", "pb_route.cu", "pb_colour.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",            # every fused op in the kernels is an explicit __fmaf_rn
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "-Xptxas", "-v",
]
for _k in ("PB_MARCH_WARPS", "PB_GENERAL_WARPS", "PB_MARCH_ROUNDS", "PB_DIRECT_WARPS", "PB_EXP", "PB_EXP_ROW_PREFETCH"):   # kernel-variant experiments
    if os.environ.get(_k):
        NVCC_FLAGS += [f"-D{_k}={int(os.environ[_k])}"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HERE, "..", "include", "phaneron_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    objs = []
    bdir = os.environ.get("PB_BUILD_DIR") or os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    nvcc = _nvcc()
    log = []
    # the translation units compile side by side (pb_march.cu, with its kernel variants, is the long pole)
    from concurrent.futures import ThreadPoolExecutor

    def compile_one(src):
        obj = os.path.join(bdir, src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        return src, obj, subprocess.run(cmd, capture_output=True, text=True)
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    for src, obj, res in results:
        log.append(res.stderr)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-ldl"]  # static cudart: self-contained .so (libnccl is dlopen'ed by pb_route.cu)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(bdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
