#!/usr/bin/env python
"""TEST INFRASTRUCTURE: pull the reference's OpenCL C kernel strings out of its TypeScript sources.

Streampunk/phaneron keeps every kernel as a template literal inside src/process/*.ts.
This script copies those literals VERBATIM from /root/reference into oracle/_ref/kernels/*.cl, which is
git-ignored (reference source never enters the history) but travels to the GPU box, where
oracle/ref_ocl/ocl_runner.c hands them to the NVIDIA OpenCL driver.  The two generated kernels
(combine_N, transition_*) are assembled from the literal fragments of their generator functions
(combine.ts:24-68, transition.ts:24-81) by interpreting exactly the control flow those functions have:
`let kernel = ...`, `kernel += ...`, `for (let i = 2; i < numLayers; ++i)`, `if (type === 'dissolve') else`.

  python oracle/ref_ocl/extract_kernels.py [/root/reference]       (run by __graft_entry__.build() when the reference is present)
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref", "kernels")


def literal_after(text, marker):
    """the template literal that starts at the first backtick after `marker`"""
    a = text.index("`", text.index(marker)) + 1
    b = text.index("`", a)
    return text[a:b]


def subst(chunk, env):
    def rep(m):
        return str(eval(m.group(1), {}, env))   # expressions are `numLayers`, `i - 1`, `type`, ...
    return re.sub(r"\$\{([^}]*)\}", rep, chunk)


def fragments(text, fn_name):
    """[(kind, literal)] in source order for `const fn_name = (...) => { ... return kernel }`"""
    a = text.index(f"const {fn_name} =")
    b = text.index("return kernel", a)
    body = text[a:b]
    out = []
    pos = 0
    while True:
        m = re.search(r"(let kernel =|kernel \+=)\s*`", body[pos:])
        if not m:
            break
        start = pos + m.end()
        end = body.index("`", start)
        # the innermost control statement that encloses this fragment
        pre = body[:pos + m.start()]
        depth_for = pre.count("for (let i = 2; i < numLayers; ++i) {") - 0
        ctx = "plain"
        last_open = max(pre.rfind("for (let i = 2; i < numLayers; ++i) {"), pre.rfind("if (type === 'dissolve') {"), pre.rfind("} else {"))
        last_close = pre.rfind("\n\t}\n")
        if last_open > last_close:
            seg = pre[last_open:]
            ctx = "for" if seg.startswith("for") else ("if" if seg.startswith("if") else "else")
        out.append((ctx, body[start:end]))
        pos = end + 1
    return out


def gen_combine(text, n):
    k = ""
    for ctx, lit in fragments(text, "getCombineKernel"):
        if ctx == "for":
            for i in range(2, n):
                k += subst(lit, {"numLayers": n, "i": i})
        else:
            k += subst(lit, {"numLayers": n})
    return k


def gen_transition(text, typ):
    k = ""
    for ctx, lit in fragments(text, "getTransitionKernel"):
        if ctx == "plain" or (ctx == "if" and typ == "dissolve") or (ctx == "else" and typ != "dissolve"):
            k += subst(lit, {"type": typ})
    return k


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    src = os.path.join(ref, "src", "process")
    if not os.path.isdir(src):
        print(f"extract_kernels: {src} not found (nothing extracted)")
        return 1
    os.makedirs(OUT, exist_ok=True)
    rd = lambda f: open(os.path.join(src, f)).read()
    files = {
        "v210.cl": literal_after(rd("v210.ts"), "const v210Kernel ="),
        "transform.cl": literal_after(rd("transform.ts"), "const transformKernel ="),
        "yadif.cl": literal_after(rd("yadifCl.ts"), "const yadifKernel ="),
        "rgba8.cl": literal_after(rd("rgba8.ts"), "const rgba8Kernel ="),
        "yuv422p10.cl": literal_after(rd("yuv422p10.ts"), "const yuv422p10leKernel ="),
        "yuv422p8.cl": literal_after(rd("yuv422p8.ts"), "const yuv422p8Kernel ="),
        "yuv420p.cl": literal_after(rd("yuv420p.ts"), "const yuv420pKernel ="),
        "nv12.cl": literal_after(rd("nv12.ts"), "const nv12Kernel ="),
        "mix.cl": literal_after(rd("mix.ts"), "const mixKernel ="),
        "wipe.cl": literal_after(rd("wipe.ts"), "const wipeKernel ="),
        "resize.cl": literal_after(rd("resize.ts"), "const resizeKernel ="),
        "transition_dissolve.cl": gen_transition(rd("transition.ts"), "dissolve"),
        "transition_wipe.cl": gen_transition(rd("transition.ts"), "wipe"),
    }
    for n in range(2, 9):
        files[f"combine_{n}.cl"] = gen_combine(rd("combine.ts"), n)
    for name, text in files.items():
        with open(os.path.join(OUT, name), "w") as f:
            f.write(text)
    print(f"extract_kernels: wrote {len(files)} kernel sources to {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
