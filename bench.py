#!/usr/bin/env python
"""bench.py -- headline benchmark of the phaneron_b200 hot path (BASELINE.json):

  metric  : 2160p50 v210 4-layer composite frames/sec (+ achieved HBM GB/s vs roofline)
  workload: BASELINE.json configs[2]: 3840x2160 v210, 4 layers (L1 identity full frame,
            L2-L4 MIXER FILL 0.5 picture-in-picture, top layer in a dissolve at mix 0.5
            with a 5th source), BT.709 sources -> BT.2020 working space -> v210 BT.2020 out

A "step" is a batch of FRAMES_PER_STEP frames.  `value` times the fused launches with
inputs resident in HBM (recorded once through the public operator API, then replayed),
rotating over enough input sets to exceed L2.  `e2e` runs every frame through the public
API (ToRGBA.loadFrame H2D from pinned host memory -> operators -> FromRGBA.saveFrame D2H).
One process per GPU; N>1 = N independent channels (weak scaling, no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--inputs ramp|noise] [--impl reference]
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, LAYERS, VARIANT = 3840, 2160, 4, "mix"
COL_READ, COL_WORK = "709", "2020"
FRAMES_PER_STEP = 240          # device-resident leg
E2E_WARM = 8                    # pipelined frames before the e2e clock starts
E2E_FRAMES_PER_STEP = 48       # public-API leg (PCIe bound: ~133 MB H2D + 22 MB D2H per frame)
L2_BYTES = 126 * 1024 * 1024
METRIC = "2160p50 v210 4-layer composite frames/sec"
LANCZOS = 0                    # > 0: the upper layers' Transforms use an N-lobe Lanczos filter (config 5)


def select_config(name: str) -> None:
    """--config: '3' = BASELINE.json configs[2] (the headline metric, default); '5' = configs[4]: 4320p, 2 layers, the upper
    one a half-size PiP resized with a Lanczos-3 filter (an extension: the reference has only the bilinear sampler)"""
    global WIDTH, HEIGHT, LAYERS, VARIANT, METRIC, LANCZOS, FRAMES_PER_STEP, CPU_BASELINE_FRAMES
    if name == "5":
        WIDTH, HEIGHT, LAYERS, VARIANT, LANCZOS = 7680, 4320, 2, "plain", 3
        METRIC = "4320p50 v210 2-layer composite with Lanczos-3 resize frames/sec"
        FRAMES_PER_STEP, CPU_BASELINE_FRAMES = 60, 4


def bench_scene(inputs: str, frame_set: int = 0):
    from phaneron_b200.scenes import layered_scene
    scene = layered_scene(WIDTH, HEIGHT, LAYERS, inputs, VARIANT, COL_READ, COL_WORK, frame_set=frame_set)
    if LANCZOS:
        for L in scene["layers"][1:]:
            L["xf"] = dict(L["xf"], filter=f"lanczos{LANCZOS}")
    return scene
CPU_BASELINE_FRAMES = 12       # cpu_baseline of the default run: whole 3840x2160 frames (about 1 s each on 16 cores)


_JSON_OUT = None   # the real stdout when fd 1 has been pointed at stderr (multi-rank runs)


def emit(line: dict) -> None:
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.stop_flag, self.proc = gpu_index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append((time.perf_counter(), [x.strip() for x in line.split(",")]))
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self, t0, t1):
        self.stop_flag = True
        time.sleep(0.15)
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.samples if t0 <= t <= t1] or [r for _, r in self.samples[-3:]]
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": len(rows)}
        try:
            sm = sorted(float(r[0]) for r in rows)
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = max(float(r[1]) for r in rows)
            out["power_w_max"] = max(float(r[2]) for r in rows)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        except Exception:
            pass
        return out


def pinned_copy(lib, arr: np.ndarray) -> np.ndarray:
    import ctypes as C
    p = lib.pb_host_alloc(arr.nbytes)
    if not p:
        raise RuntimeError("pb_host_alloc failed")
    out = np.ctypeslib.as_array((C.c_uint8 * arr.nbytes).from_address(p))
    out[:] = arr.view(np.uint8).reshape(-1)
    return out


def pin_scene(lib, scene):
    for L in scene["layers"]:
        L["src"] = pinned_copy(lib, L["src"])
        t = L.get("transition")
        if t:
            t["src"] = pinned_copy(lib, t["src"])
            if "mask" in t:
                t["mask"] = pinned_copy(lib, t["mask"])
    return scene


# ------------------------------------------------------------------------------------------
def cpu_reference_fps(steps, warmup, inputs, threads):
    """the reference's unfused stage sequence as restated in oracle/ on the host cores: WHOLE frames of the bench scene
    (5x v210 read, 5x transform, dissolve, combine_4, v210 write with RGBA-f32 intermediates in host memory), one per step,
    compiled -O3 -march=native on this machine"""
    import oracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from scene_oracle import SceneOracle
    from phaneron_b200.scenes import layered_scene
    native = oracle.use_native()
    oracle.set_threads(threads)
    scene = bench_scene(inputs)
    so = SceneOracle(scene)
    for _ in range(warmup):
        so.packed()
    t0 = time.perf_counter()
    for _ in range(steps):
        so.packed()
    dt = time.perf_counter() - t0
    return steps / dt, dt, native


def yadif_outputs_match_oracle(window, pip_frame, pip_xf, main_xf, w, h, produced):
    """--config yadif, untimed: the two output frames of one input frame against the oracle's unfused chain (v210 read x3 + PiP
    source, yadif for both fields, transform, combine_2, v210 write), byte for byte"""
    import oracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from scene_oracle import xf_matrix
    oracle.use_native()
    oracle.set_threads(os.cpu_count() or 1)
    cm_r, lut_r, gam = oracle.ycbcr2rgb_matrix("709"), oracle.gamma2linear_lut("709"), oracle.rgb2rgb_matrix("709", "2020")
    cm_w, lut_w = oracle.rgb2ycbcr_matrix("2020"), oracle.linear2gamma_lut("2020")
    win = [oracle.v210_read(f, w, h, cm_r, lut_r, gam) for f in window]
    lb = oracle.transform(oracle.v210_read(pip_frame, w, h, cm_r, lut_r, gam), xf_matrix(w, h, pip_xf), w, h)
    ok = len(produced) == 2
    for k, second in enumerate((False, True)):
        parity = 1 ^ (0 if second else 1)   # tff: (tff ? 1 : 0) ^ (!isSecond ? 1 : 0), yadif.ts:104
        deint = oracle.yadif(win[0], win[1], win[2], parity, True, False)
        la = oracle.transform(deint, xf_matrix(w, h, main_xf), w, h)
        ref = oracle.v210_write(oracle.combine([la, lb]), w, h, 0, cm_w, lut_w)
        ok = ok and np.array_equal(produced[k], ref)
    return ok


def reference_kernels_on_gpu(inputs, frames=6):
    """Baseline A: the reference's OWN OpenCL kernels (extracted from its .ts sources into the
    git-ignored oracle/_ref/) launched in the reference's unfused sequence on the same B200 through NVIDIA's
    OpenCL driver, buffers resident.  None when the driver or the extracted kernels are absent."""
    if LANCZOS:
        return {"unavailable": "the reference has no Lanczos filter (transform.ts:26-29 samples with CLK_FILTER_LINEAR)"}
    try:
        from oracle import ref_ocl
        if not ref_ocl.available():
            return {"unavailable": ref_ocl.why_unavailable()}
        import oracle
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from scene_oracle import xf_matrix
        from phaneron_b200.scenes import layered_scene
        scene = bench_scene(inputs)
        consts = (oracle.ycbcr2rgb_matrix(COL_READ), oracle.gamma2linear_lut(COL_READ), oracle.rgb2rgb_matrix(COL_READ, COL_WORK),
                  oracle.rgb2ycbcr_matrix(COL_WORK), oracle.linear2gamma_lut(COL_WORK))
        chain = ref_ocl.ReferenceChain(scene, consts, xf_matrix)
        chain.run_frames(2)
        dt = chain.run_frames(frames)
        return {"value": frames / dt, "unit": "frames/s", "device": ref_ocl.device_name() + " (NVIDIA OpenCL)", "launches_per_frame": len(chain.launches),
                "sample": f"{frames} frames of the same scene: 5x read, 5x transform, transition_dissolve, combine_4, write, RGBA-f32 intermediates resident, one clFinish"}
    except Exception as e:   # a baseline must never take the bench down
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    fps, dt, native = cpu_reference_fps(args.steps, max(args.warmup, 3), args.inputs, threads)   # (>= 3: the first frames page in 1.5 GB of intermediates)
    sample = (f"{args.steps} steps, each ONE whole {WIDTH}x{HEIGHT} frame through the full unfused chain (a v210 read + transform per source, "
              f"{'dissolve, ' if VARIANT == 'mix' else ''}combine_{LAYERS}, v210 write, RGBA-f32 intermediates), oracle/ built {'-O3 -march=native on this host' if native else '-O3'}, {dt:.1f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.inputs), "host": "CPU restatement of the reference's OpenCL kernels (oracle/), not POCL"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def ncu_traffic(kernel, inputs):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the NEWEST committed
    `ncu --set full` capture of this same scene (profiles/rNN_march_ncu_summary.json); None if not captured"""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_march_ncu_summary.json")), reverse=True):
        try:
            with open(path) as f:
                v = json.load(f).get(f"{kernel}_{inputs}" + (f"_lanczos{LANCZOS}" if LANCZOS else ""), {}).get("dram_bytes_per_launch")
            if v:
                return v
        except Exception:
            pass
    return None


def workload_name(inputs):
    if LANCZOS:
        return (f"{WIDTH}x{HEIGHT} v210, {LAYERS}-layer composite (L1 identity, L2 MIXER FILL 0.5 PiP resized with a Lanczos-{LANCZOS} filter), "
                f"{COL_READ}->{COL_WORK}, inputs={inputs}")
    return (f"{WIDTH}x{HEIGHT} v210, {LAYERS}-layer composite (L1 identity, L2-L4 MIXER FILL 0.5 PiP, top layer dissolve mix=0.5 "
            f"with a 5th source), {COL_READ}->{COL_WORK}, inputs={inputs}")


# ------------------------------------------------------------------------------------------
async def run_ours(args, rank, world, local_rank):
    from phaneron_b200 import _lib, clContext
    from phaneron_b200.harness import ChannelHarness
    from phaneron_b200.scenes import layered_scene

    # one process per GPU, on the GPU's own CPU cores / NUMA node (pinned frame buffers are first-touched after this)
    from phaneron_b200.affinity import bind_to_gpu
    cpus = None if os.environ.get("PB_NO_AFFINITY") else bind_to_gpu(local_rank, world, local_rank)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local_rank)
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_

    def barrier():
        if dist:
            dist.barrier()

    lib = _lib.lib()
    ctx = clContext({"platformIndex": 0, "deviceIndex": local_rank, "overlapping": True,
                     "marchKernel": args.kernel != "generic", "rawLut": args.kernel == "march_raw",
                     "occlusionCulling": not args.no_culling, "footprint": True})
    await ctx.initialise()

    # ---- scenes: enough distinct input sets that a replay never finds its inputs in L2 ----
    frame_bytes = (WIDTH // 48) * 128 * HEIGHT
    n_in = LAYERS + (1 if VARIANT == "mix" else 2 if VARIANT == "wipe" else 0)
    set_bytes = (n_in + 1) * frame_bytes
    n_sets = max(3, -(-2 * L2_BYTES // set_bytes) + 1)
    harnesses, chains, keep, chain_dests = [], [], [], []
    chains_nocull = []   # the same frames with occlusion culling off (reported beside the default)
    for s in range(n_sets):
        scene = bench_scene(args.inputs, s + rank * n_sets)
        h = ChannelHarness(ctx, scene, chanID=f"ch{rank}s{s}")
        await h.init()
        chain, dests = await h.record_chain()
        if not chain.complete:
            raise RuntimeError("recorded chain is not replayable")
        harnesses.append(h)
        chains.append(chain)
        keep.append(dests)
        chain_dests.append(dests)
        if not args.no_culling and args.kernel != "generic":
            ctx.setOcclusionCulling(False)
            chain2, dests2 = await h.record_chain()
            ctx.setOcclusionCulling(True)
            chains_nocull.append(chain2)
            keep.append(dests2)
    st_k = ctx.stats()
    src_bytes_read = st_k["march_src_bytes"]   # distinct packed source bytes of the last march launch recorded
    if chains_nocull:   # the last launch recorded was a no-culling one: account a culled launch again
        chain3, dests3 = await harnesses[-1].record_chain()
        src_bytes_read = ctx.stats()["march_src_bytes"]
        keep.append(dests3)
    ctx.footprint = False   # accounting costs a host pass per launch: not inside any timed region
    ctx.setOcclusionCulling(not args.no_culling)
    kernel_name = ("k_fused_march (pb_march.cu), gamma tables as 1-byte deltas in shared memory" if st_k["march_launches"] and args.kernel == "march"
                   else "k_fused_march, raw gamma tables from global memory" if st_k["march_launches"] else "k_fused_generic (pb_fused.cu)")
    alg_bytes = harnesses[0].algorithmic_bytes()
    launches_per_frame = chains[0].launches
    await ctx.waitFinish(ctx.queue.process)

    # ---- parity, untimed: the frames the timed region replays are the oracle's frames, byte for byte (rank 0) ----
    async def replayed_frame(s):
        chains[s].replay()
        await ctx.waitFinish(ctx.queue.process)
        await harnesses[s].fromRGBA.saveFrame(chain_dests[s], ctx.queue.unload)
        await ctx.waitFinish(ctx.queue.unload)
        return chain_dests[s][0].host

    parity = {"checked": False}
    ref0 = None
    if rank == 0 and not args.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from scene_oracle import SceneOracle
        ref0 = SceneOracle(harnesses[0].scene).packed()
        got = await replayed_frame(0)
        if not np.array_equal(got, ref0):
            raise RuntimeError(f"bench: the replayed frame differs from the oracle in {int((got != ref0).sum())} bytes")
        parity = {"checked": True, "what": "replayed frame of input set 0 == oracle (unfused CPU chain), byte for byte, before and after the timed region"}

    fps_n = args.frames_per_step

    def replay_step(step_index, which=None):
        base = step_index * fps_n
        cs = which or chains
        for f in range(fps_n):
            cs[(base + f) % n_sets].replay()

    # ---- the same frames without occlusion culling (untimed by the contract; reported in config) ----
    nocull_fps = None
    if chains_nocull:
        for w in range(args.warmup):
            replay_step(w, chains_nocull)
        await ctx.waitFinish(ctx.queue.process)
        e0, e1 = ctx.createEvent(), ctx.createEvent()
        e0.record()
        for k in range(args.steps):
            replay_step(k, chains_nocull)
        e1.record()
        e1.synchronize()
        nocull_fps = args.steps * fps_n / (e0.elapsed_ms(e1) * 1e-3)

    # ---- device-resident leg ----------------------------------------------------------------
    for w in range(args.warmup):
        replay_step(w)
    await ctx.waitFinish(ctx.queue.process)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    ev0, ev1 = ctx.createEvent(), ctx.createEvent()
    st0 = ctx.stats()
    barrier()
    await ctx.waitFinish(ctx.queue.process)
    t0 = time.perf_counter()
    ev0.record()
    for k in range(args.steps):
        replay_step(k)
    ev1.record()
    ev1.synchronize()
    t1 = time.perf_counter()
    barrier()
    st1 = ctx.stats()
    ms = ev0.elapsed_ms(ev1)
    clocks = sampler.finish(t0, t1)
    if dist:
        import torch
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    frames = args.steps * fps_n
    fps_rank = frames / (ms * 1e-3)
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    if ref0 is not None:   # what the timed replays left in set 0's destination (then once more, freshly replayed)
        await harnesses[0].fromRGBA.saveFrame(chain_dests[0], ctx.queue.unload)
        await ctx.waitFinish(ctx.queue.unload)
        if not np.array_equal(chain_dests[0][0].host, ref0) or not np.array_equal(await replayed_frame(0), ref0):
            raise RuntimeError("bench: a frame written inside the timed region differs from the oracle")

    # ---- end-to-end leg through the public API ------------------------------------------------
    e2e_scene = pin_scene(lib, bench_scene(args.inputs, rank * n_sets))
    he = ChannelHarness(ctx, e2e_scene, chanID=f"e2e{rank}")
    await he.init()
    for _ in range(3):
        await he.run_frame()
    e2e_steps = max(1, min(args.steps, 5))
    barrier()
    await ctx.waitFinish(ctx.queue.process)
    e2e_first = e2e_last = None
    # Frames are pipelined the way phaneron's redioactive pipes run them: while frame i is composed, packed and
    # read back, the sources of the next frames are already being copied in (every call is async work, see nodencl.py;
    # PB_E2E_DEPTH frame times of uploads in flight keep the H2D copy engine busy across the host-side hand-over).
    # The pipeline first runs E2E_WARM frames untimed (buffer pools reach their steady-state depth), then the timed frames.
    depth = int(os.environ.get('PB_E2E_DEPTH', '3'))
    n_e2e = e2e_steps * E2E_FRAMES_PER_STEP
    n_all = E2E_WARM + n_e2e
    pending = [asyncio.ensure_future(he.upload_all(1000 + j)) for j in range(min(depth, n_all))]   # H2D of every source frame, from pinned host memory
    s0, te0 = None, None
    for i in range(n_all):
        if i == E2E_WARM:
            barrier()
            s0 = ctx.stats()
            te0 = time.perf_counter()
        ups = await pending.pop(0)
        if i + depth < n_all:
            pending.append(asyncio.ensure_future(he.upload_all(1000 + i + depth)))
        frame = await he.compose(ups, 1000 + i)              # operators + job queue (records the expression)
        dests = await he.consume(frame, download=True)       # fused launch + D2H of the packed result
        if i == n_all - 1:
            te1 = time.perf_counter()
            e2e_last = dests[0].host.copy()                  # (after the clock: the last timed frame)
        elif i == E2E_WARM - 1:
            e2e_first = dests[0].host.copy()                 # (before the clock: the last warm-up frame)
        for d in dests:
            d.release()
    s1 = ctx.stats()
    e2e_dt = te1 - te0
    e2e_parity = False
    if ref0 is not None:   # rank 0's e2e scene is input set 0: the downloaded frames must be the oracle's frame
        for nm, fr in (("first", e2e_first), ("last", e2e_last)):
            if fr is None or not np.array_equal(fr, ref0):
                raise RuntimeError(f"bench: the {nm} frame downloaded by the end-to-end leg differs from the oracle")
        e2e_parity = True

    # ---- host cost of the public-API path with device-resident inputs (no H2D, no D2H): what pb_chain_replay skips ----
    host_frames = 40
    ups = await he.upload_all(5000)
    await ctx.waitFinish(ctx.queue.process)
    hs0 = ctx.stats()
    th0 = time.perf_counter()
    for i in range(host_frames):
        for per_layer in ups:   # the sources stay resident: one more reference per frame for the release callbacks
            for bufs in per_layer.values():
                for b in bufs:
                    b.addRef()
                    b.timestamp = 6000 + i
        frame = await he.compose(ups, 6000 + i)
        dests = await he.consume(frame, download=False)
        for d in dests:
            d.release()
    th1 = time.perf_counter()
    await ctx.waitFinish(ctx.queue.process)
    th2 = time.perf_counter()
    hs1 = ctx.stats()
    for per_layer in ups:
        for bufs in per_layer.values():
            for b in bufs:
                b.release()
    host_cost = {"frames": host_frames,
                 "api_us_per_frame": (th1 - th0) / host_frames * 1e6,          # Python mirror of the TS operators + C ABI, launches issued
                 "c_abi_us_per_frame": (hs1["run_program_ns"] - hs0["run_program_ns"]) / host_frames * 1e-3,   # inside pb_run_program only
                 "run_program_calls_per_frame": (hs1["run_program_calls"] - hs0["run_program_calls"]) / host_frames,
                 "wall_us_per_frame_incl_gpu": (th2 - th0) / host_frames * 1e6}
    if dist:
        import torch
        t = torch.tensor([e2e_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_fps = world * e2e_steps * E2E_FRAMES_PER_STEP / e2e_dt

    if rank == 0:
        peak, peak_kind = measured_peak()
        launch_ms = ms / (frames * launches_per_frame)
        # algorithmic bytes = (distinct packed inputs + 1 output) x frame bytes.  With occlusion culling the
        # kernel does not read source rows that lie under opaque layers: the roofline figure counts only the bytes the
        # launch has to move (what it reads after culling + the output frame), never more than the 8(d) figure.
        moved = alg_bytes // launches_per_frame
        if st_k["march_launches"] and src_bytes_read:
            moved = min(moved, int(src_bytes_read) + frame_bytes)
        achieved = moved / (launch_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": fps_rank * world, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args.inputs), "frames_per_step": fps_n, "kernel": kernel_name, "input_sets": n_sets,
                       "l2_policy": f"inputs larger than L2: {n_sets} rotating sets x {set_bytes} B = {n_sets * set_bytes} B > 126 MiB",
                       "channels": world, "parallelism": f"{world} independent channel(s), one per GPU",
                       "cpu_affinity_rank0": (f"{len(cpus)} CPUs local to the GPU" if cpus else "unchanged"),
                       "launches_per_frame": launches_per_frame,
                       "occlusion_culling": bool(st_k["march_launches"]) and not args.no_culling,
                       "occlusion_culling_note": "exact: ops under a layer whose alpha is 1.0f bit for bit over a whole strip line are skipped "
                                                 "(combine.ts multiplies them by 1 - 1 = 0); output bytes identical, tests/test_gpu_chain.py",
                       "frames_per_s_without_culling": nocull_fps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.kernel, args.inputs), "peak_kind": f"of {peak_kind}", "frac_of_nominal_8TBps": achieved / 8000.0,
                         "algorithmic_bytes_per_launch": alg_bytes // launches_per_frame, "bytes_moved_per_launch": moved,
                         "launch_us": launch_ms * 1e3, "launches_per_frame": launches_per_frame, "frame_us": launch_ms * 1e3 * launches_per_frame,
                         # SURVEY 8(d)'s own figure (every distinct packed input + the output, whether or not culling lets the launch skip
                         # rows hidden under opaque layers); `achieved` / `frac` above are the conservative ones, on the bytes actually moved
                         "achieved_on_algorithmic_bytes": (alg_bytes // launches_per_frame) / (launch_ms * 1e-3) / 1e9,
                         "frac_on_algorithmic_bytes": (alg_bytes // launches_per_frame) / (launch_ms * 1e-3) / 1e9 / peak},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": n_in * frame_bytes * E2E_FRAMES_PER_STEP,
                    "d2h_bytes_per_step": (s1["d2h_bytes"] - s0["d2h_bytes"]) // e2e_steps, "frames_per_step": E2E_FRAMES_PER_STEP,
                    "steps": e2e_steps, "parity_checked": e2e_parity},
            "parity_checked": parity["checked"], "parity": parity,
            "host_us_per_frame": host_cost,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n = CPU_BASELINE_FRAMES   # whole frames: ~10-20 core-seconds of CPU work
            fps, dt, native = cpu_reference_fps(n, 1, args.inputs, threads)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                                    "sample": f"{n} whole {WIDTH}x{HEIGHT} frames of the same scene, oracle/ unfused chain "
                                              f"({'-O3 -march=native' if native else '-O3'}), {dt:.1f} s",
                                    "reference_kernels_on_gpu": reference_kernels_on_gpu(args.inputs)}
        emit(line)
    barrier()
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--inputs", default="noise", choices=["ramp", "noise"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-culling", action="store_true", help="evaluate layers hidden under opaque ones too (A/B)")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed oracle comparison of the replayed and downloaded frames")
    ap.add_argument("--frames-per-step", type=int, default=FRAMES_PER_STEP, help="frames per device-resident step (profiling runs use a few)")
    ap.add_argument("--config", default="3", choices=["3", "route", "5", "yadif"],
                    help="3: BASELINE.json configs[2], the 2160p 4-layer composite (default, the headline metric); route: configs[3], 1080p "
                         "channels one per GPU with ROUTE cross-feed over NCCL; 5: configs[4], 4320p 2-layer composite with a Lanczos-3 PiP; "
                         "yadif: the reference's own operating point, a 1080i50 source de-interlaced and composited")
    ap.add_argument("--route-frames-per-step", type=int, default=200)
    ap.add_argument("--yadif-inputs-per-step", type=int, default=100)
    ap.add_argument("--yadif-size", default="1920x1080")
    ap.add_argument("--kernel", default="march", choices=["march", "march_raw", "generic"],
                    help="march: fused kernel, gamma tables in shared memory (default); march_raw: same kernel gathering "
                         "from the raw tables; generic: the fallback fused kernel")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun like the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if world > 1 or "route" in sys.argv or "yadif" in sys.argv:
        # libraries (NCCL's version banner, torchrun notices) write to fd 1: everything but the JSON line goes to stderr
        global _JSON_OUT
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    select_config(args.config)
    if args.impl == "reference":
        if args.config in ("route", "yadif"):   # (the CPU arm exists for the composite configurations: 3 and 5)
            if rank == 0:
                emit({"impl": "reference", "unavailable": f"no CPU arm for --config {args.config}: run it for the default config or --config 5"})
            return
        run_reference(args, rank, world)
        return
    if args.frames_per_step == 240:
        args.frames_per_step = FRAMES_PER_STEP
    if args.config == "route":
        from phaneron_b200 import bench_route
        asyncio.run(bench_route.run(args, rank, world, local_rank, emit, ClockSampler, measured_peak))
        return
    if args.config == "yadif":
        from phaneron_b200 import bench_yadif
        yw, yh = (int(v) for v in args.yadif_size.split("x"))
        asyncio.run(bench_yadif.run(args, rank, world, local_rank, emit, ClockSampler, measured_peak, yadif_outputs_match_oracle, yw, yh))
        return
    asyncio.run(run_ours(args, rank, world, local_rank))


if __name__ == "__main__":
    main()
