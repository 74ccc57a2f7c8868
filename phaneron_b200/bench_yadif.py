"""bench.py --config yadif: the reference's own operating point (index.ts:45-71 configures 1080i50 channels): an interlaced v210
source de-interlaced to 50 full frames per second and composited.

Per input frame and rank:  ToRGBA(interlaced v210 frame)  ->  Yadif send_field (yadif.ts:88-145: a three-frame window, TWO output
frames per input frame)  ->  per output frame: Mixer Transform (identity) ; a second layer = ToRGBA + 0.5x PiP Transform of a
progressive v210 source ; Combine_2 ; FromRGBA v210.

What runs: one direct-kernel launch makes the new ToRGBA output real (RGBA-f32); per output frame a pre-pass
computes the interpolated lines of the field once and ONE march launch composites, reading the field's own lines from the current
frame in place.  Timed by replaying the recorded launches of one input frame (no Python in the loop).

`value` = output frames (fields) per second over all ranks.  Parity: the two output frames of the recorded input frame are
compared with the oracle's stage-by-stage chain (v210 read x3, yadif, transform, combine, v210 write) before the clock starts.
"""
from __future__ import annotations

import time

N_SOURCES = 4


async def run(args, rank: int, world: int, local_rank: int, emit, clock_sampler_cls, measured_peak, check_outputs=None, width: int = 1920, height: int = 1080):
    from . import ClProcessJobs, clContext, _lib
    from .process import v210
    from .process.combine import Combine
    from .process.image_process import ImageProcess
    from .process.io import FromRGBA, ToRGBA
    from .process.packer import Interlace
    from .process.transform import Transform
    from .process.yadif import Yadif
    from .scenes import IDENTITY_XF, make_frame, pip

    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist_.init_process_group("gloo")   # host-side only (barriers, max over ranks): the channels share nothing
        dist = dist_

    def barrier():
        if dist:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if not dist:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    w, h = width, height
    ctx = clContext({"deviceIndex": local_rank})
    await ctx.initialise()
    jobs = ClProcessJobs(ctx).getJobs()
    to_a = ToRGBA(ctx, "709", "2020", v210.Reader(w, h), jobs)
    to_b = ToRGBA(ctx, "709", "2020", v210.Reader(w, h), jobs)
    frm = FromRGBA(ctx, "2020", v210.Writer(w, h, False), jobs)
    xa = ImageProcess(ctx, Transform(ctx, w, h), jobs)
    xb = ImageProcess(ctx, Transform(ctx, w, h), jobs)
    comb = ImageProcess(ctx, Combine(2, w, h), jobs)
    yad = Yadif(ctx, jobs, w, h, {"mode": "send_field", "tff": True}, True)
    for o in (to_a, to_b, frm, xa, xb, comb, yad):
        await o.init()
    dims = {"width": w, "height": h}
    pip_xf = pip(0.5, 0.3, 0.2)
    pip_frame = make_frame(args.inputs, w, h, 9 + 100 * rank)
    pip_srcs = await to_b.createSources("pip")
    await to_b.loadFrame(pip_frame, pip_srcs, ctx.queue.load)
    frames = [make_frame(args.inputs, w, h, 20 + i + 100 * rank) for i in range(N_SOURCES)]
    src_bufs = []
    for f in frames:
        s = await to_a.createSources("src")
        await to_a.loadFrame(f, s, ctx.queue.load)
        src_bufs.append(s)
    await ctx.waitFinish(ctx.queue.load)
    state = {"t": 0}
    all_dests = []

    async def input_frame(keep_outputs: bool):
        """one interlaced frame arrives: producer side (ToRGBA, Yadif) + the channel's work for every frame it yields"""
        t = state["t"]
        state["t"] += 1
        srcs = src_bufs[t % N_SOURCES]
        for s in srcs:
            s.addRef()
            s.timestamp = t * 2
        rgba = await to_a.createDest(dims, "src")
        rgba.timestamp = t * 2
        to_a.processFrame("src", srcs, rgba)
        if len(yad.in_) >= 2:
            await jobs.runQueue({"source": "src", "timestamp": t * 2})
        outs, produced = [], []
        await yad.processFrame(rgba, outs, "src")
        for deint in outs:
            ts = deint.timestamp
            xfa = await ctx.createBuffer(w * h * 16, "readwrite", "coarse", dims, "mixer a")
            await xa.run(dict(input=deint, output=xfa, **IDENTITY_XF), {"source": "L0", "timestamp": ts}, lambda d=deint: d.release())
            await jobs.runQueue({"source": "L0", "timestamp": ts})
            for s in pip_srcs:
                s.addRef()
                s.timestamp = ts
            rgb = await to_b.createDest(dims, "pip")
            to_b.processFrame("pip", pip_srcs, rgb)
            xfb = await ctx.createBuffer(w * h * 16, "readwrite", "coarse", dims, "mixer b")
            await xb.run(dict(input=rgb, output=xfb, **pip_xf), {"source": "pip", "timestamp": ts}, lambda r_=rgb: r_.release())
            await jobs.runQueue({"source": "pip", "timestamp": ts})
            cdest = await ctx.createBuffer(w * h * 16, "readwrite", "coarse", dims, "comb")
            cdest.timestamp = ts
            await comb.run({"inputs": [xfa, xfb], "output": cdest}, {"source": "ch", "timestamp": ts}, lambda: None)
            await jobs.runQueue({"source": "ch", "timestamp": ts})
            xfa.release()
            xfb.release()
            dests = await frm.createDests("out")
            all_dests.append(dests)
            frm.processFrame("out", cdest, dests, Interlace.Progressive)
            await jobs.runQueue({"source": "out", "timestamp": ts})
            if keep_outputs:
                await frm.saveFrame(dests, ctx.queue.unload)
                await ctx.waitFinish(ctx.queue.unload)
                produced.append(dests[0].host.copy())
        return len(outs), produced

    for _ in range(4):   # fill the window (the first two input frames yield nothing)
        await input_frame(False)
    await ctx.waitFinish(ctx.queue.process)
    t_rec = state["t"]
    st0 = ctx.stats()
    ctx.beginChain()
    fields, produced = await input_frame(not args.no_parity)
    chain = ctx.endChain()
    st1 = ctx.stats()
    if not chain.complete or fields != 2:
        raise RuntimeError(f"yadif bench: the recorded input frame is not replayable (complete={chain.complete}, fields={fields})")

    parity = {"checked": False}
    if not args.no_parity and rank == 0 and check_outputs is not None:
        # the window at input frame t is (t-2, t-1, t); the de-interlaced frames are frame t-1's two fields (yadif.ts:104-113).
        # The checker is bench.py's (the product never touches the oracle).
        ok = check_outputs([frames[(t_rec - 2 + k) % N_SOURCES] for k in range(3)], pip_frame, pip_xf, dict(IDENTITY_XF), w, h, produced)
        parity = {"checked": True, "ok": bool(ok), "what": "the two output frames of the recorded input frame vs the oracle's unfused chain, byte for byte"}
        if not ok:
            raise RuntimeError("yadif bench: output differs from the oracle")

    def timed(n_inputs: int) -> float:
        e0, e1 = ctx.createEvent(), ctx.createEvent()
        barrier()
        _lib.lib().pb_wait_finish(ctx._need(), _lib.QUEUE_PROCESS)
        e0.record()
        for _ in range(n_inputs):
            chain.replay()
        e1.record()
        e1.synchronize()
        return e0.elapsed_ms(e1)

    n_inputs = args.steps * args.yadif_inputs_per_step
    timed(max(args.warmup, 3) * 8)
    sampler = clock_sampler_cls(local_rank)
    sampler.start()
    time.sleep(0.25)
    t0 = time.perf_counter()
    ms = max_over_ranks(timed(n_inputs))
    t1 = time.perf_counter()
    clocks = sampler.finish(t0, t1)
    if rank == 0:
        peak, peak_kind = measured_peak()
        packed = v210.getPitchBytes(w) * h
        field_us = ms * 1e3 / (2 * n_inputs)
        alg = packed // 2 + packed + packed   # per output frame: half an interlaced input frame, the PiP source, the output
        rgba_bytes = w * h * 16
        line = {
            "metric": f"{h}i50 v210 source -> Yadif -> 2-layer composite -> v210, output frames/sec", "value": world * 2 * n_inputs / (ms * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"the reference's operating point (index.ts:45-71): {world} x {w}x{h} interlaced v210 channel(s), one per GPU: ToRGBA -> Yadif send_field "
                                   f"(2 output frames per input frame) -> Mixer Transform -> Combine_2 with a 0.5x PiP of a second v210 source -> FromRGBA v210, 709->2020, "
                                   f"inputs={args.inputs}",
                       "input_frames_per_step": args.yadif_inputs_per_step, "launches_per_input_frame": chain.launches,
                       "materialised_per_input_frame": st1["materialised"] - st0["materialised"], "march_launches_per_input_frame": st1["march_launches"] - st0["march_launches"],
                       "l2_policy": f"working set of one input frame larger than L2 at every size: the RGBA-f32 window read by both fields (3 x {rgba_bytes} B), the new "
                                    f"RGBA-f32 frame ({rgba_bytes} B), two half-frame blocks of interpolated lines, the packed frames = {4 * rgba_bytes + rgba_bytes + 4 * (v210.getPitchBytes(w) * h)} B > 126 MiB"},
            "field_us": field_us,
            "roofline": {"bound": "hbm", "achieved": alg / (field_us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (field_us * 1e-6) / 1e9 / peak,
                         "traffic": None, "peak_kind": f"of {peak_kind}", "algorithmic_bytes_per_output_frame": alg,
                         "note": "packed bytes only (SURVEY 8d); the RGBA-f32 window the de-interlacer reads is working state, not algorithmic traffic"},
            "e2e": {"value": world * 2 * n_inputs / (ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "device-resident sources; the host<->device leg of a frame is measured by the default config"},
            "parity_checked": parity["checked"], "parity": parity,
            "gpu_launches": int(n_inputs * chain.launches),
            "clocks": clocks,
        }
        emit(line)
    barrier()
    yad.release()
    ctx.close()
    if dist:
        dist.destroy_process_group()
