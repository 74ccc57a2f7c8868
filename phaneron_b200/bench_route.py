"""bench.py --config route: BASELINE.json configs[3] -- N independent 1080p50 v210 channels, one per GPU, each with a
second layer that is the ROUTE of its neighbour channel's output, moved point-to-point over NCCL through the library's own
C ABI (pb_route_*, csrc/pb_route.cu).

Per frame and rank r: ToRGBA(own v210 source) ; Transform(routed RGBA frame of channel r+1 -> 0.5 PiP) ; Combine_2 ; the
combined frame is materialised once as RGBA-f32 (it is both the FromRGBA input and the ROUTE payload, exactly the buffer
the reference shares between channels: routeProducer.ts:63-73, channel.ts:289-300), packed to v210, and sent to channel
r-1 while channel r+1's frame arrives.

Steady state without Python operators in the loop: the launches of a frame are recorded once per buffer slot
(pb_chain_*), and a frame period is  pb_route_wait_age(1) ; pb_chain_replay ; pb_route_begin/send/recv/end  -- a frame is
composed from what the exchange before last delivered, so exchange n runs on the side stream while frame n + 1's kernels
run on the process queue (three landing / payload slots).
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

W, H = 1920, 1080
SLOTS = 3


async def run(args, rank: int, world: int, local_rank: int, emit, clock_sampler_cls, measured_peak):
    from . import ClProcessJobs, clContext, _lib
    from .process import v210
    from .process.combine import Combine
    from .process.image_process import ImageProcess
    from .process.io import FromRGBA, ToRGBA
    from .process.transform import Transform
    from .route import RouteComm
    from .scenes import make_frame, pip

    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist_.init_process_group("gloo")   # host-side rendezvous only (NCCL id, barriers, max over ranks): no torch on the data path
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

    ctx = clContext({"deviceIndex": local_rank})
    await ctx.initialise()
    jobs = ClProcessJobs(ctx).getJobs()
    toRGBA = ToRGBA(ctx, "709", "709", v210.Reader(W, H), jobs)
    fromRGBA = FromRGBA(ctx, "709", v210.Writer(W, H, False), jobs)
    xform = ImageProcess(ctx, Transform(ctx, W, H), jobs)
    comb = ImageProcess(ctx, Combine(2, W, H), jobs)
    for o in (toRGBA, fromRGBA, xform, comb):
        await o.init()

    ids = [RouteComm.unique_id() if rank == 0 else None]
    if dist:
        dist.broadcast_object_list(ids, src=0)
    comm = RouteComm(ctx, rank, world, ids[0])
    # channel r, layer 2 = ROUTE of channel r+1: r receives from r+1 and sends to r-1
    peer_in, peer_out = (rank + 1) % world, (rank - 1) % world

    frame_bytes = W * H * 16
    dims = {"width": W, "height": H}
    srcs = await toRGBA.createSources(f"ch{rank}")
    await toRGBA.loadFrame(make_frame(args.inputs, W, H, rank), srcs)
    landing = [await ctx.createBuffer(frame_bytes, "readwrite", "coarse", dims, f"route landing {k}") for k in range(SLOTS)]
    zero = np.zeros(frame_bytes, np.uint8)   # first periods: black / transparent routed frames
    for b in landing:
        await b.hostAccess("writeonly", ctx.queue.load, zero)
    await ctx.waitFinish(ctx.queue.load)
    # between the GPUs of the node the frames cross with the copy engines (CUDA IPC + stream memory operations); NCCL carries
    # the handles once, and the frames themselves where mapping is impossible (one GPU: rank 0 routes to itself)
    transport = "copy engines (CUDA IPC push + cuStreamWaitValue32 flow control)" if comm.attach(landing, peer_in, peer_out) else "NCCL point-to-point"
    xfp = dict(pip(0.5, 0.25, 0.25))

    # ---- record one chain per slot: chain k composes from landing[(k + 1) % 3] (filled two exchanges ago) into out[k] ----
    chains, outs, dests_all = [], [], []
    for k in range(SLOTS):
        dests = await fromRGBA.createDests(f"ch{rank}")
        ctx.beginChain()
        ts = k
        own = await toRGBA.createDest(dims, f"ch{rank}")
        own.timestamp = ts
        for s in srcs:
            s.addRef()
            s.timestamp = ts
        toRGBA.processFrame(f"ch{rank}", srcs, own)
        await jobs.runQueue({"source": f"ch{rank}", "timestamp": ts})
        routed = landing[(k + 1) % SLOTS]
        routed.addRef()
        pipd = await ctx.createBuffer(frame_bytes, "readwrite", "coarse", dims, "route pip")
        pipd.timestamp = ts
        await xform.run(dict(input=routed, output=pipd, **xfp), {"source": f"r{rank}", "timestamp": ts}, lambda r=routed: r.release())
        await jobs.runQueue({"source": f"r{rank}", "timestamp": ts})
        out = await ctx.createBuffer(frame_bytes, "readwrite", "coarse", dims, f"chan out {k}")
        out.timestamp = ts
        await comb.run({"inputs": [own, pipd], "output": out}, {"source": f"c{rank}", "timestamp": ts}, lambda: None)
        await jobs.runQueue({"source": f"c{rank}", "timestamp": ts})
        own.release()
        pipd.release()
        out.devicePointer()          # the channel frame as RGBA-f32: ONE fused launch, recorded (payload and FromRGBA input)
        out.addRef()
        fromRGBA.processFrame(f"o{rank}", out, dests, None)
        await jobs.runQueue({"source": f"o{rank}", "timestamp": ts})
        chain = ctx.endChain()
        if not chain.complete:
            raise RuntimeError("route bench: the recorded frame is not replayable")
        chains.append(chain)
        outs.append(out)
        dests_all.append(dests)
    await ctx.waitFinish(ctx.queue.process)
    launches_per_frame = chains[0].launches

    lib = _lib.lib()

    counter = {"ex": 0}   # exchange periods so far: the slot rotation continues across timed runs

    def period(n: int, exchange: bool) -> None:
        if exchange:
            n = counter["ex"]
            counter["ex"] += 1
        k = n % SLOTS
        if exchange:
            comm.wait(_lib.QUEUE_PROCESS, age=1)   # landing[(k+1)%3] was filled by exchange n-2; out[k] was sent by exchange n-3
        chains[k].replay()
        if exchange and world >= 1:
            comm.begin()
            comm.send(outs[k], peer_out)
            comm.recv(landing[k], peer_in)
            comm.end()

    def timed(n_frames: int, exchange: bool) -> float:
        e0, e1 = ctx.createEvent(), ctx.createEvent()
        barrier()
        lib.pb_wait_finish(ctx._need(), _lib.QUEUE_PROCESS)
        e0.record()
        for n in range(n_frames):
            period(n, exchange)
        if exchange:
            comm.wait(_lib.QUEUE_PROCESS, age=0)
        e1.record()
        e1.synchronize()
        comm.sync()
        return e0.elapsed_ms(e1)

    frames = args.steps * args.route_frames_per_step
    timed(max(args.warmup, 3) * 8, True)
    sampler = clock_sampler_cls(local_rank)
    sampler.start()
    time.sleep(0.25)
    t0 = time.perf_counter()
    ms = max_over_ranks(timed(frames, True))
    t1 = time.perf_counter()
    clocks = sampler.finish(t0, t1)
    timed(24, False)
    ms_local = max_over_ranks(timed(frames, False))   # the same frames with no exchange at all (routed layer = a resident frame)
    info = comm.info()

    # parity of the plumbing: after the run, frame `last` of this rank must equal the oracle's composite of (own source, PiP
    # of the peer's frame two periods earlier) -- checked in tests/test_route.py at small size; here only that bytes moved
    if rank == 0:
        peak, peak_kind = measured_peak()
        per_frame_us = ms * 1e3 / frames
        # bytes through HBM per frame period: own v210 in, routed RGBA in, channel RGBA out + in again for the writer, v210 out, + the exchange's read and write
        alg = v210.getPitchBytes(W) * H * 2 + frame_bytes * 3
        line = {
            "metric": "1080p50 v210 channels with ROUTE cross-feed, frames/sec over all channels", "value": world * frames / (ms * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE.json configs[3]: {world} x 1920x1080 v210 channel(s), one per GPU, 2 layers each (own source + 0.5x PiP of the "
                                   f"neighbour channel's ROUTEd RGBA-f32 frame), 709, inputs={args.inputs}",
                       "frames_per_step": args.route_frames_per_step, "launches_per_frame": launches_per_frame,
                       "route": "pb_route_* (C ABI) on a side stream, one exchange per frame period, exchange n overlaps frame n+1; transport: " + transport,
                       "route_bytes_per_frame_per_gpu": frame_bytes, "l2_policy": "each frame period streams 110 MB through HBM (> L2 with the exchange buffers rotating over 3 slots)"},
            "frame_period_us": per_frame_us, "frame_period_us_without_route": ms_local * 1e3 / frames,
            "route_overhead": per_frame_us / (ms_local * 1e3 / frames),
            "nvlink_GBps_per_gpu_per_direction": frame_bytes / (per_frame_us * 1e-6) / 1e9,
            "nccl_bytes_sent_rank0": info["bytes_sent"],
            "roofline": {"bound": "hbm", "achieved": alg / (per_frame_us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (per_frame_us * 1e-6) / 1e9 / peak, "traffic": None, "peak_kind": f"of {peak_kind}",
                         "algorithmic_bytes_per_frame": alg},
            "e2e": {"value": world * frames / (ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "device-resident sources (the config measures the cross-GPU hand-off); the host<->device leg of a frame is measured by the default config"},
            "gpu_launches": int(frames * launches_per_frame),
            "clocks": clocks,
        }
        emit(line)
    barrier()
    comm.close()
    ctx.close()
    if dist:
        dist.destroy_process_group()
