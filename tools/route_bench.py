"""BASELINE.json config 4: N independent 1080p50 v210 channels, one per GPU, each with a second layer that is the ROUTE
of its neighbour channel's output, moved point-to-point over NCCL (phaneron_b200/route.py).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/route_bench.py [--frames K]

Per frame and rank: ToRGBA(own v210 source) ; Transform(routed RGBA frame of channel r+1 -> 0.5 PiP) ; Combine_2 ; the
combined frame is materialised once as RGBA-f32 (it is both the FromRGBA input and the ROUTE payload, exactly the buffer the
reference shares between channels), packed to v210, and sent to channel r-1 while channel r+1's frame arrives for the next
frame period.  Prints one JSON line from rank 0: whole-job frames/s (max time over ranks), bytes routed per frame."""
import argparse, asyncio, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phaneron_b200 import ClProcessJobs, clContext
from phaneron_b200.process import v210
from phaneron_b200.process.combine import Combine
from phaneron_b200.process.image_process import ImageProcess
from phaneron_b200.process.io import FromRGBA, ToRGBA
from phaneron_b200.process.transform import Transform
from phaneron_b200.route import RouteExchange, RouteTable, buffer_as_tensor, tensor_as_buffer
from phaneron_b200.scenes import make_frame, pip

W, H = 1920, 1080


async def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = clContext({"deviceIndex": local})
    await ctx.initialise()
    jobs = ClProcessJobs(ctx).getJobs()
    toRGBA = ToRGBA(ctx, "709", "709", v210.Reader(W, H), jobs)
    fromRGBA = FromRGBA(ctx, "709", v210.Writer(W, H, False), jobs)
    xform = ImageProcess(ctx, Transform(ctx, W, H), jobs)
    comb = ImageProcess(ctx, Combine(2, W, H), jobs)
    for o in (toRGBA, fromRGBA, xform, comb):
        await o.init()
    # channel r, layer 2 = ROUTE of channel r+1
    routes = [((r + 1) % world, r) for r in range(world)]
    ex = RouteExchange(RouteTable(routes), W * H * 16, dev)
    my_in = [i for i, (s, d) in enumerate(routes) if d == rank][0]
    my_out = [i for i, (s, d) in enumerate(routes) if s == rank][0]
    local_route = world == 1
    srcs = await toRGBA.createSources(f"ch{rank}")
    await toRGBA.loadFrame(make_frame("noise", W, H, rank), srcs)
    for s in srcs:
        s.addRef()   # the same source frame is replayed every period
    routed_t = torch.zeros(W * H * 16, dtype=torch.uint8, device=dev)   # first period: black / transparent
    dests = await fromRGBA.createDests(f"ch{rank}")
    xfp = dict(pip(0.5, 0.25, 0.25))

    async def frame(ts, routed_tensor):
        own = await toRGBA.createDest({"width": W, "height": H}, f"ch{rank}")
        own.timestamp = ts
        for s in srcs:
            s.addRef(); s.timestamp = ts
        toRGBA.processFrame(f"ch{rank}", srcs, own)
        await jobs.runQueue({"source": f"ch{rank}", "timestamp": ts})
        routed = tensor_as_buffer(ctx, routed_tensor, W, H, "route in")
        pipd = await ctx.createBuffer(W * H * 16, "readwrite", "coarse", {"width": W, "height": H}, "route pip")
        pipd.timestamp = ts
        await xform.run(dict(input=routed, output=pipd, **xfp), {"source": f"r{rank}", "timestamp": ts}, lambda: routed.release())
        await jobs.runQueue({"source": f"r{rank}", "timestamp": ts})
        out = await ctx.createBuffer(W * H * 16, "readwrite", "coarse", {"width": W, "height": H}, "chan")
        out.timestamp = ts
        await comb.run({"inputs": [own, pipd], "output": out}, {"source": f"c{rank}", "timestamp": ts}, lambda: None)
        await jobs.runQueue({"source": f"c{rank}", "timestamp": ts})
        own.release(); pipd.release()
        payload = buffer_as_tensor(out, dev)          # materialises the channel frame once (one fused launch)
        out.addRef()
        fromRGBA.processFrame(f"o{rank}", out, dests, None)
        await jobs.runQueue({"source": f"o{rank}", "timestamp": ts})
        await ctx.waitFinish(ctx.queue.process)
        return out, payload

    async def run(n, routed_tensor):
        for i in range(n):
            out, payload = await frame(i, routed_tensor)
            if local_route:
                routed_tensor = payload.clone()
            else:
                ex.start({my_out: payload})
                routed_tensor = ex.finish()[my_in]
            out.release()
        return routed_tensor

    routed_t = await run(a.warmup, routed_t)
    torch.cuda.synchronize(); dist.barrier()
    st0 = ctx.stats()
    t0 = time.perf_counter()
    routed_t = await run(a.frames, routed_t)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    st1 = ctx.stats()
    if rank == 0:
        print(json.dumps({"config": "8x1080p50 channels one per GPU with ROUTE cross-feed (BASELINE.json configs[3])", "n_gpus": world,
                          "value": world * a.frames / float(dt.item()), "unit": "frames/s (all channels)", "ms_per_frame_period": float(dt.item()) / a.frames * 1e3,
                          "route_bytes_per_frame_per_gpu": 0 if local_route else W * H * 16, "route": "torch.distributed P2P over NCCL (RGBA-f32 channel frame)",
                          "kernel_launches_per_frame": (st1["kernel_launches"] - st0["kernel_launches"]) / a.frames, "frames": a.frames}), flush=True)
    dist.barrier()
    dist.destroy_process_group()

asyncio.run(main())
