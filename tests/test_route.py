"""ROUTE exchange (phaneron_b200/route.py): host logic on CPU with gloo, world size 2 and 3; the routed frame as a
layer of another channel on the GPU (single device, pb_buf_wrap + RGBA leaf)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from phaneron_b200.route import RouteExchange, RouteTable, channel_rank


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _frame(channel, period, n):
    return torch.from_numpy(np.random.default_rng(1000 * channel + period).integers(0, 256, n, dtype=np.uint8))


def _worker(rank, world, port, routes, n, periods, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ex = RouteExchange(RouteTable(routes), n, torch.device("cpu"))
        ok = True
        for p in range(periods):
            outgoing = {i: _frame(src, p, n) for i, (src, dst) in enumerate(routes) if channel_rank(src, world) == rank}
            ex.start(outgoing)
            got = ex.finish()
            for i, t in got.items():
                ok &= bool(torch.equal(t, _frame(routes[i][0], p, n)))
            # exactly the routes that end on this rank and start elsewhere arrive
            want = {i for i, (s, d) in enumerate(routes) if channel_rank(d, world) == rank and channel_rank(s, world) != rank}
            ok &= set(got) == want
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,routes", [(2, [(1, 0), (0, 1)]), (3, [(1, 0), (2, 1), (0, 2), (0, 3)])])
def test_route_exchange_gloo(world, routes):
    """ring cross-feed as BASELINE config 4 wires it (channel i layer 2 = ROUTE of channel i+1), plus a local route"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, routes, 4096, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def test_route_plan_keeps_local_routes_local():
    t = RouteTable([(0, 2), (1, 0), (3, 1)])   # world 2: channels 0,2 on rank 0; 1,3 on rank 1
    assert t.plan(0, 2) == ([], [(1, 1)])
    assert t.plan(1, 2) == ([(1, 0)], [])
    assert t.plan(0, 1) == ([], [])


@pytest.mark.gpu
def test_routed_frame_enters_another_channel_as_a_layer():
    """channel A's combined RGBA frame, handed over as device memory (what RouteExchange delivers), is layer 2 of
    channel B: result == oracle combine of B's own source with A's frame"""
    import oracle
    from gpu_util import Env, run
    from scene_oracle import SceneOracle
    from phaneron_b200.harness import ChannelHarness
    from phaneron_b200.process import v210
    from phaneron_b200.process.combine import Combine
    from phaneron_b200.process.image_process import ImageProcess
    from phaneron_b200.process.io import FromRGBA
    from phaneron_b200.route import buffer_as_tensor, tensor_as_buffer
    from phaneron_b200.scenes import layered_scene, single_layer_scene

    w, h = 480, 270
    scene_a = layered_scene(w, h, 2, "noise", "plain", "709", "709")
    scene_b = single_layer_scene(w, h, "ramp", False, "709", "709")

    async def go():
        async with Env() as env:
            ha = ChannelHarness(env.ctx, scene_a, env.pj, chanID="A")
            await ha.init()
            frame_a = await ha.compose(await ha.upload_all(0), 0)           # deferred RGBA expression
            sent = buffer_as_tensor(frame_a, torch.device("cuda", 0)).clone()   # materialise + "send"
            frame_a.release()
            routed = tensor_as_buffer(env.ctx, sent, w, h)                    # "receive" side
            hb = ChannelHarness(env.ctx, scene_b, env.pj, chanID="B")
            await hb.init()
            own = await hb.compose(await hb.upload_all(0), 0)
            comb = ImageProcess(env.ctx, Combine(2, w, h), hb.clJobs)
            await comb.init()
            dest = await env.ctx.createBuffer(w * h * 16, "readwrite", "coarse", {"width": w, "height": h}, "B comb")
            dest.timestamp = 0
            await comb.run({"inputs": [own, routed], "output": dest}, {"source": "B", "timestamp": 0}, lambda: None)
            await hb.clJobs.runQueue({"source": "B", "timestamp": 0})
            out = (await hb.consume(dest))[0].host.copy()
            return out
    ours = run(go())
    a = SceneOracle(scene_a).composite()
    b = SceneOracle(scene_b).composite()
    so = SceneOracle(scene_b)
    ref = oracle.v210_write(oracle.combine([b, a]), w, h, 0, so.cm_w, so.lut_w)
    assert np.array_equal(ours, ref)
