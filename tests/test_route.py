"""ROUTE (phaneron_b200/route.py, csrc/pb_route.cu): the exchange plan on CPU with gloo (world size 2 and 3), the
RouteProducer mirror, and on the GPU the C-ABI NCCL path: a frame routed through pb_route_send / pb_route_recv enters
another channel as a layer (one GPU: rank 0 to itself; two GPUs when the box has them)."""
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


def test_route_producer_mirror():
    """routeProducer.ts:44-73,106-126,139-160: url parsing and errors, fork reference counting, release"""
    import asyncio
    from phaneron_b200.route import InvalidProducerError, RemoteChannel, RouteProducer, chanLayerFromString

    assert chanLayerFromString("2-10") == {"valid": True, "channel": 2, "layer": 10}
    assert chanLayerFromString("1") == {"valid": True, "channel": 1, "layer": 0}
    assert chanLayerFromString("x")["valid"] is False

    class Frame:
        def __init__(self):
            self.refs = 1

        def addRef(self):
            self.refs += 1

    frames = []

    def fetch():
        frames.append(Frame())
        return frames[-1]

    async def go():
        with pytest.raises(InvalidProducerError):
            RouteProducer(1, {"url": "file://clip.mxf", "layer": 10}, [])
        with pytest.raises(RuntimeError, match="failed to find route source"):
            await RouteProducer(1, {"url": "ROUTE 1", "layer": 10}, []).initialise()
        with pytest.raises(RuntimeError, match="failed to find source of channel 3"):
            await RouteProducer(1, {"url": "route://3", "layer": 10}, [None, None]).initialise()
        chans = [RemoteChannel(fetch, {"width": 1920, "height": 1080})]
        with pytest.raises(RuntimeError, match="Failed to find source pipes for layer 10"):
            await RouteProducer(1, {"url": "route://1-10", "layer": 10}, chans).initialise()
        rp = RouteProducer(7, {"url": "ROUTE://1", "layer": 20}, chans)
        with pytest.raises(RuntimeError, match="failed to find source pipes"):
            rp.getSourcePipes()
        await rp.initialise()
        assert rp.srcID() == "P7 ROUTE://1 L20"
        a, b = rp.getSourcePipes(), rp.getSourcePipes()
        assert a["format"] == {"width": 1920, "height": 1080} and rp.numForks == 2
        rp.setPaused(False)
        f = await a["video"]()
        assert f.refs == 1 + 1 + 1     # the landing frame's own reference, the consumer's (RemoteChannel), one extra fork
        rp.setPaused(True)
        f = await a["video"]()
        assert f.refs == 4             # a paused producer holds one more (routeProducer.ts:123)
        b["release"]()
        b["release"]()
        assert rp.numForks == 1
        rp.release()
        assert await a["video"]() is None
    asyncio.run(go())


def _compose_with_routed(env, w, h, scene_b, routed_buf):
    """channel B = its own source with the routed frame combined on top -> packed v210 bytes"""
    from phaneron_b200.harness import ChannelHarness
    from phaneron_b200.process.combine import Combine
    from phaneron_b200.process.image_process import ImageProcess

    async def go():
        hb = ChannelHarness(env.ctx, scene_b, env.pj, chanID="B")
        await hb.init()
        own = await hb.compose(await hb.upload_all(0), 0)
        comb = ImageProcess(env.ctx, Combine(2, w, h), hb.clJobs)
        await comb.init()
        dest = await env.ctx.createBuffer(w * h * 16, "readwrite", "coarse", {"width": w, "height": h}, "B comb")
        dest.timestamp = 0
        routed_buf.addRef()
        await comb.run({"inputs": [own, routed_buf], "output": dest}, {"source": "B", "timestamp": 0}, lambda: None)
        await hb.clJobs.runQueue({"source": "B", "timestamp": 0})
        own.release()
        routed_buf.release()
        return (await hb.consume(dest))[0].host.copy()
    return go()


@pytest.mark.gpu
def test_routed_frame_through_the_c_abi_enters_another_channel_as_a_layer():
    """channel A's combined (still deferred) RGBA frame goes through pb_route_send / pb_route_recv -- NCCL, rank 0 to itself --
    into a landing buffer that is layer 2 of channel B: result == oracle combine of B's own source with A's frame"""
    import oracle
    from gpu_util import Env, run
    from scene_oracle import SceneOracle
    from phaneron_b200.harness import ChannelHarness
    from phaneron_b200.route import GpuRouteExchange, RouteComm, RouteTable
    from phaneron_b200.scenes import layered_scene, single_layer_scene

    w, h = 480, 270
    scene_a = layered_scene(w, h, 2, "noise", "plain", "709", "709")
    scene_b = single_layer_scene(w, h, "ramp", False, "709", "709")

    async def go():
        async with Env() as env:
            comm = RouteComm(env.ctx, 0, 1, RouteComm.unique_id())
            try:
                ha = ChannelHarness(env.ctx, scene_a, env.pj, chanID="A")
                await ha.init()
                frame_a = await ha.compose(await ha.upload_all(0), 0)           # deferred RGBA expression
                assert frame_a.deferred
                landing = await env.ctx.createBuffer(w * h * 16, "readwrite", "coarse", {"width": w, "height": h}, "landing")
                before = env.ctx.stats()
                comm.begin()
                comm.send(frame_a, 0)      # materialises the frame (one fused launch), then ncclSend on the side stream
                comm.recv(landing, 0)
                comm.end()
                frame_a.release()          # the exchange holds its own reference until it has completed
                comm.wait()                # the process queue waits on the device; the host goes on
                after = env.ctx.stats()
                assert after["materialised"] - before["materialised"] == 1
                out = await _compose_with_routed(env, w, h, scene_b, landing)
                landing.release()
                assert comm.info()["bytes_sent"] == w * h * 16 == comm.info()["bytes_received"]
                return out
            finally:
                comm.close()
    ours = run(go())
    a = SceneOracle(scene_a).composite()
    b = SceneOracle(scene_b).composite()
    so = SceneOracle(scene_b)
    ref = oracle.v210_write(oracle.combine([b, a]), w, h, 0, so.cm_w, so.lut_w)
    assert np.array_equal(ours, ref)


def _nccl_worker(rank, world, uid, q, w, h, attach=False):
    """rank r composes channel r and routes its frame to rank 1 - r; -> channel r with the peer's frame as layer 2"""
    import asyncio
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from gpu_util import Env
    from phaneron_b200 import ClProcessJobs, clContext
    from phaneron_b200.harness import ChannelHarness
    from phaneron_b200.route import GpuRouteExchange, RouteComm, RouteTable
    from phaneron_b200.scenes import layered_scene

    async def go():
        ctx = clContext({"deviceIndex": rank})
        await ctx.initialise()

        class E:
            pass
        env = E()
        env.ctx, env.pj = ctx, ClProcessJobs(ctx)
        comm = RouteComm(ctx, rank, world, uid)
        ex = GpuRouteExchange(ctx, comm, RouteTable([(1, 0), (0, 1)]), w, h)
        await ex.init()
        if attach:   # copy-engine transport: the peer pushes straight into this rank's landing buffers
            my_in = [i for i, (s, d) in enumerate(ex.table.routes) if d == rank][0]
            assert comm.attach(ex.landing[my_in], 1 - rank, 1 - rank), "CUDA IPC mapping between the two GPUs failed"

        mine = layered_scene(w, h, 2, "noise", "plain", "709", "709", frame_set=rank)
        ha = ChannelHarness(ctx, mine, env.pj, chanID=f"A{rank}")
        await ha.init()
        frame = await ha.compose(await ha.upload_all(0), 0)
        my_out = [i for i, (s, d) in enumerate(ex.table.routes) if s == rank][0]
        ex.start({my_out: frame})
        frame.release()
        got = ex.finish()
        routed = got[[i for i, (s, d) in enumerate(ex.table.routes) if d == rank][0]]
        own_scene = dict(mine, layers=[mine["layers"][0]])
        out = await _compose_with_routed(env, w, h, own_scene, routed)
        ex.close()
        comm.close()
        ctx.close()
        return out
    q.put((rank, asyncio.run(go())))


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["nccl", "copy_engines"])
def test_route_between_two_gpus(transport):
    """two processes, two GPUs: each channel's frame crosses (pb_route_*: NCCL point-to-point, or pushed by the copy engines into
    the peer's IPC-mapped landing buffer after pb_route_attach) and is composited by the other"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import oracle
    from scene_oracle import SceneOracle
    from phaneron_b200.route import RouteComm
    from phaneron_b200.scenes import layered_scene
    w, h = 480, 270
    uid = RouteComm.unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, uid, q, w, h, transport == "copy_engines")) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r in range(2):
        mine = layered_scene(w, h, 2, "noise", "plain", "709", "709", frame_set=r)
        peer = layered_scene(w, h, 2, "noise", "plain", "709", "709", frame_set=1 - r)
        so = SceneOracle(mine)
        own = SceneOracle(dict(mine, layers=[mine["layers"][0]])).composite()
        ref = oracle.v210_write(oracle.combine([own, SceneOracle(peer).composite()]), w, h, 0, so.cm_w, so.lut_w)
        assert np.array_equal(res[r], ref), r
