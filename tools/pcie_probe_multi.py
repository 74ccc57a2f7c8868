"""aggregate H2D bandwidth of N ranks, one per GPU, copying 5 x 22 MB pinned frames each at the same time -- the
ceiling of bench.py's e2e leg at N GPUs.  PB_NO_AFFINITY=1: leave CPU placement to the OS (A/B for phaneron_b200/affinity.py)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe_multi.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
from phaneron_b200.affinity import bind_to_gpu
cpus = None if os.environ.get("PB_NO_AFFINITY") else bind_to_gpu(lr, world, lr)
import torch, torch.distributed as dist
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 22118400
src = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(5)]
for s in src: s.fill_(rank)   # first touch after the affinity is in force
dst = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(5)]
back = torch.empty(n, dtype=torch.uint8).pin_memory()
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
def run(frames):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(frames):
        with torch.cuda.stream(s_up):
            for a, b in zip(src, dst): b.copy_(a, non_blocking=True)
        with torch.cuda.stream(s_dn):
            back.copy_(dst[0], non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t0
run(5)
dt = run(100)
t = torch.tensor([dt], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"[pcie x{world}] affinity={'off' if cpus is None else f'{len(cpus)} cpus/rank'}: {world * 100 / t.item():.0f} frames/s aggregate, "
          f"H2D {world * 100 * 5 * n / t.item() / 1e9:.1f} GB/s total, {100 * 5 * n / t.item() / 1e9:.1f} GB/s per GPU", file=sys.stderr, flush=True)
dist.destroy_process_group()
