"""H2D / D2H bandwidth of pinned host memory on this box (the ceiling of bench.py's e2e leg):
5 x 22 MB uploads + 1 x 22 MB download per 2160p frame."""
import torch, time
n = 22118400
src = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(5)]
dst = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(5)]
back = torch.empty(n, dtype=torch.uint8).pin_memory()
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
def run(frames, both):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(frames):
        with torch.cuda.stream(s_up):
            for a, b in zip(src, dst):
                b.copy_(a, non_blocking=True)
        if both:
            with torch.cuda.stream(s_dn):
                back.copy_(dst[0], non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t0
run(5, True)
for both in (False, True):
    dt = run(100, both)
    print(f"{'H2D+D2H' if both else 'H2D only'}: {100 / dt:.0f} frames/s, H2D {100 * 5 * n / dt / 1e9:.1f} GB/s" + (f", D2H {100 * n / dt / 1e9:.1f} GB/s" if both else ""))

# latency of ONE 22 MB D2H copy (issued and waited for, as a consumer does) while the H2D engine is kept busy
torch.cuda.synchronize()
for busy in (False, True):
    lat = []
    for it in range(20):
        if busy:
            with torch.cuda.stream(s_up):
                for _ in range(2):
                    for a, b in zip(src, dst):
                        b.copy_(a, non_blocking=True)
        t0 = time.perf_counter()
        with torch.cuda.stream(s_dn):
            back.copy_(dst[0], non_blocking=True)
        s_dn.synchronize()
        lat.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
    lat.sort()
    print(f"D2H 22 MB latency, H2D {'busy' if busy else 'idle'}: median {lat[len(lat)//2]*1e3:.3f} ms, min {lat[0]*1e3:.3f} ms")
