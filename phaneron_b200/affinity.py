"""One process per GPU: keep the process -- and therefore the pinned frame buffers it first-touches -- on
the CPU cores / NUMA node the GPU hangs off.  Without it eight channels' worth of H2D traffic crosses the socket
interconnect and the public-API path stops scaling long before the kernels do."""
from __future__ import annotations

import os
from typing import List, Optional


def gpu_cpu_affinity(gpu_index: int) -> Optional[List[int]]:
    """CPUs NVML reports as local to the GPU (what `nvidia-smi topo -m` prints), None if unknown"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpus = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpus + 63) // 64)
        cpus = [w * 64 + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1]
        return cpus or None
    except Exception:
        return None


def bind_to_gpu(gpu_index: int, n_local_ranks: int = 1, local_rank: int = 0) -> Optional[List[int]]:
    """Restrict this process to the GPU's local CPUs; ranks sharing a CPU set take disjoint slices of it so that their
    copy / event threads do not pile onto the same cores.  Returns the CPU list in force (None: left unchanged)."""
    cpus = gpu_cpu_affinity(gpu_index)
    if not cpus or not hasattr(os, "sched_setaffinity"):
        return None
    allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
    if not allowed:
        return None
    if n_local_ranks > 1:
        # ranks whose GPUs report the same CPU set split it evenly (typically 4 GPUs per socket)
        try:
            import pynvml
            same = [i for i in range(n_local_ranks) if gpu_cpu_affinity(i) == cpus]
        except Exception:
            same = list(range(n_local_ranks))
        if local_rank in same and len(allowed) >= 2 * len(same):
            per = len(allowed) // len(same)
            k = same.index(local_rank)
            allowed = allowed[k * per:(k + 1) * per]
    try:
        os.sched_setaffinity(0, allowed)
    except OSError:
        return None
    return allowed
