"""Multi-GPU fan-out of independent video streams: one process per GPU, one CVideoProcessor instance and one CUDA
stream each, NO collective on the data path (the reference has no multi-GPU notion at all).
torch.distributed is used only as the job's clock: a barrier around the timed region and a MAX over ranks."""
import os


def rank_info(env=os.environ):
    return int(env.get("RANK", "0")), int(env.get("WORLD_SIZE", "1")), int(env.get("LOCAL_RANK", "0"))


def assign_streams(n_streams, world_size):
    """stream s -> rank s mod world_size (BASELINE config 5: 8 streams on 8 GPUs -> one each). Returns list per rank."""
    out = [[] for _ in range(world_size)]
    for s in range(n_streams):
        out[s % world_size].append(s)
    return out


def barrier(dist, sync=None):
    if sync:
        sync()
    if dist is not None and dist.is_initialized():
        dist.barrier()
    if sync:
        sync()


def max_over_ranks(value_ms, dist, device="cpu"):
    """the job's elapsed time is the slowest rank's"""
    if dist is None or not dist.is_initialized():
        return float(value_ms)
    import torch
    t = torch.tensor([float(value_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_fps(frames_per_rank, world_size, elapsed_ms_max):
    """whole-job throughput: every rank rendered frames_per_rank frames within the slowest rank's interval"""
    return frames_per_rank * world_size / (elapsed_ms_max * 1e-3)


def bind_to_gpu_numa(local_rank):
    """Pin this process (the host thread that feeds GPU `local_rank`) and its future allocations -- the pinned staging buffers in
    particular -- to the NUMA node the GPU hangs off, so that the H2D / D2H copies of 8 concurrent streams do not cross the socket
    interconnect or pile onto one memory controller (SCALE_r01: per-GPU D2H fell from 52 to 33 GB/s at N = 8 without this).
    Topology from sysfs (/sys/bus/pci/devices/<bdf>/numa_node, local_cpulist), NVML as the fallback.  Returns a small dict for the
    bench line; never raises (an unbound process still works, just slower)."""
    info = {"numa_node": None, "cpus": None, "bound": False, "source": None}
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
    except Exception as e:                                     # no torch / no GPU: nothing to bind to
        info["error"] = repr(e); return info
    node, cpus = None, None
    try:
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read().strip())
        cpus = _parse_cpulist(open(base + "/local_cpulist").read().strip())
        info["source"] = "sysfs"
    except Exception:
        node = None
    if node is None or node < 0 or not cpus:
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByPciBusId(bdf.encode())
            words = (os.cpu_count() + 63) // 64
            mask = N.nvmlDeviceGetCpuAffinityWithinScope(h, words, getattr(N, "NVML_AFFINITY_SCOPE_NODE", 0))
            cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1]
            nm = N.nvmlDeviceGetMemoryAffinity(h, 1, getattr(N, "NVML_AFFINITY_SCOPE_NODE", 0))
            node = next((b for b in range(64) if (int(nm[0]) >> b) & 1), None)
            info["source"] = "nvml"
        except Exception as e:
            info["error"] = repr(e)
    try:
        allowed = os.sched_getaffinity(0)
        if cpus:
            want = set(cpus) & set(allowed)
            if want:
                os.sched_setaffinity(0, want)
                info["cpus"] = len(want); info["bound"] = True
        if node is not None and node >= 0:
            info["numa_node"] = node
            _set_mempolicy_preferred(node)
    except Exception as e:
        info["error"] = repr(e)
    return info


def _parse_cpulist(s):
    out = []
    for part in s.split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def _set_mempolicy_preferred(node):
    """set_mempolicy(MPOL_PREFERRED, {node}) through the raw syscall (x86-64: 238); best effort"""
    import ctypes
    import platform
    if platform.machine() != "x86_64" or node >= 64:
        return False
    libc = ctypes.CDLL(None, use_errno=True)
    mask = ctypes.c_ulong(1 << node)
    return libc.syscall(238, 1, ctypes.byref(mask), 65) == 0
