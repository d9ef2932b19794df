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
