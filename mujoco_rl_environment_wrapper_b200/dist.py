"""Multi-GPU plumbing: environments shard by index, one process per GPU, and NOTHING is exchanged
on the step path.  The only collective is an optional end-of-rollout all-gather of a few episode
statistics per rank (NCCL on GPUs, gloo in CPU tests)."""
import os

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world: int):
    """Contiguous env index range [lo, hi) owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(total_envs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), rank=rank, world_size=world)
    return rank, local, world


def allgather_episode_stats(stats: torch.Tensor):
    """[k] per-rank statistics (returns, lengths, done counts ...) -> [world, k] on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return stats.unsqueeze(0)
    out = [torch.empty_like(stats) for _ in range(dist.get_world_size())]
    dist.all_gather(out, stats)
    return torch.stack(out)


def max_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
