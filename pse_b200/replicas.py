"""Multi-GPU host logic, round 1: independent replicas (one process per GPU, one suspension each).

The PSE step has no data-path collective in this mode; torch.distributed is used only to agree on the
timing (max over ranks) and to add up the work done.  Works with any backend (NCCL on GPUs, gloo in the
CPU tests)."""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def replica_seeds(base_seed, rank):
    """(position seed, force seed, engine seed) of a rank's suspension: distinct, reproducible streams."""
    return base_seed + rank, base_seed + 100 + rank, base_seed + 1 + rank


def max_over_ranks(x, device="cpu"):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, device="cpu"):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank, seconds_this_rank, device="cpu"):
    """Whole-job throughput: all units processed by all ranks divided by the slowest rank's time."""
    return sum_over_ranks(units_this_rank, device) / max_over_ranks(seconds_this_rank, device)
