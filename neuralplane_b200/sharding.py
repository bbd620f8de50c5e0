"""Host-side sharding of an aircraft population over ranks (one process per GPU, torch.distributed).

Control / heading / tracking / planning aircraft never interact (reference: envs/tasks/task_base.py:70-72, every
termination condition is elementwise), so the population is split into contiguous global index ranges and NO
data-path collective exists: each rank runs its own env with `index_base` = first global index, and the in-kernel
Philox streams are keyed by global index, so results do not depend on the world size.  The only collectives are
bookkeeping: termination counters (SUM) and the benchmark's timing (MAX).  Works on any backend (nccl on the GPUs,
gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Contiguous split of [0, n_total) into `world` ranges whose sizes differ by at most one aircraft PAIR
    (the step kernel moves aircraft in pairs, so every range but the last starts and ends on an even index).
    Returns (index_base, n_local)."""
    if world < 1 or not 0 <= rank < world or n_total < 0:
        raise ValueError(f"bad shard request: n_total={n_total} rank={rank} world={world}")
    pairs = (n_total + 1) // 2
    lo = (pairs * rank // world) * 2
    hi = min((pairs * (rank + 1) // world) * 2, n_total)
    return lo, max(hi - lo, 0)


def reduce_counters(counters, group=None):
    """Sum per-rank termination counters (dict name -> int) over the group; identity without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(counters)
    names = sorted(counters)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([int(counters[k]) for k in names], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return {k: int(v) for k, v in zip(names, t.tolist())}


def max_over_ranks(value, group=None):
    """MAX of a python float over the group (device-timed milliseconds in bench.py)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_rows(local, group=None):
    """All-gather equally shaped per-rank row blocks [n_local, k] into [world * n_local, k] (rank order).
    This is the exchange step of the combat tasks (8-float records per aircraft)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out
