"""Instance sharding across ranks and the only collectives the path has (result gathering).

Instances are independent: rank ``r`` of ``G`` owns the contiguous block
``[r*ceil(B/G), min(B, (r+1)*ceil(B/G)))`` and nothing is exchanged while solving.  After a run the
per-rank trajectory slabs are all-gathered and solver statistics all-reduced (``torch.distributed``:
NCCL on GPUs, gloo in the CPU tests).  The reference has no counterpart (single process, one instance).
"""
from __future__ import annotations

from typing import Dict, Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    per = -(-total // world)
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def gather_instances(local, total: int):
    """All-gather ``[..., B_local, n]`` slabs (instance axis = -2) into ``[..., total, n]`` on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    per = -(-total // world)
    pad = per - local.shape[-2]
    if pad:
        local = torch.cat([local, local.new_zeros(*local.shape[:-2], pad, local.shape[-1])], dim=-2)
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous())
    return torch.cat(parts, dim=-2)[..., :total, :]


def reduce_stats(status, iters) -> Dict[str, float]:
    """Status histogram and iteration totals over all ranks (IPOPT status codes, see include/mpcb.h)."""
    import torch
    import torch.distributed as dist
    v = torch.stack([(status == 0).sum(), (status == 1).sum(), (status == 2).sum(), (status < 0).sum(),
                     iters.sum(), torch.as_tensor(status.numel(), device=status.device)]).to(torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(v)
    n = max(v[5].item(), 1.0)
    return {"solve_succeeded": v[0].item() / n, "acceptable": v[1].item() / n, "infeasible": v[2].item() / n,
            "failed": v[3].item() / n, "mean_ipm_iterations": v[4].item() / n, "instances": v[5].item()}
