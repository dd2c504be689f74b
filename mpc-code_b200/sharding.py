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


def solve_ocp_in_chunks(cp, par, w0, chunk: int, opts_dyn=None):
    """OCP solves for more instances than one solver workspace should hold (BASELINE configs[4] asks for batches up to 1M;
    the workspace is ~`OcpLayout::total` doubles per instance - 72 kB for Ex_NMPC, 2.4 MB for 20 states and N = 200):
    the batch is cut into chunks of ``chunk`` instances that reuse ONE workspace, queued back to back on the current
    stream.  ``par`` [B, npar] and ``w0`` [B, nw] may be host arrays (only one chunk at a time is resident) or device
    tensors.  Returns (w [B, nw], f [B], status [B], iters [B]) on the host."""
    import numpy as np
    import torch
    from .solvers import BatchedNlpSolver, MpcbHandle
    B = par.shape[0]
    chunk = min(int(chunk), B)
    h = MpcbHandle(cp.library, chunk, dict(max_iter=100), dict(max_iter=100, **(opts_dyn or {})))
    solver = BatchedNlpSolver("ocp", cp.ocp_spec).attach(h)
    W = np.empty((B, cp.ocp_spec.nw)); F = np.empty(B); ST = np.empty(B, dtype=np.int32); IT = np.empty(B, dtype=np.int32)
    for lo in range(0, B, chunk):
        hi = min(B, lo + chunk)
        p_c, w_c = par[lo:hi], w0[lo:hi]
        if hi - lo < chunk:                         # last, shorter chunk: pad with copies of its first instance
            pad = chunk - (hi - lo)
            p_c = np.concatenate([np.asarray(p_c), np.repeat(np.asarray(p_c[:1]), pad, 0)]) if not isinstance(p_c, torch.Tensor) else torch.cat([p_c, p_c[:1].expand(pad, -1)])
            w_c = np.concatenate([np.asarray(w_c), np.repeat(np.asarray(w_c[:1]), pad, 0)]) if not isinstance(w_c, torch.Tensor) else torch.cat([w_c, w_c[:1].expand(pad, -1)])
        sol = solver(x0=w_c, p=p_c)
        st = solver.stats()
        W[lo:hi] = sol["x"][:hi - lo].cpu().numpy(); F[lo:hi] = sol["f"][:hi - lo].cpu().numpy()
        ST[lo:hi] = st["status"][:hi - lo].cpu().numpy(); IT[lo:hi] = st["iter_count"][:hi - lo].cpu().numpy()
    h.close()
    return W, F, ST, IT
