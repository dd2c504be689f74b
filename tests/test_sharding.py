"""N > 1 host logic on the CPU: two gloo ranks shard instances, gather trajectories, reduce statistics."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from mpc_code_b200.sharding import shard_range


def test_shard_ranges_partition_the_batch():
    for total in (1, 7, 4096, 32768, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == -(-total // world)


def _worker(rank, world, port, total, out):
    sys.path.insert(0, ROOT)
    import mpc_code_b200  # noqa: F401
    from mpc_code_b200.sharding import gather_instances, reduce_stats, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    full = torch.arange(3 * total * 2, dtype=torch.float64).reshape(3, total, 2)      # [Nsim, B, n]
    g = gather_instances(full[:, lo:hi, :].clone(), total)
    status = torch.tensor([0] * (hi - lo - 1) + [2], dtype=torch.int32); iters = torch.full((hi - lo,), 9 + rank, dtype=torch.int32)
    stats = reduce_stats(status, iters)
    if rank == 0:
        torch.save(dict(ok=bool(torch.equal(g, full)), stats=stats), out)
    dist.destroy_process_group()


def test_two_ranks_gather_and_reduce(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    total, out = 7, str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, port, total, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["ok"]
    st = res["stats"]
    assert st["instances"] == 7 and abs(st["infeasible"] - 2 / 7) < 1e-12 and abs(st["mean_ipm_iterations"] - (4 * 9 + 3 * 10) / 7) < 1e-12
