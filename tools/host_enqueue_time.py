#!/usr/bin/env python
"""Host time spent inside one `step_fused` call (enqueue only, no synchronisation) against the number of instance groups:
how far the single enqueueing thread staggers the groups.  python tools/host_enqueue_time.py"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mpc_code_b200.mpc_loop import CompiledProblem

prob, ss, ocp = bench._problem()
B, K = bench.BATCH_PER_GPU, 40
cp = CompiledProblem(prob, "nmpc_cstr")
x0, noise = bench._workload(prob, B, K + 5)
for G in (1, 2, 4):
    ctl = cp.controller(B)
    ctl.reset(x0_p=x0, x0_m=x0)
    ctl.h.set_groups(G)
    nz = torch.as_tensor(noise, device=ctl.h.device)
    for k in range(5):
        ctl.step_fused(nz[k])
    torch.cuda.synchronize()
    host = []
    t0 = time.perf_counter()
    for k in range(5, 5 + K):
        a = time.perf_counter(); ctl.step_fused(nz[k]); host.append(time.perf_counter() - a)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("groups %d: host time per step_fused call median %.0f us (max %.0f), all %d calls enqueued after %.1f ms, GPU done after %.1f ms"
          % (G, 1e6 * np.median(host), 1e6 * max(host), K, 1e3 * (t1 - t0), 1e3 * (t2 - t0)), flush=True)
