#!/usr/bin/env python
"""Histogram of OCP return statuses over the full-length bench workload (Ex_NMPC, 4096 instances, 200 steps):
which codes occur, at which steps, and how many distinct instances are involved.  python tools/status_hist.py"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mpc_code_b200.mpc_loop import CompiledProblem

prob, ss, ocp = bench._problem()
B, K = bench.BATCH_PER_GPU, 200
x0, noise = bench._workload(prob, B, K)
cp = CompiledProblem(prob, "nmpc_cstr")


def run(hold):
    ctl = cp.controller(B, hold_on_failure=hold)
    ctl.reset(x0_p=x0, x0_m=x0)
    ctl.h.set_groups(bench.DEFAULT_GROUPS)
    nz = torch.as_tensor(noise, device=ctl.h.device)
    torch.cuda.synchronize(); t0 = time.time()
    rec = ctl.run(K, noise=nz, fused=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    st = rec["STATUS_DYN"].cpu().numpy(); it = rec["ITER_DYN"].cpu().numpy()
    codes, counts = np.unique(st, return_counts=True)
    print("hold_on_failure=%s: %d steps x %d instances in %.2f s = %.0f steps/s (wall clock, recording included)" % (hold, K, B, dt, K * B / dt))
    print("status histogram:", dict(zip(codes.tolist(), counts.tolist())), "of", st.size)
    for c in codes:
        if c == 0: continue
        k, i = np.nonzero(st == c)
        print("status %d: %d solves, %d distinct instances, steps %d..%d, iterations there: min %d median %d max %d"
              % (c, k.size, np.unique(i).size, k.min(), k.max(), it[k, i].min(), np.median(it[k, i]), it[k, i].max()))
        for j in np.unique(i)[:3]:
            ks = np.nonzero(st[:, j] == c)[0]
            print("   instance %d: steps %s, level estimate before: %s" % (j, ks[:8].tolist(), np.round(rec["X_HAT"][ks[:4], j, 2].cpu().numpy(), 5).tolist()))
    print("dead instances:", int(ctl.dead.sum()))


run(False)
run(True)
