#!/usr/bin/env python
"""Histogram of OCP return statuses over the full-length bench workload (Ex_NMPC, 4096 instances, 200 steps):
which codes occur, at which steps, and how many distinct instances are involved.  python tools/status_hist.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mpc_code_b200.mpc_loop import CompiledProblem

prob, ss, ocp = bench._problem()
B, K = bench.BATCH_PER_GPU, 200
x0, noise = bench._workload(prob, B, K)
ctl = CompiledProblem(prob, "nmpc_cstr").controller(B)
ctl.reset(x0_p=x0, x0_m=x0)
ctl.h.set_groups(8)
rec = ctl.run(K, noise=torch.as_tensor(noise, device=ctl.h.device), fused=True)
st = rec["STATUS_DYN"].cpu().numpy(); it = rec["ITER_DYN"].cpu().numpy()
codes, counts = np.unique(st, return_counts=True)
print("status histogram:", dict(zip(codes.tolist(), counts.tolist())), "of", st.size)
for c in codes:
    if c == 0: continue
    k, i = np.nonzero(st == c)
    print("status %d: %d solves, %d distinct instances, steps %d..%d, iterations there: min %d median %d max %d"
          % (c, k.size, np.unique(i).size, k.min(), k.max(), it[k, i].min(), np.median(it[k, i]), it[k, i].max()))
    first = np.unique(i)[:3]
    for j in first:
        ks = np.nonzero(st[:, j] == c)[0]
        print("   instance %d: steps %s, level estimate before: %s" % (j, ks[:8].tolist(), np.round(rec["X_HAT"][ks[:4], j, 2].cpu().numpy(), 5).tolist()))
print("dead instances:", int(ctl.dead.sum()))
