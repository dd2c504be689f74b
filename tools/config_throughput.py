#!/usr/bin/env python
"""Closed-loop throughput of the four BASELINE.json configurations on one GPU, fused step with
the built-in plant:  python tools/config_throughput.py [groups]
C1 Ex_LMPC_CSTR (single instance as shipped, and 4096 copies), C2 Ex_NMPC (4096), C3 Ex_LMPC_WB (16384),
C4 Ex_ENMPC/EKF (4096 = one GPU's share of 32768)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
from mpc_code_b200.mpc_loop import CompiledProblem

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rows = []
for name, B, steps, x0fun in (
        ("lmpc_cstr", 1, 100, None), ("lmpc_cstr", 4096, 100, None),
        ("nmpc_cstr", 4096, 60, lambda p, B: np.tile(p.x0_p, (B, 1)) * (1 + np.array([0.02, 0.002, 0.02]) * np.random.default_rng(1).uniform(-1, 1, (B, 3)))),
        ("lmpc_wb", 16384, 40, lambda p, B: 0.1 * np.random.default_rng(3).uniform(-1, 1, (B, p.nx))),
        ("enmpc_reactor", 4096, 21, lambda p, B: np.clip(np.array([0.9, 0.1]) + 0.05 * np.random.default_rng(4).uniform(-1, 1, (B, 2)), 0, 1))):
    prob, ss, ocp = entry._problem(name)
    ctl = CompiledProblem(prob, name).controller(B)
    x0 = None if x0fun is None else x0fun(prob, B)
    x0m = None if x0 is None else (np.tile(prob.x0_m, (B, 1)) if name == "enmpc_reactor" else x0)
    def run():
        if x0 is None: ctl.reset()
        else: ctl.reset(x0_p=x0, x0_m=x0m)
        ctl.h.set_groups(min(G, B))
        st, it = [], []
        torch.cuda.synchronize(); t0 = time.time()
        for k in range(steps):
            o = ctl.step_fused(); st.append(o["STATUS_DYN"].clone()); it.append(o["ITER_DYN"].clone())      # outputs are reused buffers
        torch.cuda.synchronize()
        return time.time() - t0, torch.stack(st), torch.stack(it)
    run()                                   # warm-up pass (same trajectory)
    dt, st, it = run()
    ok = float((st == 0).float().mean()); inf = float((st == 2).float().mean())
    rows.append((name, B, steps, B * steps / dt, 1e3 * dt / steps, ok, inf, float(it.float().mean())))
    print("%-14s B=%6d steps=%3d  %10.0f steps/s  %7.3f ms/step  solved %.4f infeasible %.4f  mean IPM iterations %.1f"
          % rows[-1], flush=True)
