#!/usr/bin/env python
"""BASELINE configs[4]: throughput of the OCP solve over the synthetic plant family.

    python tools/synthetic_sweep.py build            # here (no GPU): compile the libraries of the sweep
    python tools/synthetic_sweep.py run [out.txt]    # on the GPU box: OCP solves per second per (nx, nu, N, B)

Models up to 8 states use the symbolically generated derivative products, larger ones the looped dense products
(devicegen.DENSE_SH_ENTRIES).  B is capped by the workspace: batches whose workspace would exceed MEM_CAP bytes are solved
in chunks that reuse one workspace (sharding.solve_ocp_in_chunks) - that is how the 1M-instance point runs."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

SIZES = [(2, 1), (4, 2), (8, 3), (12, 4), (20, 6)]
HORIZONS = [20, 50, 200]
BATCHES = [1024, 8192, 65536]
MEM_CAP = 40e9


def ws_bytes(p):
    nz = p.nx + p.nu
    rec = nz * p.nx + p.nx + nz * nz + 7 * nz + 2 * (nz + 10) + 12
    return 8 * (p.N * (rec + 30 + 4 * p.nx) + 6 * p.nw + p.npar)


def main():
    mode = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else None

    def emit(line):
        print(line, flush=True)
        if out:
            out.write(line + "\n"); out.flush()
    if mode == "build":
        from mpc_code_b200.build import build_library
        for nx, nu in SIZES:
            for N in HORIZONS:
                name = "syn_%d_%d_%d" % (nx, nu, N)
                t0 = time.time()
                prob, ss, ocp = entry._problem(name)
                res = build_library(name, prob, ss, ocp)
                print("%-16s built in %6.1f s  dense=%d  %s" % (name, time.time() - t0, res["gen"]["defines"].get("MPCB_DENSE_SH", 0), os.path.basename(res["so"])), flush=True)
        return
    import torch
    from mpc_code_b200.mpc_loop import CompiledProblem
    from mpc_code_b200.sharding import solve_ocp_in_chunks
    emit("# OCP solves per second, synthetic family (x0 ~ U(-1,1)^nx, cold start, |u| <= 1); B200, device-driven solve")
    emit("%-4s %-3s %-4s %-8s %-8s %12s %10s %8s %8s" % ("nx", "nu", "N", "B", "chunk", "solves/s", "ms/batch", "solved", "iters"))
    points = [(nx, nu, N, B) for nx, nu in SIZES for N in HORIZONS for B in BATCHES]
    points.append((2, 1, 20, 1048576))              # the 1M-instance corner of configs[4]
    for nx, nu, N, B in points:
        name = "syn_%d_%d_%d" % (nx, nu, N)
        prob, ss, ocp = entry._problem(name)
        per = ws_bytes(prob)
        chunk = B
        while chunk * per > MEM_CAP:
            chunk //= 2
        if nx >= 12 and B > 8192:
            continue                                # minutes per point: the dense path is the functional one, not the fast one
        cp = CompiledProblem(prob, name)
        rng = np.random.default_rng(nx * 100 + N)
        x0 = rng.uniform(-1, 1, (B, nx))
        par = np.zeros((B, prob.npar)); par[:, :nx] = x0
        w0 = np.zeros((B, prob.nw))
        solve_ocp_in_chunks(cp, par[:min(B, chunk)], w0[:min(B, chunk)], chunk)          # warm-up (graph build, lazy loading)
        torch.cuda.synchronize(); t0 = time.time()
        W, F, ST, IT = solve_ocp_in_chunks(cp, par, w0, chunk)
        torch.cuda.synchronize(); dt = time.time() - t0
        emit("%-4d %-3d %-4d %-8d %-8d %12.0f %10.2f %8.4f %8.1f" % (nx, nu, N, B, chunk, B / dt, 1e3 * dt, float((ST == 0).mean()), float(IT.mean())))


if __name__ == "__main__":
    main()
