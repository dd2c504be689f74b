#!/usr/bin/env python
"""Build variants of the stage-derivative kernel (block size / register cap) and time them on the GPU.

    python tools/tune_eval.py build     # here (no GPU needed): compiles every variant into mpc-code_b200/_build/
    python tools/tune_eval.py run       # on the GPU box: times mpcb_stage_derivs for B = 4096
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = [(128, 2), (128, 3), (128, 4), (64, 4), (64, 6), (64, 8), (32, 12), (256, 1)]


def libs():
    import __graft_entry__ as e
    from mpc_code_b200.build import build_library
    prob, ss, ocp = e._problem()
    out = {}
    for blk, mb in VARIANTS:
        res = build_library("nmpc_cstr", prob, ss, ocp, extra_flags=("-DMPCB_EVAL_BLOCK=%d" % blk, "-DMPCB_EVAL_MINBLOCKS=%d" % mb))
        out[(blk, mb)] = res["so"]
    return prob, out


def run():
    import numpy as np
    import torch
    from mpc_code_b200.solvers import MpcbHandle, MpcbLibrary
    prob, so = libs()
    B = 4096
    n, m, N = prob.nx, prob.nu, prob.N
    rng = np.random.default_rng(0)
    w = np.tile(np.concatenate([np.tile(np.concatenate([prob.x0_m, prob.u0]), N), prob.x0_m]), (B, 1)) * (1 + 0.01 * rng.standard_normal((B, prob.nw)))
    par = np.tile(np.concatenate([prob.x0_m, prob.x0_m, prob.u0, prob.dhat0, prob.u0, [0.0], np.zeros(4), np.zeros(5 * N)]), (B, 1))
    lam = rng.standard_normal((B, N * n))
    res = {}
    for key, path in so.items():
        h = MpcbHandle(MpcbLibrary(path), B)
        args = [h.tensor(par, prob.npar), h.tensor(w, prob.nw), h.tensor(lam, N * n)]
        for _ in range(3):
            h.stage_derivs(*args)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            h.stage_derivs(*args)
        e1.record(); torch.cuda.synchronize()
        res["block%d_minblocks%d" % key] = e0.elapsed_time(e1) / 20
        print(key, "%.1f us" % (1e3 * res["block%d_minblocks%d" % key]), flush=True)
        h.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tune_eval.json"), "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        for k, v in libs()[1].items():
            print(k, v)
    else:
        run()
