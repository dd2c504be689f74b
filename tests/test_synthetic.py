"""BASELINE configs[4]: random stable nonlinear plants of several sizes (mpc_code_b200.synthetic).

Small members are checked live against the oracle; the 8-state member against a committed fixture (its oracle
needs ~8 minutes of symbolic differentiation).  Horizons above 64 exercise the multi-round lane loops."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, _bundle

SYN = np.load(os.path.join(GOLDEN, "synthetic_oracle.npz"))


def _cases(b, n):
    p = b.prob
    rng = np.random.default_rng(5)
    par = np.stack([b.ocp_par(rng.uniform(-1, 1, p.nx), np.zeros(p.nx), np.zeros(p.nu), np.zeros(0)) for _ in range(n)])
    return par, np.tile(np.zeros(p.nw), (n, 1))


@pytest.mark.parametrize("name", ["syn_2_1_20", "syn_4_2_70"])
def test_device_code_on_cpu_matches_oracle(name):
    from oracle.ipm import IpmOptions
    from oracle.nlp import OcpNlp
    b = _bundle(name)
    p = b.prob
    par, w0 = _cases(b, 3)
    w, f, st, it, _ = b.harness_ocp(par, w0)
    on = OcpNlp(b.ocp, b.oracle)
    for i in range(3):
        lb, ub = b.ocp.w_lb.copy(), b.ocp.w_ub.copy(); lb[:p.nx] = ub[:p.nx] = par[i, :p.nx]
        r = on.solve(w0[i], par[i], lb, ub, opts=IpmOptions(max_iter=100))
        assert r.status == st[i] == 0 and r.iters == it[i]
        assert np.abs(r.x - w[i]).max() < 1e-9 and abs(r.f - f[i]) <= 1e-10 * max(1.0, abs(r.f))


def test_eight_state_member_matches_fixture_on_cpu():
    b = _bundle("syn_8_3_50")
    par = SYN["syn_8_3_50_par"]
    w, f, st, it, _ = b.harness_ocp(par, np.zeros((par.shape[0], b.prob.nw)))
    assert np.array_equal(st, SYN["syn_8_3_50_status"]) and np.array_equal(it, SYN["syn_8_3_50_iters"])
    assert np.abs(w - SYN["syn_8_3_50_w"]).max() < 1e-9
    assert np.all(np.abs(f - SYN["syn_8_3_50_f"]) <= 1e-10 * np.maximum(1.0, np.abs(SYN["syn_8_3_50_f"])))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["syn_2_1_20", "syn_4_2_70", "syn_8_3_50"])
def test_gpu_matches_oracle(name):
    from mpc_code_b200.mpc_loop import CompiledProblem
    from mpc_code_b200.solvers import BatchedNlpSolver, MpcbHandle
    b = _bundle(name)
    p = b.prob
    if name == "syn_8_3_50":
        par = SYN[name + "_par"]; ref_w, ref_f, ref_it = SYN[name + "_w"], SYN[name + "_f"], SYN[name + "_iters"]
    else:
        from oracle.ipm import IpmOptions
        from oracle.nlp import OcpNlp
        par, _ = _cases(b, 3)
        on = OcpNlp(b.ocp, b.oracle)
        res = []
        for i in range(3):
            lb, ub = b.ocp.w_lb.copy(), b.ocp.w_ub.copy(); lb[:p.nx] = ub[:p.nx] = par[i, :p.nx]
            res.append(on.solve(np.zeros(p.nw), par[i], lb, ub, opts=IpmOptions(max_iter=100)))
        ref_w = np.array([r.x for r in res]); ref_f = np.array([r.f for r in res]); ref_it = np.array([r.iters for r in res])
    B = par.shape[0]
    cp = CompiledProblem(p, name)
    h = MpcbHandle(cp.library, B, dict(max_iter=100), dict(max_iter=100))
    solver = BatchedNlpSolver("ocp", cp.ocp_spec).attach(h)
    sol = solver(x0=np.zeros((B, p.nw)), p=par)
    assert np.all(solver.stats()["status"].cpu().numpy() == 0)
    assert np.array_equal(solver.stats()["iter_count"].cpu().numpy(), ref_it)
    assert np.abs(sol["x"].cpu().numpy() - ref_w).max() < 1e-6
    assert np.all(np.abs(sol["f"].cpu().numpy() - ref_f) <= 1e-8 * np.maximum(1.0, np.abs(ref_f)))


@pytest.mark.gpu
def test_closed_loop_large_batch_of_synthetic_plants():
    """1k instances of the 4-state plant, 12 closed-loop steps through the fused step: all solved, states regulated."""
    from mpc_code_b200.mpc_loop import CompiledProblem
    b = _bundle("syn_4_2_70")
    p = b.prob
    B = 1024
    x0 = np.random.default_rng(9).uniform(-1, 1, (B, p.nx))
    ctl = CompiledProblem(p, "syn_4_2_70").controller(B)
    ctl.reset(x0_p=x0, x0_m=x0)
    rec = ctl.run(12, fused=True)
    assert int((rec["STATUS_DYN"] != 0).sum()) == 0
    assert float(rec["Xp"][-1].abs().max()) < float(rec["Xp"][0].abs().max())
    assert float(rec["U"].abs().max()) <= 1.0 + 1e-7
