"""Edge cases of the OCP solve, device code (CPU harness build) against the oracle with the SAME options:
iteration limit, option variants (bound handling, barrier start, tolerance), a start on the bounds, identical
instances in a ragged batch."""
import numpy as np
import pytest

from oracle.ipm import IpmOptions
from oracle.nlp import OcpNlp


def _case(nmpc, scale=0.01, seed=0):
    p = nmpc.prob
    rng = np.random.default_rng(seed)
    xh = p.x0_m * (1 + scale * np.array([1.0, 0.1, 1.0]) * rng.uniform(-1, 1, p.nx))
    par = nmpc.ocp_par(xh, p.x0_m, p.u0, p.dhat0)
    lb, ub = nmpc.ocp.w_lb.copy(), nmpc.ocp.w_ub.copy(); lb[:p.nx] = ub[:p.nx] = xh
    return par, lb, ub


def test_iteration_limit_reports_maximum_iterations_exceeded(nmpc):
    par, lb, ub = _case(nmpc)
    w, f, st, it, _ = nmpc.harness_ocp(par, nmpc.cold_guess(), max_iter=3)
    r = OcpNlp(nmpc.ocp, nmpc.oracle).solve(nmpc.cold_guess(), par, lb, ub, opts=IpmOptions(max_iter=3))
    assert st[0] == -1 and r.status == -1 and it[0] == r.iters == 3
    assert np.abs(w[0] - r.x).max() < 1e-9                  # the same (unfinished) iterate


@pytest.mark.parametrize("kw,okw", [
    (dict(honor=1), dict(honor_original_bounds=True)),
    (dict(mu_init=1e-2), dict(mu_init=1e-2)),
    (dict(tol=1e-6), dict(tol=1e-6)),
    (dict(relax=0.0), dict(bound_relax_factor=0.0)),
])
def test_option_variants_follow_the_oracle(nmpc, kw, okw):
    par, lb, ub = _case(nmpc, seed=3)
    w, f, st, it, _ = nmpc.harness_ocp(par, nmpc.cold_guess(), **kw)
    r = OcpNlp(nmpc.ocp, nmpc.oracle).solve(nmpc.cold_guess(), par, lb, ub, opts=IpmOptions(max_iter=100, **okw))
    assert st[0] == r.status == 0 and it[0] == r.iters
    assert np.abs(w[0] - r.x).max() < 1e-6 and abs(f[0] - r.f) <= 1e-8 * max(1.0, abs(r.f))
    if "honor" in kw:                                       # IPOPT 3.12 behaviour: the solution is clipped into the original bounds
        assert np.all(w[0] >= lb - 1e-15) and np.all(w[0] <= ub + 1e-15)


def test_guess_on_the_bounds_is_pushed_inside(nmpc):
    """IPOPT's bound_push: a warm start sitting exactly on umin / xmax must give the same solve as the oracle's."""
    p = nmpc.prob
    par, lb, ub = _case(nmpc, seed=5)
    g = nmpc.cold_guess().copy()
    nz = p.nx + p.nu
    for k in range(p.N):
        g[k * nz + p.nx] = nmpc.ocp.bounds["umin"][0]       # every u_k[0] on its lower bound
        g[(k + 1) * nz + 2] = nmpc.ocp.bounds["xmax"][2]    # every level on its upper bound
    w, f, st, it, _ = nmpc.harness_ocp(par, g)
    r = OcpNlp(nmpc.ocp, nmpc.oracle).solve(g, par, lb, ub, opts=IpmOptions(max_iter=100))
    assert st[0] == r.status == 0 and it[0] == r.iters and np.abs(w[0] - r.x).max() < 1e-6


def test_ragged_batch_of_identical_and_distinct_instances(nmpc):
    """37 instances (not a multiple of any block size): identical inputs give bit-identical outputs wherever they sit
    in the batch, and a distinct one in the middle does not disturb its neighbours."""
    par, _, _ = _case(nmpc, seed=7)
    other, _, _ = _case(nmpc, scale=0.02, seed=8)
    P = np.tile(par, (37, 1)); P[17] = other
    w, f, st, it, _ = nmpc.harness_ocp(P, np.tile(nmpc.cold_guess(), (37, 1)))
    same = [i for i in range(37) if i != 17]
    assert np.all(st == 0)
    assert np.all(w[same] == w[0]) and np.all(f[same] == f[0]) and np.all(it[same] == it[0])
    w1, f1, _, _, _ = nmpc.harness_ocp(other, nmpc.cold_guess())
    assert np.array_equal(w[17], w1[0]) and f[17] == f1[0]


@pytest.mark.gpu
def test_gpu_single_instance_and_iteration_limit(nmpc):
    from mpc_code_b200.mpc_loop import CompiledProblem
    from mpc_code_b200.solvers import MpcbHandle, BatchedNlpSolver
    cp = CompiledProblem(nmpc.prob, "nmpc_cstr")
    par, lb, ub = _case(nmpc)
    on = OcpNlp(nmpc.ocp, nmpc.oracle)
    for max_iter, status in ((100, 0), (3, -1)):
        h = MpcbHandle(cp.library, 1, dict(max_iter=100), dict(max_iter=max_iter))
        solver = BatchedNlpSolver("ocp", cp.ocp_spec).attach(h)
        sol = solver(x0=nmpc.cold_guess()[None, :], p=par[None, :])
        r = on.solve(nmpc.cold_guess(), par, lb, ub, opts=IpmOptions(max_iter=max_iter))
        assert int(solver.stats()["status"][0]) == status == r.status
        assert np.abs(sol["x"].cpu().numpy()[0] - r.x).max() < 1e-6


def _enmpc_hard_start(enmpc):
    """Cold-started first OCP of Ex_ENMPC from the plant state of instance 3677 of the full-batch workload: the filter
    line search fails at iteration 12 (step limited to 3e-3 with the violation growing along it) - the restoration case."""
    p = enmpc.prob
    eps = np.random.default_rng(20240419 + 3677).uniform(-1, 1, p.nxp)
    return np.clip(np.array([0.9, 0.1]) + 0.05 * eps, 0.0, 1.0)


def test_failed_line_search_is_recovered_by_the_feasibility_restoration(enmpc):
    from harness_loop import HarnessLoop
    from oracle.closed_loop import OracleLoop
    x0 = _enmpc_hard_start(enmpc)
    p = enmpc.prob
    ref = OracleLoop(p, enmpc.ss, enmpc.ocp, enmpc.oracle).run(Nsim=2, x0_p=x0, x0_m=p.x0_m)
    rec = HarnessLoop(enmpc, 1).run(2, x0=x0[None, :], x0_m=p.x0_m[None, :])
    assert ref["STATUS_DYN"].tolist() == [0, 0] and rec["STATUS_DYN"][:, 0].tolist() == [0, 0]
    assert ref["ITER_DYN"][0] == rec["ITER_DYN"][0, 0] == 26          # 12 regular + 1 restoration + 13 regular iterations
    assert np.abs(rec["U"][:, 0, :] - ref["U"]).max() < 1e-6


@pytest.mark.gpu
def test_gpu_restoration_and_failure_policy(enmpc, nmpc):
    """(1) the restoration case on the device: same status / iterations / input as the oracle; (2) the hold policy: with
    an iteration limit of 3 every solve fails (-1): the reference applies those iterates, the hold policy keeps u0."""
    import torch
    from mpc_code_b200.mpc_loop import CompiledProblem
    from oracle.closed_loop import OracleLoop
    x0 = _enmpc_hard_start(enmpc)
    p = enmpc.prob
    ref = OracleLoop(p, enmpc.ss, enmpc.ocp, enmpc.oracle).run(Nsim=2, x0_p=x0, x0_m=p.x0_m)
    ctl = CompiledProblem(p, "enmpc_reactor").controller(4)
    ctl.reset(x0_p=np.tile(x0, (4, 1)), x0_m=np.tile(p.x0_m, (4, 1)))
    rec = ctl.run(2, fused=True)
    assert rec["STATUS_DYN"].cpu().numpy().tolist() == [[0] * 4] * 2
    assert rec["ITER_DYN"].cpu().numpy()[0].tolist() == [int(ref["ITER_DYN"][0])] * 4
    assert np.abs(rec["U"].cpu().numpy()[:, 0, :] - ref["U"]).max() < 1e-6
    pn = nmpc.prob
    cpn = CompiledProblem(pn, "nmpc_cstr")
    x0n = pn.x0_p * (1 + np.array([0.01, 0.001, 0.01]))
    for hold, fused in ((False, True), (True, True), (True, False)):
        c = cpn.controller(3, opts_dyn=dict(max_iter=3), hold_on_failure=hold)
        c.reset(x0_p=np.tile(x0n, (3, 1)), x0_m=np.tile(x0n, (3, 1)))
        r = c.run(2, fused=fused)
        assert (r["STATUS_DYN"] == -1).all()
        moved = (r["U"] - torch.as_tensor(pn.u0, device=r["U"].device)).abs().max().item() > 1e-9
        assert moved != hold


def _lp_case(b):
    xh = np.array([3.0, 2.5, 2.0]); xs = np.array([0.2, 0.1, 0.0]); us = np.array([0.5, -0.3])
    par = b.ocp_par(xh, xs, us, np.zeros(3))
    lb, ub = b.ocp.w_lb.copy(), b.ocp.w_ub.copy(); lb[:3] = ub[:3] = xh
    return par, lb, ub


def test_lp_form_costs_device_code_follows_the_oracle():
    """Stage cost r_x |x - xs| + r_u |u - us| (`r_x` / `r_u`, Utilities.py:341-352): `fabs` / `sign` through the code
    generator.  The problem is nonsmooth at its solution, so interior-point iterates are compared, not a converged point:
    the same unfinished iterate at iteration limits 3 and 6, and the same verdict (line search and restoration fail, status
    2 after 6 iterations) without a limit."""
    from conftest import _bundle
    b = _bundle("lmpc_cstr_lp")
    par, lb, ub = _lp_case(b)
    on = OcpNlp(b.ocp, b.oracle)
    for max_iter, status in ((3, -1), (6, -1), (100, 2)):
        w, f, st, it, _ = b.harness_ocp(par, b.cold_guess(), max_iter=max_iter)
        r = on.solve(b.cold_guess(), par, lb, ub, opts=IpmOptions(max_iter=max_iter))
        assert st[0] == r.status == status and it[0] == r.iters
        assert np.abs(w[0] - r.x).max() < 1e-9 and abs(f[0] - r.f) <= 1e-10 * max(1.0, abs(r.f))


@pytest.mark.gpu
@pytest.mark.parametrize("kw,okw", [
    (dict(honor_original_bounds=1), dict(honor_original_bounds=True)),
    (dict(mu_init=1e-2), dict(mu_init=1e-2)),
    (dict(tol=1e-6), dict(tol=1e-6)),
    (dict(bound_relax_factor=0.0), dict(bound_relax_factor=0.0)),
])
def test_gpu_option_variants_follow_the_oracle(nmpc, kw, okw):
    """The option variants of `test_option_variants_follow_the_oracle` on the device (IPOPT 3.12 / 3.14 bound handling,
    barrier start, tolerance, no bound relaxation)."""
    from mpc_code_b200.mpc_loop import CompiledProblem
    from mpc_code_b200.solvers import MpcbHandle, BatchedNlpSolver
    cp = CompiledProblem(nmpc.prob, "nmpc_cstr")
    par, lb, ub = _case(nmpc, seed=3)
    h = MpcbHandle(cp.library, 2, dict(max_iter=100), dict(max_iter=100, **kw))
    solver = BatchedNlpSolver("ocp", cp.ocp_spec).attach(h)
    sol = solver(x0=np.tile(nmpc.cold_guess(), (2, 1)), p=np.tile(par, (2, 1)))
    r = OcpNlp(nmpc.ocp, nmpc.oracle).solve(nmpc.cold_guess(), par, lb, ub, opts=IpmOptions(max_iter=100, **okw))
    w = sol["x"].cpu().numpy()
    assert int(solver.stats()["status"][0]) == r.status == 0 and int(solver.stats()["iter_count"][0]) == r.iters
    assert np.abs(w[0] - r.x).max() < 1e-6 and abs(float(sol["f"][0]) - r.f) <= 1e-8 * max(1.0, abs(r.f))
    if "honor_original_bounds" in kw:
        assert np.all(w[0] >= lb - 1e-15) and np.all(w[0] <= ub + 1e-15)


@pytest.mark.gpu
def test_gpu_lp_form_costs():
    from conftest import _bundle
    from mpc_code_b200.mpc_loop import CompiledProblem
    from mpc_code_b200.solvers import MpcbHandle, BatchedNlpSolver
    b = _bundle("lmpc_cstr_lp")
    par, lb, ub = _lp_case(b)
    cp = CompiledProblem(b.prob, "lmpc_cstr_lp")
    on = OcpNlp(b.ocp, b.oracle)
    for max_iter, status in ((6, -1), (100, 2)):
        h = MpcbHandle(cp.library, 2, dict(max_iter=100), dict(max_iter=max_iter))
        solver = BatchedNlpSolver("ocp", cp.ocp_spec).attach(h)
        sol = solver(x0=np.tile(b.cold_guess(), (2, 1)), p=np.tile(par, (2, 1)))
        r = on.solve(b.cold_guess(), par, lb, ub, opts=IpmOptions(max_iter=max_iter))
        assert int(solver.stats()["status"][0]) == r.status == status and int(solver.stats()["iter_count"][0]) == r.iters
        assert np.abs(sol["x"].cpu().numpy()[0] - r.x).max() < 1e-8
