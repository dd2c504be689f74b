"""BASELINE configs[4]: random stable nonlinear plants of several sizes (mpc_code_b200.synthetic).

Small members are checked live against the oracle, the 8- and 12-state members against committed fixtures (their
oracles need minutes of symbolic differentiation and compilation), the 12-state member also against an independent
single-shooting solution.  Horizons above 64 exercise the multi-round lane loops."""
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


def _single_shooting_reference(ns, xhat):
    """Independent solution of the synthetic OCP: the plant matrices from the namespace, RK4 in NumPy, the states
    eliminated (single shooting), exact gradients by the complex step, SciPy L-BFGS-B on |u| <= 1.  Shares no code
    with the package below the three matrices."""
    import scipy.optimize as sopt
    m_ = ns["_synthetic"]
    Ac, Bc, W, nx, nu, N = m_["Ac"], m_["Bc"], m_["W"], m_["nx"], m_["nu"], m_["N"]
    h, Mx = 0.1, 4

    def f(x, u):
        return Ac @ x + Bc @ u + 0.1 * np.tanh(W @ x)

    def rollout(U):
        x = xhat.astype(U.dtype); J = 0.0; X = [x]
        for k in range(N):
            u = U[k * nu:(k + 1) * nu]
            J = J + 0.5 * (x @ x + 0.1 * (u @ u))                # F_obj = 1/2 (x'Qx + u'Ru), Q = I, R = 0.1 I; no terminal cost
            hs = h / Mx
            for _ in range(Mx):
                k1 = f(x, u); k2 = f(x + 0.5 * hs * k1, u); k3 = f(x + 0.5 * hs * k2, u); k4 = f(x + hs * k3, u)
                x = x + hs / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
            X.append(x)
        return J, X

    def fun(U):
        J, _ = rollout(U)
        g = np.empty(U.size)
        for i in range(U.size):
            Uc = U.astype(complex); Uc[i] += 1e-30j
            g[i] = rollout(Uc)[0].imag / 1e-30
        return float(J), g
    res = sopt.minimize(fun, np.zeros(N * nu), jac=True, method="L-BFGS-B", bounds=[(-1.0, 1.0)] * (N * nu),
                        options=dict(maxiter=2000, ftol=1e-16, gtol=1e-11, maxcor=50))
    J, X = rollout(res.x)
    return res.x.reshape(N, nu), np.array(X), float(J)


def _twelve_state_case():
    b = _bundle("syn_12_4_20")
    p = b.prob
    xhat = np.random.default_rng(12).uniform(-1, 1, p.nx)
    par = b.ocp_par(xhat, np.zeros(p.nx), np.zeros(p.nu), np.zeros(0))
    return b, p, xhat, par


def test_twelve_state_member_matches_an_independent_single_shooting_solution():
    """12 states (generated second-order products of 15 000 operations, compiled out of line -
    codegen.BIG_FUNCTION_FLOPS): besides the oracle fixture, an independent NumPy / SciPy solution of the same OCP."""
    b, p, xhat, par = _twelve_state_case()
    w, f, st, it, _ = b.harness_ocp(par, np.zeros(p.nw))
    U, X, J = _single_shooting_reference(p.ns, xhat)
    nz = p.nx + p.nu
    Ud = w[0, :nz * p.N].reshape(p.N, nz)[:, p.nx:]
    Xd = np.vstack([w[0, :nz * p.N].reshape(p.N, nz)[:, :p.nx], w[0, nz * p.N:]])
    assert st[0] == 0
    assert np.abs(Ud - U).max() < 1e-5 and np.abs(Xd - X).max() < 1e-5
    assert abs(f[0] - J) <= 1e-8 * max(1.0, abs(J))
    assert np.abs(Ud).max() <= 1.0 + 1e-7


@pytest.mark.parametrize("name", ["syn_8_3_50", "syn_12_4_20"])
def test_larger_members_match_oracle_fixtures_on_cpu(name):
    """Fixtures: tests/golden/make_golden.py --only-synthetic=<name> (the 12-state oracle needs ORACLE_CFLAGS=-O0 and
    23 minutes of gcc)."""
    b = _bundle(name)
    par = SYN[name + "_par"]
    w, f, st, it, _ = b.harness_ocp(par, np.zeros((par.shape[0], b.prob.nw)))
    assert np.array_equal(st, SYN[name + "_status"]) and np.array_equal(it, SYN[name + "_iters"])
    assert np.abs(w - SYN[name + "_w"]).max() < 1e-9
    assert np.all(np.abs(f - SYN[name + "_f"]) <= 1e-10 * np.maximum(1.0, np.abs(SYN[name + "_f"])))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["syn_2_1_20", "syn_4_2_70", "syn_8_3_50", "syn_12_4_20"])
def test_gpu_matches_oracle(name):
    from mpc_code_b200.mpc_loop import CompiledProblem
    from mpc_code_b200.solvers import BatchedNlpSolver, MpcbHandle
    b = _bundle(name)
    p = b.prob
    if name in ("syn_8_3_50", "syn_12_4_20"):
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
def test_gpu_twelve_state_member():
    from mpc_code_b200.mpc_loop import CompiledProblem
    from mpc_code_b200.solvers import BatchedNlpSolver, MpcbHandle
    b, p, xhat, par = _twelve_state_case()
    U, X, J = _single_shooting_reference(p.ns, xhat)
    cp = CompiledProblem(p, "syn_12_4_20")
    h = MpcbHandle(cp.library, 2, dict(max_iter=100), dict(max_iter=100))
    solver = BatchedNlpSolver("ocp", cp.ocp_spec).attach(h)
    sol = solver(x0=np.zeros((2, p.nw)), p=np.tile(par, (2, 1)))
    w = sol["x"].cpu().numpy()
    nz = p.nx + p.nu
    assert np.all(solver.stats()["status"].cpu().numpy() == 0) and np.array_equal(w[0], w[1])
    assert np.abs(w[0, :nz * p.N].reshape(p.N, nz)[:, p.nx:] - U).max() < 1e-5
    assert abs(float(sol["f"][0]) - J) <= 1e-8 * max(1.0, abs(J))


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


# ---- large-state members: dense derivative products (devicegen.DENSE_SH_ENTRIES, MPCB_DENSE_SH) ---------------------
def _single_shooting_reference_batched(ns, xhat):
    """As `_single_shooting_reference`, with the complex-step gradient evaluated for all inputs at once."""
    import scipy.optimize as sopt
    m_ = ns["_synthetic"]
    Ac, Bc, W, nx, nu, N = m_["Ac"], m_["Bc"], m_["W"], m_["nx"], m_["nu"], m_["N"]
    h, Mx = 0.1, 4

    def f(x, u):                                   # x [..., nx], u [..., nu]
        return x @ Ac.T + u @ Bc.T + 0.1 * np.tanh(x @ W.T)

    def rollout(U):                                # U [..., N*nu] -> cost [...], states
        x = np.broadcast_to(xhat.astype(U.dtype), U.shape[:-1] + (nx,)); J = 0.0; X = [x]
        hs = h / Mx
        for k in range(N):
            u = U[..., k * nu:(k + 1) * nu]
            J = J + 0.5 * ((x * x).sum(-1) + 0.1 * (u * u).sum(-1))
            for _ in range(Mx):
                k1 = f(x, u); k2 = f(x + 0.5 * hs * k1, u); k3 = f(x + 0.5 * hs * k2, u); k4 = f(x + hs * k3, u)
                x = x + hs / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
            X.append(x)
        return J, X

    def fun(U):
        Uc = np.tile(U.astype(complex), (U.size, 1)) + 1e-30j * np.eye(U.size)
        g = rollout(Uc)[0].imag / 1e-30
        return float(rollout(U)[0]), g
    res = sopt.minimize(fun, np.zeros(N * nu), jac=True, method="L-BFGS-B", bounds=[(-1.0, 1.0)] * (N * nu),
                        options=dict(maxiter=5000, ftol=1e-16, gtol=1e-11, maxcor=60))
    J, X = rollout(res.x)
    return res.x.reshape(N, nu), np.array(X), float(J)


def test_dense_derivative_products_equal_the_generated_ones(monkeypatch):
    """The looped dense products used for large models against the symbolically generated ones (which the oracle pins)."""
    from conftest import Bundle
    b = _bundle("syn_4_2_70")
    par, w0 = _cases(b, 3)
    w, f, st, it, _ = b.harness_ocp(par, w0)
    monkeypatch.setenv("MPCB_DENSE_SH", "1")
    bd = Bundle("syn_4_2_70")
    assert bd.lib["gen"]["defines"]["MPCB_DENSE_SH"] == 1 and b.lib["gen"]["defines"]["MPCB_DENSE_SH"] == 0
    wd, fd, std, itd, _ = bd.harness_ocp(par, w0)
    assert np.array_equal(st, std) and np.array_equal(it, itd)
    assert np.abs(w - wd).max() < 1e-10 and np.all(np.abs(f - fd) <= 1e-11 * np.maximum(1.0, np.abs(f)))


def _twenty_state_case():
    b = _bundle("syn_20_6_20")
    p = b.prob
    xhat = np.random.default_rng(20).uniform(-1, 1, p.nx)
    return b, p, xhat, b.ocp_par(xhat, np.zeros(p.nx), np.zeros(p.nu), np.zeros(0))


def test_twenty_state_member_matches_an_independent_single_shooting_solution():
    """BASELINE configs[4] at its largest model size (20 states, 6 inputs): device code on the CPU."""
    b, p, xhat, par = _twenty_state_case()
    assert b.lib["gen"]["defines"]["MPCB_DENSE_SH"] == 1
    w, f, st, it, _ = b.harness_ocp(par, np.zeros(p.nw))
    U, X, J = _single_shooting_reference_batched(p.ns, xhat)
    nz = p.nx + p.nu
    Ud = w[0, :nz * p.N].reshape(p.N, nz)[:, p.nx:]
    assert st[0] == 0
    assert np.abs(Ud - U).max() < 1e-5 and abs(f[0] - J) <= 1e-8 * max(1.0, abs(J))
    assert np.abs(Ud).max() <= 1.0 + 1e-7


@pytest.mark.gpu
def test_gpu_twenty_state_member():
    from mpc_code_b200.mpc_loop import CompiledProblem
    from mpc_code_b200.solvers import BatchedNlpSolver, MpcbHandle
    b, p, xhat, par = _twenty_state_case()
    wc, fc, stc, itc, _ = b.harness_ocp(par, np.zeros(p.nw))
    cp = CompiledProblem(p, "syn_20_6_20")
    h = MpcbHandle(cp.library, 3, dict(max_iter=100), dict(max_iter=100))
    solver = BatchedNlpSolver("ocp", cp.ocp_spec).attach(h)
    sol = solver(x0=np.zeros((3, p.nw)), p=np.tile(par, (3, 1)))
    w = sol["x"].cpu().numpy()
    assert np.all(solver.stats()["status"].cpu().numpy() == 0) and np.array_equal(w[0], w[2])
    assert np.array_equal(solver.stats()["iter_count"].cpu().numpy(), np.repeat(itc, 3))
    assert np.abs(w[0] - wc[0]).max() < 1e-8 and abs(float(sol["f"][0]) - fc[0]) <= 1e-9 * max(1.0, abs(fc[0]))
    # closed loop through the fused step: estimator (20 x 20 covariance), target (50 x 50 KKT), OCP
    B = 64
    x0 = np.random.default_rng(21).uniform(-1, 1, (B, p.nx))
    ctl = cp.controller(B)
    ctl.reset(x0_p=x0, x0_m=x0)
    rec = ctl.run(4, fused=True)
    assert int((rec["STATUS_DYN"] != 0).sum()) == 0 and float(rec["U"].abs().max()) <= 1.0 + 1e-7
    assert float(rec["Xp"][-1].abs().max()) < float(rec["Xp"][0].abs().max())
