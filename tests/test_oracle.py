"""The CPU oracle against its committed fixtures and against independent SciPy solvers."""
import os

import numpy as np
import pytest
import scipy.optimize as sopt

from conftest import GOLDEN
from oracle.ipm import IpmOptions, solve_nlp
from oracle.nlp import OcpNlp, TargetNlp

G = np.load(os.path.join(GOLDEN, "nmpc_oracle.npz"))


def test_oracle_reproduces_committed_ocp_fixtures(nmpc):
    on = OcpNlp(nmpc.ocp, nmpc.oracle)
    n = nmpc.prob.nx
    for b in (0, 1, 5):
        lb, ub = nmpc.ocp.w_lb.copy(), nmpc.ocp.w_ub.copy()
        lb[:n] = ub[:n] = G["ocp_par"][b, :n]
        r = on.solve(G["ocp_w0"], G["ocp_par"][b], lb, ub, opts=IpmOptions(max_iter=100))
        assert r.status == G["ocp_status"][b] and r.iters == G["ocp_iters"][b]
        assert np.abs(r.x - G["ocp_w"][b]).max() < 1e-9
        assert abs(r.f - G["ocp_f"][b]) <= 1e-10 * max(1.0, abs(G["ocp_f"][b]))


def test_oracle_kkt_point_satisfies_first_order_conditions(nmpc):
    """Independent of the solver's own bookkeeping: re-evaluate the NLP at the fixture solution."""
    on = OcpNlp(nmpc.ocp, nmpc.oracle)
    b = 3
    w = G["ocp_w"][b]
    fun = on.make_fun(G["ocp_par"][b])
    r = fun(w, np.zeros(on.m_total), 0)
    n, N = nmpc.prob.nx, nmpc.prob.N
    assert np.abs(r["g"][:n * (N + 1)]).max() < 1e-8                         # dynamics rows
    assert np.all(w[n:] >= nmpc.ocp.w_lb[n:] - 1e-7) and np.all(w[n:] <= nmpc.ocp.w_ub[n:] + 1e-7)
    yrows = r["g"][n * (N + 1):]
    assert np.all(yrows >= nmpc.ocp.g_lb[n * (N + 1):] - 1e-7) and np.all(yrows <= nmpc.ocp.g_ub[n * (N + 1):] + 1e-7)


def test_target_solution_matches_scipy_slsqp(nmpc):
    """SLSQP (a different algorithm and code base) must land on the oracle's target solution."""
    tn = TargetNlp(nmpc.ss, nmpc.oracle)
    for b in (1, 3):
        par = G["ss_par"][b]
        fun = tn.make_fun(par)
        res = sopt.minimize(lambda w: fun(w, np.zeros(5), 1)["f"], G["ss_w0"][b], jac=lambda w: fun(w, np.zeros(5), 1)["grad"],
                            constraints=[dict(type="eq", fun=lambda w: fun(w, np.zeros(5), 0)["g"],
                                              jac=lambda w: fun(w, np.zeros(5), 2)["J"])],
                            bounds=list(zip(nmpc.ss.w_lb, nmpc.ss.w_ub)), method="SLSQP", options=dict(ftol=1e-15, maxiter=500))
        assert res.success
        assert np.abs(res.x - G["ss_w"][b]).max() < 5e-6


def test_short_horizon_ocp_matches_scipy_trust_constr(nmpc):
    """Same NLP restricted to its first 4 stages (the rest of w frozen) solved by trust-constr."""
    on = OcpNlp(nmpc.ocp, nmpc.oracle)
    p = nmpc.prob
    n, m, N = p.nx, p.nu, p.N
    nxu = n + m
    b = 1
    par = G["ocp_par"][b]
    wstar = G["ocp_w"][b]
    K = 4
    free = np.arange(n, nxu * K)                      # u_0, x_1, ..., u_{K-1}; x_K and beyond frozen at the optimum
    fun = on.make_fun(par)
    lb, ub = nmpc.ocp.w_lb.copy(), nmpc.ocp.w_ub.copy()

    def full(v):
        w = wstar.copy(); w[free] = v
        return w
    rows = np.arange(n, n * (K + 1))                  # dynamics rows of stages 0..K-1
    cons = sopt.NonlinearConstraint(lambda v: fun(full(v), np.zeros(on.m_total), 0)["g"][rows], 0.0, 0.0,
                                    jac=lambda v: fun(full(v), np.zeros(on.m_total), 2)["J"][np.ix_(rows, free)])
    x0 = np.clip(wstar[free] * (1 + 1e-3), lb[free], ub[free])
    res = sopt.minimize(lambda v: fun(full(v), np.zeros(on.m_total), 1)["f"], x0,
                        jac=lambda v: fun(full(v), np.zeros(on.m_total), 1)["grad"][free],
                        constraints=[cons], bounds=sopt.Bounds(lb[free], ub[free]), method="trust-constr",
                        options=dict(gtol=1e-10, xtol=1e-12, maxiter=2000))
    assert np.abs(res.x - wstar[free]).max() < 2e-5


def test_ipm_on_a_textbook_problem():
    """Hock-Schittkowski 71: known solution (1, 4.743, 3.8211, 1.3794), f = 17.0140173."""
    def fun(x, lam, need):
        f = x[0] * x[3] * (x[0] + x[1] + x[2]) + x[2]
        g = np.array([x[0] * x[1] * x[2] * x[3], np.sum(x ** 2)])
        out = dict(f=f, g=g)
        if need >= 1:
            out["grad"] = np.array([x[3] * (2 * x[0] + x[1] + x[2]), x[0] * x[3], x[0] * x[3] + 1, x[0] * (x[0] + x[1] + x[2])])
        if need >= 2:
            out["J"] = np.array([[x[1] * x[2] * x[3], x[0] * x[2] * x[3], x[0] * x[1] * x[3], x[0] * x[1] * x[2]], 2 * x])
            H = np.zeros((4, 4))
            H[0, 0] = 2 * x[3]; H[0, 1] = H[1, 0] = x[3]; H[0, 2] = H[2, 0] = x[3]
            H[0, 3] = H[3, 0] = 2 * x[0] + x[1] + x[2]; H[1, 3] = H[3, 1] = x[0]; H[2, 3] = H[3, 2] = x[0]
            prod = lambda i, j: np.prod([x[k] for k in range(4) if k not in (i, j)])  # noqa: E731
            for i in range(4):
                for j in range(4):
                    if i != j:
                        H[i, j] += lam[0] * prod(i, j)
            H += lam[1] * 2 * np.eye(4)
            out["H"] = H
        return out
    r = solve_nlp(4, 2, fun, [1, 5, 5, 1], [1] * 4, [5] * 4, [25, 40], [np.inf, 40])
    assert r.status == 0
    assert np.allclose(r.x, [1.0, 4.74299963, 3.82114998, 1.37940829], atol=1e-6) and abs(r.f - 17.0140173) < 1e-6


def test_infeasible_initial_output_is_reported_like_ipopt(nmpc):
    """Quirk D7: the Y_0 range row only involves the fixed x_0; outside its bounds the OCP is infeasible."""
    on = OcpNlp(nmpc.ocp, nmpc.oracle)
    p = nmpc.prob
    xhat = p.x0_m.copy(); xhat[2] = 0.4995            # level below ymin = 0.5
    par = nmpc.ocp_par(xhat, p.x0_m, p.u0, p.dhat0)
    lb, ub = nmpc.ocp.w_lb.copy(), nmpc.ocp.w_ub.copy(); lb[:3] = ub[:3] = xhat
    r = on.solve(nmpc.cold_guess(), par, lb, ub, opts=IpmOptions(max_iter=100))
    assert r.return_status == "Infeasible_Problem_Detected"
