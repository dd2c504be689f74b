"""An anchor for the oracle that shares NOTHING with the product's host code: no `sx.py` tracer, no `OcpSpec`, no
generated C.  The Ex_NMPC optimal-control problem (Control_Calc.py:20-260 for the constants of Ex_NMPC.py:114-150,
217-246) is written out by hand in NumPy - CSTR right-hand side, classic RK4 with 10 sub-steps, quadratic stage cost -
and differentiated by the complex-step method.  Two checks per fixture instance, full horizon N = 50:

1. KKT certificate: at the ORACLE's solution the hand-written problem is feasible to 1e-8 and stationary - multipliers
   fitted by least squares leave a residual below 1e-6 of the gradient scale, with the right signs on the active bounds.
2. Independent solve: SciPy SLSQP on the multiple-shooting form of the hand-written problem, started from the cold
   guess, reaches the same cost and input trajectory.

A wrong cost term, bound, parameter offset or AD rule in the host package would shift the oracle's solution away from
this problem's KKT point and fail here (the device-vs-oracle tests cannot see such an error: both sides inherit it)."""
import os

import numpy as np
import pytest
import scipy.optimize as so

from conftest import GOLDEN

G = np.load(os.path.join(GOLDEN, "nmpc_oracle.npz"))
N, MX, H, NX, NU = 50, 10, 0.2, 3, 2
T0, C0, RAD, K0, EOR = 350.0, 1.0, 0.219, 7.2e10, 8750.0
U0C, RHO, CP, DH = 915.6 * 60 / 1000, 1000.0, 0.239, -5.0e4
AREA = np.pi * RAD ** 2
QW, RW = np.eye(3), 0.1 * np.eye(2)
XMIN, XMAX = np.array([0.0, 315.0, 0.50]), np.array([1.0, 375.0, 0.75])
UMIN, UMAX = np.array([295.0, 0.0]), np.array([305.0, 0.25])
YMIN, YMAX = np.array([0.0, 0.5]), np.array([1.0, 1.0])


def rhs(x, u, F0):
    c, T, lvl = x
    k = K0 * np.exp(-EOR / T)
    return np.array([F0 * (C0 - c) / (AREA * lvl) - k * c,
                     F0 * (T0 - T) / (AREA * lvl) - DH / (RHO * CP) * k * c + 2 * U0C / (RAD * RHO * CP) * (u[0] - T),
                     (F0 - u[1]) / AREA])


def step(x, u, F0):
    hs = H / MX
    for _ in range(MX):
        k1 = rhs(x, u, F0); k2 = rhs(x + 0.5 * hs * k1, u, F0); k3 = rhs(x + 0.5 * hs * k2, u, F0); k4 = rhs(x + hs * k3, u, F0)
        x = x + hs / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    return x


def unpack(par):
    return dict(x0=par[0:3], xs=par[3:6], us=par[6:8], F0=par[9])        # par = [x0|xs|us|d|um1|t|lam|px|py], F0 = d[1]


def cost_and_defects(w, P):
    """w = [x0,u0,...,xN] (Control_Calc.py:31-37).  Returns the cost and the N x 3 multiple-shooting defects (stage-major).
    All stages are integrated at once: `rhs` / `step` act on 3 x N arrays."""
    Z = w[:5 * N].reshape(N, 5)
    X, U = Z[:, :3].T, Z[:, 3:].T
    Xnext = np.concatenate([X[:, 1:], w[5 * N:].reshape(3, 1)], axis=1)
    dx, du = X - P["xs"][:, None], U - P["us"][:, None]
    f = 0.5 * (np.sum(dx * (QW @ dx)) + np.sum(du * (RW @ du)))
    return f, (step(X, U, P["F0"]) - Xnext).T.ravel()


def complex_step_jac(fun, w, m):
    J = np.zeros((m, w.size))
    for i in range(w.size):
        wc = w.astype(complex); wc[i] += 1e-30j
        J[:, i] = np.imag(fun(wc)) / 1e-30
    return J


@pytest.mark.parametrize("case", [0, 2, 5])
def test_oracle_solution_is_a_kkt_point_of_the_hand_written_problem(case):
    par, w = G["ocp_par"][case], G["ocp_w"][case]
    assert G["ocp_status"][case] == 0
    P = unpack(par)
    f, c = cost_and_defects(w, P)
    assert abs(f - G["ocp_f"][case]) <= 1e-9 * max(1.0, abs(f))                      # the same objective value
    assert np.abs(c).max() < 1e-8 and np.abs(w[:3] - P["x0"]).max() < 1e-12          # feasible: dynamics, x0
    lb = np.concatenate([np.concatenate([XMIN, UMIN])] * N + [XMIN]); ub = np.concatenate([np.concatenate([XMAX, UMAX])] * N + [XMAX])
    relax = 1e-8 * np.maximum(1.0, np.abs(np.concatenate([lb, ub]))).max()
    assert np.all(w >= lb - 2 * relax) and np.all(w <= ub + 2 * relax)
    Y = np.stack([[w[5 * k], w[5 * k + 2]] for k in range(N)])                       # y = (c, level), Ex_NMPC.py:160-166
    assert np.all(Y >= YMIN - 2e-8) and np.all(Y <= YMAX + 2e-8)
    g = complex_step_jac(lambda z: np.array([cost_and_defects(z, P)[0]]), w, 1)[0]
    J = complex_step_jac(lambda z: cost_and_defects(z, P)[1], w, 3 * N)
    free = np.arange(3, w.size)                                                       # x0 is fixed
    act_lo = [i for i in free if w[i] - lb[i] < 1e-5 * max(1.0, abs(lb[i]))]
    act_hi = [i for i in free if ub[i] - w[i] < 1e-5 * max(1.0, abs(ub[i]))]
    # the level rows of Y_k duplicate the bound x_k[2] >= 0.5: their multipliers merge with the bound multipliers here
    E = np.zeros((w.size, len(act_lo) + len(act_hi)))
    for j, i in enumerate(act_lo):
        E[i, j] = -1.0
    for j, i in enumerate(act_hi):
        E[i, len(act_lo) + j] = 1.0
    A = np.hstack([J.T, E])[free]
    sol, *_ = np.linalg.lstsq(A, -g[free], rcond=None)
    res = A @ sol + g[free]
    assert np.abs(res).max() <= 1e-6 * max(1.0, np.abs(g).max()), np.abs(res).max()
    z = sol[3 * N:]
    if z.size:
        assert np.all(z >= -1e-6 * max(1.0, np.abs(z).max()))                         # multipliers of active bounds push inwards


def _stage_jacs(w, F0):
    """d step / d (x_k, u_k) of every stage by the complex-step method: N x 3 x 5."""
    Z = w[:5 * N].reshape(N, 5)
    J = np.zeros((N, NX, NX + NU))
    for i in range(NX + NU):
        Zc = Z.astype(complex); Zc[:, i] += 1e-30j
        J[:, :, i] = (np.imag(step(Zc[:, :3].T, Zc[:, 3:].T, F0)) / 1e-30).T
    return J


@pytest.mark.parametrize("case", [2, 5])
def test_slsqp_on_the_hand_written_problem_reaches_the_oracle_solution(case):
    """SciPy's SQP on the multiple-shooting form of the hand-written problem, from the cold guess of the fixture."""
    par, w_or = G["ocp_par"][case], G["ocp_w"][case]
    P = unpack(par)
    nw = w_or.size

    def cost(w):
        return float(cost_and_defects(w, P)[0])

    def cost_grad(w):
        g = np.zeros(nw)
        for k in range(N):
            g[5 * k:5 * k + 3] = QW @ (w[5 * k:5 * k + 3] - P["xs"]); g[5 * k + 3:5 * k + 5] = RW @ (w[5 * k + 3:5 * k + 5] - P["us"])
        return g

    def defects(w):
        return np.concatenate([w[:3] - P["x0"], cost_and_defects(w, P)[1]])

    def defects_jac(w):
        J = np.zeros((3 * (N + 1), nw))
        J[:3, :3] = np.eye(3)
        Jk = _stage_jacs(w, P["F0"])
        for k in range(N):
            J[3 + 3 * k:6 + 3 * k, 5 * k:5 * k + 5] = Jk[k]
            J[3 + 3 * k:6 + 3 * k, 5 * (k + 1):5 * (k + 1) + 3] -= np.eye(3)
        return J

    lb = np.concatenate([np.concatenate([XMIN, UMIN])] * N + [XMIN]); ub = np.concatenate([np.concatenate([XMAX, UMAX])] * N + [XMAX])
    r = so.minimize(cost, G["ocp_w0"], jac=cost_grad, method="SLSQP", bounds=list(zip(lb, ub)),
                    constraints=[dict(type="eq", fun=defects, jac=defects_jac)], options=dict(maxiter=200, ftol=1e-14))
    # SLSQP (quasi-Newton, 253 variables) is still creeping towards the minimiser when it stops: the tolerances below are
    # ITS accuracy after 200 iterations, not the oracle's - the sharp statement is the KKT certificate above
    assert np.abs(defects(r.x)).max() < 1e-7
    assert 0.0 <= r.fun - G["ocp_f"][case] <= 2e-6 * max(1.0, abs(r.fun)), (r.fun, G["ocp_f"][case])
    u_or = np.stack([w_or[5 * k + 3:5 * k + 5] for k in range(N)]); u_sq = np.stack([r.x[5 * k + 3:5 * k + 5] for k in range(N)])
    print("SLSQP: nit", r.nit, "cost", r.fun, "oracle", G["ocp_f"][case], "max |du|", np.abs(u_sq - u_or).max())
    assert np.abs(u_sq - u_or).max() < 5e-5
