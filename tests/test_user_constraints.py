"""User-defined stage inequalities `User_g_ineq(x,u,y,d,t,px,py) <= 0` (Control_Calc.py:94-96,132-137,244-245): extra
nonlinear range rows of every stage, after the Y_k and DU_k rows.  Device code against the oracle."""
import numpy as np
import pytest

from conftest import _bundle
from harness_loop import HarnessLoop


@pytest.fixture(scope="module")
def gin():
    return _bundle("nmpc_cstr_gineq")


def _x0(p):
    return p.x0_p * (1 + np.array([0.01, 0.001, 0.01]))


def _oracle(b, Ns, x0):
    from oracle.closed_loop import OracleLoop
    return OracleLoop(b.prob, b.ss, b.ocp, b.oracle).run(Nsim=Ns, x0_p=x0, x0_m=x0)


def _compare(rec, ref):
    assert np.array_equal(np.asarray(rec["STATUS_DYN"]).ravel(), np.asarray(ref["STATUS_DYN"]).ravel())
    assert np.array_equal(np.asarray(rec["ITER_DYN"]).ravel(), np.asarray(ref["ITER_DYN"]).ravel())
    for key in ("U", "X_HAT", "D_HAT", "XS", "US", "Xp"):
        diff = np.abs(np.asarray(rec[key]).reshape(np.asarray(ref[key]).shape) - np.asarray(ref[key])).max()
        assert diff < 1e-6, (key, diff)


def test_row_layout_and_bounds(gin):
    o = gin.ocp
    n_dyn, ny, ndu = o.n * (o.N + 1), (0 if o.yFree else o.p * o.N), (0 if o.DuFree else o.m * o.N)
    assert o.n_gin == 2 and o.ng == n_dyn + ny + ndu + 2 * o.N                     # g = [g, g1, g2, g4]  (Control_Calc.py:254)
    assert np.all(np.isneginf(o.g_lb[n_dyn + ny + ndu:])) and np.all(o.g_ub[n_dyn + ny + ndu:] == 0.0)
    lbg, ubg = gin.ocp_range_bounds()                                              # stage-interleaved [Y | DU | G]
    rows = (0 if o.yFree else o.p) + (0 if o.DuFree else o.m) + 2
    assert lbg.size == rows * o.N and np.all(np.isneginf(lbg.reshape(o.N, rows)[:, -2:]))


def test_closed_loop_matches_oracle_and_the_constraint_is_active(gin):
    p = gin.prob
    x0 = _x0(p)
    rec = {k: v[:, 0] for k, v in HarnessLoop(gin, 1).run(5, x0=x0[None, :]).items()}
    _compare(rec, _oracle(gin, 5, x0))
    free = {k: v[:, 0] for k, v in HarnessLoop(_bundle("nmpc_cstr"), 1).run(5, x0=x0[None, :]).items()}
    assert np.abs(rec["U"] - free["U"]).max() > 1e-2                               # it changes the optimal inputs
    par = gin.ocp_par(x0, p.x0_m, p.u0, p.dhat0)
    w, f, st, it, _ = gin.harness_ocp(par, gin.cold_guess())
    nz = p.nx + p.nu
    X = w[0, :nz * p.N].reshape(p.N, nz)
    duty = (X[:, 1] - X[:, 3]) * X[:, 2] - 16.8
    assert st[0] == 0 and duty.max() <= 1e-6 and duty.max() > -1e-3                # satisfied, and tight somewhere


@pytest.mark.gpu
def test_gpu_closed_loop_matches_oracle(gin):
    from mpc_code_b200.mpc_loop import CompiledProblem
    p = gin.prob
    x0 = _x0(p)
    ctl = CompiledProblem(p, "nmpc_cstr_gineq").controller(3)
    ctl.reset(x0_p=np.tile(x0, (3, 1)), x0_m=np.tile(x0, (3, 1)))
    rec = {k: v.cpu().numpy()[:, 0] for k, v in ctl.run(5, fused=True).items()}
    _compare(rec, _oracle(gin, 5, x0))
