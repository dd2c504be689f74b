"""TermCons = True (terminal equality X_N = x_s, Control_Calc.py:194-198): the Riccati sweep carries the terminal
multiplier (Pi / Gamma / Gramian recursions of csrc/mpcb_ocp.cuh).  Device code against the oracle, which states the
same rows as plain equality constraints of its dense KKT system."""
import numpy as np
import pytest

from conftest import _bundle
from harness_loop import HarnessLoop


@pytest.fixture(scope="module")
def lin():
    return _bundle("lmpc_cstr_termcons")


@pytest.fixture(scope="module")
def nl():
    return _bundle("nmpc_cstr_termcons")


def _oracle(b, Ns, x0=None):
    from oracle.closed_loop import OracleLoop
    kw = {} if x0 is None else dict(x0_p=x0, x0_m=x0)
    return OracleLoop(b.prob, b.ss, b.ocp, b.oracle).run(Nsim=Ns, **kw)


def _compare(rec, ref, tol=1e-6):
    assert np.array_equal(np.asarray(rec["STATUS_DYN"]).ravel(), np.asarray(ref["STATUS_DYN"]).ravel())
    assert np.array_equal(np.asarray(rec["ITER_DYN"]).ravel(), np.asarray(ref["ITER_DYN"]).ravel())
    for key in ("U", "X_HAT", "D_HAT", "XS", "US", "Xp"):
        diff = np.abs(np.asarray(rec[key]).reshape(np.asarray(ref[key]).shape) - np.asarray(ref[key])).max()
        assert diff < tol, (key, diff)
    f, fr = np.asarray(rec["F_DYN"]).ravel(), np.asarray(ref["F_DYN"]).ravel()
    assert np.all(np.abs(f - fr) <= 1e-8 * np.maximum(1.0, np.abs(fr)))


def test_row_layout(lin, nl):
    for b in (lin, nl):
        o = b.ocp
        assert b.prob.flags["TermCons"] is True and o.term_eq is not None
        assert o.ng == o.n * (o.N + 1) + o.n + (0 if o.yFree else o.p * o.N) + (0 if o.DuFree else o.m * o.N)   # Control_Calc.py:200-204


def test_linear_closed_loop_matches_oracle(lin):
    Ns = 24                                     # spans the end of the plant disturbance (t = 10) and the set-point change (t = 20)
    _compare({k: v[:, 0] for k, v in HarnessLoop(lin, 1).run(Ns).items()}, _oracle(lin, Ns))


def test_nonlinear_closed_loop_matches_oracle_and_pins_the_terminal_state(nl):
    p = nl.prob
    x0 = p.x0_p * (1 + np.array([0.01, 0.001, 0.01]))
    _compare({k: v[:, 0] for k, v in HarnessLoop(nl, 1).run(5, x0=x0[None, :]).items()}, _oracle(nl, 5, x0))
    par = nl.ocp_par(x0, p.x0_m, p.u0, p.dhat0)
    w, f, st, it, _ = nl.harness_ocp(par, nl.cold_guess())
    assert st[0] == 0 and np.abs(w[0, -p.nx:] - p.x0_m).max() < 1e-9
    free = _bundle("nmpc_cstr")
    w0, f0, _, _, _ = free.harness_ocp(free.ocp_par(x0, p.x0_m, p.u0, p.dhat0), free.cold_guess())
    assert np.abs(w0[0, -p.nx:] - p.x0_m).max() > 1e-6 and f[0] > f0[0]        # the constraint binds and costs something


@pytest.mark.gpu
@pytest.mark.parametrize("name,Ns", [("lmpc_cstr_termcons", 24), ("nmpc_cstr_termcons", 5)])
def test_gpu_closed_loop_matches_oracle(name, Ns):
    from mpc_code_b200.mpc_loop import CompiledProblem
    b = _bundle(name)
    ref = _oracle(b, Ns)
    ctl = CompiledProblem(b.prob, name).controller(3)
    rec = {k: v.cpu().numpy()[:, 0] for k, v in ctl.run(Ns, fused=True).items()}
    _compare(rec, ref)
