"""User constraints of the steady-state target problem: `User_g_ineq_SS(xs,us,ys,d,t,px,py) <= 0` and
`User_h_eq_SS(...) = 0` (Target_Calc.py:87-109, bounds :146-150).  The device solver carries one slack variable per
inequality row (g - s = 0, s in (-inf, 0]), as IPOPT and the oracle do."""
import numpy as np
import pytest

from conftest import _bundle
from harness_loop import HarnessLoop


@pytest.fixture(scope="module")
def ssu():
    return _bundle("nmpc_cstr_ss_user")


def _par(b, d=None, usprev=None):
    p, ss = b.prob, b.ss
    off = ss.off
    par = np.zeros(ss.npar)
    ysp, usp, xsp = p.defSP(0.0)
    par[off["usp"]:off["usp"] + p.nu] = usp
    par[off["ysp"]:off["ysp"] + p.ny] = ysp
    par[off["d"]:off["d"] + p.nd] = p.dhat0 if d is None else d
    par[off["usprev"]:off["usprev"] + p.nu] = p.u0 if usprev is None else usprev
    return par


def _guess(b):
    p = b.prob
    return np.concatenate([p.x0_m, p.u0, [p.x0_m[0], p.x0_m[2]]])


def test_row_layout(ssu):
    ss, p = ssu.ss, ssu.prob
    assert ss.ng_ss == 1 and ss.nh_ss == 1
    n_model = p.nx + p.ny
    assert ss.g_lb.size == n_model + 2                                   # g = [model rows, g_SS, h_SS]  (Target_Calc.py:152)
    assert np.isneginf(ss.g_lb[n_model]) and ss.g_ub[n_model] == 0.0     # (:149-150)
    assert ss.g_lb[n_model + 1] == 0.0 and ss.g_ub[n_model + 1] == 0.0
    assert np.all(ss.g_lb[:n_model] == 0.0) and np.all(ss.g_ub[:n_model] == 0.0)


def test_oracle_target_against_slsqp(ssu):
    """Independent check of the oracle on this NLP: SciPy's SLSQP on the same functions."""
    from scipy.optimize import minimize
    from oracle.nlp import TargetNlp
    ss = ssu.ss
    par, w0 = _par(ssu), _guess(ssu)
    r = TargetNlp(ss, ssu.oracle).solve(w0, par)
    assert r.status == 0
    fun = TargetNlp(ss, ssu.oracle).make_fun(par)
    n_eq = ssu.prob.nx + ssu.prob.ny
    eq_rows = list(range(n_eq)) + [n_eq + 1]
    scale = np.maximum(1.0, np.abs(w0))
    res = minimize(lambda z: fun(z * scale, np.zeros(ss.g_lb.size), 0)["f"], w0 / scale, method="SLSQP",
                   bounds=list(zip(ss.w_lb / scale, ss.w_ub / scale)),
                   constraints=[dict(type="eq", fun=lambda z: fun(z * scale, np.zeros(ss.g_lb.size), 0)["g"][eq_rows]),
                                dict(type="ineq", fun=lambda z: -fun(z * scale, np.zeros(ss.g_lb.size), 0)["g"][n_eq])],
                   options=dict(ftol=1e-14, maxiter=500))
    assert res.success
    assert np.abs(res.x * scale - r.x).max() < 2e-4 * np.abs(r.x).max()  # SLSQP's accuracy; IPOPT-style interior solution
    assert abs(res.fun - r.f) < 1e-6 * max(1.0, abs(r.f))
    x = r.x
    assert abs((x[1] - x[3]) * x[2] - 15.9) < 1e-5                       # the duty limit is active at the target
    assert abs(x[6] - 0.6528 - 5.0 * (x[4] ** 2 - 0.01)) < 1e-9
    free = _bundle("nmpc_cstr")
    rf = TargetNlp(free.ss, free.oracle).solve(w0, par)
    assert np.abs(rf.x - r.x).max() > 1e-2                               # and it moves the target


def test_device_target_matches_oracle(ssu):
    from oracle.nlp import TargetNlp
    p = ssu.prob
    rng = np.random.default_rng(3)
    pars, refs = [], []
    for i in range(6):
        d = p.dhat0 * (1.0 + (0.3 * rng.uniform(-1, 1, p.nd) if i else 0.0))
        par = _par(ssu, d=d)
        pars.append(par)
        refs.append(TargetNlp(ssu.ss, ssu.oracle).solve(_guess(ssu), par))
    w, f, st, it = ssu.harness_target(np.array(pars), np.tile(_guess(ssu), (len(pars), 1)))
    for i, r in enumerate(refs):
        assert st[i] == r.status and it[i] == r.iters, (i, st[i], r.status, it[i], r.iters)
        assert np.abs(w[i] - r.x).max() < 1e-9 * np.abs(r.x).max()
        assert abs(f[i] - r.f) < 1e-10 * max(1.0, abs(r.f))


def _compare(rec, ref):
    assert np.array_equal(np.asarray(rec["STATUS_DYN"]).ravel(), np.asarray(ref["STATUS_DYN"]).ravel())
    assert np.array_equal(np.asarray(rec["ITER_DYN"]).ravel(), np.asarray(ref["ITER_DYN"]).ravel())
    for key in ("U", "X_HAT", "D_HAT", "XS", "US", "Xp"):
        diff = np.abs(np.asarray(rec[key]).reshape(np.asarray(ref[key]).shape) - np.asarray(ref[key])).max()
        assert diff < 1e-6, (key, diff)


def _x0(p):
    return p.x0_p * (1 + np.array([0.01, 0.001, 0.01]))


def _oracle(b, Ns, x0):
    from oracle.closed_loop import OracleLoop
    return OracleLoop(b.prob, b.ss, b.ocp, b.oracle).run(Nsim=Ns, x0_p=x0, x0_m=x0)


def test_closed_loop_matches_oracle(ssu):
    x0 = _x0(ssu.prob)
    rec = {k: v[:, 0] for k, v in HarnessLoop(ssu, 1).run(4, x0=x0[None, :]).items()}
    ref = _oracle(ssu, 4, x0)
    _compare(rec, ref)
    xs, us = np.asarray(ref["XS"]).reshape(4, -1), np.asarray(ref["US"]).reshape(4, -1)
    assert np.all((xs[:, 1] - us[:, 0]) * xs[:, 2] - 15.9 < 1e-5)        # every target respects the duty limit


@pytest.mark.gpu
def test_gpu_closed_loop_matches_oracle(ssu):
    from mpc_code_b200.mpc_loop import CompiledProblem
    p = ssu.prob
    x0 = _x0(p)
    ctl = CompiledProblem(p, "nmpc_cstr_ss_user").controller(3)
    ctl.reset(x0_p=np.tile(x0, (3, 1)), x0_m=np.tile(x0, (3, 1)))
    rec = {k: v.cpu().numpy()[:, 0] for k, v in ctl.run(5, fused=True).items()}
    _compare(rec, _oracle(ssu, 5, x0))
    # and through the reference-style call solver_ss(lbx=, ubx=, x0=, p=, lbg=, ubg=) with the extended g bounds
    from oracle.nlp import TargetNlp
    ss = ssu.ss
    par, w0 = np.tile(_par(ssu), (3, 1)), np.tile(_guess(ssu), (3, 1))
    sol = ctl.solver_ss(lbx=ss.w_lb, ubx=ss.w_ub, x0=w0, p=par, lbg=ss.g_lb, ubg=ss.g_ub)
    r = TargetNlp(ss, ssu.oracle).solve(_guess(ssu), _par(ssu))
    assert np.array_equal(ctl.solver_ss.stats()["status"].cpu().numpy(), np.full(3, r.status))
    assert np.abs(sol["x"].cpu().numpy() - r.x).max() < 1e-9 * np.abs(r.x).max()
    with pytest.raises(ValueError):                       # the user rows' bounds are part of the problem's construction
        ctl.solver_ss(lbx=ss.w_lb, ubx=ss.w_ub, x0=w0, p=par, lbg=np.zeros_like(ss.g_lb), ubg=ss.g_ub)


def test_infeasible_user_constraint_is_reported_like_the_oracle(tmp_path, ssu):
    """A duty limit no point inside the bounds can meet: restoration ends at a stationary point of the violation, status 2
    (Infeasible_Problem_Detected, the status `MPC_code.py:714` gates the target on) from the oracle and the device code alike."""
    import os
    import __graft_entry__ as entry
    from conftest import Bundle, ROOT
    from oracle.nlp import TargetNlp
    src = open(os.path.join(ROOT, "examples", "nmpc_cstr_ss_user.py")).read()
    assert "- 15.9)" in src
    src = src.replace("- 15.9)", "- 1.0)").replace("os.path.dirname(__file__)", repr(os.path.join(ROOT, "examples")))
    path = tmp_path / "nmpc_cstr_ss_infeas.py"
    path.write_text(src)
    entry.EXAMPLES["nmpc_cstr_ss_infeas"] = str(path)
    try:
        b = Bundle("nmpc_cstr_ss_infeas")
        par, w0 = _par(b), _guess(b)
        r = TargetNlp(b.ss, b.oracle).solve(w0, par)
        w, f, st, it = b.harness_target(par, w0)
    finally:
        entry.EXAMPLES.pop("nmpc_cstr_ss_infeas", None)
    assert r.status == 2 and st[0] == 2 and it[0] == r.iters
    assert np.abs(w[0] - r.x).max() < 1e-8 * np.abs(r.x).max()
