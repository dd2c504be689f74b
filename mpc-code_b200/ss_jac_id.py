"""``ssjacid = True``: linearise the user's nonlinear model at a steady state found near ``(x0_m, u0)``.

Host-side set-up tool, the counterpart of the reference's ``SS_JAC_ID.py``: ``ss_p_jac_id`` (``:14-121``) builds the
nonlinear model without its linear disturbance terms, looks for a steady state - the reference solves
``min |Fx(x,u)-x|^2 + |Fy(x,u)-y|^2  s.t.  Fx(x,u)-x = 0, Fy(x,u)-y = 0`` and the bounds with IPOPT (``opt_ss_id``,
``:124-201``) from the guess ``(x0_m, u0, Fy(x0_m,u0))`` with ``d = 0, t = 0`` - and returns the Jacobians
``A = dFx/dx, B = dFx/du, C = dFy/dx, D = dFy/du`` there together with ``(xlin, ulin, ylin)``; the driver then builds a
LINEAR model about that point (``MPC_code.py:84-91``).  One-off work of a few variables: done on the host by
minimum-norm Gauss-Newton steps on the same residuals with exact (symbolic) Jacobians.  The steady-state equations have
``nu`` degrees of freedom, so which steady state is returned depends on the solver's path from the guess: like IPOPT's it
is the guess itself when that is a steady state, and a nearby one otherwise (not guaranteed to be IPOPT's).
"""
from __future__ import annotations

from typing import Any, Dict, List

import numpy as np

from .model_factory import defF_model
from .sx import SX, Function, jacobian, vertcat


def ss_p_jac_id(ns: Dict[str, Any], k: SX, t: SX, px: SX, py: SX) -> List[np.ndarray]:
    x, u, y, d = ns["x"], ns["u"], ns["y"], ns["d"]
    nx, nu, ny, nd = x.size1(), u.size1(), y.size1(), d.size1()
    offree = "no" if ns.get("offree", "no") == "lin" else ns.get("offree", "no")          # SS_JAC_ID.py:19-22
    kw: Dict[str, Any] = {}
    if "User_fxm_Cont" in ns:
        kw.update(fx=ns["User_fxm_Cont"], Mx=ns["Mx"])
    elif "User_fxm_Dis" in ns:
        kw.update(Fx=ns["User_fxm_Dis"])
    elif "A" in ns and "User_fym" in ns:
        kw.update(A=ns["A"], B=ns["B"])
    else:
        raise ValueError("ssjacid needs a nonlinear model (User_fxm_Cont / User_fxm_Dis) or A, B with User_fym")
    if ns.get("StateFeedback") is True:
        kw.update(SF=True)
    elif "User_fym" in ns:
        kw.update(fy=ns["User_fym"])
    else:
        kw.update(C=ns["C"])
    Fx_model, Fy_model = defF_model(x, u, y, d, k, t, px, py, offree, ns["LinPar"], **kw)
    h = float(ns["h"])
    d0, t0 = np.zeros(nd), 0.0
    px0, py0 = np.zeros(px.size1()), np.zeros(py.size1())
    Fx = Fx_model(x, u, h, d, t, px); Fy = Fy_model(x, u, d, t, py)
    jac = Function("ssjac", [x, u, d, t, px, py], [Fx, Fy, jacobian(Fx, x), jacobian(Fx, u), jacobian(Fy, x), jacobian(Fy, u)])

    def ev(xv, uv):
        return [np.asarray(v, dtype=float) for v in jac(xv, uv, d0, t0, px0, py0)]

    def bound(name, n, sign):
        v = ns.get(name)
        return np.full(n, sign * np.inf) if v is None else np.asarray(v, dtype=float).reshape(n)
    lo = np.concatenate([bound("xmin", nx, -1), bound("umin", nu, -1), bound("ymin", ny, -1)])
    hi = np.concatenate([bound("xmax", nx, +1), bound("umax", nu, +1), bound("ymax", ny, +1)])
    x0 = np.asarray(ns["x0_m"], dtype=float).reshape(nx); u0 = np.asarray(ns["u0"], dtype=float).reshape(nu)
    y0 = ev(x0, u0)[1].reshape(ny)
    w0 = np.clip(np.concatenate([x0, u0, y0]), lo, hi)

    def res(w):
        F, Y = ev(w[:nx], w[nx:nx + nu])[:2]
        return np.concatenate([F.reshape(nx) - w[:nx], Y.reshape(ny) - w[nx + nu:]])

    def res_jac(w):
        _, _, A, B, C, D = ev(w[:nx], w[nx:nx + nu])
        J = np.zeros((nx + ny, nx + nu + ny))
        J[:nx, :nx] = A - np.eye(nx); J[:nx, nx:nx + nu] = B
        J[nx:, :nx] = C; J[nx:, nx:nx + nu] = D; J[nx:, nx + nu:] = -np.eye(ny)
        return J
    # minimum-norm Gauss-Newton steps (the steady-state equations leave nu degrees of freedom: every step moves to the
    # nearest point of the linearised solution set), halved until the residual decreases, kept inside the bounds
    w = w0.copy()
    r = res(w)
    for _ in range(60):
        if np.abs(r).max() < 1e-13 * max(1.0, np.abs(w).max()):
            break
        dw = -np.linalg.lstsq(res_jac(w), r, rcond=None)[0]
        a = 1.0
        while a > 1e-8:
            wt = np.clip(w + a * dw, lo, hi)
            rt = res(wt)
            if np.all(np.isfinite(rt)) and np.linalg.norm(rt) < (1.0 - 1e-4 * a) * np.linalg.norm(r):
                break
            a *= 0.5
        else:
            break
        w, r = wt, rt
    if not np.abs(r).max() < 1e-9 * max(1.0, np.abs(w).max()):
        raise RuntimeError("ssjacid: no steady state found near (x0_m, u0): residual %.2e" % np.abs(r).max())
    xlin, ulin, ylin = w[:nx], w[nx:nx + nu], w[nx + nu:]
    _, _, A, B, C, D = ev(xlin, ulin)
    return [A, B, C, D, xlin, ulin, ylin]


def linear_model_from_ssjacid(ns: Dict[str, Any], x, u, y, d, k, t, px, py):
    """``MPC_code.py:84-91``: the model the loop uses when ``ssjacid`` is set."""
    A, B, C, D, xlin, ulin, ylin = ss_p_jac_id(ns, k, t, px, py)
    ns.update(A=A, B=B, C=C, xlin=xlin, ulin=ulin, ylin=ylin)
    dist = dict(Bd=ns["Bd"], Cd=ns["Cd"]) if ns["offree"] == "lin" else {}
    return defF_model(x, u, y, d, k, t, px, py, ns["offree"], ns["LinPar"], A=A, B=B, C=C, xlin=xlin, ulin=ulin, ylin=ylin, **dist)
