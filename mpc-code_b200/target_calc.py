"""``opt_ss``: the steady-state target problem (``Target_Calc.py:20-161``).

Variables ``wss = [Xs, Us, Ys]``; parameters ``par_ss = [usp|ysp|xsp|d|Us_prev|vec(lam)|t|px|py]``;
equalities ``Fx_model(Xs,Us,h,d,t,px) - Xs = 0`` and ``Fy_model(Xs,Us,d,t,py) + lam (Us - Us_prev) - Ys = 0``,
then the user rows ``User_g_ineq_SS(xs,us,ys,d,t,px,py) <= 0`` and ``User_h_eq_SS(...) = 0`` (``:87-109,149-153``);
box bounds on all three blocks.  Recorded symbolically in `TargetSpec`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict

import numpy as np

from .sx import SX, Function, mtimes, vertcat


@dataclass
class TargetSpec:
    n: int; m: int; p: int; nd: int; npx: int; npy: int; h: float
    nw: int; npar: int
    off: Dict[str, int]
    wss: SX; par: SX
    Xs: SX; Us: SX; Ys: SX
    Xnext: SX            # Fx_model(Xs,Us,h,d,t,px)                  (:75)
    Ynext: SX            # Fy_model(Xs,Us,d,t,py) + lam (Us-Us_prev) (:80)
    cost: SX             # Fss_obj(dx,du,dy,xsp,usp,ysp)             (:112-124)
    Fx_model: Function; Fy_model: Function
    flags: Dict[str, Any]
    w_lb: np.ndarray; w_ub: np.ndarray; g_lb: np.ndarray; g_ub: np.ndarray
    sol_opts: Dict[str, Any] = field(default_factory=dict)
    Gss: Any = None      # User_g_ineq_SS(Xs,Us,Ys,d,t,px,py) <= 0   (:87-98), SX column or None
    Hss: Any = None      # User_h_eq_SS(...) = 0                     (:91-103)
    ng_ss: int = 0; nh_ss: int = 0


def par_ss_offsets(n, m, p, nd, npx, npy) -> Dict[str, int]:
    """Offsets inside ``par_ss`` (``Target_Calc.py:41-50``)."""
    off = dict(usp=0, ysp=m, xsp=m + p, d=m + p + n, usprev=m + p + n + nd, lam=2 * m + p + n + nd)
    off["t"] = off["lam"] + p * m
    off["px"] = off["t"] + 1
    off["py"] = off["px"] + npx
    off["end"] = off["py"] + npy
    return off


def _inf_or(v, n, sign):
    return np.full(n, sign * np.inf) if v is None else np.asarray(v, dtype=float).reshape(n)


def build_target_spec(n, m, p, nd, npx, npy, Fx_model, Fy_model, Fss_obj, QForm_ss, DUssForm, sol_opts,
                      G_ineq_SS, H_eq_SS, umin=None, umax=None, w_s=None, z_s=None, ymin=None, ymax=None,
                      xmin=None, xmax=None, h=None) -> TargetSpec:
    nxu, nxuy = n + m, n + m + p
    off = par_ss_offsets(n, m, p, nd, npx, npy)
    wss = SX.sym("wss", nxuy)
    par = SX.sym("par_ss", off["end"])
    Xs, Us, Ys = wss[0:n], wss[n:nxu], wss[nxu:nxuy]
    usp = par[off["usp"]:off["usp"] + m]
    ysp = par[off["ysp"]:off["ysp"] + p]
    xsp = par[off["xsp"]:off["xsp"] + n]
    d = par[off["d"]:off["d"] + nd]
    usprev = par[off["usprev"]:off["usprev"] + m]
    lam = par[off["lam"]:off["lam"] + p * m].reshape((p, m))
    t = par[off["t"]:off["t"] + 1]
    px = par[off["px"]:off["px"] + npx]
    py = par[off["py"]:off["py"] + npy]
    if h is None:
        h = 0.1  # (:68-69)
    Xnext = Fx_model(Xs, Us, h, d, t, px)
    Ynext = Fy_model(Xs, Us, d, t, py) + mtimes(lam, Us - usprev)
    dx, du, dy = Xs, Us, Ys
    if QForm_ss is True:
        dx, dy, du = dx - xsp, dy - ysp, du - usp
    if DUssForm is True:
        du = Us - usprev
    cost = Fss_obj(dx, du, dy, xsp, usp, ysp)
    w_lb = np.concatenate([_inf_or(xmin, n, -1), _inf_or(umin, m, -1), _inf_or(ymin, p, -1)])
    w_ub = np.concatenate([_inf_or(xmax, n, +1), _inf_or(umax, m, +1), _inf_or(ymax, p, +1)])
    Gss = Hss = None
    if G_ineq_SS is not None:
        Gss = SX(vertcat(G_ineq_SS(Xs, Us, Ys, d, t, px, py)))
    if H_eq_SS is not None:
        Hss = SX(vertcat(H_eq_SS(Xs, Us, Ys, d, t, px, py)))
    ng1 = 0 if Gss is None else Gss.numel()
    ng2 = 0 if Hss is None else Hss.numel()
    g_lb = np.zeros(n + p + ng1 + ng2); g_ub = np.zeros(n + p + ng1 + ng2)
    g_lb[n + p:n + p + ng1] = -np.inf                                   # (:149-150)
    return TargetSpec(Gss=Gss, Hss=Hss, ng_ss=ng1, nh_ss=ng2, n=n, m=m, p=p, nd=nd, npx=npx, npy=npy, h=float(h), nw=nxuy, npar=off["end"], off=off,
                      wss=wss, par=par, Xs=Xs, Us=Us, Ys=Ys, Xnext=SX(Xnext), Ynext=SX(Ynext), cost=SX(cost),
                      Fx_model=Fx_model, Fy_model=Fy_model, flags=dict(QForm_ss=QForm_ss, DUssForm=DUssForm),
                      w_lb=w_lb, w_ub=w_ub, g_lb=g_lb, g_ub=g_ub, sol_opts=dict(sol_opts or {}))


def opt_ss(*args, **kwargs):
    """Reference-compatible entry point: ``[solver_ss, wss_lb, wss_ub, gss_lb, gss_ub]`` (``Target_Calc.py:161``)."""
    from .solvers import BatchedNlpSolver
    spec = build_target_spec(*args, **kwargs)
    solver = BatchedNlpSolver("target", spec)
    return [solver, spec.w_lb.copy(), spec.w_ub.copy(), spec.g_lb.copy(), spec.g_ub.copy()]
