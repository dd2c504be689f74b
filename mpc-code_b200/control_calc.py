"""``opt_dyn``: the dynamic optimal-control problem, stated stage by stage.

Same call signature, variable order, parameter order, constraint order and bounds as the
reference builder (``Control_Calc.py:20-260``).  Instead of one unrolled NLP graph handed to
IPOPT, the builder records the *stage* expressions (dynamics, output rows, stage cost,
terminal cost) once; the device solver and the CPU oracle both work from that record
(`OcpSpec`).  The returned ``solver`` is called like the CasADi one
(``solver(lbx=, ubx=, x0=, p=, lbg=, ubg=)``) but on a whole batch of instances.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import numpy as np

from .sx import SX, Function, mtimes, vertcat


@dataclass
class OcpSpec:
    """Stage-wise record of the OCP of ``Control_Calc.py:20-260`` (multiple-shooting path)."""
    n: int; m: int; p: int; nd: int; npx: int; npy: int; N: int; h: float
    nw: int; npar: int
    off: Dict[str, int]                 # offsets of x0|xs|us|d|um1|t|lam|px|py inside par (:43-57)
    X: SX; U: SX; Uprev: SX; par: SX; pxk: SX; pyk: SX   # stage symbols
    Xnext: SX                           # Fx_model(X,U,h,d,t,pxk)                     (:161)
    Y: SX                               # Fy_model(X,U,d,t,pyk) + lam (U - us)        (:130)
    DU: SX                              # U - Uprev                                   (:163-166)
    stage_cost: SX                      # F_obj(dx,du,dy,xs,us_obj,ys)                (:173-188)
    XN: SX; term_cost: SX               # Vfin(dxN, xs)                               (:194-196,209)
    term_eq: Optional[SX]               # X_N - xs when TermCons                      (:197-198)
    yFree: bool; DuFree: bool
    Fx_model: Function; Fy_model: Function
    flags: Dict[str, Any]
    w_lb: np.ndarray; w_ub: np.ndarray; g_lb: np.ndarray; g_ub: np.ndarray
    bounds: Dict[str, np.ndarray]
    sol_opts: Dict[str, Any] = field(default_factory=dict)
    G: Optional[SX] = None              # user stage inequalities G_ineq(X,U,Y,d,t,pxk,pyk) <= 0   (:94-96,132-137,244-245)
    quad_cost: Optional[SX] = None      # ContForm: integrand of the stage cost     (:102-111)
    cont_rhs: Optional[SX] = None       # ContForm: ode right-hand side fx(...)+px  (:103)
    cont_substeps: int = 0

    @property
    def uses_uprev(self) -> bool:
        from . import symbolic as S
        ids = {e.uid for e in self.Uprev.elements()}
        exprs = list(self.stage_cost.elements())
        return (not self.DuFree) or any(s.uid in ids for s in S.symbols_of(exprs))

    @property
    def n_gin(self) -> int:
        return 0 if self.G is None else self.G.numel()

    @property
    def ng(self) -> int:
        return self.g_lb.size


def par_offsets(n, m, p, nd, npx, npy, N) -> Dict[str, int]:
    """Offsets inside ``par = [x0|xs|us|d|um1|t|vec(lam)|vec(px)|vec(py)]`` (``Control_Calc.py:43-52``)."""
    nxu = n + m
    off = dict(x0=0, xs=n, us=2 * n, d=n + nxu, um1=n + nxu + nd, t=2 * nxu + nd, lam=2 * nxu + nd + 1)
    off["px"] = off["lam"] + p * m
    off["py"] = off["px"] + npx * N
    off["end"] = off["py"] + npy * N
    return off


def _inf_or(v, n, sign):
    return np.full(n, sign * np.inf) if v is None else np.asarray(v, dtype=float).reshape(n)


def build_ocp_spec(xSX, uSX, ySX, dSX, tSX, pxSX, pySX, n, m, p, nd, npx, npy, ng_v, nh_v, Fx_model, Fy_model,
                   F_obj, Vfin, N, QForm, DUForm, DUFormEcon, ContForm, TermCons, slacks, slacksG, slacksH, nw,
                   sol_opts, G_ineq, H_eq, umin=None, umax=None, W=None, Z=None, ymin=None, ymax=None,
                   xmin=None, xmax=None, Dumin=None, Dumax=None, h=None, fx=None, xstat=None, ustat=None,
                   Ws=None) -> OcpSpec:
    if slacks is True or H_eq is not None:
        raise NotImplementedError("slack variables and user equality constraints h are outside the accelerated path")
    nxu = n + m
    if nw != nxu * N + n:
        raise ValueError("nw must be n*(N+1)+m*N without slacks (Control_Calc.py:28)")
    off = par_offsets(n, m, p, nd, npx, npy, N)
    par = SX.sym("par", off["end"])
    xs = par[off["xs"]:off["xs"] + n]
    us = par[off["us"]:off["us"] + m]
    d = par[off["d"]:off["d"] + nd]
    t = par[off["t"]:off["t"] + 1]
    lam = par[off["lam"]:off["lam"] + p * m].reshape((p, m))
    py0 = par[off["py"]:off["py"] + npy]
    X = SX.sym("X", n); U = SX.sym("U", m); Uprev = SX.sym("Uprev", m)
    pxk = SX.sym("pxk", npx); pyk = SX.sym("pyk", npy)
    if h is None:
        h = 0.1  # (:91-92)

    yFree = ymin is None and ymax is None          # (:60-73)
    DuFree = Dumin is None and Dumax is None       # (:82-89)
    ymin_v, ymax_v = _inf_or(ymin, p, -1), _inf_or(ymax, p, +1)
    xmin_v, xmax_v = _inf_or(xmin, n, -1), _inf_or(xmax, n, +1)
    umin_v, umax_v = _inf_or(umin, m, -1), _inf_or(umax, m, +1)
    Dumin_v, Dumax_v = _inf_or(Dumin, m, -1), _inf_or(Dumax, m, +1)

    ys = Fy_model(xs, us, d, t, py0)                                # (:124)
    Y = Fy_model(X, U, d, t, pyk) + mtimes(lam, U - us)            # (:130)
    DU = U - Uprev                                                  # (:163-166)
    G = None
    if G_ineq is not None:                                          # (:94-96,132-137): evaluated with the corrected Y_k
        G = SX(G_ineq(X, U, Y, d, t, pxk, pyk))
        if G.numel() == 0:
            G = None
    quad_cost = cont_rhs = None
    if ContForm is True:                                            # (:102-111,153-158)
        # The reference integrates  xdot = fx(x,u,d,t,px) + px  together with the quadrature of the stage cost over
        # [0, h] with SUNDIALS IDAS (variable-step BDF, rel. tol 1e-6).  DEVIATION: a fixed-step classic RK4 with the
        # model's Mx sub-steps is used instead (device and oracle alike) - IDAS is neither vendored nor smooth to 1e-8.
        from .sx import simpleRK, vertcat as _vc
        cont_rhs = fx(X, U, d, t, pxk) + pxk
        ystat = Fy_model(xs, us, d, t, pyk)
        quad_cost = F_obj(X, U, Fy_model(X, U, d, t, pyk), xs, us, ystat)
        substeps = int(Fx_model.meta.get("substeps", 10))
        qsym = SX.sym("q", 1)
        frozen = _vc(U, par, pxk, pyk)
        f_aug = Function("ocq_rhs", [_vc(X, qsym), frozen], [_vc(SX(cont_rhs), SX(quad_cost))])
        end = simpleRK(f_aug, substeps)(_vc(X, SX(0.0)), frozen, h)
        Xnext = end[0:n, :]
        stage_cost = end[n, :]
    else:
        Xnext = Fx_model(X, U, h, d, t, pxk)                        # (:161)
        dx, du, dy = X, U, Y                                        # (:173-185)
        if QForm is True:
            dx, du, dy = dx - xs, du - us, dy - ys
        if DUForm is True:
            du = DU
        us_obj = DU if DUFormEcon is True else us
        stage_cost = F_obj(dx, du, dy, xs, us_obj, ys)
    XN = SX.sym("XN", n)
    dxN = XN - xs if QForm is True else XN                          # (:194-196)
    term_cost = Vfin(dxN, xs)                                       # (:209)
    term_eq = dxN if TermCons is True else None                     # (:197-198)

    # bounds (:213-252)
    w_lb = np.full(nw, -np.inf); w_ub = np.full(nw, np.inf)
    for k in range(N + 1):
        w_lb[k * nxu:k * nxu + n] = xmin_v; w_ub[k * nxu:k * nxu + n] = xmax_v
    for k in range(1, N + 1):
        w_lb[k * nxu - m:k * nxu] = umin_v; w_ub[k * nxu - m:k * nxu] = umax_v
    ng = n * (N + 1) + (n if TermCons is True else 0)
    ng1 = 0 if yFree else p * N
    if ContForm is True:
        DuFree = True                  # the ContForm branch never forms DU_k: no g2 rows, Dumin/Dumax ignored (Control_Calc.py:153-158)
    ng2 = 0 if DuFree else m * N
    ng4 = 0 if G is None else G.numel() * N
    g_lb = np.zeros(ng + ng1 + ng2 + ng4); g_ub = np.zeros(ng + ng1 + ng2 + ng4)
    if ng1:
        g_lb[ng:ng + ng1] = np.tile(ymin_v, N); g_ub[ng:ng + ng1] = np.tile(ymax_v, N)
    if ng2:
        g_lb[ng + ng1:ng + ng1 + ng2] = np.tile(Dumin_v, N); g_ub[ng + ng1:ng + ng1 + ng2] = np.tile(Dumax_v, N)
    if ng4:
        g_lb[ng + ng1 + ng2:] = -np.inf                             # (:244-245)

    flags = dict(QForm=QForm, DUForm=DUForm, DUFormEcon=DUFormEcon, ContForm=ContForm, TermCons=TermCons)
    return OcpSpec(n=n, m=m, p=p, nd=nd, npx=npx, npy=npy, N=N, h=float(h), nw=nw, npar=off["end"], off=off,
                   X=X, U=U, Uprev=Uprev, par=par, pxk=pxk, pyk=pyk, Xnext=SX(Xnext), Y=SX(Y), DU=DU,
                   stage_cost=SX(stage_cost), XN=XN, term_cost=SX(term_cost), term_eq=term_eq,
                   yFree=yFree, DuFree=DuFree, Fx_model=Fx_model, Fy_model=Fy_model, flags=flags,
                   w_lb=w_lb, w_ub=w_ub, g_lb=g_lb, g_ub=g_ub,
                   bounds=dict(xmin=xmin_v, xmax=xmax_v, umin=umin_v, umax=umax_v, ymin=ymin_v, ymax=ymax_v,
                               Dumin=Dumin_v, Dumax=Dumax_v),
                   sol_opts=dict(sol_opts or {}), G=G, quad_cost=quad_cost, cont_rhs=cont_rhs,
                   cont_substeps=(int(Fx_model.meta.get("substeps", 10)) if ContForm is True else 0))


def opt_dyn(*args, **kwargs):
    """Reference-compatible entry point: returns ``[solver, w_lb, w_ub, g_lb, g_ub]`` (``Control_Calc.py:260``).

    ``solver`` is a `solvers.BatchedNlpSolver`; it compiles the device code on first use (or
    when `MpcProblem`-level compilation attaches a shared handle) and fails loudly if the CUDA
    library cannot be built or loaded - there is no CPU fallback.
    """
    from .solvers import BatchedNlpSolver
    spec = build_ocp_spec(*args, **kwargs)
    solver = BatchedNlpSolver("ocp", spec)
    return [solver, spec.w_lb.copy(), spec.w_ub.copy(), spec.g_lb.copy(), spec.g_ub.copy()]
