"""Model, plant and objective factories with the reference's names and argument meaning.

Each factory returns `sx.Function` objects that behave like the CasADi ones the reference
builds, and additionally carry ``meta`` describing their structure (linear map / RK4 of a
continuous right-hand side / discrete map, additive terms) so that the code generator can
hand the device kernels the *pieces* (right-hand side, its derivatives) instead of one
unrolled graph.

Reference: ``Utilities.py:21-100`` (defF_p), ``:102-245`` (defF_model), ``:247-265`` (xQx),
``:267-321`` (defFss_obj), ``:323-381`` (defF_obj), ``:383-420`` (defVfin).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as scla

from .sx import SX, DM, Function, fabs, mtimes, simpleRK, vertcat

__all__ = ["defF_p", "defF_model", "xQx", "defFss_obj", "defF_obj", "defVfin"]


def _depends_on(expr: SX, var: SX) -> bool:
    from . import symbolic as S
    ids = {e.uid for e in var.elements()}
    return any(s.uid in ids for s in S.symbols_of(expr.elements()))


def _rk4_of(rhs_expr: SX, state: SX, t: SX, frozen: SX, substeps: int):
    """RK4 of ``[rhs; 1]`` on the state ``[state; t]`` with ``frozen`` held over the step.

    Time is carried as an extra state exactly as ``Utilities.py:163-169`` does, so a
    time-dependent right-hand side sees ``t_k + c_i * dt`` at the Runge-Kutta stage points.
    """
    aug_rhs = vertcat(rhs_expr, SX(1.0))
    aug_state = vertcat(state, t)
    f_aug = Function("rhs_aug", [aug_state, frozen], [aug_rhs])
    return simpleRK(f_aug, substeps), aug_state


def defF_p(x, u, y, k, t, pxp, pyp, pxmp, pymp, LinPar, **plant):
    """Plant maps ``Fx_p(x,u,pxp,t,k,pxmp)`` and ``Fy_p(x,u,pyp,t,pymp)`` (``Utilities.py:21-100``)."""
    nx = x.size1()
    Fx_p = Fy_p = None
    for key in plant:
        if key == "Ap":
            fx = mtimes(plant["Ap"], x) + mtimes(plant["Bp"], u) + pxp + pxmp
            Fx_p = Function("Fx_p", [x, u, pxp, t, k, pxmp], [fx])
            Fx_p.meta.update(kind="linear")
        elif key == "Fx":
            fx = plant["Fx"](x, t, u, pxp, pxmp)
            if LinPar is True:
                fx = fx + pxp + pxmp
            Fx_p = Function("Fx_p", [x, u, pxp, t, k, pxmp], [fx])
            Fx_p.meta.update(kind="discrete")
        elif key == "fx":
            substeps = plant["Mx"]
            rhs = plant["fx"](x, t, u, pxp, pxmp)
            frozen = vertcat(u, pxp, pxmp)
            rk, aug = _rk4_of(rhs, x, t, frozen, substeps)
            step = rk(aug, frozen, k)[:nx, :]
            post = SX.zeros(nx, 1)
            if LinPar is True:
                post = pxp + pxmp
            Fx_p = Function("Fx_p", [x, u, pxp, t, k, pxmp], [step + post])
            Fx_p.meta.update(kind="rk4", substeps=substeps,
                             rhs=Function("fxp", [x, u, pxp, t, pxmp], [rhs]),
                             post=Function("fxp_post", [pxp, pxmp], [post]))
        if key == "SF":
            Fy_p = Function("Fy_p", [x, u, pyp, t, pymp], [x])
        elif key == "Cp":
            Fy_p = Function("Fy_p", [x, u, pyp, t, pymp], [mtimes(plant["Cp"], x) + pyp + pymp])
        elif key == "fy":
            fy = plant["fy"](x, u, t, pyp, pymp)
            if LinPar is True:
                fy = fy + pyp + pymp
            Fy_p = Function("Fy_p", [x, u, pyp, t, pymp], [fy])
    return [Fx_p, Fy_p]


def defF_model(x, u, y, d, k, t, px, py, offree, LinPar, **model):
    """Model maps ``Fx_model(x,u,k,d,t,px)`` and ``Fy_model(x,u,d,t,py)`` (``Utilities.py:102-245``)."""
    nx = x.size1()
    Bd = Cd = None
    if offree == "lin":
        Bd, Cd = model["Bd"], model["Cd"]
    Fx_model = None
    fy_model = None

    if "A" in model:  # linear state map, optionally about (xlin, ulin)  (:135-155)
        A, B = model["A"], model["B"]
        if "xlin" in model:
            xlin, ulin = model["xlin"], model["ulin"]
            fx = mtimes(A, x - xlin) + mtimes(B, u - ulin) + xlin
        else:
            fx = mtimes(A, x) + mtimes(B, u)
        if offree == "lin":
            fx = fx + mtimes(Bd, d)
        fx = fx + px
        Fx_model = Function("Fx_model", [x, u, k, d, t, px], [fx])
        Fx_model.meta.update(kind="linear")
    elif "fx" in model:  # continuous nonlinear -> RK4 with Mx sub-steps  (:157-183)
        substeps = model["Mx"]
        rhs = model["fx"](x, u, d, t, px)
        frozen = vertcat(u, d, px) if offree == "nl" else vertcat(u, px)
        if offree != "nl" and _depends_on(SX(rhs), d):
            raise ValueError("the continuous model depends on d but offree != 'nl' "
                             "(d would be a free variable of the integrator, Utilities.py:127-130,167)")
        rk, aug = _rk4_of(rhs, x, t, frozen, substeps)
        step = rk(aug, frozen, k)[:nx, :]
        post = SX.zeros(nx, 1)
        if offree == "lin":
            post = post + mtimes(Bd, d)
        if LinPar is True:
            post = post + px
        Fx_model = Function("Fx_model", [x, u, k, d, t, px], [step + post])
        Fx_model.meta.update(kind="rk4", substeps=substeps,
                             rhs=Function("fxm", [x, u, d, t, px], [rhs]),
                             post=Function("fxm_post", [d, px], [post]))
    elif "Fx" in model:  # discrete nonlinear  (:186-198)
        fx = model["Fx"](x, u, d, t, px)
        if offree == "lin":
            fx = fx + mtimes(Bd, d)
        if LinPar is True:
            fx = fx + px
        Fx_model = Function("Fx_model", [x, u, k, d, t, px], [fx])
        Fx_model.meta.update(kind="discrete")

    if "SF" in model:  # state feedback: y = x  (:201-205)
        fy_model = x
        if offree == "lin":
            fy_model = fy_model + mtimes(Cd, d)
    elif "C" in model:  # (:208-230)
        C = model["C"]
        if "ylin" in model and "xlin" in model:
            fy_model = mtimes(C, x - model["xlin"]) + model["ylin"]
        elif "ylin" in model:
            fy_model = mtimes(C, x) + model["ylin"]
        else:
            fy_model = mtimes(C, x)
        if offree == "lin":
            fy_model = fy_model + mtimes(Cd, d)
    elif "fy" in model:  # (:232-238)
        fy_model = model["fy"](x, u, d, t, py)
        if offree == "lin":
            fy_model = fy_model + mtimes(Cd, d)
    if LinPar is True:
        fy_model = fy_model + py
    Fy_model = Function("Fy_model", [x, u, d, t, py], [fy_model])
    return [Fx_model, Fy_model]


def xQx(x, Q):
    """``x' Q x`` (``Utilities.py:247-265``)."""
    return mtimes(SX(x).T if not isinstance(x, SX) else x.T, mtimes(Q, x))


def defFss_obj(x, u, y, xsp, usp, ysp, **kwargs):
    """Target-problem objective ``Fss_obj(x,u,y,xsp,usp,ysp)`` (``Utilities.py:267-321``)."""
    if "r_y" in kwargs:
        r_u = kwargs["r_u"] if "r_u" in kwargs else kwargs["r_Du"]
        fss = mtimes(DM(kwargs["r_y"]).reshape(1, -1), y) + mtimes(DM(r_u).reshape(1, -1), fabs(u))
    elif "Q" in kwargs:
        Ru = kwargs["R"] if "R" in kwargs else kwargs["S"]
        fss = 0.5 * (xQx(y, kwargs["Q"]) + xQx(u, Ru))
    elif "f_obj" in kwargs:
        fss = kwargs["f_obj"](x, u, y, xsp, usp, ysp)
    else:
        raise ValueError("defFss_obj needs r_y, Q or f_obj")
    return Function("Fss_obj", [x, u, y, xsp, usp, ysp], [fss])


def defF_obj(x, u, y, xs, us, ys, **kwargs):
    """Stage cost ``F_obj(x,u,y,xs,us,ys)`` (``Utilities.py:323-381``)."""
    if "r_x" in kwargs:
        r_u = kwargs["r_u"] if "r_u" in kwargs else kwargs["r_Du"]
        f = mtimes(DM(kwargs["r_x"]).reshape(1, -1), fabs(x)) + mtimes(DM(r_u).reshape(1, -1), fabs(u))
    elif "Q" in kwargs:
        Ru = kwargs["R"] if "R" in kwargs else kwargs["S"]
        f = 0.5 * (xQx(x, kwargs["Q"]) + xQx(u, Ru))
    elif "f_Cont" in kwargs:
        f = kwargs["f_Cont"](x, u, y, xs, us, ys)
    elif "f_Dis" in kwargs:
        f = kwargs["f_Dis"](x, u, y, xs, us, ys)
    else:
        raise ValueError("defF_obj needs r_x, Q, f_Cont or f_Dis (collocation is out of scope)")
    return Function("F_obj", [x, u, y, xs, us, ys], [f])


def defVfin(x, xs, **Tcost):
    """Terminal cost ``Vfin(x,xs)``: zero, DARE-weighted quadratic, or user (``Utilities.py:383-420``)."""
    P = None
    if not Tcost:
        vfin = SX(0.0)
    elif "A" in Tcost:
        P = scla.solve_discrete_are(np.array(Tcost["A"], dtype=float), np.array(Tcost["B"], dtype=float),
                                    np.array(Tcost["Q"], dtype=float), np.array(Tcost["R"], dtype=float))
        vfin = 0.5 * xQx(x, P)
    elif "vfin_F" in Tcost:
        vfin = Tcost["vfin_F"](x, xs)
    else:
        raise ValueError("defVfin: unknown terminal cost specification")
    V = Function("Vfin", [x, xs], [vfin])
    V.meta["P"] = P
    return V
