"""Generated-C model maps for the oracle (TEST INFRASTRUCTURE - see ``oracle/__init__.py``).

The reference obtains every derivative by CasADi AD over the *fully unrolled* graph of a stage
(``Fx_model`` = ``Mx`` RK4 sub-steps expanded symbolically, ``Utilities.py:157-183``; the NLP
Jacobian/Hessian are then derived from that graph by ``nlpsol``, ``Control_Calc.py:258``).  The
oracle does the same: it differentiates the unrolled stage expressions symbolically and compiles
the result with gcc.  (The device path instead chains hand-written RK4 sensitivities - the two
derivative computations share no code below the user's right-hand side.)
"""
from __future__ import annotations

import os

import numpy as np

import mpc_code_b200  # noqa: F401  (host-side tracer / code generator only)
from mpc_code_b200.codegen import CFunction, CModule
from mpc_code_b200.sx import SX, hessian, jacobian, mtimes, vertcat

BUILD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build")


def _scalar_hess(expr, wrt):
    H, g = hessian(expr, wrt)
    return H, g


def ocp_functions(spec):
    """Stage maps of the OCP with first and second derivatives (``Control_Calc.py:128-210``)."""
    X, U, Up, par, pxk, pyk = spec.X, spec.U, spec.Uprev, spec.par, spec.pxk, spec.pyk
    n, m, p = spec.n, spec.m, spec.p
    z = vertcat(X, U)
    lam = SX.sym("lam", n)
    mult = SX.sym("mult", max(p, 1))
    fns = []
    ins = [("X", X), ("U", U), ("par", par), ("pxk", pxk)]
    fns.append(CFunction("orc_dyn", ins, [("Xn", spec.Xnext)]))
    Hd, _ = _scalar_hess(mtimes(lam.T, spec.Xnext), z)
    fns.append(CFunction("orc_dyn_d", ins + [("lam", lam)],
                         [("Xn", spec.Xnext), ("J", jacobian(spec.Xnext, z)), ("H", Hd)]))
    ins_y = [("X", X), ("U", U), ("par", par), ("pyk", pyk)]
    fns.append(CFunction("orc_out", ins_y, [("Y", spec.Y)]))
    Hy, _ = _scalar_hess(mtimes(mult[0:p].T, spec.Y) if p else SX(0.0), z)
    fns.append(CFunction("orc_out_d", ins_y + [("mult", mult)],
                         [("Y", spec.Y), ("JY", jacobian(spec.Y, z)), ("HY", Hy)]))
    if getattr(spec, "G", None) is not None:        # user stage inequalities (Control_Calc.py:132-137)
        ng_in = spec.G.numel()
        multg = SX.sym("multg", ng_in)
        ins_g = [("X", X), ("U", U), ("par", par), ("pxk", pxk), ("pyk", pyk)]
        Hg, _ = _scalar_hess(mtimes(multg.T, spec.G), z)
        fns.append(CFunction("orc_gin", ins_g, [("G", spec.G)]))
        fns.append(CFunction("orc_gin_d", ins_g + [("multg", multg)], [("G", spec.G), ("JG", jacobian(spec.G, z)), ("HG", Hg)]))
    zc = vertcat(X, U, Up)
    ins_c = [("X", X), ("U", U), ("Up", Up), ("par", par), ("pxk", pxk), ("pyk", pyk)]
    Hc, gc = _scalar_hess(spec.stage_cost, zc)
    fns.append(CFunction("orc_cost", ins_c, [("l", spec.stage_cost)]))
    fns.append(CFunction("orc_cost_d", ins_c, [("l", spec.stage_cost), ("g", gc), ("H", Hc)]))
    ins_t = [("XN", spec.XN), ("par", par)]
    Ht, gt = _scalar_hess(spec.term_cost, spec.XN)
    fns.append(CFunction("orc_term", ins_t, [("V", spec.term_cost)]))
    fns.append(CFunction("orc_term_d", ins_t, [("V", spec.term_cost), ("g", gt), ("H", Ht)]))
    return fns


def target_functions(spec):
    """Target-problem maps (``Target_Calc.py:75-124``)."""
    w, par = spec.wss, spec.par
    con = vertcat(spec.Xnext - spec.Xs, spec.Ynext - spec.Ys)
    for extra in (getattr(spec, "Gss", None), getattr(spec, "Hss", None)):      # user rows (Target_Calc.py:87-109)
        if extra is not None:
            con = vertcat(con, extra)
    mult = SX.sym("mult", con.numel())
    Hc, _ = _scalar_hess(mtimes(mult.T, con), w)
    Hf, gf = _scalar_hess(spec.cost, w)
    ins = [("wss", w), ("par", par)]
    return [
        CFunction("orc_ss_con", ins, [("c", con)]),
        CFunction("orc_ss_con_d", ins + [("mult", mult)], [("c", con), ("J", jacobian(con, w)), ("H", Hc)]),
        CFunction("orc_ss_cost", ins, [("f", spec.cost)]),
        CFunction("orc_ss_cost_d", ins, [("f", spec.cost), ("g", gf), ("H", Hf)]),
    ]


def model_functions(prob):
    """Model / plant maps and the estimator Jacobians (``MPC_code.py:546-575``, ``Estimator.py:288-291,343-345,372-373``)."""
    s = prob.sym
    x, u, d, k, t, px, py = s["x"], s["u"], s["d"], s["k"], s["t"], s["px"], s["py"]
    xp, pxp, pyp, pxmp, pymp = s["xp"], s["pxp"], s["pyp"], s["pxmp"], s["pymp"]
    nx, nd = prob.nx, prob.nd
    Fx = prob.Fx_model(x, u, k, d, t, px)
    Fy = prob.Fy_model(x, u, d, t, py)
    fns = [
        CFunction("orc_fx", [("x", x), ("u", u), ("k", k), ("d", d), ("t", t), ("px", px)], [("xn", Fx)]),
        CFunction("orc_fy", [("x", x), ("u", u), ("d", d), ("t", t), ("py", py)], [("y", Fy)]),
    ]
    if prob.flags["offree"] != "no":   # augmented estimator model, xi = [x; d]
        xi = vertcat(x, d)
        Fxes = vertcat(Fx, d)
    else:
        xi = x
        Fxes = Fx
    # the estimator functions take xi split as (x, d) so that inputs stay purely symbolic
    fns.append(CFunction("orc_fxes_d", [("x", x), ("d", d), ("u", u), ("k", k), ("t", t), ("px", px)],
                         [("F", Fxes), ("A", jacobian(Fxes, xi))]))
    fns.append(CFunction("orc_fyes_d", [("x", x), ("d", d), ("u", u), ("t", t), ("py", py)],
                         [("y", Fy), ("C", jacobian(Fy, xi))]))
    if prob.flags["Fp_nominal"] is not True:   # nominal case: the plant *is* the model (MPC_code.py:172-174)
        Fxp = prob.Fx_p(xp, u, pxp, t, k, pxmp)
        Fyp = prob.Fy_p(xp, u, pyp, t, pymp)
        fns.append(CFunction("orc_fxp", [("x", xp), ("u", u), ("pxp", pxp), ("t", t), ("k", k), ("pxmp", pxmp)],
                             [("xn", Fxp)]))
        fns.append(CFunction("orc_fyp", [("x", xp), ("u", u), ("pyp", pyp), ("t", t), ("pymp", pymp)],
                             [("y", Fyp)]))
    return fns


def build(name, prob, ocp_spec=None, ss_spec=None) -> CModule:
    fns = model_functions(prob)
    if ocp_spec is not None:
        fns += ocp_functions(ocp_spec)
    if ss_spec is not None:
        fns += target_functions(ss_spec)
    # ORACLE_CFLAGS=-O0 for the largest members of the synthetic family: their unrolled second-derivative functions are
    # tens of MB of C and gcc -O1 does not finish on them in an hour
    cflags = tuple(os.environ.get("ORACLE_CFLAGS", "-O1").split())
    return CModule("orc_" + name, fns, BUILD_DIR, cflags=cflags)
