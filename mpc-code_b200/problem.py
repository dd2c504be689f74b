"""Turn a loaded problem namespace into the objects the step solvers are built from.

This is the setup half of the reference driver (``MPC_code.py:30-483``): dimensions from the
symbol sizes, the name-presence ladder that decides which model / plant / objective factory
call is made, bound overrides for the target and dynamic problems, estimator selection and the
initial loop state.  The logic is reproduced rule by rule so that an unmodified ``Ex_*.py``
selects the same problem; the outcome is one `MpcProblem` instead of module globals.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Dict, Optional

import numpy as np

from .sx import SX, Function
from .model_factory import defF_model, defF_p, defF_obj, defFss_obj, defVfin


def _col(v, n=None) -> Optional[np.ndarray]:
    if v is None:
        return None
    a = np.asarray(v, dtype=float).reshape(-1)
    if n is not None and a.size != n:
        raise ValueError("expected %d entries, got %d" % (n, a.size))
    return a


@dataclass
class MpcProblem:
    """Everything the builders (`opt_ss`, `opt_dyn`, `defEstimator`) and the loop need."""
    ns: Dict[str, Any]
    nx: int; nxp: int; nu: int; ny: int; nd: int; npx: int; npy: int; npxp: int; npyp: int
    N: int; h: float; Nsim: int
    sym: Dict[str, SX]
    Fx_model: Function; Fy_model: Function
    Fx_p: Function; Fy_p: Function
    Fss_obj: Optional[Function]; F_obj: Optional[Function]; Vfin: Optional[Function]
    flags: Dict[str, Any]
    bounds_ss: Dict[str, Optional[np.ndarray]]
    bounds_dyn: Dict[str, Optional[np.ndarray]]
    sol_optss: Dict[str, Any]; sol_optdyn: Dict[str, Any]
    estimator: Dict[str, Any]
    x0_p: np.ndarray; x0_m: np.ndarray; u0: np.ndarray; dhat0: np.ndarray
    defSP: Optional[Callable] = None
    R_wn: Optional[np.ndarray] = None
    extra: Dict[str, Any] = field(default_factory=dict)

    @property
    def nw(self): return self.nx * (self.N + 1) + self.nu * self.N
    @property
    def npar(self): return 2 * (self.nx + self.nu) + self.nd + 1 + self.ny * self.nu + (self.npx + self.npy) * self.N
    @property
    def npar_ss(self): return 2 * self.nu + self.ny + self.nd + self.ny * self.nu + self.nx + 1 + self.npx + self.npy
    @property
    def nxi(self): return self.nx + (self.nd if self.flags["offree"] != "no" else 0)


def build_problem(ns: Dict[str, Any]) -> MpcProblem:
    """Apply the reference driver's setup rules to a namespace from `loader.load_example`."""
    has = lambda name: name in ns and ns[name] is not None  # noqa: E731  ('X' in locals())
    x, xp, u, y, d = ns["x"], ns["xp"], ns["u"], ns["y"], ns["d"]
    nx, nxp, nu, ny, nd = x.size1(), xp.size1(), u.size1(), y.size1(), d.size1()
    LinPar = ns["LinPar"]
    if LinPar is False:  # MPC_code.py:36-44
        npx, npy = ns["px"].size1(), ns["py"].size1()
        if ns["Fp_nominal"] is True:
            npxp, npyp = npx, npy
        else:
            npxp, npyp = ns["pxp"].size1(), ns["pyp"].size1()
    else:  # :45-48
        npx, npxp, npy, npyp = nx, nxp, ny, ny
    N, h = int(ns["N"]), float(ns["h"])
    if ns.get("Adaptation") is True or ns.get("Collocation") is True:
        raise NotImplementedError("Adaptation / Collocation are outside the accelerated path")
    if ns.get("slacks") is True:
        raise NotImplementedError("soft-constraint slacks are outside the accelerated path (no shipped example enables them)")

    # fixed symbols (:65-81)
    k = SX.sym("k", 1); t = SX.sym("t", 1)
    px = ns["px"] if LinPar is False else SX.sym("px", npx)
    py = ns["py"] if LinPar is False else SX.sym("py", npy)
    pxp = SX.sym("pxp", npxp); pyp = SX.sym("pyp", npyp)
    pxmp = SX.sym("pxmp", npxp); pymp = SX.sym("pymp", npyp)
    xs = SX.sym("xs", nx); us = SX.sym("us", nu); ys = SX.sym("ys", ny)
    xsp = SX.sym("xsp", nx); usp = SX.sym("usp", nu); ysp = SX.sym("ysp", ny)
    sym = dict(x=x, xp=xp, u=u, y=y, d=d, k=k, t=t, px=px, py=py, pxp=pxp, pyp=pyp, pxmp=pxmp, pymp=pymp,
               xs=xs, us=us, ys=ys, xsp=xsp, usp=usp, ysp=ysp)

    offree, SF = ns["offree"], ns["StateFeedback"]
    dist = dict(Bd=ns["Bd"], Cd=ns["Cd"]) if offree == "lin" else {}

    # model ladder (:93-167) - the branches differ only in which keywords are forwarded
    kw: Dict[str, Any] = dict(dist)
    if ns.get("ssjacid") is True:                                   # (:84-91) linearise the nonlinear model at a steady state
        from .ss_jac_id import linear_model_from_ssjacid
        Fx_model, Fy_model = linear_model_from_ssjacid(ns, x, u, y, d, k, t, px, py)
    elif "User_fxm_Cont" in ns:
        kw.update(fx=ns["User_fxm_Cont"], Mx=ns["Mx"])
        if SF is True: kw.update(SF=SF)
        elif "User_fym" in ns: kw.update(fy=ns["User_fym"])
        else: kw.update(C=ns["C"])
    elif "User_fxm_Dis" in ns:
        kw.update(Fx=ns["User_fxm_Dis"])
        if SF is True: kw.update(SF=SF)
        elif "User_fym" in ns: kw.update(fy=ns["User_fym"])
        else: kw.update(C=ns["C"])
    elif "A" in ns:
        kw.update(A=ns["A"], B=ns["B"])
        if SF is True:
            kw.update(SF=SF)
        elif "User_fym" in ns:
            kw.update(fy=ns["User_fym"])
            if "xlin" in ns: kw.update(xlin=ns["xlin"], ulin=ns["ulin"])
        else:
            kw.update(C=ns["C"])
            if "ylin" in ns: kw.update(ylin=ns["ylin"])
            if "xlin" in ns: kw.update(xlin=ns["xlin"], ulin=ns["ulin"])
    else:
        raise ValueError("no model: define User_fxm_Cont, User_fxm_Dis or A/B")
    if ns.get("ssjacid") is not True:
        Fx_model, Fy_model = defF_model(x, u, y, d, k, t, px, py, offree, LinPar, **kw)

    # plant (:171-196)
    if ns["Fp_nominal"] is True:
        Fx_p, Fy_p = Fx_model, Fy_model
    else:
        pk: Dict[str, Any] = {}
        if "Ap" in ns:
            pk.update(Ap=ns["Ap"], Bp=ns["Bp"])
            if SF is True: pk.update(SF=SF)
            elif "User_fyp" in ns: pk.update(fyp=ns["User_fyp"])
            else: pk.update(Cp=ns["Cp"])
        elif "User_fxp_Dis" in ns:
            pk.update(Fx=ns["User_fxp_Dis"])
            if SF is True: pk.update(SF=SF)
            elif "User_fyp" in ns: pk.update(fy=ns["User_fyp"])
            elif "Cp" in ns: pk.update(Cp=ns["Cp"])
        elif "User_fxp_Cont" in ns:
            pk.update(fx=ns["User_fxp_Cont"], Mx=ns["Mx"])
            if SF is True: pk.update(SF=SF)
            elif "User_fyp" in ns: pk.update(fy=ns["User_fyp"])
            else: pk.update(Cp=ns["Cp"])
        else:
            raise ValueError("no plant: define Ap/Bp, User_fxp_Dis or User_fxp_Cont (or Fp_nominal)")
        if "fyp" in pk:  # the reference passes fyp= here, which defF_p silently ignores (:180 vs Utilities.py:93)
            pk["fy"] = pk.pop("fyp")
        Fx_p, Fy_p = defF_p(xp, u, y, k, t, pxp, pyp, pxmp, pymp, LinPar, **pk)

    flags = {n: ns[n] for n in ("QForm_ss", "DUssForm", "ContForm", "TermCons", "QForm", "DUForm", "DUFormEcon",
                                "offree", "StateFeedback", "Fp_nominal", "LinPar", "estimating")}
    flags["Sol_Hess_constss"], flags["Sol_Hess_constdyn"] = ns["Sol_Hess_constss"], ns["Sol_Hess_constdyn"]
    Fss_obj = F_obj = Vfin = None
    lin_AC = "A" in ns and "C" in ns
    if ns["estimating"] is False:
        # target objective (:202-220)
        if "rss_y" in ns:
            if lin_AC: flags["Sol_Hess_constss"] = "yes"
            if "rss_u" in ns:
                Fss_obj = defFss_obj(x, u, y, xsp, usp, ysp, r_y=ns["rss_y"], r_u=ns["rss_u"])
            else:
                Fss_obj = defFss_obj(x, u, y, xsp, usp, ysp, r_y=ns["rss_y"], r_Du=ns["rss_Du"]); flags["DUssForm"] = True
        elif "Qss" in ns:
            flags["QForm_ss"] = True
            if lin_AC: flags["Sol_Hess_constss"] = "yes"
            if "Rss" in ns:
                Fss_obj = defFss_obj(x, u, y, xsp, usp, ysp, Q=ns["Qss"], R=ns["Rss"])
            else:
                Fss_obj = defFss_obj(x, u, y, xsp, usp, ysp, Q=ns["Qss"], S=ns["Sss"]); flags["DUssForm"] = True
        elif "User_fssobj" in ns:
            Fss_obj = defFss_obj(x, u, y, xsp, usp, ysp, f_obj=ns["User_fssobj"])
        # dynamic objective (:222-246)
        if "r_x" in ns:
            flags["QForm"] = True
            if lin_AC: flags["Sol_Hess_constdyn"] = "yes"
            if "r_u" in ns:
                F_obj = defF_obj(x, u, y, xs, us, ys, r_x=ns["r_x"], r_u=ns["r_u"])
            else:
                F_obj = defF_obj(x, u, y, xs, us, ys, r_x=ns["r_x"], r_Du=ns["r_Du"]); flags["DUForm"] = True
        elif "Q" in ns:
            flags["QForm"] = True
            if lin_AC: flags["Sol_Hess_constdyn"] = "yes"
            if "R" in ns:
                F_obj = defF_obj(x, u, y, xs, us, ys, Q=ns["Q"], R=ns["R"])
            else:
                F_obj = defF_obj(x, u, y, xs, us, ys, Q=ns["Q"], S=ns["S"]); flags["DUForm"] = True
        elif "User_fobj_Cont" in ns:
            flags["ContForm"] = True
            F_obj = defF_obj(x, u, y, xs, us, ys, f_Cont=ns["User_fobj_Cont"])
        elif "User_fobj_Dis" in ns:
            F_obj = defF_obj(x, u, y, xs, us, ys, f_Dis=ns["User_fobj_Dis"])
        # terminal cost (:248-257)
        if "User_vfin" in ns:
            Vfin = defVfin(x, xs, vfin_F=ns["User_vfin"])
        elif "A" in ns:
            if "Q" in ns:
                R_are = ns["S"] if "S" in ns else ns["R"]  # (:253-254) DARE uses S when given
                Vfin = defVfin(x, xs, A=ns["A"], B=ns["B"], Q=ns["Q"], R=R_are)
        else:
            Vfin = defVfin(x, xs)
        if Vfin is None:
            raise ValueError("no terminal cost can be formed (linear model without Q and without User_vfin)")

    itmax = ns["Sol_itmax"]
    sol_optss = {"ipopt.max_iter": itmax, "ipopt.hessian_constant": flags["Sol_Hess_constss"],
                 "ipopt.print_level": 0, "ipopt.sb": "yes", "print_time": 0}
    sol_optdyn = {"ipopt.max_iter": itmax, "ipopt.hessian_constant": flags["Sol_Hess_constdyn"],
                  "ipopt.print_level": 0, "ipopt.sb": "yes", "print_time": 0}

    def pick(base, special, n):  # (:291-293, :302-304)
        return _col(ns[base] if ns.get(special) is None else ns[special], n)
    bounds_ss = dict(xmin=pick("xmin", "xmin_ss", nx), xmax=pick("xmax", "xmax_ss", nx),
                     umin=pick("umin", "umin_ss", nu), umax=pick("umax", "umax_ss", nu),
                     ymin=pick("ymin", "ymin_ss", ny), ymax=pick("ymax", "ymax_ss", ny))
    bounds_dyn = dict(xmin=pick("xmin", "xmin_dyn", nx), xmax=pick("xmax", "xmax_dyn", nx),
                      umin=pick("umin", "umin_dyn", nu), umax=pick("umax", "umax_dyn", nu),
                      ymin=pick("ymin", "ymin_dyn", ny), ymax=pick("ymax", "ymax_dyn", ny),
                      Dumin=_col(ns["Dumin"], nu), Dumax=_col(ns["Dumax"], nu))

    # estimator selection (:577-650)
    nxi = nx + (nd if offree != "no" else 0)
    if offree == "no" and nd != 0:
        raise SystemExit("The disturbance dimension is not zero but no disturbance model has been selected")
    est: Dict[str, Any] = {}
    if ns["kalss"] is True or ns["lue"] is True:
        if ns["kalss"] is True:                                     # steady-state gain (MPC_code.py:339-363)
            from .estimator_setup import Kkalss
            have_A, have_C = "A" in ns, "C" in ns
            linmod = "full" if (have_A and have_C) else ("onlyA" if have_A else ("onlyC" if have_C else "no"))
            kwk = {}
            if offree == "lin": kwk.update(Bd=ns["Bd"], Cd=ns["Cd"])
            if have_A: kwk["A"] = ns["A"]
            else: kwk["Fx"] = Fx_model
            if have_C: kwk["C"] = ns["C"]
            else: kwk["Fy"] = Fy_model
            var = () if linmod == "full" else (x, u, k, d, t, h, px, py, ns.get("x_ss"), ns.get("u_ss"), ns.get("px_ss"), ns.get("py_ss"))
            if linmod != "full" and ("x_ss" not in ns or "u_ss" not in ns):
                raise SystemExit("kalss with a nonlinear model needs the linearisation point x_ss, u_ss")
            ns["K"] = Kkalss(ny, nd, nx, ns["Q_kf"], ns["R_kf"], offree, linmod, *var, **kwk)
        K = np.eye(nxi) if (SF is True and offree == "no") else np.asarray(ns["K"], dtype=float)
        est = dict(type="kalss", K=K.reshape(nxi, ny))
    elif ns["mhe"] is True:
        raise NotImplementedError("MHE is outside the accelerated path; use the EKF variant of the problem")
    else:
        if ns["kal"] is True:
            if "A" not in ns:
                raise SystemExit("You cannot use the kalman filter if the model you have chosen is not linear")
            etype = "kal"
        elif ns["ekf"] is True:
            etype = "ekf"
        else:
            raise ValueError("no estimator selected (kalss / lue / kal / ekf)")
        est = dict(type=etype, Q=np.asarray(ns["Q_kf"], dtype=float).reshape(nxi, nxi),
                   R=np.asarray(ns["R_kf"], dtype=float).reshape(ny, ny))
    P0 = np.asarray(ns["P0"], dtype=float).reshape(nxi, nxi) if "P0" in ns else np.zeros((nxi, nxi))
    est["P0"] = P0
    est["dmin"], est["dmax"] = _col(ns["dmin"], nd), _col(ns["dmax"], nd)

    dhat0 = _col(ns["dhat0"], nd) if "dhat0" in ns else np.zeros(nd)
    for name in ("def_px", "def_py", "def_pxp", "def_pyp", "def_pxmp", "def_pymp"):
        pass  # time-varying parameters are evaluated by the loop from ns directly
    return MpcProblem(
        ns=ns, nx=nx, nxp=nxp, nu=nu, ny=ny, nd=nd, npx=npx, npy=npy, npxp=npxp, npyp=npyp,
        N=N, h=h, Nsim=int(ns["Nsim"]), sym=sym, Fx_model=Fx_model, Fy_model=Fy_model, Fx_p=Fx_p, Fy_p=Fy_p,
        Fss_obj=Fss_obj, F_obj=F_obj, Vfin=Vfin, flags=flags, bounds_ss=bounds_ss, bounds_dyn=bounds_dyn,
        sol_optss=sol_optss, sol_optdyn=sol_optdyn, estimator=est,
        x0_p=_col(ns["x0_p"], nxp), x0_m=_col(ns["x0_m"], nx), u0=_col(ns["u0"], nu), dhat0=dhat0,
        defSP=ns.get("defSP"), R_wn=(np.asarray(ns["R_wn"], dtype=float) if "R_wn" in ns else None),
        extra=dict(G_wn=ns.get("G_wn"), Q_wn=ns.get("Q_wn")),
    )


def make_specs(prob: MpcProblem):
    """Build the target and OCP records with the argument lists of ``MPC_code.py:300`` and ``:331-335``."""
    from .control_calc import build_ocp_spec
    from .target_calc import build_target_spec
    if prob.flags["estimating"] is True:
        return None, None
    s, b = prob.sym, prob.bounds_ss
    ss = build_target_spec(prob.nx, prob.nu, prob.ny, prob.nd, prob.npx, prob.npy, prob.Fx_model, prob.Fy_model,
                           prob.Fss_obj, prob.flags["QForm_ss"], prob.flags["DUssForm"], prob.sol_optss,
                           prob.ns.get("User_g_ineq_SS"), prob.ns.get("User_h_eq_SS"),
                           umin=b["umin"], umax=b["umax"], w_s=None, z_s=None, ymin=b["ymin"], ymax=b["ymax"],
                           xmin=b["xmin"], xmax=b["xmax"], h=prob.h)
    b = prob.bounds_dyn
    extra = {}
    if "User_fobj_Cont" in prob.ns:
        extra = dict(fx=prob.ns["User_fxm_Cont"], xstat=s["xs"], ustat=s["us"])
    f = prob.flags
    if prob.ns.get("User_h_eq") is not None:
        raise NotImplementedError("User_h_eq is outside the accelerated path (User_g_ineq, User_g_ineq_SS and "
                                  "User_h_eq_SS are supported)")
    ocp = build_ocp_spec(s["x"], s["u"], s["y"], s["d"], s["t"], s["px"], s["py"], prob.nx, prob.nu, prob.ny, prob.nd,
                         prob.npx, prob.npy, 0, 0, prob.Fx_model, prob.Fy_model, prob.F_obj, prob.Vfin, prob.N,
                         f["QForm"], f["DUForm"], f["DUFormEcon"], f["ContForm"], f["TermCons"], False, True, True,
                         prob.nw, prob.sol_optdyn, prob.ns.get("User_g_ineq"), None, umin=b["umin"], umax=b["umax"], W=None, Z=None,
                         ymin=b["ymin"], ymax=b["ymax"], xmin=b["xmin"], xmax=b["xmax"], Dumin=b["Dumin"],
                         Dumax=b["Dumax"], h=prob.h, Ws=[], **extra)
    return ss, ocp
