"""One-instance closed loop on the CPU (TEST INFRASTRUCTURE - see ``oracle/__init__.py``).

Restates the loop body of the reference driver, ``MPC_code.py:485-875`` (branches taken by the
kal / ekf / kalss-lue estimators, no MHE, no adaptation, no collocation):
parameters ``:492-515`` -> model/plant output ``:524-541`` -> estimator ``:546-668`` -> target
``:690-731`` -> OCP warm start, solve, extraction ``:734-810`` -> plant step ``:813-827``.
Measurement noise is an explicit input instead of the unseeded draw at ``:540``.
"""
from __future__ import annotations

import numpy as np

from . import estimators
from .ipm import IpmOptions
from .nlp import OcpNlp, TargetNlp

INFEASIBLE = 2


class OracleLoop:
    def __init__(self, prob, ss_spec, ocp_spec, cmod, opts_ss: IpmOptions = None, opts_dyn: IpmOptions = None):
        self.prob, self.ss, self.ocp, self.c = prob, ss_spec, ocp_spec, cmod
        self.tn, self.on = TargetNlp(ss_spec, cmod), OcpNlp(ocp_spec, cmod)
        itmax = prob.sol_optss["ipopt.max_iter"]
        self.opts_ss = opts_ss or IpmOptions(max_iter=itmax)
        self.opts_dyn = opts_dyn or IpmOptions(max_iter=itmax)

    def _params(self, t_k):
        p, ns = self.prob, self.prob.ns
        p_xk = np.zeros((p.npx, p.N)); p_yk = np.zeros((p.npy, p.N))
        if "def_px" in ns:
            for i in range(p.N):
                p_xk[:, i] = np.asarray(ns["def_px"](t_k + i)[0], dtype=float).ravel()
        if "def_py" in ns:
            for i in range(p.N):
                p_yk[:, i] = np.asarray(ns["def_py"](t_k + i)[0], dtype=float).ravel()
        p_xmp = np.zeros(p.npxp); p_ymp = np.zeros(p.npyp)
        if "def_px" in ns:
            p_xmp = np.asarray(ns["def_pxmp"](t_k)[0], dtype=float).ravel() if "def_pxmp" in ns else p_xk[:, 0].copy()
        if "def_py" in ns:
            p_ymp = np.asarray(ns["def_pymp"](t_k)[0], dtype=float).ravel() if "def_pymp" in ns else p_yk[:, 0].copy()
        p_xp = np.asarray(ns["def_pxp"](t_k)[0], dtype=float).ravel() if "def_pxp" in ns else np.zeros(p.npxp)
        p_yp = np.asarray(ns["def_pyp"](t_k)[0], dtype=float).ravel() if "def_pyp" in ns else np.zeros(p.npyp)
        return p_xk, p_yk, p_xmp, p_ymp, p_xp, p_yp

    def run(self, Nsim=None, x0_p=None, x0_m=None, noise=None, state_noise=None, on_step=None):
        """``on_step(k)`` (optional) is called at the start of every step - bench.py's CPU leg uses it for time stamps."""
        p, c = self.prob, self.c
        nx, nu, ny, nd, N, h = p.nx, p.nu, p.ny, p.nd, p.N, p.h
        nxu = nx + nu
        Nsim = p.Nsim if Nsim is None else Nsim
        x_k = (p.x0_p if x0_p is None else np.asarray(x0_p, dtype=float)).copy()
        u_k = p.u0.copy()
        x0_m = p.x0_m if x0_m is None else np.asarray(x0_m, dtype=float)
        xhat_k = x0_m.copy()
        dhat_k = p.dhat0.copy()
        P_k = p.estimator["P0"].copy()
        est = p.estimator
        lam = np.zeros(ny * nu)
        nominal = p.flags["Fp_nominal"] is True
        offree = p.flags["offree"]
        out = {k: [] for k in ("Xp", "Yp", "U", "XS", "US", "YS", "X_HAT", "Y_HAT", "D_HAT", "F_DYN", "F_SS",
                               "STATUS_SS", "STATUS_DYN", "ITER_SS", "ITER_DYN", "W_OPT")}
        w_opt = None
        w_guess = None
        last_dyn_status = 0
        xs_k = us_k = None
        for ksim in range(Nsim):
            if on_step is not None:
                on_step(ksim)
            t_k = ksim * h
            p_xk, p_yk, p_xmp, p_ymp, p_xp, p_yp = self._params(t_k)
            p_x_k, p_y_k = p_xk[:, 0], p_yk[:, 0]
            out["Xp"].append(x_k.copy()); out["X_HAT"].append(xhat_k.copy())
            yhat_k = c.orc_fy(xhat_k, u_k, dhat_k, t_k, p_y_k).ravel()                       # :524
            if nominal:
                y_k = c.orc_fy(x_k, u_k, dhat_k, t_k, p_y_k).ravel()                         # :532
            else:
                y_k = c.orc_fyp(x_k, u_k, p_yp, t_k, p_ymp).ravel()                          # :534
            if noise is not None:
                y_k = y_k + noise[ksim]                                                      # :538-541
            out["Yp"].append(y_k.copy()); out["Y_HAT"].append(yhat_k.copy())
            # estimator (:546-668)
            xi = np.concatenate([xhat_k, dhat_k]) if offree != "no" else xhat_k.copy()
            if est["type"] == "kalss":
                xi = estimators.kalss(p, c, y_k, u_k, est["K"], xi, t_k, p_y_k)
            elif est["type"] == "kal":
                P_k, _, xi = estimators.kalman(p, c, y_k, u_k, est["Q"], est["R"], P_k, xi, t_k, p_y_k, p_x_k, h)
            else:
                P_k, _, xi = estimators.ekf(p, c, y_k, u_k, est["Q"], est["R"], P_k, xi, h, t_k, p_y_k, p_x_k)
            if offree != "no":
                xhat_k, dhat_k = xi[:nx].copy(), xi[nx:].copy()
                if est["dmin"] is not None:                                                  # :659-665
                    dhat_k = np.minimum(np.maximum(dhat_k, est["dmin"]), est["dmax"])
            else:
                xhat_k = xi
            out["D_HAT"].append(dhat_k.copy())
            if np.any(np.isnan(xhat_k)):
                raise FloatingPointError("xhat_k has NaN components (MPC_code.py:671-673)")
            if p.flags["estimating"] is False:
                if p.defSP is not None:
                    ysp_k, usp_k, xsp_k = [np.asarray(v, dtype=float).ravel() for v in p.defSP(t_k)]
                else:
                    ysp_k, usp_k, xsp_k = np.zeros(ny), np.zeros(nu), np.zeros(nx)
                if ksim == 0:
                    us_k, xs_k = u_k.copy(), x0_m.copy()                                     # :682-684
                us_prev, xs_prev = us_k.copy(), xs_k.copy()
                par_ss = np.concatenate([usp_k, ysp_k, xsp_k, dhat_k, us_prev, lam, [t_k], p_x_k, p_y_k])  # :693
                y0 = c.orc_fy(x0_m, p.u0, dhat_k, t_k, p_y_k).ravel()
                wss_guess = np.concatenate([x0_m, p.u0, y0])                                 # :696-700
                r_ss = self.tn.solve(wss_guess, par_ss, opts=self.opts_ss)
                if r_ss.status != INFEASIBLE:                                                # :714-718
                    xs_k, us_k = r_ss.x[:nx].copy(), r_ss.x[nx:nxu].copy()
                out["XS"].append(xs_k.copy()); out["US"].append(us_k.copy())
                out["YS"].append(c.orc_fy(xs_k, us_k, dhat_k, t_k, p_y_k).ravel())           # :730
                out["F_SS"].append(r_ss.f); out["STATUS_SS"].append(r_ss.status); out["ITER_SS"].append(r_ss.iters)
                w_lb, w_ub = self.ocp.w_lb.copy(), self.ocp.w_ub.copy()
                w_lb[:nx] = w_ub[:nx] = xhat_k                                               # :734
                if ksim == 0:                                                                # :740-756
                    w_guess = np.zeros(self.ocp.nw)
                    for key in range(1, N + 1):
                        w_guess[key * nxu - nu:key * nxu] = p.u0
                        w_guess[key * nxu:key * nxu + nx] = x0_m
                    w_guess[:nx] = x0_m
                elif last_dyn_status != INFEASIBLE:                                          # :763-764
                    w_guess = np.concatenate([w_opt[nxu:], us_prev, xs_prev])
                par = np.concatenate([xhat_k, xs_k, us_k, dhat_k, u_k, [t_k], lam,
                                      p_xk.reshape(-1, order="F"), p_yk.reshape(-1, order="F")])  # :769-772
                r = self.on.solve(w_guess, par, w_lb, w_ub, opts=self.opts_dyn)
                last_dyn_status = r.status
                if r.status != INFEASIBLE:                                                   # :786-800
                    w_opt = r.x.copy()
                    u_k = w_opt[nx:nxu].copy()
                    xhat_k = w_opt[nxu:nxu + nx].copy()
                else:                                                                        # :804-805
                    xhat_k = c.orc_fx(xhat_k, u_k, h, dhat_k, t_k, p_x_k).ravel()
                out["U"].append(u_k.copy()); out["F_DYN"].append(r.f)
                out["STATUS_DYN"].append(r.status); out["ITER_DYN"].append(r.iters)
                out["W_OPT"].append(r.x.copy())
            if nominal:
                x_k = c.orc_fx(x_k, u_k, h, dhat_k, t_k, p_xmp).ravel()                      # :814
            else:
                x_k = c.orc_fxp(x_k, u_k, p_xp, t_k, h, p_xmp).ravel()                       # :816
            if state_noise is not None:
                x_k = x_k + state_noise[ksim]                                                # :824-827
        return {k: np.array(v) for k, v in out.items()}
