"""Dense primal-dual interior-point NLP solver (TEST INFRASTRUCTURE - see ``oracle/__init__.py``).

Stands in for ``nlpsol(..., 'ipopt', ...)`` (reference call sites ``Control_Calc.py:258``,
``Target_Calc.py:159``; options ``MPC_code.py:262-263``).  IPOPT itself is not vendored by the
reference and cannot be installed here, so this follows its published algorithm
(A. Waechter, L. T. Biegler, Math. Program. 106 (2006) 25-57, and the IPOPT option reference):

* problem  min f(x)  s.t.  gL <= g(x) <= gU,  xL <= x <= xU ; rows with gL == gU are equalities,
  the others get a slack ``g(x) - s = 0, gL <= s <= gU``; variables with xL == xU are removed
  (``fixed_variable_treatment = make_parameter``);
* all finite bounds relaxed by ``bound_relax_factor * max(1, |b|)`` (default 1e-8);
* starting point pushed inside by ``bound_push = bound_frac = 1e-2``; bound multipliers 1;
  constraint multipliers 0;
* monotone barrier update (mu_init 0.1, kappa_mu 0.2, theta_mu 1.5, kappa_eps 10, tau_min 0.99);
* scaled optimality error E_mu with s_max = 100; tol 1e-8 plus the absolute caps
  dual_inf_tol 1, constr_viol_tol 1e-4, compl_inf_tol 1e-4; "acceptable" exit after 15
  consecutive iterations at 1e-6;
* filter line search (gamma_theta 1e-5, gamma_phi 1e-8, eta_phi 1e-8, s_theta 1.1, s_phi 2.3,
  delta 1) with backtracking by 1/2; no second-order correction, no watchdog; restoration by proximal Gauss-Newton
  phase: a failed line search ends with ``Infeasible_Problem_Detected`` when the constraint
  violation is above ``constr_viol_tol`` (what IPOPT's restoration reports when it cannot reduce
  it) and with ``Restoration_Failed`` otherwise;
* inertia correction of the augmented system by the delta_w ladder (1e-4 first, x100 / x8 up,
  /3 down) using a dense symmetric-indefinite factorisation (LAPACK ``dsytrf``) for the inertia.

The linear algebra is deliberately *unstructured* (one dense KKT matrix): the device solver
exploits the stage structure with a Riccati recursion, and the two must agree on the answer.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Optional

import numpy as np
from scipy.linalg import lapack

STATUS = {0: "Solve_Succeeded", 1: "Solved_To_Acceptable_Level", 2: "Infeasible_Problem_Detected",
          -1: "Maximum_Iterations_Exceeded", -2: "Restoration_Failed", -3: "Error_In_Step_Computation",
          -13: "Invalid_Number_Detected"}


@dataclass
class IpmOptions:
    max_iter: int = 3000            # IPOPT default; the reference sets Sol_itmax (100 / 200)
    tol: float = 1e-8
    dual_inf_tol: float = 1.0
    constr_viol_tol: float = 1e-4
    compl_inf_tol: float = 1e-4
    acceptable_tol: float = 1e-6
    acceptable_iter: int = 15
    acceptable_dual_inf_tol: float = 1e10
    acceptable_constr_viol_tol: float = 1e-2
    acceptable_compl_inf_tol: float = 1e-2
    mu_init: float = 0.1
    kappa_mu: float = 0.2
    theta_mu: float = 1.5
    kappa_eps: float = 10.0
    tau_min: float = 0.99
    s_max: float = 100.0
    bound_push: float = 1e-2
    bound_frac: float = 1e-2
    bound_relax_factor: float = 1e-8
    honor_original_bounds: bool = False   # IPOPT >= 3.14 default 'no' (3.12: 'yes')
    kappa_sigma: float = 1e10
    gamma_theta: float = 1e-5
    gamma_phi: float = 1e-8
    eta_phi: float = 1e-8
    s_theta: float = 1.1
    s_phi: float = 2.3
    delta: float = 1.0
    gamma_alpha: float = 0.05
    delta_w0: float = 1e-4
    delta_w_min: float = 1e-20
    delta_w_max: float = 1e40
    kappa_w_plus_bar: float = 100.0
    kappa_w_plus: float = 8.0
    kappa_w_minus: float = 1.0 / 3.0
    delta_c_bar: float = 1e-8
    kappa_c: float = 0.25
    hessian_constant: bool = False
    # feasibility restoration (entered when the filter line search fails; see `solve_nlp`)
    resto_kappa: float = 0.9            # required reduction of the violation (IPOPT required_infeasibility_reduction)
    resto_max_calls: int = 3
    resto_armijo: float = 1e-4
    resto_alpha_min: float = 1e-5
    verbose: bool = False


@dataclass
class IpmResult:
    x: np.ndarray
    f: float
    status: int
    iters: int
    lam_g: np.ndarray
    info: Dict[str, float] = field(default_factory=dict)

    @property
    def return_status(self) -> str:
        return STATUS.get(self.status, "Internal_Error")


def _factor(K):
    ldu, ipiv, info = lapack.dsytrf(K, lower=1)
    n = K.shape[0]
    npos = nneg = nzero = 0
    i = 0
    while i < n:
        if ipiv[i] > 0:
            dval = ldu[i, i]
            if dval > 0: npos += 1
            elif dval < 0: nneg += 1
            else: nzero += 1
            i += 1
        else:  # 2x2 pivot block: one positive and one negative eigenvalue unless degenerate
            a, b, c = ldu[i, i], ldu[i + 1, i], ldu[i + 1, i + 1]
            ev = np.linalg.eigvalsh(np.array([[a, b], [b, c]]))
            for e in ev:
                if e > 0: npos += 1
                elif e < 0: nneg += 1
                else: nzero += 1
            i += 2
    if info > 0:
        nzero = max(nzero, 1)
    return ldu, ipiv, (npos, nneg, nzero)


def solve_nlp(n: int, m: int, fun: Callable, x0, xL, xU, gL, gU, opts: Optional[IpmOptions] = None) -> IpmResult:
    """Solve the NLP.  ``fun(x, lam, need)`` returns a dict with keys ``f, grad, g`` and, when
    ``need == 2``, ``J`` (m x n) and ``H`` (n x n Hessian of f + lam' g)."""
    o = opts or IpmOptions()
    x0 = np.asarray(x0, dtype=float).copy()
    xL = np.asarray(xL, dtype=float).copy(); xU = np.asarray(xU, dtype=float).copy()
    gL = np.asarray(gL, dtype=float).copy(); gU = np.asarray(gU, dtype=float).copy()
    fixed = xL == xU
    free = np.where(~fixed)[0]
    xfull = x0.copy()
    xfull[fixed] = xL[fixed]
    eq = np.where(gL == gU)[0]
    ineq = np.where(gL != gU)[0]
    # Equality rows that involve no free variable (the reference's `x0 - X[0]` row once X[0] is
    # fixed through its bounds, Control_Calc.py:126 + MPC_code.py:734) are constants: IPOPT keeps
    # them and regularises the singular row with delta_c; here they are checked and dropped.
    if eq.size and free.size:
        r0 = fun(xfull, np.zeros(m), 2)
        J0 = np.asarray(r0["J"], dtype=float)[np.ix_(eq, free)]
        const_rows = np.abs(J0).max(axis=1) == 0.0
        if const_rows.any():
            viol = np.abs(np.asarray(r0["g"], dtype=float)[eq[const_rows]] - gL[eq[const_rows]])
            if viol.size and viol.max() > 1e-8:
                return IpmResult(x=xfull.copy(), f=float(r0["f"]), status=2, iters=0, lam_g=np.zeros(m),
                                 info=dict(reason="constant equality row violated"))
            eq = eq[~const_rows]
    # Range rows that involve no free variable (the reference's Y_0 row when the output map does not
    # depend on u: its x is the fixed X[0], Control_Calc.py:128-151) cannot be influenced: if such a
    # row lies outside its relaxed bounds the NLP is infeasible - IPOPT ends in restoration with
    # Infeasible_Problem_Detected, which is the one status the reference loop acts on (MPC_code.py:786).
    if ineq.size and free.size:
        r0 = fun(xfull, np.zeros(m), 2)
        Ji = np.asarray(r0["J"], dtype=float)[np.ix_(ineq, free)]
        crow = np.abs(Ji).max(axis=1) == 0.0
        if crow.any():
            v = np.asarray(r0["g"], dtype=float)[ineq[crow]]
            lo_c = gL[ineq[crow]]; hi_c = gU[ineq[crow]]
            lo_c = np.where(np.isfinite(lo_c), lo_c - o.bound_relax_factor * np.maximum(1.0, np.abs(lo_c)), lo_c)
            hi_c = np.where(np.isfinite(hi_c), hi_c + o.bound_relax_factor * np.maximum(1.0, np.abs(hi_c)), hi_c)
            if np.any(v < lo_c - o.tol) or np.any(v > hi_c + o.tol):
                return IpmResult(x=xfull.copy(), f=float(r0["f"]), status=2, iters=0, lam_g=np.zeros(m),
                                 info=dict(reason="constant range row outside its bounds"))
    nf, me, mi = free.size, eq.size, ineq.size
    nv = nf + mi                       # primal variables: free x, then slacks
    mc = me + mi
    # bounds of the extended variable vector, relaxed
    lo = np.concatenate([xL[free], gL[ineq]]); hi = np.concatenate([xU[free], gU[ineq]])
    lo_orig, hi_orig = lo.copy(), hi.copy()
    hasL, hasU = np.isfinite(lo), np.isfinite(hi)
    lo = np.where(hasL, lo - o.bound_relax_factor * np.maximum(1.0, np.abs(lo)), lo)
    hi = np.where(hasU, hi + o.bound_relax_factor * np.maximum(1.0, np.abs(hi)), hi)
    nb = int(hasL.sum() + hasU.sum())

    def push(v):
        v = v.copy()
        both = hasL & hasU
        pL = np.where(hasL, np.minimum(o.bound_push * np.maximum(1.0, np.abs(lo)),
                                       np.where(both, o.bound_frac * (hi - lo), np.inf)), 0.0)
        pU = np.where(hasU, np.minimum(o.bound_push * np.maximum(1.0, np.abs(hi)),
                                       np.where(both, o.bound_frac * (hi - lo), np.inf)), 0.0)
        v = np.where(hasL, np.maximum(v, lo + pL), v)
        v = np.where(hasU, np.minimum(v, hi - pU), v)
        return v

    def evaluate(v, y, need):
        xfull[free] = v[:nf]
        lam_full = np.zeros(m)
        lam_full[eq] = y[:me]; lam_full[ineq] = y[me:]
        r = fun(xfull, lam_full, need)
        g = np.asarray(r["g"], dtype=float)
        C = np.concatenate([g[eq] - gL[eq], g[ineq] - v[nf:]])
        out = dict(f=float(r["f"]), C=C)
        if need >= 1:
            out["grad"] = np.concatenate([np.asarray(r["grad"], dtype=float)[free], np.zeros(mi)])
        if need >= 2:
            Jfull = np.asarray(r["J"], dtype=float)
            J = np.zeros((mc, nv))
            J[:me, :nf] = Jfull[np.ix_(eq, free)]
            J[me:, :nf] = Jfull[np.ix_(ineq, free)]
            J[me:, nf:] = -np.eye(mi)
            W = np.zeros((nv, nv))
            W[:nf, :nf] = np.asarray(r["H"], dtype=float)[np.ix_(free, free)]
            out["J"], out["W"] = J, W
        return out

    # starting point
    v = np.concatenate([xfull[free], np.zeros(mi)])
    v[:nf] = push(np.concatenate([v[:nf], np.zeros(mi)]))[:nf]
    xfull[free] = v[:nf]
    g0 = np.asarray(fun(xfull, np.zeros(m), 0)["g"], dtype=float)
    v[nf:] = g0[ineq]
    v = push(v)
    y = np.zeros(mc)
    zL = np.where(hasL, 1.0, 0.0); zU = np.where(hasU, 1.0, 0.0)
    mu = o.mu_init
    tau = max(o.tau_min, 1.0 - mu)
    filt = []
    delta_w_last = 0.0
    acceptable_count = 0
    theta0 = None
    status = -1
    it = 0
    resto, resto_calls, theta_entry = False, 0, 0.0
    ev = evaluate(v, y, 2)
    info: Dict[str, float] = {}

    def errors(ev, v, y, zL, zU, mu_):
        dL = np.where(hasL, v - lo, 1.0); dU = np.where(hasU, hi - v, 1.0)
        rd = ev["grad"] + ev["J"].T @ y - zL + zU
        dual = np.abs(rd).max() if nv else 0.0
        prim = np.abs(ev["C"]).max() if mc else 0.0
        cL = np.where(hasL, dL * zL - mu_, 0.0); cU = np.where(hasU, dU * zU - mu_, 0.0)
        comp = max(np.abs(cL).max() if nv else 0.0, np.abs(cU).max() if nv else 0.0)
        zsum = np.abs(zL).sum() + np.abs(zU).sum()
        s_d = max(o.s_max, (np.abs(y).sum() + zsum) / max(mc + nb, 1)) / o.s_max
        s_c = max(o.s_max, zsum / max(nb, 1)) / o.s_max
        return max(dual / s_d, prim, comp / s_c), dual, prim, comp

    while True:
        E0, dual0, prim0, comp0 = errors(ev, v, y, zL, zU, 0.0)
        info.update(E0=E0, dual_inf=dual0, constr_viol=prim0, compl=comp0, mu=mu)
        if not np.isfinite(E0):
            status = -13
            break
        if not resto and E0 <= o.tol and dual0 <= o.dual_inf_tol and prim0 <= o.constr_viol_tol and comp0 <= o.compl_inf_tol:
            status = 0
            break
        if resto:
            pass
        elif (E0 <= o.acceptable_tol and dual0 <= o.acceptable_dual_inf_tol and prim0 <= o.acceptable_constr_viol_tol
                and comp0 <= o.acceptable_compl_inf_tol):
            acceptable_count += 1
            if acceptable_count >= o.acceptable_iter:
                status = 1
                break
        else:
            acceptable_count = 0
        if it >= o.max_iter:
            status = -1
            break
        # barrier parameter update
        mu_min = o.tol / 10.0
        changed = False
        while not resto and mu > mu_min and errors(ev, v, y, zL, zU, mu)[0] <= o.kappa_eps * mu:
            mu = max(mu_min, min(o.kappa_mu * mu, mu ** o.theta_mu))
            tau = max(o.tau_min, 1.0 - mu)
            changed = True
        if changed:
            filt = []
        # Newton system
        dL = np.where(hasL, v - lo, 1.0); dU = np.where(hasU, hi - v, 1.0)
        sigma = np.where(hasL, zL / dL, 0.0) + np.where(hasU, zU / dU, 0.0)
        J, W = ev["J"], ev["W"]
        r1 = ev["grad"] + J.T @ y - np.where(hasL, mu / dL, 0.0) + np.where(hasU, mu / dU, 0.0)
        if resto:
            # Feasibility restoration step (proximal Gauss-Newton on the constraint violation): minimise
            #   zeta/2 |D_R dv|^2 - mu sum log(slacks to the bounds)   s.t.   C + J dv = 0,
            # zeta = sqrt(mu), D_R = diag(1 / max(1, |v|)) - IPOPT's restoration objective with the reference point reset
            # to the current iterate every iteration, the l1 penalty replaced by the linearised constraints.
            sigma = np.where(hasL, mu / dL ** 2, 0.0) + np.where(hasU, mu / dU ** 2, 0.0)
            W = np.diag(np.sqrt(mu) / np.maximum(1.0, np.abs(v)) ** 2)
            r1 = -np.where(hasL, mu / dL, 0.0) + np.where(hasU, mu / dU, 0.0)
        rhs = -np.concatenate([r1, ev["C"]])
        K0 = np.zeros((nv + mc, nv + mc))
        K0[:nv, :nv] = W + np.diag(sigma)
        K0[nv:, :nv] = J
        K0[:nv, nv:] = J.T
        delta_w, delta_c = 0.0, 0.0
        first = True
        sol = None
        while True:
            K = K0.copy()
            K[np.arange(nv), np.arange(nv)] += delta_w
            K[np.arange(nv, nv + mc), np.arange(nv, nv + mc)] -= delta_c
            ldu, ipiv, (npos, nneg, nzero) = _factor(K)
            if npos == nv and nneg == mc and nzero == 0:
                sol, sinfo = lapack.dsytrs(ldu, ipiv, rhs, lower=1)
                if np.all(np.isfinite(sol)):
                    break
            if nzero > 0:
                delta_c = o.delta_c_bar * mu ** o.kappa_c
            if first:
                delta_w = o.delta_w0 if delta_w_last == 0.0 else max(o.delta_w_min, o.kappa_w_minus * delta_w_last)
                first = False
            else:
                delta_w *= o.kappa_w_plus_bar if delta_w_last == 0.0 else o.kappa_w_plus
            if delta_w > o.delta_w_max:
                sol = None
                break
        if sol is None:
            status = -3
            break
        if delta_w > 0.0:
            delta_w_last = delta_w
        dv, dy = sol[:nv], sol[nv:]
        dzL = np.where(hasL, mu / dL - zL - zL / dL * dv, 0.0)
        dzU = np.where(hasU, mu / dU - zU + zU / dU * dv, 0.0)

        def ftb(val, dval, mask):
            neg = mask & (dval < 0)
            return min(1.0, float(np.min(-tau * val[neg] / dval[neg]))) if neg.any() else 1.0
        alpha_max = min(ftb(dL, dv, hasL), ftb(dU, -dv, hasU))
        alpha_z = min(ftb(zL, dzL, hasL), ftb(zU, dzU, hasU))

        # filter line search
        def barrier(f, vv):
            return f - mu * (np.log(vv - lo)[hasL].sum() + np.log(hi - vv)[hasU].sum())
        theta = np.abs(ev["C"]).sum()
        phi = barrier(ev["f"], v)
        if theta0 is None:
            theta0 = theta
        theta_min = 1e-4 * max(1.0, theta0); theta_max = 1e4 * max(1.0, theta0)
        gphi_d = float(ev["grad"] @ dv - mu * (dv / dL)[hasL].sum() + mu * (dv / dU)[hasU].sum())
        if gphi_d < 0 and theta <= theta_min:
            a_min = min(o.gamma_theta, o.gamma_phi * theta / (-gphi_d),
                        o.delta * theta ** o.s_theta / (-gphi_d) ** o.s_phi)
        elif gphi_d < 0:
            a_min = min(o.gamma_theta, o.gamma_phi * theta / (-gphi_d))
        else:
            a_min = o.gamma_theta
        a_min *= o.gamma_alpha
        alpha = alpha_max
        accepted = False
        ftype = False
        ev_t = None
        if resto:
            # restoration iteration: Armijo backtracking on the violation alone.  A step shorter than resto_alpha_min of
            # the Gauss-Newton step (jammed against the bounds, or no descent) means the violation cannot be reduced.
            while alpha > o.resto_alpha_min:
                vt = v + alpha * dv
                ev_t = evaluate(vt, y, 0)
                th_t = np.abs(ev_t["C"]).sum()
                if np.isfinite(th_t) and th_t <= (1.0 - o.resto_armijo * alpha) * theta:
                    accepted = True
                    break
                alpha *= 0.5
            if o.verbose:
                print("it %3d RESTO mu %.1e theta %.3e -> %.3e (entry %.3e) amax %.2e alpha %.2e %s" % (
                    it, mu, theta, th_t, theta_entry, alpha_max, alpha, "acc" if accepted else "REJ"))
            if not accepted:             # the violation cannot be reduced: a stationary point of the infeasibility
                status = 2 if theta > o.constr_viol_tol else -2
                break
            v = vt
            it += 1
            # multipliers are not iterated during restoration; they restart from zero / mu over the slack
            y = np.zeros(mc)
            dLn = np.where(hasL, v - lo, 1.0); dUn = np.where(hasU, hi - v, 1.0)
            zL = np.where(hasL, mu / dLn, 0.0); zU = np.where(hasU, mu / dUn, 0.0)
            ph_t = barrier(ev_t["f"], v)
            if th_t <= o.resto_kappa * theta_entry and not any(th_t >= tf_ and ph_t >= pf_ for tf_, pf_ in filt):
                resto = False            # sufficiently more feasible and acceptable to the filter: resume
            ev = evaluate(v, y, 2)
            continue
        while alpha >= a_min * (1 - 1e-12) and alpha > 1e-16:
            vt = v + alpha * dv
            ev_t = evaluate(vt, y, 0)
            th_t = np.abs(ev_t["C"]).sum()
            ph_t = barrier(ev_t["f"], vt)
            ok = np.isfinite(th_t) and np.isfinite(ph_t) and th_t <= theta_max
            if o.verbose and o.verbose > 1:
                print("      trial alpha %.3e theta_t %.9e (theta %.9e) phi_t %.9e (phi %.9e)" % (alpha, th_t, theta, ph_t, phi))
            if ok:
                for (tf_, pf_) in filt:
                    if th_t >= tf_ and ph_t >= pf_:
                        ok = False
                        break
            if ok:
                switching = gphi_d < 0 and theta <= theta_min and \
                    alpha * (-gphi_d) ** o.s_phi > o.delta * theta ** o.s_theta
                slack_eps = 10.0 * np.finfo(float).eps * abs(phi)
                if switching:
                    if ph_t - phi - slack_eps <= o.eta_phi * alpha * gphi_d:
                        accepted, ftype = True, True
                else:
                    if th_t <= (1 - o.gamma_theta) * theta or ph_t - slack_eps <= phi - o.gamma_phi * theta:
                        accepted = True
            if accepted:
                break
            alpha *= 0.5
        if o.verbose:
            print("it %3d mu %.1e E0 %.2e theta %.3e phi %.8e gphid %.2e amax %.2e az %.2e alpha %.2e dw %.1e %s nfilt %d a_min %.1e"
                  % (it, mu, E0, theta, phi, gphi_d, alpha_max, alpha_z, alpha, delta_w, "acc" if accepted else "REJ", len(filt), a_min))
        if not accepted:
            # IPOPT enters its restoration phase here: it looks for a point with a violation reduced to kappa_resto
            # times the current one that the filter (augmented by the current point) accepts, and reports
            # Infeasible_Problem_Detected when it ends at a stationary point of the violation instead.  Restated with a
            # simpler sub-solver (proximal Gauss-Newton steps, above) but the same entry, goal and exits.  Called at an
            # (almost) feasible point, or too often, it gives up as IPOPT does: Restoration_Failed.
            if theta < 1e-2 * o.constr_viol_tol or resto_calls >= o.resto_max_calls:
                status = 2 if theta > o.constr_viol_tol else -2
                break
            filt.append(((1 - o.gamma_theta) * theta, phi - o.gamma_phi * theta))
            resto, theta_entry = True, theta
            resto_calls += 1
            continue
        if not ftype:
            filt.append(((1 - o.gamma_theta) * theta, phi - o.gamma_phi * theta))
        v = v + alpha * dv
        y = y + alpha * dy
        zL = zL + alpha_z * dzL; zU = zU + alpha_z * dzU
        # keep bound multipliers within kappa_sigma of mu / slack
        dLn = np.where(hasL, v - lo, 1.0); dUn = np.where(hasU, hi - v, 1.0)
        zL = np.where(hasL, np.maximum(np.minimum(zL, o.kappa_sigma * mu / dLn), mu / (o.kappa_sigma * dLn)), 0.0)
        zU = np.where(hasU, np.maximum(np.minimum(zU, o.kappa_sigma * mu / dUn), mu / (o.kappa_sigma * dUn)), 0.0)
        it += 1
        ev = evaluate(v, y, 2)

    xfull[free] = v[:nf]
    if o.honor_original_bounds:
        xfull[free] = np.minimum(np.maximum(xfull[free], lo_orig[:nf]), hi_orig[:nf])
    lam_full = np.zeros(m)
    lam_full[eq] = y[:me]; lam_full[ineq] = y[me:]
    fval = float(fun(xfull, lam_full, 0)["f"])
    info.update(zL=zL, zU=zU, slack=v[nf:], lo=lo, hi=hi, free=free)
    return IpmResult(x=xfull.copy(), f=fval, status=status, iters=it, lam_g=lam_full, info=info)
