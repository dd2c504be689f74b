"""Dense assembly of the reference's two NLPs (TEST INFRASTRUCTURE - see ``oracle/__init__.py``).

`OcpNlp` follows ``Control_Calc.py:20-260`` literally: ``w = [x0,u0,x1,...,u_{N-1},x_N]``,
constraint vector ``[x0 - X0 ; Fx(Xk,Uk) - X_{k+1} (k=0..N-1) ; (X_N - xs) ; Y_k rows ; DU_k rows]``
with the bounds ``g_lb/g_ub`` of ``:224-243``; `TargetNlp` follows ``Target_Calc.py:20-161``.
Derivatives come from the generated C of `oracle.cmodel` (symbolic AD of the unrolled stage).
"""
from __future__ import annotations

import numpy as np

from .ipm import IpmOptions, IpmResult, solve_nlp


class OcpNlp:
    def __init__(self, spec, cmod):
        self.s, self.c = spec, cmod
        s = spec
        self.nxu = s.n + s.m
        self.n_dyn = s.n * (s.N + 1) + (s.n if s.term_eq is not None else 0)
        self.n_y = 0 if s.yFree else s.p * s.N
        self.n_du = 0 if s.DuFree else s.m * s.N
        self.n_gin = getattr(s, "n_gin", 0)                  # user stage inequalities, rows after the DU rows (Control_Calc.py:254)
        self.m_total = self.n_dyn + self.n_y + self.n_du + self.n_gin * s.N
        assert self.m_total == s.g_lb.size

    def _slices(self, par, k):
        s = self.s
        pxk = par[s.off["px"] + k * s.npx: s.off["px"] + (k + 1) * s.npx]
        pyk = par[s.off["py"] + k * s.npy: s.off["py"] + (k + 1) * s.npy]
        return pxk, pyk

    def make_fun(self, par):
        s, c = self.s, self.c
        n, m, p, N, nxu = s.n, s.m, s.p, s.N, self.nxu
        par = np.asarray(par, dtype=float).reshape(-1)
        x0 = par[s.off["x0"]:s.off["x0"] + n]
        xs = par[s.off["xs"]:s.off["xs"] + n]
        um1 = par[s.off["um1"]:s.off["um1"] + m]
        oy, odu = self.n_dyn, self.n_dyn + self.n_y
        ogi, ngi = self.n_dyn + self.n_y + self.n_du, self.n_gin

        def fun(w, lam, need):
            g = np.zeros(self.m_total)
            f = 0.0
            grad = np.zeros(s.nw) if need >= 1 else None
            J = np.zeros((self.m_total, s.nw)) if need >= 2 else None
            H = np.zeros((s.nw, s.nw)) if need >= 2 else None
            g[0:n] = x0 - w[0:n]
            if need >= 2:
                J[0:n, 0:n] = -np.eye(n)
            for k in range(N):
                X = w[nxu * k:nxu * k + n]; U = w[nxu * k + n:nxu * (k + 1)]
                Xn = w[nxu * (k + 1):nxu * (k + 1) + n]
                Up = um1 if k == 0 else w[nxu * (k - 1) + n:nxu * k]
                pxk, pyk = self._slices(par, k)
                iz = slice(nxu * k, nxu * (k + 1))
                rows = slice(n * (k + 1), n * (k + 2))
                if need >= 2:
                    F, Jd, Hd = c.orc_dyn_d(X, U, par, pxk, lam[rows])
                    J[rows, iz] = Jd
                    J[rows, nxu * (k + 1):nxu * (k + 1) + n] -= np.eye(n)
                    H[iz, iz] += Hd
                else:
                    F = c.orc_dyn(X, U, par, pxk)
                g[rows] = F.ravel() - Xn
                if self.n_y:
                    yr = slice(oy + p * k, oy + p * (k + 1))
                    if need >= 2:
                        Y, JY, HY = c.orc_out_d(X, U, par, pyk, lam[yr])
                        J[yr, iz] = JY
                        H[iz, iz] += HY
                    else:
                        Y = c.orc_out(X, U, par, pyk)
                    g[yr] = Y.ravel()
                if self.n_du:
                    dr = slice(odu + m * k, odu + m * (k + 1))
                    g[dr] = U - Up
                    if need >= 2:
                        J[dr, nxu * k + n:nxu * (k + 1)] = np.eye(m)
                        if k > 0:
                            J[dr, nxu * (k - 1) + n:nxu * k] = -np.eye(m)
                if ngi:
                    gr = slice(ogi + ngi * k, ogi + ngi * (k + 1))
                    if need >= 2:
                        Gv, JG, HG = c.orc_gin_d(X, U, par, pxk, pyk, lam[gr])
                        J[gr, iz] = JG
                        H[iz, iz] += HG
                    else:
                        Gv = c.orc_gin(X, U, par, pxk, pyk)
                    g[gr] = np.asarray(Gv).ravel()
                if need >= 1:
                    l, gc, Hc = c.orc_cost_d(X, U, Up, par, pxk, pyk)
                    f += float(l[0, 0])
                    gc = gc.ravel()
                    grad[iz] += gc[:nxu]
                    if k > 0:
                        ip = slice(nxu * (k - 1) + n, nxu * k)
                        grad[ip] += gc[nxu:]
                    if need >= 2:
                        H[iz, iz] += Hc[:nxu, :nxu]
                        if k > 0:
                            H[iz, ip] += Hc[:nxu, nxu:]
                            H[ip, iz] += Hc[nxu:, :nxu]
                            H[ip, ip] += Hc[nxu:, nxu:]
                else:
                    f += float(c.orc_cost(X, U, Up, par, pxk, pyk)[0, 0])
            XN = w[nxu * N:nxu * N + n]
            iN = slice(nxu * N, nxu * N + n)
            if s.term_eq is not None:
                rows = slice(n * (N + 1), n * (N + 2))
                g[rows] = XN - xs if s.flags["QForm"] else XN
                if need >= 2:
                    J[rows, iN] = np.eye(n)
            if need >= 1:
                V, gt, Ht = c.orc_term_d(XN, par)
                f += float(V[0, 0])
                grad[iN] += gt.ravel()
                if need >= 2:
                    H[iN, iN] += Ht
            else:
                f += float(c.orc_term(XN, par)[0, 0])
            out = dict(f=f, g=g)
            if need >= 1:
                out["grad"] = grad
            if need >= 2:
                out["J"], out["H"] = J, 0.5 * (H + H.T)
            return out
        return fun

    def solve(self, w_guess, par, w_lb=None, w_ub=None, g_lb=None, g_ub=None, opts: IpmOptions = None) -> IpmResult:
        s = self.s
        w_lb = s.w_lb if w_lb is None else w_lb
        w_ub = s.w_ub if w_ub is None else w_ub
        g_lb = s.g_lb if g_lb is None else g_lb
        g_ub = s.g_ub if g_ub is None else g_ub
        return solve_nlp(s.nw, self.m_total, self.make_fun(par), w_guess, w_lb, w_ub, g_lb, g_ub, opts)


class TargetNlp:
    def __init__(self, spec, cmod):
        self.s, self.c = spec, cmod

    def make_fun(self, par):
        s, c = self.s, self.c
        par = np.asarray(par, dtype=float).reshape(-1)

        def fun(w, lam, need):
            if need >= 2:
                con, J, Hc = c.orc_ss_con_d(w, par, lam)
                f, gf, Hf = c.orc_ss_cost_d(w, par)
                H = Hc + Hf
                return dict(f=float(f[0, 0]), g=con.ravel(), grad=gf.ravel(), J=J, H=0.5 * (H + H.T))
            con = c.orc_ss_con(w, par)
            if need >= 1:
                f, gf, _ = c.orc_ss_cost_d(w, par)
                return dict(f=float(f[0, 0]), g=con.ravel(), grad=gf.ravel())
            return dict(f=float(c.orc_ss_cost(w, par)[0, 0]), g=con.ravel())
        return fun

    def solve(self, w_guess, par, w_lb=None, w_ub=None, opts: IpmOptions = None) -> IpmResult:
        s = self.s
        w_lb = s.w_lb if w_lb is None else w_lb
        w_ub = s.w_ub if w_ub is None else w_ub
        return solve_nlp(s.nw, s.g_lb.size, self.make_fun(par), w_guess, w_lb, w_ub, s.g_lb, s.g_ub, opts)
