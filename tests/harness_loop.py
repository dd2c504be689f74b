"""TEST INFRASTRUCTURE: the closed loop of mpc_loop.BatchedMpc driven through the CPU harness
(tests/cpu_harness.cpp) instead of the CUDA library, so the loop glue + device arithmetic can be
checked against the oracle without a GPU.  Mirrors MPC_code.py:485-875 like the other two loops."""
import ctypes

import numpy as np

VP = ctypes.c_void_p


def _p(a):
    return a.ctypes.data_as(VP)


class HarnessLoop:
    def __init__(self, bundle, B):
        self.b, self.B = bundle, B
        self.H = bundle.harness
        self.H.h_estimate.argtypes = [ctypes.c_int, ctypes.c_int] + [VP] * 12 + [ctypes.c_int]

    def _params(self, t_k):
        p, ns = self.b.prob, self.b.prob.ns
        get = lambda name, n: np.asarray(ns[name](t_k)[0], dtype=float).ravel() if name in ns else np.zeros(n)  # noqa: E731
        p_xk = np.zeros((p.npx, p.N)); p_yk = np.zeros((p.npy, p.N))
        for i in range(p.N):
            if "def_px" in ns: p_xk[:, i] = np.asarray(ns["def_px"](t_k + i)[0], dtype=float).ravel()
            if "def_py" in ns: p_yk[:, i] = np.asarray(ns["def_py"](t_k + i)[0], dtype=float).ravel()
        p_xmp = get("def_pxmp", p.npxp) if "def_pxmp" in ns else (p_xk[:, 0].copy() if "def_px" in ns else np.zeros(p.npxp))
        p_ymp = get("def_pymp", p.npyp) if "def_pymp" in ns else (p_yk[:, 0].copy() if "def_py" in ns else np.zeros(p.npyp))
        return p_xk, p_yk, p_xmp, p_ymp, get("def_pxp", p.npxp), get("def_pyp", p.npyp)

    def run(self, Nsim, x0=None, noise=None, x0_m=None):
        b, B, H = self.b, self.B, self.H
        p = b.prob
        nx, nu, ny, nd, N = p.nx, p.nu, p.ny, p.nd, p.N
        nxu = nx + nu
        x_k = np.ascontiguousarray(np.tile(p.x0_p, (B, 1)) if x0 is None else x0, dtype=float).copy()
        x0_m = np.ascontiguousarray(np.tile(p.x0_m, (B, 1)) if x0 is None else (x0 if x0_m is None else x0_m), dtype=float).copy()
        u_k = np.tile(p.u0, (B, 1)); xhat = x0_m.copy(); dhat = np.tile(p.dhat0, (B, 1))
        est = p.estimator
        est_type = 0 if est["type"] == "kalss" else 1
        P = np.tile(est["P0"].reshape(1, -1), (B, 1))
        Q = np.ascontiguousarray(est.get("Q", np.zeros((p.nxi, p.nxi)))); R = np.ascontiguousarray(est.get("R", np.zeros((ny, ny))))
        Kz = np.ascontiguousarray(est.get("K", np.zeros((p.nxi, ny))), dtype=float).reshape(-1)
        has_db = est["dmin"] is not None
        dmin = est["dmin"].copy() if has_db else np.zeros(max(nd, 1)); dmax = est["dmax"].copy() if has_db else np.zeros(max(nd, 1))
        us_k, xs_k = u_k.copy(), x0_m.copy()
        w_opt = w_guess = None
        st_dyn = np.zeros(B, dtype=np.int32)
        rec = {k: [] for k in ("U", "X_HAT", "D_HAT", "XS", "US", "Xp", "Yp", "STATUS_DYN", "ITER_DYN", "F_DYN", "STATUS_SS")}
        rows = lambda v: np.ascontiguousarray(np.tile(np.asarray(v, dtype=float).ravel(), (B, 1)))  # noqa: E731
        for k in range(Nsim):
            t_k = k * p.h
            tt = np.full(B, t_k)
            p_xk, p_yk, p_xmp, p_ymp, p_xp, p_yp = self._params(t_k)
            zx, zy = rows(p_xk[:, 0]), rows(p_yk[:, 0])
            rec["Xp"].append(x_k.copy()); rec["X_HAT"].append(xhat.copy())
            y = np.zeros((B, ny))
            H.h_plant_meas(B, _p(x_k), _p(u_k), _p(tt), _p(rows(p_yp)), _p(rows(p_ymp)), _p(y))
            if noise is not None:
                y = y + noise[k]
            rec["Yp"].append(y.copy())
            xi = np.ascontiguousarray(np.hstack([xhat, dhat]))
            H.h_estimate(B, est_type, _p(y), _p(u_k), _p(tt), _p(zx), _p(zy), _p(xi), _p(P), _p(Q), _p(R), _p(Kz), _p(dmin), _p(dmax),
                         1 if has_db else 0)
            xhat, dhat = np.ascontiguousarray(xi[:, :nx]), np.ascontiguousarray(xi[:, nx:])
            rec["D_HAT"].append(dhat.copy())
            ysp, usp, xsp = [rows(v) for v in (p.defSP(t_k) if p.defSP is not None else (np.zeros(ny), np.zeros(nu), np.zeros(nx)))]
            us_prev, xs_prev = us_k.copy(), xs_k.copy()
            par_ss = np.hstack([usp, ysp, xsp, dhat, us_prev, np.zeros((B, ny * nu)), tt[:, None], zx, zy])
            y0 = np.stack([b.oracle.orc_fy(x0_m[i], p.u0, dhat[i], t_k, p_yk[:, 0]).ravel() for i in range(B)])
            wss, fss, st_ss, it_ss = b.harness_target(par_ss, np.hstack([x0_m, np.tile(p.u0, (B, 1)), y0]))
            ok = (st_ss != 2)[:, None]
            xs_k = np.where(ok, wss[:, :nx], xs_k); us_k = np.where(ok, wss[:, nx:nxu], us_k)
            rec["XS"].append(xs_k.copy()); rec["US"].append(us_k.copy()); rec["STATUS_SS"].append(st_ss.copy())
            if k == 0:
                w_guess = np.hstack([np.tile(np.hstack([x0_m, np.tile(p.u0, (B, 1))]), (1, N)), x0_m])
            else:
                shifted = np.hstack([w_opt[:, nxu:], us_prev, xs_prev])
                w_guess = np.where((st_dyn == 2)[:, None], w_guess, shifted)
            par = np.hstack([xhat, xs_k, us_k, dhat, u_k, tt[:, None], np.zeros((B, ny * nu)),
                             rows(p_xk.reshape(-1, order="F")), rows(p_yk.reshape(-1, order="F"))])
            w_new, f, st_dyn, it, _ = b.harness_ocp(par, w_guess)
            okd = (st_dyn != 2)[:, None]
            w_opt = w_new if w_opt is None else np.where(okd, w_new, w_opt)
            xpred = np.zeros((B, nx))
            H.h_model_step(B, _p(xhat), _p(u_k), _p(dhat), _p(tt), _p(zx), _p(xpred))
            xhat = np.ascontiguousarray(np.where(okd, w_new[:, nxu:nxu + nx], xpred))
            u_k = np.ascontiguousarray(np.where(okd, w_new[:, nx:nxu], u_k))
            rec["U"].append(u_k.copy()); rec["STATUS_DYN"].append(st_dyn.copy()); rec["ITER_DYN"].append(it.copy()); rec["F_DYN"].append(f.copy())
            H.h_plant_step(B, _p(x_k), _p(u_k), _p(tt), _p(rows(p_xp)), _p(rows(p_xmp)))
        return {k: np.array(v) for k, v in rec.items()}
