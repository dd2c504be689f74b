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

    def run(self, Nsim, x0, noise=None):
        b, B, H = self.b, self.B, self.H
        p = b.prob
        nx, nu, ny, nd, N = p.nx, p.nu, p.ny, p.nd, p.N
        nxu = nx + nu
        x_k = np.ascontiguousarray(x0, dtype=float).copy(); x0_m = x_k.copy()
        u_k = np.tile(p.u0, (B, 1)); xhat = x0_m.copy(); dhat = np.tile(p.dhat0, (B, 1))
        P = np.tile(p.estimator["P0"].reshape(1, -1), (B, 1))
        Q, R = np.ascontiguousarray(p.estimator["Q"]), np.ascontiguousarray(p.estimator["R"])
        Kz = np.zeros(p.nxi * ny); dmin, dmax = p.estimator["dmin"].copy(), p.estimator["dmax"].copy()
        zx, zy = np.zeros((B, p.npx)), np.zeros((B, p.npy)); zpx = np.zeros((B, p.npxp)); zpy = np.zeros((B, p.npyp))
        us_k, xs_k = u_k.copy(), x0_m.copy()
        w_opt = w_guess = None
        st_dyn = np.zeros(B, dtype=np.int32)
        rec = {k: [] for k in ("U", "X_HAT", "D_HAT", "XS", "US", "Xp", "Yp", "STATUS_DYN", "ITER_DYN", "F_DYN", "STATUS_SS")}
        ysp, usp, xsp = [np.tile(np.asarray(v, dtype=float), (B, 1)) for v in p.defSP(0.0)]
        for k in range(Nsim):
            t_k = k * p.h
            tt = np.full(B, t_k)
            rec["Xp"].append(x_k.copy()); rec["X_HAT"].append(xhat.copy())
            y = np.zeros((B, ny))
            H.h_plant_meas(B, _p(x_k), _p(u_k), _p(tt), _p(zpy), _p(zpy), _p(y))
            if noise is not None:
                y = y + noise[k]
            rec["Yp"].append(y.copy())
            xi = np.ascontiguousarray(np.hstack([xhat, dhat]))
            H.h_estimate(B, 1, _p(y), _p(u_k), _p(tt), _p(zx), _p(zy), _p(xi), _p(P), _p(Q), _p(R), _p(Kz), _p(dmin), _p(dmax), 1)
            xhat, dhat = np.ascontiguousarray(xi[:, :nx]), np.ascontiguousarray(xi[:, nx:])
            rec["D_HAT"].append(dhat.copy())
            us_prev, xs_prev = us_k.copy(), xs_k.copy()
            par_ss = np.hstack([usp, ysp, xsp, dhat, us_prev, np.zeros((B, ny * nu)), tt[:, None], zx, zy])
            y0 = np.stack([b.oracle.orc_fy(x0_m[i], p.u0, dhat[i], t_k, np.zeros(p.npy)).ravel() for i in range(B)])
            wss, fss, st_ss, it_ss = b.harness_target(par_ss, np.hstack([x0_m, np.tile(p.u0, (B, 1)), y0]))
            ok = (st_ss != 2)[:, None]
            xs_k = np.where(ok, wss[:, :nx], xs_k); us_k = np.where(ok, wss[:, nx:nxu], us_k)
            rec["XS"].append(xs_k.copy()); rec["US"].append(us_k.copy()); rec["STATUS_SS"].append(st_ss.copy())
            if k == 0:
                w_guess = np.hstack([np.tile(np.hstack([x0_m, np.tile(p.u0, (B, 1))]), (1, N)), x0_m])
            else:
                shifted = np.hstack([w_opt[:, nxu:], us_prev, xs_prev])
                w_guess = np.where((st_dyn == 2)[:, None], w_guess, shifted)
            par = np.hstack([xhat, xs_k, us_k, dhat, u_k, tt[:, None], np.zeros((B, ny * nu)), np.zeros((B, (p.npx + p.npy) * N))])
            w_new, f, st_dyn, it, _ = b.harness_ocp(par, w_guess)
            okd = (st_dyn != 2)[:, None]
            w_opt = w_new if w_opt is None else np.where(okd, w_new, w_opt)
            xpred = np.zeros((B, nx))
            H.h_model_step(B, _p(xhat), _p(u_k), _p(dhat), _p(tt), _p(zx), _p(xpred))
            xhat = np.ascontiguousarray(np.where(okd, w_new[:, nxu:nxu + nx], xpred))
            u_k = np.ascontiguousarray(np.where(okd, w_new[:, nx:nxu], u_k))
            rec["U"].append(u_k.copy()); rec["STATUS_DYN"].append(st_dyn.copy()); rec["ITER_DYN"].append(it.copy()); rec["F_DYN"].append(f.copy())
            H.h_plant_step(B, _p(x_k), _p(u_k), _p(tt), _p(zpx), _p(zpx))
        return {k: np.array(v) for k, v in rec.items()}
