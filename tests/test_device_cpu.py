"""The device code's arithmetic, executed on the CPU through tests/cpu_harness.cpp, against the oracle.

These run without a GPU: the kernel bodies are host/device functions, compiled here with g++.
The `-m gpu` tests repeat the comparisons through the C ABI on the real device.
"""
import ctypes
import os

import numpy as np

from conftest import GOLDEN

G = np.load(os.path.join(GOLDEN, "nmpc_oracle.npz"))
VP = ctypes.c_void_p


def _unpack(Hp, nz):
    H = np.zeros((nz, nz))
    for i in range(nz):
        for j in range(i + 1):
            H[i, j] = H[j, i] = Hp[i * (i + 1) // 2 + j]
    return H


def test_rk4_sensitivities_and_exact_hessian_match_symbolic_ad(nmpc):
    p, mod, H = nmpc.prob, nmpc.oracle, nmpc.harness
    n, m, N = p.nx, p.nu, p.N
    nz = n + m
    rng = np.random.default_rng(0)
    B = 2
    par = np.stack([nmpc.ocp_par(p.x0_m, p.x0_m, p.u0, np.array([0, 0.1])), nmpc.ocp_par(p.x0_m * 1.01, p.x0_m, p.u0, np.array([0.2, 0.12]))])
    w = np.zeros((B, p.nw))
    for b in range(B):
        for k in range(N + 1):
            w[b, nz * k:nz * k + n] = p.x0_m * (1 + 0.02 * rng.uniform(-1, 1, n))
        for k in range(N):
            w[b, nz * k + n:nz * (k + 1)] = p.u0 * (1 + 0.01 * rng.uniform(-1, 1, m))
    lam = rng.standard_normal((B, N, n))
    A = np.zeros((B, N, n * n)); Bm = np.zeros((B, N, n * m)); c = np.zeros((B, N, n)); Hh = np.zeros((B, N, nz * (nz + 1) // 2))
    H.h_stage_derivs.argtypes = [ctypes.c_int] + [VP] * 7
    H.h_stage_derivs(B, par.ctypes.data, w.ctypes.data, lam.ctypes.data, A.ctypes.data, Bm.ctypes.data, c.ctypes.data, Hh.ctypes.data)
    for b in range(B):
        for k in range(0, N, 7):
            X, U = w[b, nz * k:nz * k + n], w[b, nz * k + n:nz * (k + 1)]
            F, J, Ho = mod.orc_dyn_d(X, U, par[b], np.zeros(p.npx), lam[b, k])
            assert np.allclose(A[b, k].reshape(n, n, order="F"), J[:, :n], rtol=1e-12, atol=1e-13)
            assert np.allclose(Bm[b, k].reshape(n, m, order="F"), J[:, n:], rtol=1e-12, atol=1e-13)
            assert np.allclose(_unpack(Hh[b, k], nz), Ho, rtol=1e-11, atol=1e-12 * np.abs(Ho).max())
            assert np.allclose(c[b, k], F.ravel() - w[b, nz * (k + 1):nz * (k + 1) + n], rtol=0, atol=1e-12)


def test_riccati_ipm_reaches_the_dense_oracle_solution(nmpc):
    par, w0 = G["ocp_par"], np.tile(G["ocp_w0"], (G["ocp_par"].shape[0], 1))
    w, f, st, it, ticks = nmpc.harness_ocp(par, w0)
    assert np.array_equal(st, G["ocp_status"])
    assert np.array_equal(it, G["ocp_iters"])            # same algorithm, different linear algebra -> same path
    assert np.abs(w - G["ocp_w"]).max() < 1e-9           # north_star: input trajectory <= 1e-6 abs
    assert np.all(np.abs(f - G["ocp_f"]) <= 1e-10 * np.maximum(1.0, np.abs(G["ocp_f"])))   # cost <= 1e-8 rel
    # same active set: distance-to-bound pattern of the inputs
    n, m, N = nmpc.prob.nx, nmpc.prob.nu, nmpc.prob.N
    lb, ub = nmpc.ocp.w_lb, nmpc.ocp.w_ub
    act = lambda W: (np.abs(W - lb) < 1e-6) | (np.abs(W - ub) < 1e-6)  # noqa: E731
    assert np.array_equal(act(w)[:, n:], act(G["ocp_w"])[:, n:])


def test_target_solver_matches_oracle(nmpc):
    w, f, st, it = nmpc.harness_target(G["ss_par"], G["ss_w0"])
    assert np.array_equal(st, G["ss_status"]) and np.array_equal(it, G["ss_iters"])
    assert np.abs(w - G["ss_w"]).max() < 1e-10 and np.abs(f - G["ss_f"]).max() < 1e-12


def test_ekf_update_matches_oracle(nmpc):
    p, H = nmpc.prob, nmpc.harness
    H.h_estimate.argtypes = [ctypes.c_int, ctypes.c_int] + [VP] * 12 + [ctypes.c_int]
    xi = np.concatenate([p.x0_m, p.dhat0]); P = p.estimator["P0"].copy().ravel()
    Q, R = np.ascontiguousarray(p.estimator["Q"]), np.ascontiguousarray(p.estimator["R"])
    K = np.zeros(p.nxi * p.ny); dmin = p.estimator["dmin"].copy(); dmax = p.estimator["dmax"].copy()
    px, py, u = np.zeros(p.npx), np.zeros(p.npy), p.u0.copy()
    for k in range(3):
        y = np.ascontiguousarray(G["ekf_y"][k]); tt = np.array([k * p.h])
        H.h_estimate(1, 1, y.ctypes.data, u.ctypes.data, tt.ctypes.data, px.ctypes.data, py.ctypes.data, xi.ctypes.data,
                     P.ctypes.data, Q.ctypes.data, R.ctypes.data, K.ctypes.data, dmin.ctypes.data, dmax.ctypes.data, 1)
        # P0 = ones(5,5) is singular and P_corr = P - K C P cancels four digits: the reference's own
        # inv()-based update (Estimator.py:354-358) carries the same conditioning, hence 1e-8.
        assert np.abs(xi - G["ekf_xi"][k]).max() < 1e-9
        assert np.abs(P.reshape(p.nxi, p.nxi) - G["ekf_P"][k]).max() < 1e-8 * np.abs(G["ekf_P"][k]).max()


def test_infeasible_initial_output_gives_ipopt_status_2(nmpc):
    p = nmpc.prob
    xhat = p.x0_m.copy(); xhat[2] = 0.4995
    w, f, st, it, _ = nmpc.harness_ocp(nmpc.ocp_par(xhat, p.x0_m, p.u0, p.dhat0), nmpc.cold_guess())
    assert st[0] == 2 and it[0] == 0


def test_plant_and_model_steps_match_oracle(nmpc):
    p, mod, H = nmpc.prob, nmpc.oracle, nmpc.harness
    rng = np.random.default_rng(5)
    for t in (0.0, 6.0, 16.4, 30.0):                      # the plant's feed-rate profile switches at t = 5, 15, 25
        x = p.x0_p * (1 + 0.02 * rng.standard_normal(3)); u = p.u0 * (1 + 0.01 * rng.standard_normal(2)); z = np.zeros(3)
        ref = mod.orc_fxp(x, u, z, t, p.h, z).ravel()
        xx = x.copy(); tt = np.array([t])
        H.h_plant_step(1, xx.ctypes.data_as(VP), u.ctypes.data_as(VP), tt.ctypes.data_as(VP), z.ctypes.data_as(VP), z.ctypes.data_as(VP))
        assert np.abs(xx - ref).max() <= 1e-13 * np.abs(ref).max()
        d = np.array([0.0, 0.11]); xn = np.zeros(3)
        H.h_model_step(1, x.ctypes.data_as(VP), u.ctypes.data_as(VP), d.ctypes.data_as(VP), tt.ctypes.data_as(VP), z.ctypes.data_as(VP), xn.ctypes.data_as(VP))
        assert np.abs(xn - mod.orc_fx(x, u, p.h, d, t, z).ravel()).max() <= 1e-13 * np.abs(xn).max()


def test_bunch_kaufman_inertia_and_solve(nmpc):
    H = nmpc.harness
    rng = np.random.default_rng(1)
    for trial in range(100):
        M = rng.standard_normal((12, 12)); M = M + M.T
        if trial % 3 == 0:
            M[6:, 6:] = 0.0                                # saddle-point structure like a KKT matrix
        b = rng.standard_normal(12)
        A = np.asfortranarray(M.copy()); x = b.copy(); inert = np.zeros(3, dtype=np.int32)
        H.h_bk_test(12, A.ctypes.data_as(VP), x.ctypes.data_as(VP), inert.ctypes.data_as(VP))
        ev = np.linalg.eigvalsh(M)
        assert inert[0] == (ev > 0).sum() and inert[1] == (ev < 0).sum() and inert[2] == 0
        assert np.abs(M @ x - b).max() < 1e-10 * np.linalg.cond(M)


def test_closed_loop_of_device_code_matches_oracle_fixture(nmpc):
    """Eight closed-loop steps, three instances: loop glue + device arithmetic (on the CPU) vs the oracle."""
    from harness_loop import HarnessLoop
    Ns, B = G["cl_noise"].shape[0], G["cl_x0"].shape[0]
    rec = HarnessLoop(nmpc, B).run(Ns, G["cl_x0"], G["cl_noise"])
    assert np.array_equal(rec["STATUS_DYN"], G["cl_STATUS_DYN"]) and np.array_equal(rec["ITER_DYN"], G["cl_ITER_DYN"])
    for key in ("U", "X_HAT", "D_HAT", "XS", "US", "Xp", "Yp"):
        assert np.abs(rec[key] - G["cl_" + key]).max() < 1e-6, key
    assert np.all(np.abs(rec["F_DYN"] - G["cl_F_DYN"]) <= 1e-8 * np.maximum(1.0, np.abs(G["cl_F_DYN"])))


L = np.load(os.path.join(GOLDEN, "lmpc_oracle.npz"))


def _check_linear_loop(bundle, tag):
    from harness_loop import HarnessLoop
    Ns = L[tag + "_U"].shape[0]
    rec = HarnessLoop(bundle, 1).run(Ns)
    assert np.array_equal(rec["STATUS_DYN"][:, 0], L[tag + "_STATUS_DYN"])
    assert np.array_equal(rec["ITER_DYN"][:, 0], L[tag + "_ITER_DYN"])
    for key in ("U", "X_HAT", "D_HAT", "XS", "US", "Xp", "Yp"):
        assert np.abs(rec[key][:, 0, :] - L["%s_%s" % (tag, key)]).max() < 1e-6, key
    ok = L[tag + "_STATUS_DYN"] == 0
    assert np.all(np.abs(rec["F_DYN"][ok, 0] - L[tag + "_F_DYN"][ok]) <= 1e-8 * np.maximum(1.0, np.abs(L[tag + "_F_DYN"][ok])))


def test_lmpc_cstr_closed_loop_including_infeasible_first_steps(lmpc_cstr):
    """BASELINE configs[0].  The shipped initial state makes x_1 <= xmax unreachable: the first OCPs are infeasible
    and the loop keeps u = u0 (MPC_code.py:804-805) until the state has decayed."""
    assert list(L["cstr_STATUS_DYN"][:4]) == [2, 2, 2, 0]
    _check_linear_loop(lmpc_cstr, "cstr")


def test_lmpc_wb_closed_loop_with_delta_u_cost(lmpc_wb):
    """BASELINE configs[2]: Delta-u penalty (DUForm) - the device carries u_{k-1} as extra stage state."""
    assert lmpc_wb.ocp.uses_uprev and lmpc_wb.prob.flags["DUForm"] is True
    _check_linear_loop(lmpc_wb, "wb")


def test_enmpc_closed_loop_with_integrated_economic_cost(enmpc):
    """BASELINE configs[3] (EKF branch): ContForm - the stage cost is a quadrature state of the RK4 sweep; the economic
    cost is indefinite, so the delta_w inertia ladder of the Riccati sweep is exercised (first solve: 30 iterations)."""
    assert enmpc.prob.flags["ContForm"] is True and enmpc.ocp.cont_substeps == 10
    assert L["enmpc_ITER_DYN"][0] >= 20
    _check_linear_loop(enmpc, "enmpc")
