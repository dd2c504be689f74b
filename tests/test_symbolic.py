"""Host-side tracer: CasADi-compatible semantics, AD correctness, C code generation."""
import os

import numpy as np
import pytest

from mpc_code_b200 import symbolic as S
from mpc_code_b200.codegen import CFunction, CModule
from mpc_code_b200.sx import SX, DM, Function, exp, fabs, gradient, hessian, if_else, jacobian, mtimes, simpleRK, sqrt, vertcat, reshape


def test_indexing_and_shapes_follow_casadi():
    x = SX.sym("x", 3)
    assert x.size1() == 3 and x.size2() == 1 and x[1].shape == (1, 1) and x[0:2].shape == (2, 1)
    M = SX.sym("M", 2, 3)
    assert M.T.shape == (3, 2) and M[:, 1].shape == (2, 1) and M[1, :].shape == (1, 3)
    v = reshape(M, 6, 1)                       # column-major like CasADi
    assert v[1]._as_scalar_expr() is M[1, 0]._as_scalar_expr() and v[2]._as_scalar_expr() is M[0, 1]._as_scalar_expr()
    z = SX.zeros(3, 1)
    z[1] = if_else(x[0] > 0, x[0], 0.0)        # in-place element assignment (Ex_NMPC_dis.py:75-77)
    F = Function("F", [x], [z])
    assert np.allclose(np.asarray(F(np.array([2.0, 0, 0]))).ravel(), [0, 2, 0])
    assert np.allclose(np.asarray(F(np.array([-2.0, 0, 0]))).ravel(), [0, 0, 0])


def test_numpy_interop_and_mtimes():
    x = SX.sym("x", 2)
    A = np.array([[1.0, 2.0], [3.0, 4.0]])
    y = mtimes(A, x) + np.array([1.0, -1.0])
    q = 0.5 * mtimes(x.T, mtimes(A, x))
    F = Function("F", [x], [y, q])
    yv, qv = F(np.array([1.0, 2.0]))
    assert np.allclose(np.asarray(yv).ravel(), [6.0, 10.0]) and np.isclose(float(qv), 0.5 * (1 + 4 + 6 + 16))
    assert (A * 2.0 @ np.ones(2)).shape == (2,)      # plain numpy unaffected
    assert isinstance(np.array([1.0, 2.0]) * x, SX)  # numpy on the left defers to SX


def _cstr_rhs(x, u, d):
    k = 7.2e10 * exp(-8750.0 / 350.0) * exp(-8750.0 * (1.0 / x[1] - 1.0 / 350.0)) * x[0]
    ar = np.pi * 0.219 ** 2
    return vertcat(d[1] * (1.0 - x[0]) / (ar * x[2]) - k,
                   d[1] * (350.0 - x[1]) / (ar * x[2]) + 5.0e4 / 239.0 * k + 2 * 54.936 / (0.219 * 239.0) * (u[0] - x[1]),
                   (d[1] - u[1]) / ar)


def test_jacobian_and_hessian_against_finite_differences():
    x, u, d = SX.sym("x", 3), SX.sym("u", 2), SX.sym("d", 2)
    f = _cstr_rhs(x, u, d)
    z = vertcat(x, u)
    lam = SX.sym("lam", 3)
    H, g = hessian(mtimes(lam.T, f), z)
    F = Function("F", [x, u, d, lam], [f, jacobian(f, z), g, H])
    rng = np.random.default_rng(0)
    x0 = np.array([0.87, 325.0, 0.65]); u0 = np.array([300.0, 0.1]); d0 = np.array([0.0, 0.1]); l0 = rng.standard_normal(3)
    f0, J0, g0, H0 = [np.asarray(v) for v in F(x0, u0, d0, l0)]
    z0 = np.concatenate([x0, u0])
    eps = 1e-6
    for j in range(5):
        dz = np.zeros(5); dz[j] = eps * max(1.0, abs(z0[j]))
        fp = np.asarray(F(z0[:3] + dz[:3], z0[3:] + dz[3:], d0, l0)[0]).ravel()
        fm = np.asarray(F(z0[:3] - dz[:3], z0[3:] - dz[3:], d0, l0)[0]).ravel()
        col = (fp - fm) / (2 * dz[j])
        assert np.allclose(col, J0[:, j], rtol=1e-6, atol=1e-8)
        gp = np.asarray(F(z0[:3] + dz[:3], z0[3:] + dz[3:], d0, l0)[2]).ravel()
        gm = np.asarray(F(z0[:3] - dz[:3], z0[3:] - dz[3:], d0, l0)[2]).ravel()
        assert np.allclose((gp - gm) / (2 * dz[j]), H0[:, j], rtol=1e-5, atol=1e-7)
    assert np.allclose(H0, H0.T) and np.allclose(g0.ravel(), J0.T @ l0)


def test_forward_and_reverse_agree_on_unrolled_rk4():
    xt, p = SX.sym("xt", 4), SX.sym("p", 4)       # [x; t], [u; d]
    rhs = vertcat(_cstr_rhs(xt[0:3], p[0:2], p[2:4]), SX(1.0))
    rk = simpleRK(Function("f", [xt, p], [rhs]), 3)
    h = SX.sym("h", 1)
    out = rk(xt, p, h)[0:3, :]
    Jf = [S.forward_derivative(out.elements(), {e.uid: S.ONE}) for e in xt.elements()[:3]]
    Jr = [S.reverse_gradient(o, xt.elements()[:3]) for o in out.elements()]
    vals = {e.uid: v for e, v in zip(xt.elements() + p.elements() + h.elements(), [0.87, 325.0, 0.65, 0.0, 300.0, 0.1, 0.0, 0.1, 0.2])}
    a = np.array([[S.evaluate([Jf[j][i]], vals)[0] for j in range(3)] for i in range(3)], dtype=float)
    b = np.array([[S.evaluate([Jr[i][j]], vals)[0] for j in range(3)] for i in range(3)], dtype=float)
    assert np.allclose(a, b, rtol=1e-12, atol=1e-14)


def test_generated_c_matches_symbolic_evaluation(tmp_path):
    x, u = SX.sym("x", 3), SX.sym("u", 2)
    e = vertcat(sqrt(x[0] ** 2 + 1.0) * fabs(u[0]), if_else(x[1] <= u[1], x[2] ** 3, -x[2]), exp(-x[0]) / (1.0 + x[1] ** 2))
    cf = CFunction("tfun", [("x", x), ("u", u)], [("e", e), ("J", jacobian(e, x))])
    mod = CModule("tmod", [cf], str(tmp_path))
    F = Function("F", [x, u], [e, jacobian(e, x)])
    rng = np.random.default_rng(3)
    for _ in range(5):
        xv, uv = rng.standard_normal(3), rng.standard_normal(2)
        ec, Jc = mod.tfun(xv, uv)
        es, Js = F(xv, uv)
        assert np.allclose(ec, np.asarray(es), rtol=1e-14, atol=1e-15) and np.allclose(Jc, np.asarray(Js), rtol=1e-14, atol=1e-15)
    assert cf.flops > 0 and "exp" in cf.counts


def test_dm_helpers():
    assert DM.zeros(3).shape == (3, 1) and np.all(np.isinf(DM.inf(2)))
    assert np.allclose(np.asarray(vertcat(np.array([1.0, 2.0]), 3.0)).ravel(), [1, 2, 3])
