"""Host-side set-up of the estimators: the steady-state Kalman gain ``Kkalss`` (``Estimator.py:103-229``).

The gain is computed once, before the loop (``MPC_code.py:339-363``), from the model linearised at a user-given
point ``(x_ss, u_ss, px_ss, py_ss)`` unless the model matrices ``A`` / ``C`` are given; the device then runs the
constant-gain update (``kalss``, ``Estimator.py:231-261``).  Same call signature and the same augmented-model
conventions as the reference, including its treatment of ``offree == 'nl'`` (the disturbance block of the augmented
``A`` is the identity and is NOT coupled to the states).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as scla

from .sx import DM, Function, jacobian


def _num(v, n):
    return np.zeros(n) if v is None else np.asarray(DM(v), dtype=float).reshape(n)


def Kkalss(ny, nd, nx, Q_kf, R_kf, offree, linmod, *var, **kwargs):
    A = np.asarray(DM(kwargs["A"]), dtype=float) if linmod in ("onlyA", "full") else None
    C = np.asarray(DM(kwargs["C"]), dtype=float) if linmod in ("onlyC", "full") else None
    if A is None or C is None:
        x, u, k, d, t, h, px, py, x_ss, u_ss, px_ss, py_ss = var
        xs_, us_ = _num(x_ss, x.numel()), _num(u_ss, u.numel())
        d0, pxs, pys = np.zeros(d.numel()), _num(px_ss, px.numel()), _num(py_ss, py.numel())
    if A is None:                                   # dFx/dx at (x_ss, u_ss, k = h, d = 0, t = 0, px_ss)   (:134-152)
        Fx = kwargs["Fx"]
        JA = Function("A_dm", [x, u, k, d, t, px], [jacobian(Fx(x, u, k, d, t, px), x)])
        A = np.asarray(JA(xs_, us_, float(h), d0, 0.0, pxs), dtype=float).reshape(nx, nx)
    if C is None:                                   # dFy/dx at (x_ss, u_ss, d = 0, t = 0, py_ss)          (:154-176)
        Fy = kwargs["Fy"]
        JC = Function("C_dm", [x, u, d, t, py], [jacobian(Fy(x, u, d, t, py), x)])
        C = np.asarray(JC(xs_, us_, d0, 0.0, pys), dtype=float).reshape(ny, nx)
    Aaug = np.eye(nx + nd); Caug = np.zeros((ny, nx + nd))              # (:178-199)
    if A.shape[1] < nx + nd or offree != "nl":
        Aaug[:nx, :nx] = A[:nx, :nx]
    else:
        Aaug = A
    if C.shape[1] < nx + nd or offree != "nl":
        Caug[:, :nx] = C[:, :nx]
    else:
        Caug = C
    if offree == "lin":
        Aaug[:nx, nx:] = np.asarray(DM(kwargs["Bd"]), dtype=float).reshape(nx, nd)
        Caug[:, nx:] = np.asarray(DM(kwargs["Cd"]), dtype=float).reshape(ny, nd)
    Ae, Be = Aaug.T, Caug.T                                             # (:201-216)
    Pe = scla.solve_discrete_are(Ae, Be, np.asarray(DM(Q_kf), dtype=float), np.asarray(DM(R_kf), dtype=float))
    Ke = Pe @ Be @ np.linalg.inv(Be.T @ Pe @ Be + np.asarray(DM(R_kf), dtype=float))
    return Ke
