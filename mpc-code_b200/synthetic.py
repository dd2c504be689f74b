"""Synthetic problem family for scaling sweeps (BASELINE.json configs[4]).

``synthetic_namespace(nx, nu, N)`` returns what executing a problem file would leave behind - the same names an
``Ex_*.py`` defines - for a random stable nonlinear plant::

    xdot = A_c x + B_c u + 0.1 tanh(W x),   A_c = -diag(U(0.2, 2)) + 0.1 G   (re-drawn until max Re(eig) < -0.05)

with h = 0.1, Mx = 4 RK4 sub-steps, Q = I, R = 0.1 I, |u| <= 1, outputs y = first ny states, EKF without disturbance
model, quadratic target.  Seeds are ``1000 nx + nu`` as in SURVEY.md.
"""
from __future__ import annotations

import numpy as np

from .defaults import default_namespace
from .sx import SX, mtimes, tanh


def synthetic_namespace(nx: int, nu: int, N: int, ny: int | None = None, Nsim: int = 20) -> dict:
    rng = np.random.default_rng(1000 * nx + nu)
    while True:
        Ac = -np.diag(rng.uniform(0.2, 2.0, nx)) + 0.1 * rng.standard_normal((nx, nx))
        if np.linalg.eigvals(Ac).real.max() < -0.05:
            break
    Bc = rng.standard_normal((nx, nu)) / np.sqrt(nx)
    W = rng.standard_normal((nx, nx)) / np.sqrt(nx)
    ny = min(nx, 2) if ny is None else ny
    C = np.eye(nx)[:ny]

    def rhs(x, u):
        return mtimes(Ac, x) + mtimes(Bc, u) + 0.1 * tanh(mtimes(W, x))

    ns = default_namespace()
    ns.update(
        Nsim=Nsim, N=N, h=0.1, Mx=4,
        xp=SX.sym("xp", nx), x=SX.sym("x", nx), u=SX.sym("u", nu), y=SX.sym("y", ny), d=SX.sym("d", 0),
        User_fxp_Cont=lambda x, t, u, pxp, pxmp: rhs(x, u),
        User_fxm_Cont=lambda x, u, d, t, px: rhs(x, u),
        C=C, Cp=C,
        offree="no",
        x0_p=np.zeros(nx), x0_m=np.zeros(nx), u0=np.zeros(nu),
        ekf=True, Q_kf=1e-4 * np.eye(nx), R_kf=1e-4 * np.eye(ny), P0=1e-2 * np.eye(nx),
        defSP=lambda t: [np.zeros(ny), np.zeros(nu), np.zeros(nx)],
        umin=-np.ones(nu), umax=np.ones(nu),
        Qss=np.eye(ny), Rss=1e-3 * np.eye(nu),
        Q=np.eye(nx), R=0.1 * np.eye(nu),
    )
    ns["_synthetic"] = dict(Ac=Ac, Bc=Bc, W=W, nx=nx, nu=nu, N=N)
    return ns
