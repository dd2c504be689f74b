"""Offset-free linear MPC of a NONLINEAR plant: a CSTR with level dynamics, linearised at an open-loop UNSTABLE point.

Problem file in the CPCLAB-UNIPI/MPC-code user format; same configuration as the reference's "Ex_LMPC_nlplant":
continuous nonlinear plant (RK4, 10 sub-steps per h = 0.2), linear model = Jacobians of that RK4 map at
(xlin, ulin), input disturbance model (Bd = B), Kalman filter, Q/S (delta-u) cost with DARE terminal weight,
N = 50.  The model has an eigenvalue of 2.15, which makes this the case that exposes any loss of symmetry in
the filter covariance recursion (tests/test_guide_curves.py).
"""
import math

import numpy as np
import scipy.linalg as scla
from casadi import SX, exp, vertcat

Nsim, N, h = 200, 50, 0.2

xp = SX.sym("xp", 3); x = SX.sym("x", 3); u = SX.sym("u", 2); y = SX.sym("y", 2); d = SX.sym("d", 2)


def User_fxp_Cont(x, t, u, pxp, pxmp):
    """x = (concentration, temperature, level); u = (coolant temperature, outlet flow)."""
    F0, T0, c0, r = 0.1, 350.0, 1.0, 0.219
    k0, EoR, U0 = 7.2e10, 8750.0, 915.6 * 60 / 1000
    rho, Cp2, DH = 1000.0, 0.239, -5.0e4
    area = math.pi * r ** 2
    rate = k0 * exp(-EoR / T0) * exp(-EoR * (1.0 / x[1] - 1.0 / T0)) * x[0]
    return vertcat(F0 * (c0 - x[0]) / (area * x[2]) - rate,
                   F0 * (T0 - x[1]) / (area * x[2]) - DH / (rho * Cp2) * rate + 2 * U0 / (r * rho * Cp2) * (u[0] - x[1]),
                   (F0 - u[1]) / area)


Mx = 10
Cp = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])

A = np.array([[0.51448, -0.00917517, -0.117995], [53.6817, 2.15004, -3.77725], [0.0, 0.0, 1.0]])
B = np.array([[-0.0017669, 0.0864569], [0.639423, 1.60696], [0.0, -1.32737]])
C = Cp.copy()
xlin = np.array([0.5, 350.0, 0.659]); ulin = np.array([300.0, 0.1])

offree = "lin"
Bd = B.copy(); Cd = np.zeros((2, 2))
x0_p = xlin.copy(); x0_m = xlin.copy(); u0 = ulin.copy()

kal = True
Q_kf = scla.block_diag(1.0e-5 * np.eye(3), np.eye(2)); R_kf = 1.0e-4 * np.eye(2); P0 = 1e-3 * Q_kf


def defSP(t):
    ysp = np.array([0.51, 0.659]) if 20 <= t < 40 else np.array([0.5, 0.659])
    return [ysp, np.array([299.963, 0.1]), np.zeros(3)]     # ysp, usp, xsp


umin = np.array([295.0, 0.0]); umax = np.array([305.0, 0.25])
xmin = np.array([0.0, 320.0, 0.45]); xmax = np.array([1.0, 375.0, 0.75])

Qss = np.diag([10.0, 0.01]); Rss = np.zeros((2, 2))
Q = np.diag([10.0, 1.0, 1.0]); S = 0.1 * np.eye(2)
