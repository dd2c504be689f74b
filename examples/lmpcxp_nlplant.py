"""Offset-free linear MPC of a nonlinear plant whose MODEL has one state more than the plant (nx = 4, nxp = 3).

Problem file in the CPCLAB-UNIPI/MPC-code user format; same configuration as the reference's "Ex_LMPCxp_nlplant":
the plant is the CSTR of `lmpc_nlplant.py` (continuous, RK4 with 10 sub-steps per h = 0.2); the model is its
linearisation at an open-loop unstable point augmented with a first-order lag of the coolant input (pole 0.01) that
feeds 0.001 of itself into the first output - so plant and model state vectors differ in length.  Input-disturbance
model (Bd = B), Kalman filter, output bounds, Q / S (Delta-u) cost with DARE terminal weight, N = 50.
"""
import math

import numpy as np
import scipy.linalg as scla
from casadi import SX, exp, vertcat

Nsim, N, h = 200, 50, 0.2

xp = SX.sym("xp", 3); x = SX.sym("x", 4); u = SX.sym("u", 2); y = SX.sym("y", 2); d = SX.sym("d", 2)


def User_fxp_Cont(x, t, u, pxp, pxmp):
    """x = (concentration, temperature, level); u = (coolant temperature, outlet flow)."""
    F0, T0, c0, r = 0.1, 350, 1.0, 0.219
    k0, EoR, U0 = 7.2e10, 8750, 915.6 * 60 / 1000
    rho, Cp2, DH = 1000.0, 0.239, -5.0e4
    pi = math.pi
    kT0 = k0 * exp(-EoR / T0)
    return vertcat(F0 * (c0 - x[0]) / (pi * r ** 2 * x[2]) - kT0 * exp(-EoR * (1.0 / x[1] - 1.0 / T0)) * x[0],
                   F0 * (T0 - x[1]) / (pi * (r ** 2) * x[2]) - DH / (rho * Cp2) * kT0 * exp(-EoR * (1.0 / x[1] - 1.0 / T0)) * x[0]
                   + 2 * U0 / (r * rho * Cp2) * (u[0] - x[1]),
                   (F0 - u[1]) / (pi * r ** 2))


Mx = 10
Cp = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])

_Alin = np.array([[0.51448, -0.00917517, -0.117995], [53.6817, 2.15004, -3.77725], [0.0, 0.0, 1]])
_Blin = np.array([[-0.0017669, 0.0864569], [0.639423, 1.60696], [0.0, -1.32737]])
_phi = 0.01                                                  # extra model state: x4+ = phi x4 + (1 - phi) u1
A = scla.block_diag(_Alin, _phi)
B = np.vstack([_Blin, np.array([[1.0 - _phi, 0.0]])])
C = np.column_stack([Cp, (_phi / 10.0) * np.array([[1.0], [0.0]])])
xlin = np.array([0.5, 350, 0.659, 0.0]); ulin = np.array([300, 0.1]); ylin = np.array([0.5, 0.659])

offree = "lin"
Bd = B; Cd = np.zeros((2, 2))
x0_p = np.array([0.5, 350, 0.659]); x0_m = np.array([0.5, 350, 0.659, 0.0]); u0 = np.array([300, 0.1])

kal = True
Q_kf = scla.block_diag(1.0e-2 * np.eye(4), np.eye(2)); R_kf = 1.0e-2 * np.eye(2); P0 = Q_kf


def defSP(t):
    ysp = np.array([0.5, 0.659]) if t < 20 else np.array([0.51, 0.659])
    return [ysp, np.array([300., 0.1]), np.zeros(4)]          # ysp, usp, xsp


umin = np.array([295, 0.00]); umax = np.array([305, 0.25])
xmin = np.array([0.0, 300, 0.45, -1.0]); xmax = np.array([1.0, 375, 0.75, 1.0])
ymin = np.array([0.0, 0.0]); ymax = np.array([1.0, 1.0])

Qss = np.eye(2); Rss = np.zeros((2, 2))
Q = np.diag([1.0, 1.0, 1.0, 0.1]); S = 0.10 * np.eye(2)
