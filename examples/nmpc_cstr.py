"""Nonlinear MPC of a jacketed CSTR with level control, EKF with a nonlinear (feed-rate) disturbance.

Problem file in the CPCLAB-UNIPI/MPC-code user format (same names the driver probes for).  It states
the configuration BASELINE.json calls "Ex_NMPC": states (c, T, level), inputs (Tc, F), measured
(c, level), disturbance estimate d[1] = feed rate F0, horizon 50, h = 0.2 min, 10 RK4 sub-steps.
"""
import math
import numpy as np
import scipy.linalg as scla
from casadi import SX, exp, if_else, vertcat

Nsim, N, h, Mx = 201, 50, 0.2, 10

xp = SX.sym("xp", 3); x = SX.sym("x", 3); u = SX.sym("u", 2); y = SX.sym("y", 2); d = SX.sym("d", 2)

# physical constants of the reactor
_T0, _c0, _r, _k0, _EoR = 350.0, 1.0, 0.219, 7.2e10, 8750.0
_U0, _rho, _Cp, _DH = 915.6 * 60 / 1000, 1000.0, 0.239, -5.0e4
_area = math.pi * _r ** 2


def _cstr(xv, uv, F0):
    rate = _k0 * exp(-_EoR / _T0) * exp(-_EoR * (1.0 / xv[1] - 1.0 / _T0)) * xv[0]
    return vertcat(F0 * (_c0 - xv[0]) / (_area * xv[2]) - rate,
                   F0 * (_T0 - xv[1]) / (_area * xv[2]) - _DH / (_rho * _Cp) * rate
                   + 2 * _U0 / (_r * _rho * _Cp) * (uv[0] - xv[1]),
                   (F0 - uv[1]) / _area)


def User_fxp_Cont(x, t, u, pxp, pxmp):          # plant: the feed rate follows a step profile
    F0 = if_else(t <= 5, 0.1, if_else(t <= 15, 0.15, if_else(t <= 25, 0.08, 0.1)))
    return _cstr(x, u, F0)


def User_fyp(x, u, t, pyp, pymp):
    return vertcat(x[0], x[2])


def User_fxm_Cont(x, u, d, t, px):              # model: the feed rate is the estimated disturbance
    return _cstr(x, u, d[1])


def User_fym(x, u, d, t, py):
    return vertcat(x[0], x[2])


R_wn = 1e-7 * np.eye(2)                          # measurement noise covariance
offree = "nl"
x0_p = np.array([0.874317, 325, 0.6528]); x0_m = x0_p.copy()
u0 = np.array([300.157, 0.1]); dhat0 = np.array([0, 0.1])

ekf = True
Q_kf = scla.block_diag(1.0e-5 * np.eye(3), np.eye(2)); R_kf = 1.0e-4 * np.eye(2); P0 = np.ones((5, 5))


def defSP(t):
    return [np.array([0.874317, 0.6528]), np.array([300.157, 0.1]), np.zeros(3)]   # ysp, usp, xsp


umin = np.array([295, 0.00]); umax = np.array([305, 0.25])
xmin = np.array([0.0, 315, 0.50]); xmax = np.array([1.0, 375, 0.75])
ymin = np.array([0.0, 0.5]); ymax = np.array([1.0, 1.0])
dmin = -100 * np.ones((2, 1)); dmax = 100 * np.ones((2, 1))

Qss = np.diag([10.0, 1.0]); Rss = np.zeros((2, 2))
Q = np.eye(3); R = 0.1 * np.eye(2)
